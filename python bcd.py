#!/usr/bin/env python
"""Drop-in entry point with the reference's command line (README.md:6-21); see
lk-s-2022-estimacija-pokreta_b200/stage2.py."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.exit(importlib.import_module("lk-s-2022-estimacija-pokreta_b200.stage2").main(sys.argv))
