"""Drop-in for the reference's `edge` module (import edge); see lk-s-2022-estimacija-pokreta_b200/edges.py."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
_m = importlib.import_module("lk-s-2022-estimacija-pokreta_b200.edges")
canny_ivice = _m.canny_ivice
sed_ivice = _m.sed_ivice
