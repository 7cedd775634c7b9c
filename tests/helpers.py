"""Shared test helpers: golden fixture loading and the importable product package."""
import importlib
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PKG_NAME = "lk-s-2022-estimacija-pokreta_b200"


def pkg(sub=None):
    return importlib.import_module(PKG_NAME + ("." + sub if sub else ""))


def load_npz(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def load_case(name):
    """Restore the reference dtypes (int64 proposals / nprop / labels, float64 lcosts / flows)."""
    z = load_npz(name)
    out = {}
    for k, v in z.items():
        base = k.split("_", 1)[1] if k[:1] == "b" and k[1:2].isdigit() else k
        if base in ("proposals", "nprop", "labels00", "draws") or base.startswith("labels"):
            out[k] = v.astype(np.int64)
        elif base == "lcosts" or base.startswith("flow"):
            out[k] = v.astype(np.float64)
        else:
            out[k] = v
    return out


def oracle_params(meta, **kw):
    from oracle.proposals import Params
    H, W, cw, ch = (int(v) for v in meta[:4])
    return Params(H, W, cw, ch, **kw)
