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


def segment_test_field(rng, A, B, kind):
    """float32 (A,B,3) = (dx, dy, valid) fields that exercise removeSmallSegments (postprocessing.py:29-76):
    small islands on piecewise constant / smooth flow, invalid pixels with zeroed or with stale flow, pure noise."""
    f = np.zeros((A, B, 3), np.float32)
    if kind == "blocks":
        ba, bb = int(rng.integers(3, 9)), int(rng.integers(3, 9))
        base = rng.integers(-20, 21, size=((A + ba - 1) // ba, (B + bb - 1) // bb, 2)) * 3
        f[..., :2] = np.repeat(np.repeat(base, ba, 0), bb, 1)[:A, :B] + rng.integers(-1, 2, size=(A, B, 2))
        f[..., 2] = 1
        f[rng.random((A, B)) < 0.08] = 0
    elif kind == "noise":
        f[..., :2] = rng.normal(0, 6, size=(A, B, 2))
        f[..., 2] = rng.random((A, B)) < 0.9
    elif kind == "smooth":
        a = np.arange(A)[:, None]
        b = np.arange(B)[None, :]
        f[..., 0] = 5 * np.sin(a / 7.0) + 0.3 * b
        f[..., 1] = 4 * np.cos(b / 5.0)
        f[..., 2] = 1
        for _ in range(int(rng.integers(3, 12)) * max(1, A * B // 2000)):
            a0, b0 = int(rng.integers(0, max(1, A - 3))), int(rng.integers(0, max(1, B - 3)))
            h, w = int(rng.integers(1, 6)), int(rng.integers(1, 6))
            f[a0:a0 + h, b0:b0 + w, :2] += rng.integers(15, 40)
        f[rng.random((A, B)) < 0.03] = 0
    elif kind == "stale":
        f[..., :2] = rng.integers(-4, 5, size=(A, B, 2)) * 2
        f[..., 2] = rng.random((A, B)) < 0.7      # invalid pixels keep their flow values
    else:
        raise ValueError(kind)
    return f


SEGMENT_KINDS = ("blocks", "noise", "smooth", "stale")
