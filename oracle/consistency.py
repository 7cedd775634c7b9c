"""Forward/backward consistency oracle (numpy).  TEST INFRASTRUCTURE — see oracle/__init__.py.

Restates `postprocessing.py`:
  FlowImage.ucitajFlow :7-17, consistencyCheck :79-110, fowardBackwardConsistency :114-117,
  postProcessing :123-135.  Pinned: tests import the reference module itself in the build
  container and store its outputs under tests/golden/.
"""
import numpy as np


def ucitaj_flow(flow_yx):
    """ucitajFlow (:7-17): npy (H,W,2) [dy,dx] -> float32 (H,W,3) = (dx, dy, 1)."""
    flow_yx = np.asarray(flow_yx)
    H, W, _ = flow_yx.shape
    out = np.zeros((H, W, 3), dtype=np.float32)
    out[..., 0] = flow_yx[..., 1]
    out[..., 1] = flow_yx[..., 0]
    out[..., 2] = 1
    return out


def forward_backward_consistency(flow1, flow2, tresh):
    """fowardBackwardConsistency (:114-117) = consistencyCheck (:79-110) at every (a, b) of flow1.

    Quirk Q5 is reproduced: channel 0 (dx) is added to the ROW index a, channel 1 (dy) to the
    column index b; bounds are shape[0] / shape[1] in that order.  All arithmetic float32
    (numpy scalar float32 + Python int stays float32).  Only flow1 is written and only flow2 is
    gathered, so evaluating all pixels at once equals the reference's sequential loop.
    Returns the modified copy of flow1."""
    f1 = np.array(flow1, dtype=np.float32, copy=True)
    f2 = np.asarray(flow2, dtype=np.float32)
    A, B, _ = f1.shape
    a = np.arange(A, dtype=np.float32)[:, None]
    b = np.arange(B, dtype=np.float32)[None, :]
    valid = f1[..., 2] > 0.5
    a2 = np.trunc(f1[..., 0] + a).astype(np.int64)          # int() truncates toward zero
    b2 = np.trunc(f1[..., 1] + b).astype(np.int64)
    oob = (a2 < 0) | (b2 < 0) | (a2 >= A) | (b2 >= B)
    a2c, b2c = np.clip(a2, 0, A - 1), np.clip(b2, 0, B - 1)
    g = f2[a2c, b2c]
    inv2 = ~(g[..., 2] > 0.5)
    du = f1[..., 0] + g[..., 0]
    dv = f1[..., 1] + g[..., 1]
    err = np.sqrt(dv * dv + du * du)                         # float32
    bad = valid & (oob | inv2 | (err > np.float32(tresh)))
    f1[bad] = 0
    return f1


def post_processing(flow_fwd_yx, flow_bwd_yx, con_tresh):
    """postProcessing (:123-135) without the file I/O: returns float32 (H,W,3) (dx,dy,valid)."""
    return forward_backward_consistency(ucitaj_flow(flow_fwd_yx), ucitaj_flow(flow_bwd_yx), con_tresh)
