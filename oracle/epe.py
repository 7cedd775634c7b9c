"""End-point-error metric (numpy).  TEST INFRASTRUCTURE — see oracle/__init__.py.

Restates visualization.errorImage (visualization.py:128-152): over pixels valid in both fields,
f_err = sqrt(dfu^2 + dfv^2) in float32; mean f_err; outliers = 100 * #(f_err > 3.0) / #valid.
(visualization.py itself cannot be imported: it runs at import time and needs matplotlib.)
"""
import numpy as np


def error_image(test_uvv, gt_uvv, abs_thresh=3.0):
    """test/gt: float32 (H,W,3) = (u, v, valid).  Returns (mean_epe, outlier_percent, n_valid)."""
    t = np.asarray(test_uvv, dtype=np.float32)
    g = np.asarray(gt_uvv, dtype=np.float32)
    m = (t[..., 2] > 0.5) & (g[..., 2] > 0.5)
    dfu = t[..., 0][m] - g[..., 0][m]
    dfv = t[..., 1][m] - g[..., 1][m]
    err = np.sqrt(dfu * dfu + dfv * dfv)
    n = int(m.sum())
    if n == 0:
        return float("nan"), float("nan"), 0
    return float(np.average(err)), float((err > abs_thresh).sum() * 100 / n), n
