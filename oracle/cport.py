"""ctypes wrapper of oracle/flow_oracle.c.  TEST INFRASTRUCTURE — see oracle/__init__.py.

The C restatement is the multi-threaded CPU baseline (`bench.py` cpu_baseline / `--impl reference`)
and a second, independent checker; it is pinned against tests/golden by tests/test_oracle_c.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libflow_oracle.so")
_lib = None


def build():
    subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.fo_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else C.c_void_p(0)


def num_threads():
    return int(lib().fo_num_threads())


def set_num_threads(n):
    lib().fo_set_num_threads(C.c_int(int(n)))


def generisi(desc1, desc2, p, want_idx=False):
    """Same outputs as oracle.proposals.generisi (reference dtypes), optionally the kNN indices."""
    d1 = np.ascontiguousarray(desc1, dtype=np.float32)
    d2 = np.ascontiguousarray(desc2, dtype=np.float32)
    H, W, K = p.H, p.W, p.maxnprop
    prop = np.empty((H, W, K, 2), dtype=np.int32)
    lc = np.empty((H, W, K), dtype=np.float64)
    npr = np.empty((H, W), dtype=np.int32)
    best = np.empty((H, W), dtype=np.int32)
    r = 2 * p.cell_radius + 1
    idx = np.empty((H, W, r * r, p.k_cell), dtype=np.int32) if want_idx else None
    rc = lib().fo_generisi(_p(d1), _p(d2), H, W, p.cellw, p.cellh, p.cell_radius, p.k_cell, K, C.c_double(p.tphi),
                           _p(prop), _p(lc), _p(npr), _p(best), _p(idx))
    if rc:
        raise RuntimeError(f"fo_generisi failed: {rc}")
    out = (prop.astype(np.int64), lc, npr.astype(np.int64), best.astype(np.int64))
    return out + (idx,) if want_idx else out


def nasumicni(desc1, desc2, proposals, lcosts, nprop, bestlabels, p, draws=None, seed=0):
    """oracle.proposals.nasumicni in C.  Returns new (proposals int64, lcosts f64, nprop int64)."""
    d1 = np.ascontiguousarray(desc1, dtype=np.float32)
    d2 = np.ascontiguousarray(desc2, dtype=np.float32)
    prop = np.array(proposals, dtype=np.int32, copy=True)
    lc = np.array(lcosts, dtype=np.float64, copy=True)
    npr = np.array(nprop, dtype=np.int32, copy=True)
    best = np.ascontiguousarray(bestlabels, dtype=np.int32)
    dr = np.ascontiguousarray(draws, dtype=np.int16) if draws is not None else None
    rc = lib().fo_nasumicni(_p(d1), _p(d2), p.H, p.W, p.cellw, p.cellh, p.k_cell, p.maxnprop, p.n_gauss,
                            C.c_double(p.sigma), C.c_double(p.tphi), _p(prop), _p(lc), _p(npr), _p(best), _p(dr),
                            C.c_uint64(int(seed)))
    if rc:
        raise RuntimeError(f"fo_nasumicni failed: {rc}")
    return prop.astype(np.int64), lc, npr.astype(np.int64)


def ceo_bcd(proposals, lcosts, nprop, labels, bcd_times, tpsi=8, lamda=0.05):
    """Same contract as oracle.bcd.ceo_bcd: list of int64 label images after each sweep."""
    prop = np.ascontiguousarray(proposals, dtype=np.int32)
    lc = np.ascontiguousarray(lcosts, dtype=np.float64)
    npr = np.ascontiguousarray(nprop, dtype=np.int32)
    lab = np.array(labels, dtype=np.int32, copy=True)
    H, W, K, _ = prop.shape
    snaps = np.empty((bcd_times, H, W), dtype=np.int32)
    rc = lib().fo_bcd(_p(prop), _p(lc), _p(npr), _p(lab), H, W, K, C.c_double(lamda), int(tpsi), int(bcd_times),
                      _p(snaps))
    if rc:
        raise RuntimeError(f"fo_bcd failed: {rc}")
    return [s.astype(np.int64) for s in snaps]


def forward_backward_consistency(flow1, flow2, tresh):
    f1 = np.array(flow1, dtype=np.float32, copy=True)
    f2 = np.ascontiguousarray(flow2, dtype=np.float32)
    A, B, _ = f1.shape
    lib().fo_consistency(_p(f1), _p(f2), A, B, C.c_float(tresh))
    return f1


def remove_small_segments(flow, tresh, min_segment_size, want_count=False):
    """removeSmallSegments (postprocessing.py:29-76) on a copy of flow; returns the modified copy."""
    f = np.array(flow, dtype=np.float32, copy=True)
    A, B, _ = f.shape
    nrem = C.c_int32(0)
    rc = lib().fo_remove_small_segments(_p(f), A, B, C.c_float(tresh), int(min_segment_size), C.byref(nrem))
    if rc:
        raise RuntimeError(f"fo_remove_small_segments failed: {rc}")
    return (f, int(nrem.value)) if want_count else f
