"""removeSmallSegments (postprocessing.py:29-76) on the GPU vs the reference's own outputs and the
oracle.  Bit-exact valid flags, flow components untouched.  (Named to run after the hot-path GPU tests.)"""
import importlib
import os
import sys

import numpy as np
import pytest

from helpers import SEGMENT_KINDS, load_npz, pkg, segment_test_field
from oracle import cport

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_remove_small_segments_golden_through_the_module():
    """import postprocessing; removeSmallSegments(flow, tresh, min) in place on a host array, returns None."""
    sys.path.insert(0, ROOT)
    pp = importlib.import_module("postprocessing")
    z = load_npz("segments")
    for i in range(int(z["n"])):
        f, (tresh, ms) = z[f"c{i}_in"].copy(), z[f"c{i}_par"]
        src = f.copy()
        assert pp.removeSmallSegments(f, float(tresh), int(ms)) is None
        assert np.array_equal(f[..., :2], src[..., :2])
        assert np.array_equal(f[..., 2], z[f"c{i}_valid_out"].astype(np.float32)), i
    fi = pp.FlowImage()
    fi.flow = z["c0_in"].copy()
    pp.removeSmallSegments(fi.flow, *z["c0_par"])
    assert np.array_equal(fi.flow[..., 2], z["c0_valid_out"].astype(np.float32))


@pytest.mark.parametrize("shape", [(37, 53), (64, 64), (1, 77), (90, 1), (1, 1), (130, 33)])
def test_remove_small_segments_vs_oracle_random(shape):
    ops = pkg("ops")
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    for it in range(8):
        f = segment_test_field(rng, *shape, SEGMENT_KINDS[it % 4])
        tresh = [10, 3, 1.5, 0][it % 4 if it < 4 else int(rng.integers(0, 4))]
        ms = [100, 10, 4, 1 << 30][int(rng.integers(0, 4))]
        want = cport.remove_small_segments(f, tresh, ms)
        got = ops.remove_small_segments(torch.from_numpy(f.copy()).cuda(), tresh, ms).cpu().numpy()
        assert np.array_equal(got, want), (it, tresh, ms, int((got != want).sum()))


def test_remove_small_segments_full_size_after_consistency():
    """1024x436 field with the structure of a checked flow field: smooth motion, islands, zeroed invalid pixels."""
    ops = pkg("ops")
    rng = np.random.default_rng(5)
    f = segment_test_field(rng, 436, 1024, "smooth")
    want, n = cport.remove_small_segments(f, 10, 100, want_count=True)
    assert n > 50
    t = torch.from_numpy(f.copy()).cuda()
    ws = torch.empty(int(pkg("_lib").load().flowb200_segments_workspace_bytes(436, 1024)), dtype=torch.uint8, device="cuda")
    got = ops.remove_small_segments(t, 10, 100, workspace=ws).cpu().numpy()
    assert np.array_equal(got, want)
    # a second pass over the result: every small segment is gone already or is rebuilt the same way by the oracle
    again = ops.remove_small_segments(torch.from_numpy(want.copy()).cuda(), 10, 100).cpu().numpy()
    assert np.array_equal(again, cport.remove_small_segments(want, 10, 100))


def test_remove_small_segments_workspace_too_small():
    lib = pkg("_lib")
    L = lib.load()
    t = torch.zeros((8, 8, 3), dtype=torch.float32, device="cuda")
    ws = torch.empty(16, dtype=torch.uint8, device="cuda")
    assert L.flowb200_remove_small_segments(t.data_ptr(), 8, 8, 1.0, 10, ws.data_ptr(), 16, None) == lib.EWORKSPACE
    assert L.flowb200_remove_small_segments(None, 8, 8, 1.0, 10, ws.data_ptr(), 16, None) == lib.EINVAL
