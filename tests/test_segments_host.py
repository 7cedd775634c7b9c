"""removeSmallSegments: the device code's replay logic (csrc/segments_core.cuh) compiled for the HOST.

The CUDA implementation labels connected components in parallel and replays the reference's scan at component
granularity (see segments_core.cuh for why that is exact).  The replay functions are plain C++ shared by the kernel
and by tests/seg_host_emul.cpp; here they are checked on the CPU against the golden fixtures (the reference's own
function) and against the C oracle on random fields.  The kernels themselves are covered by tests/test_z_gpu_segments.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import SEGMENT_KINDS, load_npz, segment_test_field
from oracle import cport

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("seg") / "libseg_emul.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-ffp-contract=off", "-o", so,
                    os.path.join(HERE, "seg_host_emul.cpp")], check=True)
    lib = C.CDLL(so)

    def run(flow, tresh, ms, prefilter=0):
        f = np.array(flow, dtype=np.float32, copy=True)
        assert lib.seg_emul(f.ctypes.data_as(C.c_void_p), f.shape[0], f.shape[1], C.c_float(tresh), int(ms),
                            int(prefilter)) == 0
        return f
    return run


@pytest.mark.parametrize("prefilter", [0, 1])
def test_replay_logic_golden(emul, prefilter):
    z = load_npz("segments")
    for i in range(int(z["n"])):
        f, (tresh, ms) = z[f"c{i}_in"], z[f"c{i}_par"]
        out = emul(f, float(tresh), int(ms), prefilter)
        assert np.array_equal(out[..., :2], f[..., :2])
        assert np.array_equal(out[..., 2], z[f"c{i}_valid_out"].astype(np.float32)), i


@pytest.mark.parametrize("prefilter", [0, 1])
def test_replay_logic_vs_oracle_random(emul, prefilter):
    rng = np.random.default_rng(77)
    removed = 0
    for it in range(120):
        kind = SEGMENT_KINDS[it % 4]
        A, B = int(rng.integers(2, 90)), int(rng.integers(2, 90))
        f = segment_test_field(rng, A, B, kind)
        tresh = [10, 3, 1.5, 0][int(rng.integers(0, 4))]
        ms = [100, 10, 4, 2, 1000, 1 << 30][int(rng.integers(0, 6))]
        want, n = cport.remove_small_segments(f, tresh, ms, want_count=True)
        removed += n
        assert np.array_equal(emul(f, tresh, ms, prefilter), want), (it, kind, A, B, tresh, ms)
    assert removed > 1000
