"""Generates tests/golden/segments.npz: outputs of the reference's own `postprocessing.removeSmallSegments`
(postprocessing.py:29-76), imported unmodified from /root/reference (this container only).

Cases: every field kind of tests/helpers.segment_test_field at several thresholds / minimum sizes, chosen so that
segments are removed (the column rebinding of :74 then takes effect) and invalid seeds absorb neighbours.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_harness as rh          # noqa: E402
from helpers import SEGMENT_KINDS, segment_test_field   # noqa: E402

CASES = [  # kind, A, B, tresh, min_segment_size
    ("blocks", 36, 44, 3, 100),
    ("blocks", 41, 29, 10, 100),
    ("noise", 30, 33, 10, 10),
    ("noise", 25, 40, 3, 4),
    ("smooth", 48, 52, 1.5, 100),
    ("smooth", 40, 37, 10, 100),
    ("stale", 33, 35, 3, 100),
    ("stale", 28, 31, 0, 1000),
    ("blocks", 1, 40, 3, 100),
    ("stale", 37, 1, 3, 100),
    ("noise", 2, 2, 100, 100),
]

if __name__ == "__main__":
    assert rh.available(), "reference source not found"
    pp = rh.reference_postprocessing()
    rng = np.random.default_rng(2022)
    data = {"n": np.array(len(CASES))}
    for i, (kind, A, B, tresh, ms) in enumerate(CASES):
        assert kind in SEGMENT_KINDS
        f = segment_test_field(rng, A, B, kind)
        out = f.copy()
        assert pp.removeSmallSegments(out, tresh, ms) is None
        assert np.array_equal(out[..., :2], f[..., :2])
        data[f"c{i}_in"] = f
        data[f"c{i}_valid_out"] = out[..., 2].astype(np.uint8)
        data[f"c{i}_par"] = np.array([tresh, ms], dtype=np.float64)
        print(i, kind, A, B, tresh, ms, "valid", int(f[..., 2].sum()), "->", int(out[..., 2].sum()))
    np.savez_compressed(os.path.join(HERE, "segments.npz"), **data)
