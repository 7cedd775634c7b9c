"""CPU oracle for the discrete-optical-flow hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy + a small C library) of the algorithm the reference
`pfe-rs/lk-s-2022-estimacija-pokreta` runs in `daisy i flann.py`, `python bcd.py` and
`postprocessing.py`.  It exists so that the CUDA product path can be checked for parity.

Rules (enforced by tests/test_layout.py):
  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
    legs may import or execute anything under `oracle/`;
  * the product package never imports it and has no CPU fallback.

Pinning status (see DESIGN.md "Oracle"):
  * BCD, K-set packing, proposal bookkeeping (generisi / nasumicni with replayed draws) and the
    consistency check are PINNED: `tests/golden/*.npz` holds outputs of the reference's own,
    unmodified source executed in the build container by `oracle/ref_harness.py`
    (script: `tests/golden/make_golden.py`), and `tests/test_oracle_golden.py` checks the
    restatements against them bit for bit.
  * removeSmallSegments (`flow_oracle.c: fo_remove_small_segments`) is PINNED by outputs of the
    reference's own function (`tests/golden/make_golden_segments.py`, `tests/test_oracle_c.py`);
    the Canny edge map (`oracle/edges.py`) is PINNED twice: step by step against cv2 itself (its
    arithmetic is OpenCV's main module, installed here) and against outputs of the reference's own
    `edge.canny_ivice` (`tests/golden/make_golden_edges.py`, `tests/test_oracle_edges.py`).
  * DAISY (third-party opencv-contrib `xfeatures2d::DAISY`, no version pinned by the reference,
    not installable offline) and FLANN (third-party `pyflann`, approximate and unseeded) are
    PARITY UNPINNED: the restatement in `oracle/daisy.py` follows the published algorithm
     and `oracle/proposals.knn_exact` is exact brute force, the search the
    north star substitutes for FLANN.
"""
