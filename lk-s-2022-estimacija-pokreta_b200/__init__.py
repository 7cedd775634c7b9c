"""flowb200: B200-native discrete optical flow hot path (DAISY -> kNN proposals -> BCD -> fwd/bwd check).

Drop-in for the hot path of pfe-rs/lk-s-2022-estimacija-pokreta (`daisy i flann.py`, `python bcd.py`,
`postprocessing.py`).  All arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI in
include/flowb200.h (libflowb200.so); there is no CPU fallback.
"""
from .params import FlowParams, for_k  # noqa: F401
