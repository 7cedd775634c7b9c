"""Drop-in for the reference's `postprocessing` module (postprocessing.py:1-142), hot-path part.

Same names, arguments, in-place semantics and output as the reference:
  FlowImage().ucitajFlow(path)                       :5-17     .npy only; [dy,dx] -> (dx, dy, 1) float32
  consistencyCheck(flow1, flow2, u1, v1, tresh)      :79-110   one pixel, in place on flow1.flow
  fowardBackwardConsistency(flow1, flow2, tresh)     :114-117  every pixel of flow1
  postProcessing(filename1, filename2, con_tresh, npysave) -> FlowImage   :123-135
The checks run on the GPU through flowb200_consistency (quirk Q5 reproduced); there is no CPU fallback.
removeSmallSegments (:29-76) is disabled in the reference (:132) and is not part of the hot path.
"""
import os
import sys

import numpy as np

from . import _lib


class FlowImage:
    def __init__(self):
        self.flow = np.zeros((1, 1, 3), dtype=np.float32)
        self.height = 1
        self.width = 1

    def ucitajFlow(self, flow_file):
        ext = os.path.splitext(flow_file)[1]
        if ext == ".npy":
            img = np.load(flow_file)
            self.height, self.width = int(img.shape[0]), int(img.shape[1])
            self.flow = np.zeros((self.height, self.width, 3), dtype=np.float32)
            self.flow[:, :, 0] = img[:, :, 1]
            self.flow[:, :, 1] = img[:, :, 0]
            self.flow[:, :, 2] = 1
        else:
            print("Unknown file format!")


def _check_region(flow1, flow2, a0, a1, b0, b1, tresh):
    f1 = flow1.flow
    if f1.dtype != np.float32 or not f1.flags.c_contiguous:
        raise TypeError("flow1.flow must be a C-contiguous float32 array")
    f2 = np.ascontiguousarray(flow2.flow, dtype=np.float32)
    A, B = f1.shape[0], f1.shape[1]
    if f2 is f1 or np.shares_memory(f1, f2):
        f2 = f2.copy()
    L = _lib.load()
    _lib.check(L.flowb200_consistency_host(f1.ctypes.data, f2.ctypes.data, A, B, float(tresh), a0, a1, b0, b1),
               "flowb200_consistency_host")


def consistencyCheck(flow1, flow2, u1, v1, tresh):
    """One pixel (u1, v1) = (row, col) of flow1, in place; returns None like the reference."""
    if u1 < 0 or v1 < 0 or u1 >= flow1.flow.shape[0] or v1 >= flow1.flow.shape[1]:
        raise IndexError("index out of bounds")
    _check_region(flow1, flow2, int(u1), int(u1) + 1, int(v1), int(v1) + 1, tresh)


def fowardBackwardConsistency(flow1, flow2, tresh):
    _check_region(flow1, flow2, 0, flow1.flow.shape[0], 0, flow1.flow.shape[1], tresh)


def postProcessing(filename1, filename2, con_tresh, npysave):
    flow1 = FlowImage()
    flow1.ucitajFlow(filename1)
    flow2 = FlowImage()
    flow2.ucitajFlow(filename2)
    fowardBackwardConsistency(flow1, flow2, con_tresh)
    np.save(npysave, flow1.flow)
    return flow1


if __name__ == "__main__":
    postProcessing(sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4])
