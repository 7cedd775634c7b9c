"""Drop-in for the reference's `postprocessing` module (import postprocessing); see
lk-s-2022-estimacija-pokreta_b200/postprocessing.py."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
_m = importlib.import_module("lk-s-2022-estimacija-pokreta_b200.postprocessing")
FlowImage = _m.FlowImage
removeSmallSegments = _m.removeSmallSegments
consistencyCheck = _m.consistencyCheck
fowardBackwardConsistency = _m.fowardBackwardConsistency
postProcessing = _m.postProcessing

if __name__ == "__main__":
    postProcessing(sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4])
