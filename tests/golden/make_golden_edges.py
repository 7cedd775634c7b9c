"""Generates tests/golden/edges.npz: outputs of the reference's own `edge.canny_ivice` (edge.py:19-35), imported
unmodified from /root/reference (this container only), on synthetic images written as PNG files."""
import importlib
import importlib.util
import os
import sys
import tempfile

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
synth = importlib.import_module("lk-s-2022-estimacija-pokreta_b200.synth")


def images():
    rng = np.random.default_rng(19)
    out = {"texture": synth.texture(96, 128, 4), "noise": rng.integers(0, 256, size=(50, 61, 3), dtype=np.uint8)}
    boxes = np.full((70, 90, 3), 40, np.uint8)
    for _ in range(9):
        a, b = int(rng.integers(0, 60)), int(rng.integers(0, 80))
        boxes[a:a + int(rng.integers(2, 30)), b:b + int(rng.integers(2, 30))] = rng.integers(0, 256, size=3)
    out["boxes"] = boxes
    ramp = np.zeros((40, 64, 3), np.uint8)
    ramp[...] = (np.arange(64) * 4)[None, :, None]              # weak gradient everywhere, strong nowhere
    ramp[10:30, 20:22] = 255
    out["ramp"] = ramp
    out["row"] = rng.integers(0, 256, size=(1, 80, 3), dtype=np.uint8)
    out["column"] = rng.integers(0, 256, size=(77, 1, 3), dtype=np.uint8)
    return out


if __name__ == "__main__":
    spec = importlib.util.spec_from_file_location("_ref_edge", "/root/reference/edge.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    data = {}
    for name, img in images().items():
        d = tempfile.mkdtemp()
        png, binf = os.path.join(d, "a.png"), os.path.join(d, "a.bin")
        assert cv2.imwrite(png, img)
        ref.canny_ivice(png, binf)
        e = np.fromfile(binf, dtype=np.float32).reshape(img.shape[:2])
        assert set(np.unique(e)) <= {0.0, 1.0}
        data[name + "_img"] = img
        data[name + "_edges"] = (e == 0).astype(np.uint8)      # 1 where canny_ivice wrote 0.0 (an edge)
        print(name, img.shape, "edge pixels", int((e == 0).sum()))
    np.savez_compressed(os.path.join(HERE, "edges.npz"), **data)
