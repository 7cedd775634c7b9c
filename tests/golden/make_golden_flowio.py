"""Generates tests/golden/flowio.npz from the reference's own codec code (visualization.py:9-29, 34-124), run here.

visualization.py executes a GUI script at import (matplotlib, cv2.imshow, sys.argv), so only the text of
`read_flo_file` and `class FlowImage` is exec'd, with cv2.imshow stubbed.  Needs /root/reference (this container only).
"""
import os
import sys
import tempfile
import types

import cv2
import numpy as np

REF = "/root/reference/visualization.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def reference_namespace():
    src = open(REF).read()
    a = src.index("def read_flo_file")
    b = src.index("def crop(")
    c = src.index("class FlowImage:")
    d = src.index("def errorImage(")
    fake_cv2 = types.SimpleNamespace(**{k: getattr(cv2, k) for k in ("imread", "imwrite", "cvtColor", "COLOR_BGR2RGB")})
    fake_cv2.imshow = lambda *a, **k: None
    ns = {"np": np, "cv2": fake_cv2, "os": os}
    exec(src[a:b] + "\n" + src[c:d], ns)
    return ns


def main():
    ns = reference_namespace()
    rng = np.random.default_rng(11)
    h, w = 9, 12
    flow = np.zeros((h, w, 3), np.float32)
    flow[:, :, 0] = rng.uniform(-40, 40, (h, w)).astype(np.float32)
    flow[:, :, 1] = rng.uniform(-12, 12, (h, w)).astype(np.float32)
    flow[:, :, 2] = rng.random((h, w)) > 0.3
    out = {"flow": flow}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            fi = ns["FlowImage"]()
            fi.flow, fi.height, fi.width = flow.copy(), h, w
            fi.writeFlowField()                                   # -> flow_field.png
            out["written_png"] = cv2.imread("flow_field.png", -1)
            # a KITTI-layout file for the reader
            rgb = np.zeros((h, w, 3), np.uint16)
            rgb[:, :, 0] = np.where(flow[:, :, 2] > 0, np.rint(flow[:, :, 0] * 64.0 + 32768.0), 0)
            rgb[:, :, 1] = np.where(flow[:, :, 2] > 0, np.rint(flow[:, :, 1] * 64.0 + 32768.0), 0)
            rgb[:, :, 2] = flow[:, :, 2]
            cv2.imwrite("gt.png", cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR))
            out["kitti_png_bgr"] = cv2.imread("gt.png", -1)
            a = ns["FlowImage"](); a.ucitajFlow("gt.png"); out["read_png"] = a.flow
            np.save("f2.npy", np.stack([flow[:, :, 1], flow[:, :, 0]], -1).astype(np.float64))
            np.save("f3.npy", flow)
            a = ns["FlowImage"](); a.ucitajFlow("f2.npy"); out["read_npy2"] = a.flow
            a = ns["FlowImage"](); a.ucitajFlow("f3.npy"); out["read_npy3"] = a.flow
            with open("f.flo", "wb") as f:
                np.array([202021.25], np.float32).tofile(f)
                np.array([w, h], np.int32).tofile(f)
                flow[:, :, :2].astype(np.float32).tofile(f)
            out["flo_bytes"] = np.frombuffer(open("f.flo", "rb").read(), np.uint8)
            out["read_flo_raw"] = ns["read_flo_file"]("f.flo")
            a = ns["FlowImage"](); a.ucitajFlow("f.flo"); out["read_flo"] = a.flow
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "flowio.npz"), **out)
    print({k: (v.shape, v.dtype) for k, v in out.items()})


if __name__ == "__main__":
    sys.exit(main())
