"""Generates two fixtures from the reference's own source, run in the build container (needs /root/reference):

  zac.npz  `daisy i flann.py <pair> 0 0` (dopython = 0): pakovanjeZaC's four chain-ordered K-set arrays
           (daisy i flann.py:321-398) together with the proposals / nprop they were packed from.  maxnprop is
           substituted to 46 (2 x 2 cells x 5 + 25 random <= 45 labels; K*K must not be a multiple of 8 for :352 to
           broadcast) so that the arrays stay small.
  epe.npz  visualization.errorImage (visualization.py:128-157) exec'd as text (the module itself runs a GUI script at
           import and needs matplotlib): mean end-point error and outlier percentage it appends to
           srednja_greska.txt / procenat_outliera.txt, for random (u, v, valid) fields.
"""
import importlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh          # noqa: E402

synth = importlib.import_module("lk-s-2022-estimacija-pokreta_b200.synth")
OUT = os.path.dirname(os.path.abspath(__file__))
VIS = os.path.join(rh.REF, "visualization.py")


def zac_case():
    H, W, cw, ch, pair, K = 22, 27, 12, 10, 4, 46
    img1, img2, _, _ = synth.make_pair(H, W, pair, max_dx=4, max_dy=3, n_rect=1)
    wd = tempfile.mkdtemp(prefix="golden_zac_")
    r = rh.run_stage1(wd, img1, img2, pair, 0, False, cw, ch, seed=21, maxnprop=K)
    assert "packedksets" not in r and all(f"zac{i}" in r for i in range(4))
    data = {"meta": np.array([H, W, cw, ch, pair, K], dtype=np.int64),
            "proposals": r["proposals"].astype(np.int16), "nprop": r["nprop"].astype(np.int16)}
    for i in range(4):
        data[f"zac{i}"] = r[f"zac{i}"]
        print("zac", i, r[f"zac{i}"].shape, int(np.unpackbits(r[f"zac{i}"]).sum()), "bits set")
    print("nprop", r["nprop"].min(), r["nprop"].max())
    np.savez_compressed(os.path.join(OUT, "zac.npz"), **data)


def epe_case():
    src = open(VIS).read()
    c = src.index("class FlowImage:")
    d = src.index("def errorImage(")
    e = src.index("# def reverse(flow)")
    ns = {"np": np, "os": os, "cv2": None, "cmap": lambda x: (0.0, 0.0, 0.0, 1.0)}
    exec(src[c:d] + "\n" + src[d:e], ns)
    rng = np.random.default_rng(5)
    data = {}
    cwd = os.getcwd()
    for k, (h, w, noise, pv) in enumerate([(9, 14, 2.0, 0.8), (17, 11, 6.0, 0.5), (5, 5, 0.2, 1.0)]):
        g = np.zeros((h, w, 3), np.float32)
        g[..., :2] = rng.normal(0, 7, (h, w, 2))
        g[..., 2] = rng.random((h, w)) < 0.9
        t = np.zeros((h, w, 3), np.float32)
        t[..., :2] = g[..., :2] + rng.normal(0, noise, (h, w, 2)).astype(np.float32)
        t[..., 2] = rng.random((h, w)) < pv
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)
            try:
                ft, fg = ns["FlowImage"](), ns["FlowImage"]()
                ft.flow, ft.height, ft.width = t.copy(), h, w
                fg.flow, fg.height, fg.width = g.copy(), h, w
                ns["errorImage"](ft, fg)
                mean = float(open("srednja_greska.txt").read().strip())
                outl = float(open("procenat_outliera.txt").read().strip())
            finally:
                os.chdir(cwd)
        data[f"c{k}_test"], data[f"c{k}_gt"] = t, g
        data[f"c{k}_mean"], data[f"c{k}_outliers"] = np.float64(mean), np.float64(outl)
        data[f"c{k}_nvalid"] = np.int64(((t[..., 2] > 0.5) & (g[..., 2] > 0.5)).sum())
        print("epe", k, mean, outl)
    np.savez_compressed(os.path.join(OUT, "epe.npz"), **data)


if __name__ == "__main__":
    assert rh.available(), "reference source not found"
    zac_case()
    epe_case()
