"""Recall of FLANN's randomized kd-forest against the exact per-cell search (north star: "reporting recall against
FLANN").  pyflann (the reference's binding, daisy i flann.py:13) is not installable offline; cv2.flann_Index is the
same FLANN code base (OpenCV's bundled fork) and is used with pyflann's defaults: kd-tree forest, 4 trees, 32 checks.
CPU only; descriptors and the exact neighbours come from the oracle.  Usage: python tests/tools/flann_recall.py [k]"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402
from oracle import daisy as od  # noqa: E402
from oracle.proposals import Params, cell_descriptors, knn_exact  # noqa: E402

synth = importlib.import_module("lk-s-2022-estimacija-pokreta_b200.synth")


def main():
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    H, W = 125, 365                                   # 5 x 5 cells of 73 x 25
    img1, img2, _, _ = synth.make_pair(H, W, 3)
    d1, d2 = od.daisy(img1), od.daisy(img2)
    p = Params(H, W)
    cd2 = cell_descriptors(d2, p)
    rng = np.random.default_rng(0)
    hits = top1 = total = 0
    t_build = t_search = 0.0
    for cell in range(p.ncellx * p.ncelly):
        tgt = np.ascontiguousarray(cd2[cell])
        t0 = time.perf_counter()
        index = cv2.flann_Index(tgt, dict(algorithm=1, trees=4))
        t_build += time.perf_counter() - t0
        ys, xs = rng.integers(0, H, 400), rng.integers(0, W, 400)
        q = np.ascontiguousarray(d1[ys, xs])
        exact = knn_exact(q, tgt, k)
        t0 = time.perf_counter()
        approx, _ = index.knnSearch(q, k, params=dict(checks=32))
        t_search += time.perf_counter() - t0
        for a, e in zip(approx, exact):
            hits += len(set(a.tolist()) & set(e.tolist()))
            top1 += int(a[0] == e[0])
        total += len(q)
    out = {"k": k, "queries": total, "recall_at_k": hits / (total * k), "top1_agreement": top1 / total,
           "index": "cv2.flann_Index(algorithm=KDTREE, trees=4), checks=32", "cells": p.ncellx * p.ncelly,
           "image": f"{W}x{H} synthetic pair 3",
           # SURVEY 8(d) baseline (3): the kd-forest itself on this host's CPU, one thread, batched queries (the
           # reference issues them one by one through Python); extrapolated to one direction of 1024x436
           # (238 cells, 9.34 M (pixel, cell) searches)
           "flann_build_ms_per_cell": round(1e3 * t_build / (p.ncellx * p.ncelly), 3),
           "flann_search_us_per_query": round(1e6 * t_search / total, 3),
           "flann_extrapolated_s_per_direction_1024x436": round(238 * t_build / (p.ncellx * p.ncelly)
                                                                + 9.34e6 * t_search / total, 2)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
