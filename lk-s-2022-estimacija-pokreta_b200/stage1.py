"""Stage 1 drop-in: `daisy i flann.py <pair> <backward> <dopython>` (reference daisy i flann.py:11-432).

Same argv, same input paths, same output files (names, shapes, dtypes, fills) as the reference script;
the arithmetic runs on the GPU through libflowb200.so.  Differences, all documented in DESIGN.md:
  * FLANN's approximate kd-tree search is replaced by the exact per-cell search (north star);
  * np.random (unseeded in the reference) is replaced by a Philox stream; seed = FLOWB200_SEED or
    2*pair+backward;
  * the hard-coded constants (picw, pich, cellw, cellh, ...) can be overridden through FLOWB200_* environment
    variables instead of editing the source.
"""
import datetime as dt
import os
import sys

import numpy as np

from . import io_contract as ioc
from .params import DEFAULT_KNN_MODE, FlowParams


def params_from_env(**kw):
    e = os.environ
    base = dict(W=int(e.get("FLOWB200_PICW", 1241)), H=int(e.get("FLOWB200_PICH", 375)),
                cellw=int(e.get("FLOWB200_CELLW", 73)), cellh=int(e.get("FLOWB200_CELLH", 25)),
                k_cell=int(e.get("FLOWB200_KCELL", 5)), n_gauss=int(e.get("FLOWB200_NGAUSS", 25)),
                maxnprop=int(e.get("FLOWB200_MAXNPROP", 150)),
                knn_mode=int(e.get("FLOWB200_KNN_MODE", DEFAULT_KNN_MODE)))
    base.update(kw)
    return FlowParams(**base).validate()


def pack_for_c(packed, H, W):
    """pakovanjeZaC (daisy i flann.py:321-398): the same K-set bits re-laid per chain in traversal order."""
    kdim = packed.shape[-1]
    s0 = np.zeros(((W + 1) // 2, H, kdim), dtype=np.uint8)
    s1 = np.zeros(((H + 1) // 2, W, kdim), dtype=np.uint8)
    s2 = np.zeros(((W + 1) // 2, H, kdim), dtype=np.uint8)
    s3 = np.zeros(((H + 1) // 2, W, kdim), dtype=np.uint8)
    ty = np.arange(H - 1)
    for tx in range(W):
        if tx % 2 == 0:
            s0[tx // 2, ty] = packed[ty, tx, 0]
        else:
            s2[(W - 1 - tx) // 2, H - 2 - ty] = packed[ty, tx, 0]
    tx = np.arange(W - 1)
    for ty_ in range(H):
        if ty_ % 2 == 0:
            s1[ty_ // 2, W - 2 - tx] = packed[ty_, tx, 1]
        else:
            s3[(H - 1 - ty_) // 2, tx] = packed[ty_, tx, 1]
    return s0, s1, s2, s3


def run(picindex, backward, dopython, p=None, data_root="..", out_dir=".", seed=None, log=print):
    import cv2
    import torch
    from . import ops
    picindex = ioc.pad2(picindex)
    backward = int(backward)
    p = p or params_from_env()
    f1, f2 = ioc.image_paths(picindex, backward, data_root)
    pic1, pic2 = cv2.imread(f1), cv2.imread(f2)          # None -> TypeError below, like the reference (:52)
    pic3 = np.ascontiguousarray(pic1[0:p.H, 0:p.W, :])
    pic4 = np.ascontiguousarray(pic2[0:p.H, 0:p.W, :])
    if pic3.shape[:2] != (p.H, p.W) or pic4.shape[:2] != (p.H, p.W):
        raise ValueError(f"image smaller than pich x picw = {p.H} x {p.W}")
    if seed is None:
        seed = int(os.environ.get("FLOWB200_SEED", 2 * int(picindex) + backward))
    log(dt.datetime.now())
    d1 = ops.daisy(torch.from_numpy(pic3).cuda())
    d2 = ops.daisy(torch.from_numpy(pic4).cuda())
    log("napravio")
    pvec, lcost, nprop, labels = ops.knn_proposals(d1, d2, p)
    log("generisao")
    log(dt.datetime.now())

    def save(name, arr):
        np.save(os.path.join(out_dir, name), arr)
    flow00, _ = ops.flow_from_labels(pvec, labels, want_uvv=False)
    save(ioc.flow_file(picindex, backward, 0), flow00.cpu().numpy())                          # :201
    save(ioc.labels_file(picindex, backward, 0), labels.cpu().numpy().astype(np.int64))       # :202
    log("sacuvao0")
    ops.random_proposals(d1, d2, p, pvec, lcost, nprop, labels, seed=seed)
    log("i ovo")
    save(ioc.stage1_file(picindex, backward, "proposals_nakon_gausa"), ioc.unpack_proposals(pvec.cpu().numpy()))
    save(ioc.stage1_file(picindex, backward, "lcosts_nakon_gausa"), lcost.cpu().numpy().astype(np.float64))
    save(ioc.stage1_file(picindex, backward, "nprop"), nprop.cpu().numpy().astype(np.int64))
    log("sacuvao1")
    if os.environ.get("FLOWB200_SKIP_KSETS") != "1":
        packed = ops.ksets_pack(pvec, nprop, p.tpsi).cpu().numpy()
        if dopython:
            save(ioc.stage1_file(picindex, backward, "packedksets"), packed)                  # :308
            log("spakovao za python")
        else:
            for i, a in enumerate(pack_for_c(packed, p.H, p.W)):                              # :394-397
                save(ioc.stage1_file(picindex, backward, f"pakovani za c {i}"), a)
            log("spakovao za C")
    log(dt.datetime.now())
    return 0


def main(argv=None):
    argv = sys.argv if argv is None else argv
    picindex, backward, dopython = argv[1], argv[2], argv[3] == "1"      # IndexError on bad argv, like the reference
    return run(picindex, backward, dopython)
