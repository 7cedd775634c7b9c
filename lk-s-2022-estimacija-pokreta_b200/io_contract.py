"""File and array contract between the reference's stages.

Names, shapes, dtypes and fill values are exactly those the reference scripts write and read:
  stage 1 (daisy i flann.py:200-202, 249-253, 308), stage 2 (python bcd.py:73-81, 279-283).
Conversions between the reference's host arrays (int64 proposals (H,W,K,2) [dy,dx], float64 lcosts)
and the compact device layout of ops.py are lossless for every value the path can produce
(|dy|,|dx| < 32768; data costs are float32 sums widened to float64).
"""
import os

import numpy as np


def pad2(n):
    """picindex formatting: zero-padded to two digits (daisy i flann.py:24-25)."""
    s = str(int(n))
    return s if len(s) > 1 else "0" + s


def image_paths(picindex, backward, root=".."):
    """Source and target frame paths (daisy i flann.py:26-27)."""
    b = int(backward)
    d = os.path.join(root, "data_scene_flow", "training", "image_2")
    return (os.path.join(d, f"0001{pad2(picindex)}_1{b}.png"), os.path.join(d, f"0001{pad2(picindex)}_1{1 - b}.png"))


def flow_file(picindex, backward, w):
    return f"Gotova flow slika 1{pad2(picindex)} backward={int(backward)} posle {pad2(w)} BCD.npy"


def labels_file(picindex, backward, w):
    return f"Bestlabels fajl slike 1{pad2(picindex)} backward={int(backward)} posle {pad2(w)} BCD.npy"


def stage1_file(picindex, backward, what):
    """what in: proposals_nakon_gausa, lcosts_nakon_gausa, nprop, packedksets, 'pakovani za c N'."""
    return f"Daisy output slike 1{pad2(picindex)} backward={int(backward)} {what}.npy"


def pack_proposals(proposals):
    """int (H,W,K,2) [dy,dx] -> int32 (H,W,K) packed (int16 dy | int16 dx << 16); (-1,-1) -> -1."""
    p = np.asarray(proposals)
    if p.min(initial=0) < -32768 or p.max(initial=0) > 32767:
        raise ValueError("proposal component outside int16")
    dy = p[..., 0].astype(np.int64) & 0xFFFF
    dx = p[..., 1].astype(np.int64) & 0xFFFF
    return (dy | (dx << 16)).astype(np.uint32).view(np.int32)


def unpack_proposals(pvec):
    """int32 (H,W,K) packed -> int64 (H,W,K,2) [dy,dx] (the reference's `proposals`)."""
    v = np.asarray(pvec, dtype=np.int32)
    dy = (v & 0xFFFF).astype(np.uint16).view(np.int16).astype(np.int64)
    dx = (v >> 16).astype(np.int64)
    return np.stack([dy, dx], axis=-1)


def lcosts_to_f32(lcosts):
    """float64 lcosts -> float32, refusing values the reference's float32 sums cannot have produced."""
    a = np.asarray(lcosts, dtype=np.float64)
    f = a.astype(np.float32)
    if not np.array_equal(f.astype(np.float64), a):
        return None          # caller falls back to the float64-cost BCD mode
    return f


def quantised_m(lcosts, lamda, shift):
    """If lamda*lcost*2^S is an exact integer for every used slot return those int32 m, else None."""
    a = np.asarray(lcosts, dtype=np.float64)
    used = a != 1000.0
    m = np.where(used, lamda * a * float(1 << shift), 0.0)
    r = np.rint(m)
    # the int32 programme holds a data cost in 16 bits (include/flowb200.h: 0 <= m < 65536)
    if not np.array_equal(r, m) or r.max(initial=0) >= 65536 or r.min(initial=0) < 0:
        return None
    return r.astype(np.int32)
