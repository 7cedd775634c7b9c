"""Drop-in for the reference's `napravi_parove` module (napravi_parove.py:3-13): the EpicFlow match list
`u v u+dx v+dy` for every valid pixel of a (H,W,3) float32 (dx,dy,valid) field, one line per match, raster order.
Host-side wire format after the hot path; text is byte-identical to the reference's."""
import numpy as np


def parovi(npyfile, txtfile):
    flow = np.load(npyfile)
    height, width, _ = flow.shape
    vs, us = np.nonzero(flow[:, :, 2] > 0.5)                 # raster order: v outer, u inner (:9-10)
    # the reference adds a Python int to a numpy float32 scalar: the sum stays float32 and str() prints its
    # shortest float32 representation
    tu = flow[vs, us, 0] + us.astype(np.float32)
    tv = flow[vs, us, 1] + vs.astype(np.float32)
    with open(txtfile, "w+") as f:
        f.write("".join(f"{u} {v} {str(a)} {str(b)}\n" for u, v, a, b in zip(us.tolist(), vs.tolist(), tu, tv)))
