"""Synthetic textured image pairs warped by a known flow (no datasets offline).

Image 1 is seeded multi-octave band-limited noise; the ground-truth forward flow is a global
translation plus a low-frequency sinusoid plus a few rectangles with their own integer offset
(motion discontinuities), all inside the reference's +-2-cell search window.  Image 2 is image 1
resampled so that I2(p + F(p)) = I1(p) away from occlusions.  Files are laid out the way the
reference reads them (`daisy i flann.py`:26-27):  <root>/training/image_2/0001NN_1{0,1}.png.
"""
import os
import numpy as np
from scipy import ndimage


def texture(H, W, seed):
    rng = np.random.default_rng(seed)
    img = np.zeros((H, W, 3), dtype=np.float64)
    for s, wgt in ((1, 1.0), (2, 1.0), (4, 1.2), (8, 1.5), (16, 2.0)):
        n = rng.random((H, W, 3))
        n = ndimage.gaussian_filter(n, sigma=(s, s, 0), mode="reflect")
        n = (n - n.mean()) / (n.std() + 1e-12)
        img += wgt * n
    lo, hi = np.percentile(img, 0.5), np.percentile(img, 99.5)
    img = np.clip((img - lo) / (hi - lo), 0, 1)
    return (img * 255.0 + 0.5).astype(np.uint8)


def gt_flow(H, W, seed, max_dx=40, max_dy=12, n_rect=4):
    """float32 (H,W,2) [dy,dx] forward flow (integer-valued on rectangles, smooth elsewhere)."""
    rng = np.random.default_rng(seed + 7919)
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    tx = rng.integers(-max_dx, max_dx + 1)
    ty = rng.integers(-max_dy, max_dy + 1)
    ax, ay = rng.uniform(2, 8), rng.uniform(1, 4)
    px, py = rng.uniform(0, 2 * np.pi, 2)
    dx = tx + ax * np.sin(2 * np.pi * ys / (1.7 * H) + px) * np.cos(2 * np.pi * xs / (2.3 * W))
    dy = ty + ay * np.cos(2 * np.pi * xs / (1.9 * W) + py)
    for _ in range(n_rect):
        h, w = rng.integers(max(4, H // 8), max(5, H // 3)), rng.integers(max(4, W // 10), max(5, W // 4))
        y0, x0 = rng.integers(0, max(1, H - h)), rng.integers(0, max(1, W - w))
        dx[y0:y0 + h, x0:x0 + w] = tx + rng.integers(-12, 13)
        dy[y0:y0 + h, x0:x0 + w] = ty + rng.integers(-6, 7)
    return np.stack([dy, dx], axis=-1).astype(np.float32)


def warp_pair(img1, flow):
    """Image 2 by forward-splatting-free inverse lookup: iterate q -> p with p + F(p) = q."""
    H, W, _ = img1.shape
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    py, px = ys.copy(), xs.copy()
    for _ in range(12):                      # fixed-point inversion of p + F(p) = q
        fy = ndimage.map_coordinates(flow[..., 0], [py, px], order=1, mode="nearest")
        fx = ndimage.map_coordinates(flow[..., 1], [py, px], order=1, mode="nearest")
        py, px = ys - fy, xs - fx
    img2 = np.empty_like(img1)
    for c in range(3):
        v = ndimage.map_coordinates(img1[..., c].astype(np.float64), [py, px], order=1, mode="reflect")
        img2[..., c] = np.clip(v + 0.5, 0, 255).astype(np.uint8)
    bwd = np.stack([py - ys, px - xs], axis=-1).astype(np.float32)     # backward flow at q
    return img2, bwd


def make_pair(H, W, pair=0, **kw):
    """Returns (img1 u8 BGR, img2 u8 BGR, fwd float32 [dy,dx], bwd float32 [dy,dx])."""
    seed = 1000 + int(pair)
    img1 = texture(H, W, seed)
    fwd = gt_flow(H, W, seed, **kw)
    img2, bwd = warp_pair(img1, fwd)
    return img1, img2, fwd, bwd


def gt_uvv(flow_yx):
    """(H,W,2) [dy,dx] -> float32 (H,W,3) (u=dx, v=dy, valid) with out-of-frame targets invalid."""
    H, W, _ = flow_yx.shape
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    ty, tx = ys + flow_yx[..., 0], xs + flow_yx[..., 1]
    valid = (ty >= 0) & (ty <= H - 1) & (tx >= 0) & (tx <= W - 1)
    return np.stack([flow_yx[..., 1], flow_yx[..., 0], valid.astype(np.float32)], axis=-1).astype(np.float32)


def write_dataset(root, H, W, pairs):
    """Write pairs as <root>/training/image_2/0001NN_1{0,1}.png (needs cv2, like the reference)."""
    import cv2
    d = os.path.join(root, "training", "image_2")
    os.makedirs(d, exist_ok=True)
    for n in pairs:
        img1, img2, fwd, bwd = make_pair(H, W, n)
        nn = f"{int(n):02d}"
        cv2.imwrite(os.path.join(d, f"0001{nn}_10.png"), img1)
        cv2.imwrite(os.path.join(d, f"0001{nn}_11.png"), img2)
        np.save(os.path.join(d, f"0001{nn}_gt_fwd.npy"), fwd)
        np.save(os.path.join(d, f"0001{nn}_gt_bwd.npy"), bwd)
