"""DAISY descriptor oracle (numpy, float32).  TEST INFRASTRUCTURE — see oracle/__init__.py.

Restates what the reference obtains from
    cv2.xfeatures2d.DAISY_create(radius=5, q_radius=4, q_theta=4, q_hist=4)     (daisy i flann.py:66)
    daisy.compute(picture, kp)  with one keypoint per pixel, row-major           (daisy i flann.py:69-77)
i.e. opencv-contrib `xfeatures2d/src/daisy.cpp` with its defaults norm=NRM_NONE, interpolation=True,
use_orientation=False.  That library is a third-party dependency which the reference does not pin
and which is absent from /root/reference and from this image: PARITY UNPINNED.  The algorithm below
is the published one (Tola et al. PAMI 2010 / SURVEY.md Appendix C steps 1-8).

Only the primitives are pinned: `tests/test_oracle_daisy.py` checks gray conversion, the Gaussian
kernels, the separable blur and the central-difference gradient against the main-module `cv2`
functions daisy.cpp calls (cvtColor / getGaussianKernel / GaussianBlur / Sobel) when cv2 is present.
"""
import numpy as np

RADIUS = 5
Q_RADIUS = 4
Q_THETA = 4
Q_HIST = 4
N_REGIONS = 1 + Q_RADIUS * Q_THETA          # 17
DESC_DIM = N_REGIONS * Q_HIST               # 68
SIGMA_INIT = 1.6                            # g_sigma_init in daisy.cpp


def gray_u8(bgr):
    """cv2.cvtColor(BGR2GRAY) on uint8 as OpenCV 4.x computes it: 15-bit fixed point
    (B*3735 + G*19235 + R*9798 + 2^14) >> 15 (checked against cv2 4.13 in tests/test_oracle_daisy.py)."""
    b = bgr[..., 0].astype(np.int32)
    g = bgr[..., 1].astype(np.int32)
    r = bgr[..., 2].astype(np.int32)
    return ((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15).astype(np.uint8)


def gaussian_kernel(ksize, sigma):
    """cv2.getGaussianKernel(ksize, sigma, CV_32F) for sigma > 0: exp(-x^2/2s^2), normalised in double."""
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    k = np.exp(-(x * x) / (2.0 * sigma * sigma))
    k /= k.sum()
    return k.astype(np.float32)


def filter_size(sigma, factor=5.0):
    """daisy.cpp filter_size(): int(5*sigma) made odd, at least 3."""
    fsz = int(factor * sigma)
    if fsz % 2 == 0:
        fsz += 1
    return max(fsz, 3)


def blur_taps():
    """The five separable blurs applied to the orientation layers, in order.

    [0] base smoothing sqrt(1.6^2 - 0.5^2); [1..4] the increments that take the layers to the
    cumulative cube sigmas (r+1)*radius/q_radius/2 = 0.625, 1.25, 1.875, 2.5 (Appendix C 4-5).
    Returns a list of float32 kernels (lengths 7, 3, 5, 7, 9).
    """
    sig = [np.sqrt(SIGMA_INIT * SIGMA_INIT - 0.25)]
    cum = [(r + 1) * RADIUS / Q_RADIUS / 2.0 for r in range(Q_RADIUS)]
    prev = 0.0
    for c in cum:
        sig.append(np.sqrt(c * c - prev * prev))
        prev = c
    return [gaussian_kernel(filter_size(s), s) for s in sig]


def blur_sep(img, k):
    """Separable symmetric blur, BORDER_REPLICATE, float32 throughout.

    Accumulation order: centre tap first, then k[i]*(left+right) for i = 1..r (the order of
    OpenCV's symmetric row/column filters); horizontal pass then vertical pass.
    """
    img = np.asarray(img, dtype=np.float32)
    r = len(k) // 2
    squeeze = img.ndim == 2
    if squeeze:
        img = img[..., None]
    H, W, _ = img.shape
    p = np.pad(img, ((0, 0), (r, r), (0, 0)), mode="edge")
    out = k[r] * p[:, r:r + W]
    for i in range(1, r + 1):
        out = out + k[r + i] * (p[:, r - i:r - i + W] + p[:, r + i:r + i + W])
    p = np.pad(out, ((r, r), (0, 0), (0, 0)), mode="edge")
    out = k[r] * p[r:r + H]
    for i in range(1, r + 1):
        out = out + k[r + i] * (p[r - i:r - i + H] + p[r + i:r + i + H])
    out = out.astype(np.float32)
    return out[..., 0] if squeeze else out


def central_gradient(img):
    """cv2.Sobel(ksize=1, scale=0.5, BORDER_REPLICATE) in x and y: 0.5*(I[+1]-I[-1])."""
    p = np.pad(img, 1, mode="edge")
    dx = np.float32(0.5) * (p[1:-1, 2:] - p[1:-1, :-2])
    dy = np.float32(0.5) * (p[2:, 1:-1] - p[:-2, 1:-1])
    return dx.astype(np.float32), dy.astype(np.float32)


def orientation_layers(dx, dy):
    """G_h = max(cos(2 pi h/4)*dx + sin(2 pi h/4)*dy, 0) with float32 cos/sin (LayeredGradientInvoker)."""
    H, W = dx.shape
    out = np.empty((H, W, Q_HIST), dtype=np.float32)
    for h in range(Q_HIST):
        ang = h * 2.0 * np.pi / Q_HIST
        kos = np.float32(np.cos(ang))
        zin = np.float32(np.sin(ang))
        out[..., h] = np.maximum(kos * dx + zin * dy, np.float32(0))
    return out


def grid_offsets():
    """17 sample points (dy, dx) in double: centre, then ring r (radius 1.25*(r+1)) x angle j*pi/2."""
    offs = np.zeros((N_REGIONS, 2), dtype=np.float64)
    for r in range(Q_RADIUS):
        rho = (r + 1) * RADIUS / Q_RADIUS
        for j in range(Q_THETA):
            ang = j * 2.0 * np.pi / Q_THETA
            offs[1 + r * Q_THETA + j] = (rho * np.sin(ang), rho * np.cos(ang))
    return offs


def region_cube():
    """cube index sampled by each of the 17 regions (centre -> cube 0, ring r -> cube r)."""
    return np.array([0] + [r for r in range(Q_RADIUS) for _ in range(Q_THETA)], dtype=np.int32)


def cubes(bgr):
    """(4, H, W, 4) float32: orientation layers after cumulative sigma 0.625/1.25/1.875/2.5."""
    g = gray_u8(bgr).astype(np.float32) / np.float32(255.0)
    g = blur_sep(g, gaussian_kernel(5, 0.5))
    dx, dy = central_gradient(g)
    lay = orientation_layers(dx, dy)
    taps = blur_taps()
    lay = blur_sep(lay, taps[0])
    out = []
    for r in range(Q_RADIUS):
        lay = blur_sep(lay, taps[1 + r])
        out.append(lay)
    return np.stack(out)


def sample(cube_stack):
    """Bilinear petal sampling (bi_get_histogram / i_get_descriptor semantics, Appendix C step 7)."""
    _, H, W, _ = cube_stack.shape
    offs = grid_offsets()
    rc = region_cube()
    desc = np.zeros((H, W, DESC_DIM), dtype=np.float32)
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    for reg in range(N_REGIONS):
        y = yy + offs[reg, 0]
        x = xx + offs[reg, 1]
        inside = (x >= 0) & (x < W - 1) & (y >= 0) & (y < H - 1)
        if reg == 0:
            inside = np.ones_like(inside)       # the centre histogram has no inside() test
        mnx = np.trunc(x).astype(np.int64)
        mny = np.trunc(y).astype(np.int64)
        ok = inside & (mnx < W - 2) & (mny < H - 2)
        mnx_c = np.clip(mnx, 0, W - 2)
        mny_c = np.clip(mny, 0, H - 2)
        alpha = mnx + 1 - x
        beta = mny + 1 - y
        w0 = (alpha * beta).astype(np.float32)
        w1 = (beta - w0.astype(np.float64)).astype(np.float32)
        w2 = (alpha - w0.astype(np.float64)).astype(np.float32)
        w3 = (1 + w0.astype(np.float64) - alpha - beta).astype(np.float32)
        cube = cube_stack[rc[reg]]
        A = cube[mny_c, mnx_c]
        C = cube[mny_c, mnx_c + 1]
        B = cube[mny_c + 1, mnx_c]
        D = cube[mny_c + 1, mnx_c + 1]
        v = w0[..., None] * A + w1[..., None] * C + w2[..., None] * B + w3[..., None] * D
        v = np.where(ok[..., None], v, np.float32(0)).astype(np.float32)
        desc[..., reg * Q_HIST:(reg + 1) * Q_HIST] = v
    return desc


def daisy(bgr):
    """uint8 BGR (H,W,3) -> float32 (H,W,68) dense DAISY, keypoints row-major (daisy i flann.py:69-77)."""
    return sample(cubes(bgr))
