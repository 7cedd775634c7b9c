// Per-pixel arithmetic of edge.canny_ivice (edge.py:19-35): cv2.cvtColor(BGR2GRAY) -> cv2.GaussianBlur(3x3, 0) ->
// cv2.Canny(100, 200), all on 8-bit data, i.e. OpenCV's integer arithmetic (restated in oracle/edges.py, where every
// step is pinned against cv2 itself).  Plain C++ without CUDA intrinsics: tests/edge_host_emul.cpp compiles the same
// functions for the host, so the arithmetic is checked against the oracle without a GPU; edges.cu wraps them in
// kernels and adds the hysteresis (union-find over the surviving pixels).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define EDGE_HD __host__ __device__ __forceinline__
#else
#define EDGE_HD static inline
#endif

namespace flowb200 {

constexpr int kCannyShift = 15;
constexpr int kCannyTg22 = 13573;   // (int)(0.41421356237309504 * (1 << 15) + 0.5)

// BORDER_REFLECT_101 (GaussianBlur's default), valid for offsets of one pixel
EDGE_HD int edge_reflect101(int i, int n) {
  if (n == 1) return 0;
  if (i < 0) return -i;
  if (i >= n) return 2 * n - 2 - i;
  return i;
}
EDGE_HD int edge_clamp(int i, int n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); }

// cv2 BGR2GRAY on uint8: 15-bit fixed point (same as the DAISY front end)
EDGE_HD int edge_gray(const uint8_t* bgr) { return (bgr[0] * 3735 + bgr[1] * 19235 + bgr[2] * 9798 + 16384) >> 15; }

// gray + GaussianBlur(3x3, sigma 0): weights 1 2 1 / 2 4 2 / 1 2 1, (sum + 8) >> 4, REFLECT_101
EDGE_HD uint8_t edge_gray_blur_at(const uint8_t* bgr, int H, int W, int y, int x) {
  int s = 0;
  for (int dy = -1; dy <= 1; ++dy) {
    const int yy = edge_reflect101(y + dy, H);
    for (int dx = -1; dx <= 1; ++dx) {
      const int xx = edge_reflect101(x + dx, W);
      s += (2 - (dy != 0)) * (2 - (dx != 0)) * edge_gray(bgr + ((size_t)yy * W + xx) * 3);
    }
  }
  return (uint8_t)((s + 8) >> 4);
}

// Sobel 3x3 with BORDER_REPLICATE (what cv::Canny requests): packed (int16 dx) | (int16 dy) << 16
EDGE_HD int32_t edge_sobel_at(const uint8_t* img, int H, int W, int y, int x) {
  int p[3][3];
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) p[dy + 1][dx + 1] = img[(size_t)edge_clamp(y + dy, H) * W + edge_clamp(x + dx, W)];
  const int gx = (p[0][2] + 2 * p[1][2] + p[2][2]) - (p[0][0] + 2 * p[1][0] + p[2][0]);
  const int gy = (p[2][0] + 2 * p[2][1] + p[2][2]) - (p[0][0] + 2 * p[0][1] + p[0][2]);
  return (int32_t)(((uint32_t)gx & 0xffffu) | ((uint32_t)gy << 16));
}
EDGE_HD int edge_gx(int32_t g) { return (int)(int16_t)(g & 0xffff); }
EDGE_HD int edge_gy(int32_t g) { return (int)(g >> 16); }
EDGE_HD int edge_abs(int v) { return v < 0 ? -v : v; }

// L1 gradient magnitude, zero outside the image (cv::Canny pads its magnitude rows with zeros)
EDGE_HD int edge_mag_at(const int32_t* grad, int H, int W, int y, int x) {
  if (y < 0 || x < 0 || y >= H || x >= W) return 0;
  const int32_t g = grad[(size_t)y * W + x];
  return edge_abs(edge_gx(g)) + edge_abs(edge_gy(g));
}

// Non-maximum suppression + double threshold: 1 = no edge, 0 = survivor (weak), 2 = survivor above `high`.
// The comparisons are OpenCV's, asymmetric on purpose (> towards left/up, >= towards right/down).
EDGE_HD uint8_t edge_nms_at(const int32_t* grad, int H, int W, int y, int x, int low, int high) {
  const int32_t g = grad[(size_t)y * W + x];
  const int xs = edge_gx(g), ys = edge_gy(g);
  const int m = edge_abs(xs) + edge_abs(ys);
  if (!(m > low)) return 1;
  const int ax = edge_abs(xs);
  const int ay = edge_abs(ys) << kCannyShift;
  const int tg22x = ax * kCannyTg22;
  bool keep;
  if (ay < tg22x) {
    keep = m > edge_mag_at(grad, H, W, y, x - 1) && m >= edge_mag_at(grad, H, W, y, x + 1);
  } else {
    const int tg67x = tg22x + (ax << (kCannyShift + 1));
    if (ay > tg67x) {
      keep = m > edge_mag_at(grad, H, W, y - 1, x) && m >= edge_mag_at(grad, H, W, y + 1, x);
    } else {
      const int s = (xs ^ ys) < 0 ? -1 : 1;
      keep = m > edge_mag_at(grad, H, W, y - 1, x - s) && m > edge_mag_at(grad, H, W, y + 1, x + s);
    }
  }
  if (!keep) return 1;
  return m > high ? 2 : 0;
}

}  // namespace flowb200
