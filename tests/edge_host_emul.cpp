// Host build of the per-pixel arithmetic of csrc/edges_core.cuh (TEST INFRASTRUCTURE): gray + blur, Sobel,
// non-maximum suppression exactly as edges.cu's kernels call them, and a sequential union-find standing in for the
// device hysteresis.  tests/test_oracle_edges.py compares it with the oracle and with cv2 on the CPU.
#include <stdlib.h>

#include <vector>

#include "../lk-s-2022-estimacija-pokreta_b200/csrc/edges_core.cuh"

using namespace flowb200;

static int find_root(std::vector<int32_t>& p, int x) {
  while (p[x] != x) x = p[x] = p[p[x]];
  return x;
}

extern "C" int edge_emul(const uint8_t* bgr, int H, int W, int low, int high, float* out, uint8_t* blur_out) {
  const int n = H * W;
  std::vector<uint8_t> blur(n), pmap(n), strong(n, 0);
  std::vector<int32_t> grad(n), parent(n);
  for (int i = 0; i < n; ++i) blur[i] = edge_gray_blur_at(bgr, H, W, i / W, i % W);
  for (int i = 0; i < n; ++i) grad[i] = edge_sobel_at(blur.data(), H, W, i / W, i % W);
  for (int i = 0; i < n; ++i) {
    pmap[i] = edge_nms_at(grad.data(), H, W, i / W, i % W, low, high);
    parent[i] = pmap[i] != 1 ? i : -1;
  }
  for (int i = 0; i < n; ++i) {
    if (parent[i] < 0) continue;
    const int y = i / W, x = i % W;
    const int ny[4] = {y, y + 1, y + 1, y + 1}, nx[4] = {x + 1, x - 1, x, x + 1};
    for (int t = 0; t < 4; ++t) {
      if (ny[t] >= H || nx[t] < 0 || nx[t] >= W) continue;
      const int j = ny[t] * W + nx[t];
      if (parent[j] < 0) continue;
      const int a = find_root(parent, i), b = find_root(parent, j);
      if (a < b) parent[b] = a; else parent[a] = b;
    }
  }
  for (int i = 0; i < n; ++i)
    if (pmap[i] == 2) strong[find_root(parent, i)] = 1;
  for (int i = 0; i < n; ++i) out[i] = (parent[i] >= 0 && strong[find_root(parent, i)]) ? 0.f : 1.f;
  if (blur_out)
    for (int i = 0; i < n; ++i) blur_out[i] = blur[i];
  return 0;
}
