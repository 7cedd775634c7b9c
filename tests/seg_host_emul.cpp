// Host build of the device replay logic of removeSmallSegments (csrc/segments_core.cuh), TEST INFRASTRUCTURE.
// tests/test_segments_host.py compiles this file with g++ and compares it with the oracle on the CPU, so that the
// component-granular replay (the part of segments.cu that is not a textbook kernel) is checked without a GPU.
// The component labelling that seg_union_kernel / seg_flatten_kernel do in parallel is done here by a sequential
// union-find with the same result (root = smallest scan index of the component, size at the root).
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../lk-s-2022-estimacija-pokreta_b200/csrc/segments_core.cuh"

using namespace flowb200;

static int find_root(std::vector<int32_t>& p, int x) {
  while (p[x] != x) x = p[x] = p[p[x]];
  return x;
}

extern "C" int seg_emul(float* flow, int A, int B, float tresh, int min_size, int prefilter) {
  const int n = A * B;
  std::vector<float2> fT(n);
  std::vector<int32_t> root(n), size(n, 0), queue(n);
  std::vector<uint8_t> cchk(n, 0), vis(n, 0), touch(n, 0);
  for (int s = 0; s < n; ++s) {
    const float* p = flow + ((size_t)(s % A) * B + s / A) * 3;
    fT[s].x = p[0];
    fT[s].y = p[1];
    root[s] = p[2] > 0.5f ? s : kSegInvalid;
  }
  for (int s = 0; s < n; ++s) {
    if (root[s] < 0) continue;
    const int a = s % A, b = s / A;
    const int cand[2] = {a + 1 < A ? s + 1 : -1, b + 1 < B ? s + A : -1};
    for (int t = 0; t < 2; ++t) {
      const int m = cand[t];
      if (m < 0 || root[m] < 0 || !seg_near(fT[s], fT[m], tresh)) continue;
      int x = find_root(root, s), y = find_root(root, m);
      if (x < y) root[y] = x; else root[x] = y;
    }
  }
  for (int s = 0; s < n; ++s)
    if (root[s] >= 0) { root[s] = find_root(root, s); }
  for (int s = 0; s < n; ++s)
    if (root[s] >= 0) ++size[root[s]];
  SegState S;
  S.A = A; S.B = B; S.tresh = tresh; S.min_size = min_size;
  S.fT = fT.data(); S.root = root.data(); S.size = size.data(); S.cchk = cchk.data(); S.vis = vis.data();
  S.queue = queue.data(); S.flow = flow; S.touch = touch.data();
  if (prefilter) {   // seg_prefilter_invalid_kernel, then seg_prefilter_component_kernel
    for (int s = 0; s < n; ++s)
      if (root[s] < 0 && !seg_invalid_touches(S, s)) root[s] = kSegInvalidChecked;
    for (int s = 0; s < n; ++s)
      if (root[s] == s && seg_component_inert(S, s)) cchk[s] = 1;
  }
  // seg_replay_kernel, one pixel per step instead of 32
  for (int bo = 0; bo < B; ++bo) {
    int b = bo;
    for (int a = 0; a < A; ++a)
      if (seg_unchecked(S, b * A + a)) b = seg_process_seed(S, b * A + a, b);
  }
  return 0;
}
