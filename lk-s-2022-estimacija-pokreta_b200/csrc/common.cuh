// Shared device/host helpers for the flowb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/flowb200.h"

#define FLOWB200_VERSION 100

namespace flowb200 {

constexpr int kDescDim = FLOWB200_DESC_DIM;   // 68
constexpr int kNumSMs = 148;                  // B200

// last CUDA error text, per host thread (flowb200_last_cuda_error)
void set_cuda_error(cudaError_t e, const char* where);

#define FB_CUDA_CHECK(expr)                                   \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) {                                  \
      ::flowb200::set_cuda_error(_e, #expr);                  \
      return FLOWB200_ECUDA;                                  \
    }                                                         \
  } while (0)

// kernel launches enqueued by this process (diagnostics: flowb200_launch_count)
void count_launches(int n);

#define FB_LAUNCH_CHECK_N(n)              \
  do {                                    \
    ::flowb200::count_launches(n);        \
    FB_CUDA_CHECK(cudaGetLastError());    \
  } while (0)
#define FB_LAUNCH_CHECK() FB_LAUNCH_CHECK_N(1)

__host__ __device__ inline int32_t pack_vec(int dy, int dx) {
  return (int32_t)(((uint32_t)dy & 0xffffu) | ((uint32_t)dx << 16));
}
__host__ __device__ inline int vec_dy(int32_t v) { return (int)(int16_t)(v & 0xffff); }
__host__ __device__ inline int vec_dx(int32_t v) { return (int)(v >> 16); }

__device__ __forceinline__ int l1_vec(int dy, int dx, int32_t v) {
  return abs(dy - vec_dy(v)) + abs(dx - vec_dx(v));
}

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

// float32 sum of 68 values in the order numpy's pairwise_sum uses for n < 128
// (8 strided accumulators, ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the n%8 tail):
// what `np.sum(...)` evaluates in daisy i flann.py:178-180 and :228-229.
// f(d) must return element d.  No FMA contraction: explicit round-to-nearest adds.
template <typename F>
__device__ __forceinline__ float sum68_numpy_order(F f) {
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = f(j);
#pragma unroll
  for (int i = 8; i < 64; i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], f(i + j));
  }
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
#pragma unroll
  for (int i = 64; i < 68; ++i) res = __fadd_rn(res, f(i));
  return res;
}

// cells in range of a source coordinate (daisy i flann.py:167-168):
// cell c is searched from coordinate v iff  cell*(c-r) <= v < cell*(c+r+1),  0 <= c < ncell.
__host__ __device__ inline void cell_range(int v, int cell, int ncell, int r, int* cmin, int* cmax) {
  int q = v / cell;
  int lo = q - r;
  int hi = q + r;
  *cmin = lo < 0 ? 0 : lo;
  *cmax = hi > ncell - 1 ? ncell - 1 : hi;   // may be < cmin when v is far in the remainder strip
}

}  // namespace flowb200
