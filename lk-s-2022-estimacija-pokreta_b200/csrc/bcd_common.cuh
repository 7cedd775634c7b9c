// Pieces shared by the two BCD implementations (bcd.cu: K-sets re-evaluated every step, float64 or int32;
// bcd_ksets.cu: K-sets compiled once per proposal set into sparse records, int32).
// Reference: `python bcd.py` ceoBCD() :261-284 (phase order and directions).
#pragma once
#include "common.cuh"

namespace flowb200 {

struct ChainGeom {
  int sy, sx, ystep, xstep, len;
};

__device__ __forceinline__ ChainGeom chain_geom(int phase, int c, int H, int W) {
  ChainGeom g;
  switch (phase) {
    case 0: g = {0, 2 * c, 1, 0, H}; break;           // even columns, downwards   (:265-266)
    case 1: g = {2 * c, W - 1, 0, -1, W}; break;      // even rows, right to left  (:270-271)
    case 2: g = {H - 1, 2 * c + 1, -1, 0, H}; break;  // odd columns, upwards      (:273-274)
    default: g = {2 * c + 1, 0, 0, 1, W}; break;      // odd rows, left to right   (:276-277)
  }
  return g;
}

static inline int phase_chains(int phase, int H, int W) {
  switch (phase) {
    case 0: return (W + 1) / 2;
    case 1: return (H + 1) / 2;
    case 2: return W / 2;
    default: return H / 2;
  }
}

// Spatial hash of flow vectors: bucket edge 2^bshift >= tpsi, 16 x 64 buckets (offset and clamped), so that every
// u with L1(v, u) < tpsi lies in the 3 x 3 buckets around v's bucket.
constexpr int kHashY = 16, kHashX = 64, kHashSize = kHashY * kHashX;
// bucket coordinates are offset and clamped (not wrapped) so that buckets adjacent in x have consecutive keys:
// the leaders of three x-adjacent buckets form ONE contiguous range.  Clamping merges far-away buckets into the
// border ones, which is harmless because membership is always decided by the exact L1 test.
__device__ __forceinline__ int bkt_y(int b) { return min(max(b + kHashY / 2, 0), kHashY - 1); }
__device__ __forceinline__ int bkt_x(int b) { return min(max(b + kHashX / 2, 0), kHashX - 1); }
__device__ __forceinline__ int bucket_key(int32_t v, int bshift) {
  return (bkt_y(vec_dy(v) >> bshift) << 6) | bkt_x(vec_dx(v) >> bshift);
}

// K-set implementation entry (bcd_ksets.cu).  CostT = int32_t (units of 2^-shift) or float (quantised on the fly);
// shift < 0: the float64 programme (python bcd.py's own arithmetic) on float / double data costs.
template <typename CostT>
int launch_sweeps_ksets(const int32_t* pvec, const CostT* cost, const int32_t* nprop, int32_t* labels, int H, int W,
                        int K, double lamda, int tpsi, int shift, int sweeps, int32_t* labels_per_sweep,
                        void* workspace, size_t workspace_bytes, cudaStream_t stream);
// the same in pieces, for callers that exchange labels between phases (single-huge-image mode): part `part` of `nparts`
// owns a contiguous range of every phase's chains
template <typename CostT>
int ksets_prepare(const int32_t* pvec, const CostT* cost, const int32_t* nprop, int H, int W, int K, double lamda,
                  int tpsi, int shift, int part, int nparts, void* workspace, size_t workspace_bytes,
                  cudaStream_t stream);
template <typename CostT>
int ksets_phase(const int32_t* pvec, const CostT* cost, const int32_t* nprop, int32_t* labels, int H, int W, int K,
                double lamda, int tpsi, int shift, int phase, int part, int nparts, void* workspace,
                size_t workspace_bytes, cudaStream_t stream);
size_t ksets_workspace_bytes(int H, int W, int K);
size_t ksets_min_workspace_bytes(int H, int W, int K);   // fixed part only: every K-set is then evaluated densely

}  // namespace flowb200
