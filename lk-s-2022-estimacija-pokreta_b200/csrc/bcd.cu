// Block-coordinate-descent MRF inference: row/column chain Viterbi over the K proposal labels.
// Reference: `python bcd.py` bcd() :101-257 (one chain), ceoBCD() :261-284 (phase order),
// sidepsi :84-88, purepsi :98-99.  See oracle/bcd.py for the recurrence in words.
//
// Mapping: one CTA per chain, one thread per label of the current pixel.  All chains of a phase are
// independent (they read and write only their own pixels), so a phase is one launch; the four phases
// and the sweeps are stream-ordered.  K-sets (the reference's packedksets cache, daisy i flann.py:256-309)
// are never materialised: membership L1(v_l, u_k) < tpsi is re-evaluated from the packed vectors held in
// shared memory.
//
// Arithmetic modes (include/flowb200.h): float64 with the reference's operation order (bit-exact for any
// data cost) or int32 in units of 2^-S (bit-exact when lcost == 20*m/2^S).
#include <math_constants.h>

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

template <typename DP>
struct DpOps;

template <>
struct DpOps<int32_t> {
  static __device__ __forceinline__ int32_t inf() { return 0x3fffffff; }
};
template <>
struct DpOps<double> {
  static __device__ __forceinline__ double inf() { return CUDART_INF; }
};

// lexicographic (value, index) minimum across the warp: lowest index wins ties (np.argmin, :155/:175/:234)
template <typename DP>
__device__ __forceinline__ void warp_argmin(DP& val, int& idx) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    DP ov = __shfl_xor_sync(0xffffffffu, val, off);
    int oi = __shfl_xor_sync(0xffffffffu, idx, off);
    if (ov < val || (ov == val && oi < idx)) {
      val = ov;
      idx = oi;
    }
  }
}

// DP = int32_t: CostT = int32_t (m, units of 2^-shift).  DP = double: CostT = float or double (lcost).
template <typename DP, typename CostT>
__global__ void __launch_bounds__(512)
bcd_chain_kernel(const int32_t* __restrict__ pvec, const CostT* __restrict__ cost, const int32_t* __restrict__ nprop,
                 int32_t* __restrict__ labels, uint16_t* __restrict__ bp, int H, int W, int K, int Kpad, int phase,
                 double lamda, int tpsi, int shift) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const ChainGeom g = chain_geom(phase, blockIdx.x, H, W);
  const int l = threadIdx.x;
  const int nwarps = blockDim.x >> 5;
  const int warp = l >> 5, lane = l & 31;

  DP* dp_s = reinterpret_cast<DP*>(smem_raw);                          // [2][Kpad]
  DP* red_val = dp_s + 2 * Kpad;                                       // [2][16]
  int32_t* vec_s = reinterpret_cast<int32_t*>(red_val + 32);           // [2][Kpad]
  int32_t* red_idx = vec_s + 2 * Kpad;                                 // [2][16]
  int32_t* oldvec = red_idx + 32;                                      // [len]

  auto pixel = [&](int i) { return (g.sy + i * g.ystep) * W + (g.sx + i * g.xstep); };

  // vectors of the chain's labels before this call (bestlabels is only rewritten by the backtrack)
  for (int i = l; i < g.len; i += blockDim.x) {
    int p = pixel(i);
    oldvec[i] = pvec[(size_t)p * K + labels[p]];
  }
  __syncthreads();

  const int s = g.ystep + g.xstep;   // +1: image coordinate grows with the step index
  uint16_t* bp_chain = bp + (size_t)blockIdx.x * g.len * Kpad;
  const DP INF = DpOps<DP>::inf();

  // software prefetch of the next pixel's label data
  int pix = pixel(0);
  int n_cur = nprop[pix];
  int32_t v_cur = (l < n_cur) ? pvec[(size_t)pix * K + l] : 0;
  CostT c_cur = (l < n_cur) ? cost[(size_t)pix * K + l] : CostT(0);
  int n_prev = 0;
  DP dpv = INF;

  for (int i = 0; i < g.len; ++i) {
    const int n = n_cur;
    const int32_t v = v_cur;
    const CostT c = c_cur;
    if (i + 1 < g.len) {
      int pn = pixel(i + 1);
      n_cur = nprop[pn];
      v_cur = (l < n_cur) ? pvec[(size_t)pn * K + l] : 0;
      c_cur = (l < n_cur) ? cost[(size_t)pn * K + l] : CostT(0);
    }
    const int cur = i & 1, prv = cur ^ 1;
    const int dy = vec_dy(v), dx = vec_dx(v);

    // side terms (sidepsi :84-88): neighbours along the chain axis, old labels; 0 when off-image
    const int ip = i + s, im = i - s;
    int psi_p = 0, psi_m = 0;
    if (ip >= 0 && ip < g.len) psi_p = min(tpsi, l1_vec(dy, dx, oldvec[ip]));
    if (im >= 0 && im < g.len) psi_m = min(tpsi, l1_vec(dy, dx, oldvec[im]));

    dpv = INF;
    if (l < n) {
      if (i == 0) {
        if constexpr (sizeof(DP) == 8) {   // (psi+ + psi-) + lamda*lcost   (:118-120)
          dpv = __dadd_rn((double)(psi_p + psi_m), __dmul_rn(lamda, (double)c));
        } else {
          dpv = (DP)c + ((psi_p + psi_m) << shift);
        }
      } else {
        // truncation candidate: min_k (tpsi + dp_prev[k]), lowest k (:152-157)
        DP tr = red_val[prv * 16];
        int tr_arg = red_idx[prv * 16];
        for (int w = 1; w < nwarps; ++w) {
          DP ov = red_val[prv * 16 + w];
          int oi = red_idx[prv * 16 + w];
          if (ov < tr || (ov == tr && oi < tr_arg)) {
            tr = ov;
            tr_arg = oi;
          }
        }
        // near candidates: k with L1(v_l, u_k) < tpsi  (the K-set, :131-142; :170-175 / :213-218)
        DP best = INF;
        int arg = 0;
        const DP* dpp = dp_s + prv * Kpad;
        const int32_t* vp = vec_s + prv * Kpad;
#pragma unroll 4
        for (int k = 0; k < n_prev; ++k) {
          int l1 = l1_vec(dy, dx, vp[k]);
          DP cand;
          if constexpr (sizeof(DP) == 8) {
            cand = __dadd_rn(dpp[k], (double)l1);
          } else {
            cand = dpp[k] + (l1 << shift);
          }
          if (l1 < tpsi && cand < best) {   // strict <: lowest k wins ties
            best = cand;
            arg = k;
          }
        }
        DP m = best;
        if (!(best < INF)) {                // quirk Q1: truncation only when the K-set is empty
          m = tr;
          arg = tr_arg;
        }
        if constexpr (sizeof(DP) == 8) {    // (lamda*lcost + psi+) + psi-, then m + that  (:161-162, :176)
          double U = __dadd_rn(__dadd_rn(__dmul_rn(lamda, (double)c), (double)psi_p), (double)psi_m);
          dpv = __dadd_rn(m, U);
        } else {
          dpv = m + (DP)c + ((psi_p + psi_m) << shift);
        }
        bp_chain[(size_t)i * Kpad + l] = (uint16_t)arg;
      }
      dp_s[cur * Kpad + l] = dpv;
      vec_s[cur * Kpad + l] = v;
    }
    // block argmin of (tpsi + dp) for the next step
    DP rv = INF;
    if (l < n) {
      if constexpr (sizeof(DP) == 8) rv = __dadd_rn((double)tpsi, dpv);
      else rv = dpv + (tpsi << shift);
    }
    int ri = l;
    warp_argmin(rv, ri);
    if (lane == 0) {
      red_val[cur * 16 + warp] = rv;
      red_idx[cur * 16 + warp] = ri;
    }
    n_prev = n;
    __syncthreads();
  }

  // final label: lowest-index argmin of dp_last (:231-237), then backtrack (:238-253)
  {
    DP rv = dpv;   // INF for l >= n
    int ri = l;
    warp_argmin(rv, ri);
    const int fin = g.len & 1;   // buffer not used by the last step's reduction (cur = (len-1)&1)
    if (lane == 0) {
      red_val[fin * 16 + warp] = rv;
      red_idx[fin * 16 + warp] = ri;
    }
    __syncthreads();
    if (l == 0) {
      DP bv = red_val[fin * 16];
      int lab = red_idx[fin * 16];
      for (int w = 1; w < nwarps; ++w) {
        DP ov = red_val[fin * 16 + w];
        int oi = red_idx[fin * 16 + w];
        if (ov < bv || (ov == bv && oi < lab)) {
          bv = ov;
          lab = oi;
        }
      }
      for (int i = g.len - 1; i >= 0; --i) {
        labels[pixel(i)] = lab;
        if (i > 0) lab = bp_chain[(size_t)i * Kpad + lab];
      }
    }
  }
}

template <typename DP, typename CostT>
static int launch_sweeps(const int32_t* pvec, const CostT* cost, const int32_t* nprop, int32_t* labels, int H, int W,
                         int K, double lamda, int tpsi, int shift, int sweeps, int32_t* labels_per_sweep, uint16_t* bp,
                         cudaStream_t stream) {
  const int Kpad = (K + 31) / 32 * 32;
  auto kern = bcd_chain_kernel<DP, CostT>;
  const int maxlen = H > W ? H : W;
  size_t smem = 2 * (size_t)Kpad * (sizeof(DP) + 4) + 32 * (sizeof(DP) + 4) + (size_t)maxlen * 4;
  if (smem > 48 * 1024) FB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int w = 0; w < sweeps; ++w) {
    for (int phase = 0; phase < 4; ++phase) {
      int nch = phase_chains(phase, H, W);
      if (nch == 0) continue;
      kern<<<nch, Kpad, smem, stream>>>(pvec, cost, nprop, labels, bp, H, W, K, Kpad, phase, lamda, tpsi, shift);
      FB_LAUNCH_CHECK();
    }
    if (labels_per_sweep)
      FB_CUDA_CHECK(cudaMemcpyAsync(labels_per_sweep + (size_t)w * H * W, labels, sizeof(int32_t) * H * W,
                                    cudaMemcpyDeviceToDevice, stream));
  }
  return FLOWB200_OK;
}

__global__ void quantise_costs_kernel(const float* __restrict__ lcost, int32_t* __restrict__ m, size_t n, double lamda,
                                      double scale) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float c = lcost[i];
  m[i] = (c == FLOWB200_UNUSED_COST) ? 0 : (int32_t)rint(__dmul_rn(__dmul_rn(lamda, (double)c), scale));
}

}  // namespace flowb200

using namespace flowb200;

extern "C" size_t flowb200_bcd_workspace_bytes(int H, int W, int K) {
  if (H <= 0 || W <= 0 || K <= 0) return 0;
  size_t Kpad = (size_t)(K + 31) / 32 * 32;
  size_t col = (size_t)((W + 1) / 2) * H, row = (size_t)((H + 1) / 2) * W;
  return align_up((col > row ? col : row) * Kpad * sizeof(uint16_t));
}

extern "C" int flowb200_bcd(const int32_t* pvec, const void* cost, const int32_t* nprop, int32_t* labels, int H, int W,
                            int K, int bcd_mode, double lamda, int tpsi, int cost_shift, int sweeps,
                            int32_t* labels_per_sweep, void* workspace, size_t workspace_bytes,
                            flowb200_stream_t stream) {
  if (!pvec || !cost || !nprop || !labels || !workspace) return FLOWB200_EINVAL;
  if (H <= 0 || W <= 0 || K <= 0 || K > 512 || sweeps < 0 || tpsi < 0) return FLOWB200_EINVAL;
  if (H > 16384 || W > 16384) return FLOWB200_EINVAL;
  if (workspace_bytes < flowb200_bcd_workspace_bytes(H, W, K)) return FLOWB200_EWORKSPACE;
  uint16_t* bp = static_cast<uint16_t*>(workspace);
  switch (bcd_mode) {
    case FLOWB200_BCD_FP64_F32COST:
      return launch_sweeps<double, float>(pvec, static_cast<const float*>(cost), nprop, labels, H, W, K, lamda, tpsi, 0,
                                          sweeps, labels_per_sweep, bp, stream);
    case FLOWB200_BCD_FP64_F64COST:
      return launch_sweeps<double, double>(pvec, static_cast<const double*>(cost), nprop, labels, H, W, K, lamda, tpsi,
                                           0, sweeps, labels_per_sweep, bp, stream);
    case FLOWB200_BCD_INT32:
      if (cost_shift < 0 || cost_shift > 14) return FLOWB200_EINVAL;
      return launch_sweeps<int32_t, int32_t>(pvec, static_cast<const int32_t*>(cost), nprop, labels, H, W, K, lamda,
                                             tpsi, cost_shift, sweeps, labels_per_sweep, bp, stream);
    default:
      return FLOWB200_EINVAL;
  }
}

extern "C" int flowb200_quantise_costs(const float* lcost, int32_t* m, size_t n, double lamda, int shift,
                                       flowb200_stream_t stream) {
  if (!lcost || !m || shift < 0 || shift > 14) return FLOWB200_EINVAL;
  if (n == 0) return FLOWB200_OK;
  quantise_costs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(lcost, m, n, lamda, (double)(1 << shift));
  FB_LAUNCH_CHECK();
  return FLOWB200_OK;
}
