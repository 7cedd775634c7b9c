// Block-coordinate-descent MRF inference: row/column chain Viterbi over the K proposal labels.
// Reference: `python bcd.py` bcd() :101-257 (one chain), ceoBCD() :261-284 (phase order),
// sidepsi :84-88, purepsi :98-99.  See oracle/bcd.py for the recurrence in words.
//
// Mapping: one CTA per chain, one thread per label of the current pixel.  All chains of a phase are
// independent (they read and write only their own pixels), so a phase is one launch; the four phases
// and the sweeps are stream-ordered.  K-sets (the reference's packedksets cache, daisy i flann.py:256-309)
// are never materialised: labels are processed in the order of a spatial hash of their flow vectors and
// membership L1(v_l, u_k) < tpsi is re-evaluated exactly on the 3 x 3 buckets around v_l (see below).
//
// Arithmetic modes (include/flowb200.h): float64 with the reference's operation order (bit-exact for any
// data cost) or int32 in units of 2^-S (bit-exact when lcost == 20*m/2^S).
#include <math_constants.h>
#include <stdlib.h>
#include <type_traits>

#include "bcd_common.cuh"

namespace flowb200 {

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
__device__ __forceinline__ void warp_argmin(int32_t& val, int& idx) {
  const int m = __reduce_min_sync(0xffffffffu, val);
  idx = __reduce_min_sync(0xffffffffu, val == m ? idx : 0x7fffffff);
  val = m;
}
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
//
// The K-set S_l = {k : L1(v_l, u_k) < tpsi} is found through a spatial hash instead of a dense K x K scan, and
// labels that carry the same flow vector are evaluated once.
//
// bcd_sort_kernel (once per proposal set; the sets are static across phases and sweeps, which is what the
// reference's packedksets cache exploits) orders every pixel's labels by (bucket of the flow vector, vector)
// with bucket edge 2^bshift >= tpsi and 16 x 64 buckets (offset and clamped).  Labels with equal vectors become a
// contiguous RUN; its first label is the run's LEADER and leaders are numbered 0..D-1 in sorted order.
//
// The chain kernel works on runs:
//   * the pairwise minimum m(l) depends on v_l only, so it is computed once per leader;
//   * of the previous pixel's labels with equal vector only the one with the smallest (dp, original index)
//     can be the (lowest-index) minimiser, so each run is represented by that entry;
//   * the leaders of one bucket are contiguous, so S_l lies in the 3 x 3 bucket ranges around v_l of the
//     previous pixel's per-leader arrays.
// The hash only prunes: membership is always decided by the exact L1 test, and ties are resolved to the lowest
// ORIGINAL label index as in the reference (np.argmin), so labels are bit-identical to the dense evaluation.
// per-pixel order entry: [0,10) original label, [10,20) leader number of the run, 20 leader, 21 first leader of
// its bucket, 22 last leader of its bucket, [23,32) run length - 1 (leaders only)
constexpr uint32_t kOrdLeader = 1u << 20, kOrdBStart = 1u << 21, kOrdBEnd = 1u << 22;

constexpr uint32_t kRngEmpty = 0x0000FFFFu;   // first = 0xFFFF, last+1 = 0

// stable counting sort of idx_in[0..n) by key(idx) into idx_out (one warp; cnt has nkeys + nkeys/32 entries:
// counter k lives at k + (k >> 5) so that the per-lane scan of 32 consecutive counters is bank-conflict free)
__device__ __forceinline__ int cidx(int k) { return k + (k >> 5); }

template <typename KeyFn>
__device__ __forceinline__ void warp_stable_sort(const uint16_t* idx_in, uint16_t* idx_out, int n, int* cnt, int nkeys,
                                                 int lane, KeyFn key) {
  const int per = nkeys >> 5;
  for (int i = 0; i < per; ++i) cnt[cidx(i * 32 + lane)] = 0;
  __syncwarp();
  for (int j = lane; j < n; j += 32) atomicAdd(&cnt[cidx(key(idx_in[j]))], 1);
  __syncwarp();
  // exclusive scan of cnt: every lane owns nkeys/32 consecutive counters
  int sum = 0;
  for (int i = 0; i < per; ++i) sum += cnt[cidx(lane * per + i)];
  int pre = sum;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, pre, off);
    if (lane >= off) pre += t;
  }
  pre -= sum;
  __syncwarp();
  for (int i = 0; i < per; ++i) {
    const int c = cnt[cidx(lane * per + i)];
    cnt[cidx(lane * per + i)] = pre;
    pre += c;
  }
  __syncwarp();
  // stable scatter, 32 elements at a time in input order
  for (int j0 = 0; j0 < n; j0 += 32) {
    const int j = j0 + lane;
    const bool act = j < n;
    const int id = act ? idx_in[j] : 0;
    const int k = act ? key(id) : -1 - lane;        // inactive lanes get unique keys
    const unsigned same = __match_any_sync(0xffffffffu, k);
    const int rank = __popc(same & ((1u << lane) - 1));
    int base = 0;
    if (act) base = cnt[cidx(k)];
    __syncwarp();
    if (act) {
      idx_out[base + rank] = (uint16_t)id;
      if (rank == __popc(same) - 1) cnt[cidx(k)] = base + rank + 1;   // last lane of the key group advances the counter
    }
    __syncwarp();
  }
}

// one warp per pixel
__global__ void __launch_bounds__(128)
bcd_sort_kernel(const int32_t* __restrict__ pvec, const int32_t* __restrict__ nprop, int npix, int K, int bshift,
                uint32_t* __restrict__ order) {
  __shared__ int cnt_s[4][kHashSize + kHashSize / 32];
  __shared__ uint16_t idx_s[4][2][512];
  __shared__ uint16_t lpos_s[4][513];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* cnt = cnt_s[warp];
  uint16_t* ia = idx_s[warp][0];
  uint16_t* ib = idx_s[warp][1];
  uint16_t* lpos = lpos_s[warp];
  const int fmask = (1 << bshift) - 1;
  for (int pix = blockIdx.x * 4 + warp; pix < npix; pix += gridDim.x * 4) {
    const int n = nprop[pix];
    const int32_t* v = pvec + (size_t)pix * K;
    uint32_t* o = order + (size_t)pix * K;
    for (int j = lane; j < n; j += 32) ia[j] = (uint16_t)j;
    __syncwarp();
    // LSD: position inside the bucket first, then the bucket
    const int nfine = max(32, 1 << (2 * bshift));
    warp_stable_sort(ia, ib, n, cnt, nfine, lane,
                     [&](int id) { return ((vec_dy(v[id]) & fmask) << bshift) | (vec_dx(v[id]) & fmask); });
    warp_stable_sort(ib, ia, n, cnt, kHashSize, lane, [&](int id) { return bucket_key(v[id], bshift); });
    // leaders, leader numbers
    int nlead = 0;
    for (int j0 = 0; j0 < n; j0 += 32) {
      const int j = j0 + lane;
      bool lead = false;
      if (j < n) lead = j == 0 || v[ia[j]] != v[ia[j - 1]];
      const unsigned m = __ballot_sync(0xffffffffu, lead);
      const int r = nlead + __popc(m & ((1u << lane) - 1));     // leaders before this position
      if (j < n) {
        ib[j] = (uint16_t)(lead ? r : r - 1);                   // leader number of the run this position is in
        if (lead) lpos[r] = (uint16_t)j;
      }
      nlead += __popc(m);
    }
    if (lane == 0) lpos[nlead] = (uint16_t)n;
    __syncwarp();
    for (int j = lane; j < n; j += 32) {
      const int id = ia[j], r = ib[j];
      uint32_t e = (uint32_t)id | ((uint32_t)r << 10);
      if (lpos[r] == j) {
        e |= kOrdLeader | ((uint32_t)(lpos[r + 1] - j - 1) << 23);
        const int key = bucket_key(v[id], bshift);
        if (j == 0 || bucket_key(v[ia[j - 1]], bshift) != key) e |= kOrdBStart;
        if (r == nlead - 1 || bucket_key(v[ia[lpos[r + 1]]], bshift) != key) e |= kOrdBEnd;
      }
      o[j] = e;
    }
    __syncwarp();
  }
}

// data term in DP units.  int32 programme on float32 costs: quantised on the fly exactly as flowb200_quantise_costs
// does, m = rint(lamda * lcost * 2^S) in float64
template <typename DP, typename CostT>
__device__ __forceinline__ DP int_cost(CostT c, double lamda, int shift) {
  if constexpr (sizeof(CostT) == 4 && !std::is_integral<CostT>::value)
    return (DP)rint(__dmul_rn(__dmul_rn(lamda, (double)c), (double)(1 << shift)));
  else
    return (DP)c;
}

template <typename DP> struct RepT;
template <> struct RepT<int32_t> { using type = unsigned long long; };   // dp << 32 | original label
template <> struct RepT<double> { using type = double2; };               // (dp, original label)

// One CTA of T threads per chain; thread l handles the sorted positions l, l+T, ... (PER of them).  Few warps per
// chain keep the two barriers of a step cheap and let one thread average over several labels (the per-step time
// is set by the slowest thread), while all chains of a phase -- and both directions -- are resident at once.
template <typename DP, typename CostT, int T, int PER, int MINB>
__global__ void __launch_bounds__(T, MINB)
bcd_chain_kernel(const int32_t* __restrict__ pvec, const CostT* __restrict__ cost, const int32_t* __restrict__ nprop,
                 const uint32_t* __restrict__ order, int32_t* __restrict__ labels, uint16_t* __restrict__ bp, int H,
                 int W, int K, int Kpad, int phase, double lamda, int tpsi, int shift, int bshift) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using Rep = typename RepT<DP>::type;
  constexpr int NW = T / 32;
  const ChainGeom g = chain_geom(phase, blockIdx.x, H, W);
  const int l = threadIdx.x;
  const int warp = l >> 5, lane = l & 31;

  Rep* rep_s = reinterpret_cast<Rep*>(smem_raw);                        // [2][Kpad] per leader: best (dp, label) of the run
  DP* dp_s = reinterpret_cast<DP*>(rep_s + 2 * Kpad);                   // [Kpad]    per position (double mode: run scan)
  DP* m_s = dp_s + Kpad;                                                // [Kpad]    per leader: pairwise minimum
  DP* red_val = m_s + Kpad;                                             // [2][16]
  int32_t* vrep_s = reinterpret_cast<int32_t*>(red_val + 32);           // [2][Kpad] per leader: flow vector
  int32_t* arg_s = vrep_s + 2 * Kpad;                                   // [Kpad]    per leader: back-pointer
  int32_t* org_s = arg_s + Kpad;                                        // [Kpad]    per position (double mode)
  int32_t* red_idx = org_s + Kpad;                                      // [2][16]
  uint32_t* rng_s = reinterpret_cast<uint32_t*>(red_idx + 32);          // [3][kHashSize] first | (last+1) << 16 leader
  int32_t* oldvec = reinterpret_cast<int32_t*>(rng_s + 3 * kHashSize);  // [len]
  __shared__ int nlead_s[2];

  auto pixel = [&](int i) { return (g.sy + i * g.ystep) * W + (g.sx + i * g.xstep); };

  // vectors of the chain's labels before this call (bestlabels is only rewritten by the backtrack)
  for (int i = l; i < g.len; i += T) {
    int p = pixel(i);
    oldvec[i] = pvec[(size_t)p * K + labels[p]];
  }
  for (int i = l; i < 3 * kHashSize; i += T) rng_s[i] = kRngEmpty;
  __syncthreads();

  const int s = g.ystep + g.xstep;   // +1: image coordinate grows with the step index
  uint16_t* bp_chain = bp + (size_t)blockIdx.x * g.len * Kpad;
  const DP INF = DpOps<DP>::inf();

  // software prefetch, three stages deep so that no load waits for another one's result inside a step:
  //   stage c (pixel i+3): nprop and the order entries (independent loads)
  //   stage b (pixel i+2): vectors and costs, addressed through the order entries fetched one step earlier
  //   stage a (pixel i+1): complete; published to shared memory during step i
  int n_a = 0, n_b = 0, n_c = 0;
  uint32_t o_a[PER], o_b[PER], o_c[PER];
  int32_t v_a[PER], v_b[PER];
  CostT c_a[PER], c_b[PER];
  const ptrdiff_t pstep = (ptrdiff_t)g.ystep * W + g.xstep;        // pixel index increment per chain step
  const int32_t* f_np = nprop + pixel(0);
  const uint32_t* f_or = order + (size_t)pixel(0) * K + l;
  const int32_t* f_pv = pvec + (size_t)pixel(0) * K;
  const CostT* f_co = cost + (size_t)pixel(0) * K;
  auto fetch_order = [&](int i, int& n, uint32_t (&o)[PER]) {      // i = 0, 1, 2, ... in order
    n = 0;
#pragma unroll
    for (int r = 0; r < PER; ++r) o[r] = 0;
    if (i < g.len) {
      n = *f_np;
#pragma unroll
      for (int r = 0; r < PER; ++r)
        if (l + r * T < K) o[r] = f_or[r * T];                     // entries >= nprop are unused (masked below)
      f_np += pstep;
      f_or += pstep * K;
    }
  };
  auto fetch_data = [&](int i, int n, const uint32_t (&o)[PER], int32_t (&v)[PER], CostT (&c)[PER]) {
#pragma unroll
    for (int r = 0; r < PER; ++r) { v[r] = 0; c[r] = CostT(0); }
    if (i < g.len) {
#pragma unroll
      for (int r = 0; r < PER; ++r)
        if (l + r * T < n) {
          v[r] = f_pv[o[r] & 1023];
          c[r] = f_co[o[r] & 1023];
        }
      f_pv += pstep * K;
      f_co += pstep * K;
    }
  };
  fetch_order(0, n_a, o_a);
  fetch_order(1, n_b, o_b);
  fetch_order(2, n_c, o_c);
  fetch_data(0, n_a, o_a, v_a, c_a);
  fetch_data(1, n_b, o_b, v_b, c_b);
  DP dpv[PER];
  int orig[PER];
  int key0[PER], key1[PER], key2[PER];   // buckets opened 0 / 1 / 2 builds ago (emptied after their query)
#pragma unroll
  for (int r = 0; r < PER; ++r) { dpv[r] = INF; orig[r] = 0; key0[r] = key1[r] = key2[r] = -1; }

  // Leaders of pixel `i` publish their flow vector, open their bucket's leader range and reset their run's
  // representative.  Runs one step ahead (from the prefetched registers), so it needs no barrier of its own.
  auto publish = [&](int i, int n, const uint32_t (&of)[PER], const int32_t (&v)[PER]) {
    const int buf = i & 1;
    uint32_t* rng_b = rng_s + (i % 3) * kHashSize;
#pragma unroll
    for (int r = 0; r < PER; ++r) {
      const int p = l + r * T;
      if (p < n && (of[r] & kOrdLeader)) {
        const int run = (int)((of[r] >> 10) & 1023);
        vrep_s[buf * Kpad + run] = v[r];
        if constexpr (sizeof(DP) == 4) rep_s[buf * Kpad + run] = ~0ull;
        const int key = bucket_key(v[r], bshift);
        uint16_t* half = reinterpret_cast<uint16_t*>(rng_b + key);
        if (of[r] & kOrdBStart) {
          half[0] = (uint16_t)run;
          key0[r] = key;
        }
        if (of[r] & kOrdBEnd) half[1] = (uint16_t)(run + 1);
      }
      if (p < n && p == n - 1) nlead_s[buf] = (int)((of[r] >> 10) & 1023) + 1;
    }
  };
  publish(0, n_a, o_a, v_a);
  __syncthreads();

  for (int i = 0; i < g.len; ++i) {
    const int n = n_a;
    uint32_t of[PER];
    int32_t v[PER];
    CostT c[PER];
#pragma unroll
    for (int r = 0; r < PER; ++r) {
      of[r] = o_a[r]; v[r] = v_a[r]; c[r] = c_a[r];
      o_a[r] = o_b[r]; v_a[r] = v_b[r]; c_a[r] = c_b[r];
      o_b[r] = o_c[r];
    }
    n_a = n_b;
    n_b = n_c;
    fetch_data(i + 2, n_b, o_b, v_b, c_b);       // addresses come from order entries loaded a step ago
    fetch_order(i + 3, n_c, o_c);
    const int cur = i & 1, prv = cur ^ 1;
    uint32_t* rng_q = rng_s + ((i + 2) % 3) * kHashSize;   // ranges of pixel i-1, queried now
    uint32_t* rng_c = rng_s + ((i + 1) % 3) * kHashSize;   // ranges of pixel i-2 (queried at step i-1): emptied now,
                                                           // refilled for pixel i+1 after this step's first barrier
#pragma unroll
    for (int r = 0; r < PER; ++r) {
      if (key2[r] >= 0) rng_c[key2[r]] = kRngEmpty;
      key2[r] = key1[r];
      key1[r] = key0[r];
      key0[r] = -1;
      orig[r] = (int)(of[r] & 1023);
    }

    // ---- phase 2: leader j's pairwise minimum against the previous pixel's runs (threads take leaders j, j+T, ..)
    if (i > 0) {
      const int D = nlead_s[cur];
      const Rep* rp = rep_s + prv * Kpad;
      const int32_t* vp = vrep_s + prv * Kpad;
#pragma unroll 1
      for (int L = l; L < D; L += T) {
        const int32_t vq = vrep_s[cur * Kpad + L];
        const int dy = vec_dy(vq), dx = vec_dx(vq);
        const int by = dy >> bshift, bx = dx >> bshift;
        DP best = INF;
        int arg = 0x7fffffff;
        unsigned long long key_a = ~0ull, key_b = ~0ull, key_c = ~0ull, key_d = ~0ull;
        (void)key_a; (void)key_b; (void)key_c; (void)key_d;
        const int kx0 = bkt_x(bx - 1), kx1 = bkt_x(bx), kx2 = bkt_x(bx + 1);
#pragma unroll 1
        for (int oy = -1; oy <= 1; ++oy) {
          // the three x-adjacent buckets of this bucket row are one contiguous leader range
          const uint32_t* row = rng_q + (bkt_y(by + oy) << 6);
          const uint32_t r0 = row[kx0], r1 = row[kx1], r2 = row[kx2];
          const int t0 = (int)min(min(r0 & 0xFFFFu, r1 & 0xFFFFu), r2 & 0xFFFFu);
          const int t1 = (int)max(max(r0 >> 16, r1 >> 16), r2 >> 16);
          if constexpr (sizeof(DP) == 4) {
            // int32 mode: (dp << 32 | label) keys make "lowest label wins ties" a plain 64-bit minimum; four
            // independent accumulators keep four candidates in flight
            for (int t = t0; t < t1; t += 4) {
              // groups of four; indices past the end are clamped (re-evaluating a candidate is harmless for a min)
              const int tl = t1 - 1;
              const int ib = min(t + 1, tl), ic = min(t + 2, tl), id = min(t + 3, tl);
              const int32_t ua = vp[t], ub = vp[ib], uc = vp[ic], ud = vp[id];
              const unsigned long long ra = rp[t], rb = rp[ib], rc = rp[ic], rd = rp[id];
              const int la = l1_vec(dy, dx, ua), lb = l1_vec(dy, dx, ub), lc = l1_vec(dy, dx, uc),
                        ld = l1_vec(dy, dx, ud);
              const unsigned long long ka = ra + ((unsigned long long)(uint32_t)(la << shift) << 32);
              const unsigned long long kb = rb + ((unsigned long long)(uint32_t)(lb << shift) << 32);
              const unsigned long long kc = rc + ((unsigned long long)(uint32_t)(lc << shift) << 32);
              const unsigned long long kd = rd + ((unsigned long long)(uint32_t)(ld << shift) << 32);
              if (la < tpsi && ka < key_a) key_a = ka;    // near candidates (the K-set, :131-142; :170-175 / :213-218)
              if (lb < tpsi && kb < key_b) key_b = kb;
              if (lc < tpsi && kc < key_c) key_c = kc;
              if (ld < tpsi && kd < key_d) key_d = kd;
            }
          } else {
            for (int t = t0; t < t1; ++t) {
              const int l1 = l1_vec(dy, dx, vp[t]);
              if (l1 < tpsi) {
                const double2 r = rp[t];
                const double cand = __dadd_rn(r.x, (double)l1);
                const int k = (int)r.y;
                if (cand < best || (cand == best && k < arg)) {   // np.argmin: lowest k wins ties
                  best = cand;
                  arg = k;
                }
              }
            }
          }
        }
        if constexpr (sizeof(DP) == 4) {
          const unsigned long long kab = key_a < key_b ? key_a : key_b, kcd = key_c < key_d ? key_c : key_d;
          const unsigned long long kk = kab < kcd ? kab : kcd;
          if (kk != ~0ull) {
            best = (DP)(kk >> 32);
            arg = (int)(kk & 0xffffffffu);
          }
        }
        if (arg == 0x7fffffff) {            // quirk Q1: truncation only when the K-set is empty
          // min_k (tpsi + dp_prev[k]), lowest k (:152-157)
          DP tr = red_val[prv * 16];
          int tr_arg = red_idx[prv * 16];
#pragma unroll
          for (int w = 1; w < NW; ++w) {
            DP ov = red_val[prv * 16 + w];
            int oi = red_idx[prv * 16 + w];
            if (ov < tr || (ov == tr && oi < tr_arg)) {
              tr = ov;
              tr_arg = oi;
            }
          }
          best = tr;
          arg = tr_arg;
        }
        m_s[L] = best;
        arg_s[L] = arg;
      }
      __syncthreads();
    }

    // ---- phase 3: every label adds its own unary term; runs elect their representative
    DP rv = INF;                       // running (tpsi + dp, label) minimum over this thread's labels
    int ri = 0x7fffffff;
#pragma unroll
    for (int r = 0; r < PER; ++r) {
      const int p = l + r * T;
      dpv[r] = INF;
      if (p < n) {
        const int run = (int)((of[r] >> 10) & 1023);
        const int dy = vec_dy(v[r]), dx = vec_dx(v[r]);
        // side terms (sidepsi :84-88): neighbours along the chain axis, old labels; 0 when off-image
        const int ip = i + s, im = i - s;
        int psi_p = 0, psi_m = 0;
        if (ip >= 0 && ip < g.len) psi_p = min(tpsi, l1_vec(dy, dx, oldvec[ip]));
        if (im >= 0 && im < g.len) psi_m = min(tpsi, l1_vec(dy, dx, oldvec[im]));
        if (i == 0) {
          if constexpr (sizeof(DP) == 8) {   // (psi+ + psi-) + lamda*lcost   (:118-120)
            dpv[r] = __dadd_rn((double)(psi_p + psi_m), __dmul_rn(lamda, (double)c[r]));
          } else {
            dpv[r] = int_cost<DP, CostT>(c[r], lamda, shift) + ((psi_p + psi_m) << shift);
          }
        } else {
          const DP m = m_s[run];
          if constexpr (sizeof(DP) == 8) {    // (lamda*lcost + psi+) + psi-, then m + that  (:161-162, :176)
            double U = __dadd_rn(__dadd_rn(__dmul_rn(lamda, (double)c[r]), (double)psi_p), (double)psi_m);
            dpv[r] = __dadd_rn(m, U);
          } else {
            dpv[r] = m + int_cost<DP, CostT>(c[r], lamda, shift) + ((psi_p + psi_m) << shift);
          }
          bp_chain[(size_t)i * Kpad + orig[r]] = (uint16_t)arg_s[run];
        }
        if constexpr (sizeof(DP) == 4) {
          const unsigned long long key = ((unsigned long long)(uint32_t)dpv[r] << 32) | (uint32_t)orig[r];
          if ((of[r] & kOrdLeader) && (of[r] >> 23) == 0) rep_s[cur * Kpad + run] = key;      // run of one label
          else atomicMin(rep_s + cur * Kpad + run, key);
        } else {
          dp_s[p] = dpv[r];
          org_s[p] = orig[r];
        }
        DP tv;
        if constexpr (sizeof(DP) == 8) tv = __dadd_rn((double)tpsi, dpv[r]);
        else tv = dpv[r] + (tpsi << shift);
        if (tv < rv || (tv == rv && orig[r] < ri)) {
          rv = tv;
          ri = orig[r];
        }
      }
    }
    if (i + 1 < g.len) publish(i + 1, n_a, o_a, v_a);
    // block argmin of (tpsi + dp) for the next step's truncation candidate, ties -> lowest original index
    warp_argmin(rv, ri);
    if (lane == 0) {
      red_val[cur * 16 + warp] = rv;
      red_idx[cur * 16 + warp] = ri;
    }
    if constexpr (sizeof(DP) == 8) {
      // float64 mode (parity / arbitrary costs): the leader scans its run for the representative
      __syncthreads();
#pragma unroll
      for (int r = 0; r < PER; ++r) {
        const int p = l + r * T;
        if (p < n && (of[r] & kOrdLeader)) {
          const int run = (int)((of[r] >> 10) & 1023);
          const int len = (int)(of[r] >> 23) + 1;
          double bd = dpv[r];
          int bo = orig[r];
          for (int t = 1; t < len; ++t) {
            const double od = dp_s[p + t];
            const int oo = org_s[p + t];
            if (od < bd || (od == bd && oo < bo)) {
              bd = od;
              bo = oo;
            }
          }
          rep_s[cur * Kpad + run] = make_double2(bd, (double)bo);
        }
      }
    }
    __syncthreads();
  }

  // final label: lowest-index argmin of dp_last (:231-237), then backtrack (:238-253)
  {
    DP rv = INF;
    int ri = 0x7fffffff;
#pragma unroll
    for (int r = 0; r < PER; ++r)
      if (dpv[r] < rv || (dpv[r] == rv && dpv[r] < INF && orig[r] < ri)) {
        rv = dpv[r];
        ri = orig[r];
      }
    if (!(rv < INF)) ri = 0x7fffffff;
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
      for (int w = 1; w < NW; ++w) {
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
                         uint32_t* order, cudaStream_t stream) {
  const int Kpad = (K + 31) / 32 * 32;
  // one label per thread: the thread count is Kpad rounded up to the next instantiated size
  void (*kern)(const int32_t*, const CostT*, const int32_t*, const uint32_t*, int32_t*, uint16_t*, int, int, int, int,
               int, double, int, int, int) = nullptr;
  int T = 0;
#define FB_BCD_CASE(TT, MB) if (!kern && K <= TT) { kern = bcd_chain_kernel<DP, CostT, TT, 1, MB>; T = TT; }
  FB_BCD_CASE(64, 8) FB_BCD_CASE(128, 6) FB_BCD_CASE(192, 5) FB_BCD_CASE(256, 4) FB_BCD_CASE(320, 4)
  FB_BCD_CASE(384, 3) FB_BCD_CASE(512, 2)
#undef FB_BCD_CASE
  if (!kern) return FLOWB200_EUNSUPPORTED;
  const int maxlen = H > W ? H : W;
  size_t smem = (size_t)Kpad * (2 * sizeof(typename RepT<DP>::type) + 2 * sizeof(DP) + 16) + 32 * (sizeof(DP) + 4) +
                3 * (size_t)kHashSize * 4 + (size_t)maxlen * 4;
  FB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int bshift = 0;
  while ((1 << bshift) < tpsi) ++bshift;
  if (bshift > 5) return FLOWB200_EUNSUPPORTED;
  if (sweeps > 0) {
    const int npix = H * W;
    bcd_sort_kernel<<<min((npix + 3) / 4, 16 * kNumSMs), 128, 0, stream>>>(pvec, nprop, npix, K, bshift, order);
    FB_LAUNCH_CHECK();
  }
  for (int w = 0; w < sweeps; ++w) {
    for (int phase = 0; phase < 4; ++phase) {
      int nch = phase_chains(phase, H, W);
      if (nch == 0) continue;
      kern<<<nch, T, smem, stream>>>(pvec, cost, nprop, order, labels, bp, H, W, K, Kpad, phase, lamda, tpsi, shift,
                                     bshift);
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

static size_t legacy_workspace_bytes(int H, int W, int K) {
  size_t Kpad = (size_t)(K + 31) / 32 * 32;
  size_t col = (size_t)((W + 1) / 2) * H, row = (size_t)((H + 1) / 2) * W;
  return align_up((col > row ? col : row) * Kpad * sizeof(uint16_t)) + align_up((size_t)H * W * K * sizeof(uint32_t));
}

// Every mode runs on compiled K-sets (bcd_ksets.cu) unless FLOWB200_BCD_LEGACY is set in the environment (read at every
// call: the tests use it to compare the two implementations) or tpsi > 8; the implementation in this file re-evaluates
// the K-sets at every chain step.
static bool use_ksets(int tpsi, int cost_shift) {
  return getenv("FLOWB200_BCD_LEGACY") == nullptr && tpsi >= 1 && tpsi <= 8 && cost_shift >= 4;
}

extern "C" size_t flowb200_bcd_workspace_bytes(int H, int W, int K) {
  if (H <= 0 || W <= 0 || K <= 0) return 0;
  const size_t a = legacy_workspace_bytes(H, W, K), b = ksets_workspace_bytes(H, W, K);
  return a > b ? a : b;
}

extern "C" size_t flowb200_bcd_min_workspace_bytes(int H, int W, int K) {
  if (H <= 0 || W <= 0 || K <= 0) return 0;
  const size_t a = legacy_workspace_bytes(H, W, K), b = ksets_min_workspace_bytes(H, W, K);
  return a > b ? a : b;
}

extern "C" int flowb200_bcd(const int32_t* pvec, const void* cost, const int32_t* nprop, int32_t* labels, int H, int W,
                            int K, int bcd_mode, double lamda, int tpsi, int cost_shift, int sweeps,
                            int32_t* labels_per_sweep, void* workspace, size_t workspace_bytes,
                            flowb200_stream_t stream) {
  if (!pvec || !cost || !nprop || !labels || !workspace) return FLOWB200_EINVAL;
  if (H <= 0 || W <= 0 || K <= 0 || K > 512 || sweeps < 0 || tpsi < 0) return FLOWB200_EINVAL;
  if (H > 16384 || W > 16384) return FLOWB200_EINVAL;
  const bool int_mode = bcd_mode == FLOWB200_BCD_INT32 || bcd_mode == FLOWB200_BCD_INT32_F32COST;
  if (int_mode && use_ksets(tpsi, cost_shift)) {
    // any workspace that holds the fixed part is accepted: records that do not fit are evaluated densely
    if (cost_shift < 0 || cost_shift > 14) return FLOWB200_EINVAL;
    if (bcd_mode == FLOWB200_BCD_INT32)
      return launch_sweeps_ksets<int32_t>(pvec, static_cast<const int32_t*>(cost), nprop, labels, H, W, K, lamda, tpsi,
                                          cost_shift, sweeps, labels_per_sweep, workspace, workspace_bytes, stream);
    return launch_sweeps_ksets<float>(pvec, static_cast<const float*>(cost), nprop, labels, H, W, K, lamda, tpsi,
                                      cost_shift, sweeps, labels_per_sweep, workspace, workspace_bytes, stream);
  }
  if (!int_mode && use_ksets(tpsi, 4)) {   // the float64 programme on the same records (shift -1)
    if (bcd_mode == FLOWB200_BCD_FP64_F32COST)
      return launch_sweeps_ksets<float>(pvec, static_cast<const float*>(cost), nprop, labels, H, W, K, lamda, tpsi, -1,
                                        sweeps, labels_per_sweep, workspace, workspace_bytes, stream);
    if (bcd_mode == FLOWB200_BCD_FP64_F64COST)
      return launch_sweeps_ksets<double>(pvec, static_cast<const double*>(cost), nprop, labels, H, W, K, lamda, tpsi, -1,
                                         sweeps, labels_per_sweep, workspace, workspace_bytes, stream);
    return FLOWB200_EINVAL;
  }
  if (workspace_bytes < legacy_workspace_bytes(H, W, K)) return FLOWB200_EWORKSPACE;
  uint16_t* bp = static_cast<uint16_t*>(workspace);
  uint32_t* order;
  {
    size_t Kpad = (size_t)(K + 31) / 32 * 32;
    size_t col = (size_t)((W + 1) / 2) * H, row = (size_t)((H + 1) / 2) * W;
    order = reinterpret_cast<uint32_t*>(static_cast<char*>(workspace) +
                                        align_up((col > row ? col : row) * Kpad * sizeof(uint16_t)));
  }
  switch (bcd_mode) {
    case FLOWB200_BCD_FP64_F32COST:
      return launch_sweeps<double, float>(pvec, static_cast<const float*>(cost), nprop, labels, H, W, K, lamda, tpsi, 0,
                                          sweeps, labels_per_sweep, bp, order, stream);
    case FLOWB200_BCD_FP64_F64COST:
      return launch_sweeps<double, double>(pvec, static_cast<const double*>(cost), nprop, labels, H, W, K, lamda, tpsi,
                                           0, sweeps, labels_per_sweep, bp, order, stream);
    case FLOWB200_BCD_INT32_F32COST:
      if (cost_shift < 0 || cost_shift > 14) return FLOWB200_EINVAL;
      return launch_sweeps<int32_t, float>(pvec, static_cast<const float*>(cost), nprop, labels, H, W, K, lamda, tpsi,
                                           cost_shift, sweeps, labels_per_sweep, bp, order, stream);
    case FLOWB200_BCD_INT32:
      if (cost_shift < 0 || cost_shift > 14) return FLOWB200_EINVAL;
      return launch_sweeps<int32_t, int32_t>(pvec, static_cast<const int32_t*>(cost), nprop, labels, H, W, K, lamda,
                                             tpsi, cost_shift, sweeps, labels_per_sweep, bp, order, stream);
    default:
      return FLOWB200_EINVAL;
  }
}

// ---- the K-set programme in pieces: one part of every phase's chains (single-huge-image mode) ----
extern "C" int flowb200_bcd_prepare(const int32_t* pvec, const void* cost, const int32_t* nprop, int H, int W, int K,
                                    int bcd_mode, double lamda, int tpsi, int cost_shift, int part, int nparts,
                                    void* workspace, size_t workspace_bytes, flowb200_stream_t stream) {
  if (!pvec || !cost || !nprop || !workspace) return FLOWB200_EINVAL;
  if (H <= 0 || W <= 0 || K <= 0 || K > 512 || H > 16384 || W > 16384) return FLOWB200_EINVAL;
  if (cost_shift < 0 || cost_shift > 14) return FLOWB200_EINVAL;
  if (bcd_mode == FLOWB200_BCD_INT32)
    return ksets_prepare<int32_t>(pvec, static_cast<const int32_t*>(cost), nprop, H, W, K, lamda, tpsi, cost_shift, part,
                                  nparts, workspace, workspace_bytes, stream);
  if (bcd_mode == FLOWB200_BCD_INT32_F32COST)
    return ksets_prepare<float>(pvec, static_cast<const float*>(cost), nprop, H, W, K, lamda, tpsi, cost_shift, part,
                                nparts, workspace, workspace_bytes, stream);
  return FLOWB200_EUNSUPPORTED;
}

extern "C" int flowb200_bcd_phase(const int32_t* pvec, const void* cost, const int32_t* nprop, int32_t* labels, int H,
                                  int W, int K, int bcd_mode, double lamda, int tpsi, int cost_shift, int phase, int part,
                                  int nparts, void* workspace, size_t workspace_bytes, flowb200_stream_t stream) {
  if (!pvec || !cost || !nprop || !labels || !workspace) return FLOWB200_EINVAL;
  if (H <= 0 || W <= 0 || K <= 0 || K > 512 || H > 16384 || W > 16384) return FLOWB200_EINVAL;
  if (cost_shift < 0 || cost_shift > 14) return FLOWB200_EINVAL;
  if (bcd_mode == FLOWB200_BCD_INT32)
    return ksets_phase<int32_t>(pvec, static_cast<const int32_t*>(cost), nprop, labels, H, W, K, lamda, tpsi, cost_shift,
                                phase, part, nparts, workspace, workspace_bytes, stream);
  if (bcd_mode == FLOWB200_BCD_INT32_F32COST)
    return ksets_phase<float>(pvec, static_cast<const float*>(cost), nprop, labels, H, W, K, lamda, tpsi, cost_shift,
                              phase, part, nparts, workspace, workspace_bytes, stream);
  return FLOWB200_EUNSUPPORTED;
}

extern "C" int flowb200_quantise_costs(const float* lcost, int32_t* m, size_t n, double lamda, int shift,
                                       flowb200_stream_t stream) {
  if (!lcost || !m || shift < 0 || shift > 14) return FLOWB200_EINVAL;
  if (n == 0) return FLOWB200_OK;
  quantise_costs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(lcost, m, n, lamda, (double)(1 << shift));
  FB_LAUNCH_CHECK();
  return FLOWB200_OK;
}
