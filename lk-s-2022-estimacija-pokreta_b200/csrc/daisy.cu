// Dense DAISY(radius=5, q_radius=4, q_theta=4, q_hist=4, NRM_NONE, interpolation on) descriptors.
// Reference call sites: daisy i flann.py:66 (DAISY_create), :69-77 (izracunajDaisy: one keypoint per
// pixel, row-major).  The arithmetic lives in third-party opencv-contrib xfeatures2d (absent offline,
// unpinned); the algorithm is restated in oracle/daisy.py and this file computes
// the same float32 operations in the same order with explicit round-to-nearest intrinsics, so the
// result is bit-identical to the oracle, not merely within 1e-4.
//
// Layout: the four orientation layers of a pixel are one float4, so every blur tap and every petal
// tap is a single 16-byte access.  Seven launches per image: gray + pre-blur + gradient + layers fused in one
// shared-memory tile kernel; every separable blur does both passes in one kernel through shared memory; the
// sampler stages the four cubes of a 32 x 8 tile (+halo) in shared memory and writes the descriptor
// (272 B/pixel, the HBM-bound part) as fully coalesced float4 stores.
#include "common.cuh"
#include "daisy_constants.h"

namespace flowb200 {

constexpr int kMaxTaps = 9;
struct Taps {
  float k[kMaxTaps];   // full symmetric kernel, length 2R+1
};

struct GridPts {
  double dy[17], dx[17];
  int cube[17];
};

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
  return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}
__device__ __forceinline__ float4 f4_scale(float k, float4 a) {
  return make_float4(__fmul_rn(k, a.x), __fmul_rn(k, a.y), __fmul_rn(k, a.z), __fmul_rn(k, a.w));
}
__device__ __forceinline__ float f4_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float f4_scale(float k, float a) { return __fmul_rn(k, a); }

// symmetric taps in the oracle's order: k[R]*c + sum_i k[R+i]*(p[-i] + p[+i])   (oracle blur_sep)
template <typename T, int R, typename F>
__device__ __forceinline__ T sym_taps(const Taps& taps, F at) {
  T acc = f4_scale(taps.k[R], at(0));
#pragma unroll
  for (int i = 1; i <= R; ++i) acc = f4_add(acc, f4_scale(taps.k[R + i], f4_add(at(-i), at(i))));
  return acc;
}

constexpr int kTX = 32, kTY = 16;   // output tile of the fused kernels

// gray (cv2 BGR2GRAY, 15-bit fixed point) / 255 -> 5x5 Gaussian (sigma 0.5) -> central differences -> the four
// rectified orientation layers, one float4 per pixel.  All borders replicate, as cv2 does on the whole image.
__global__ void __launch_bounds__(kTX* kTY)
pre_layers_kernel(const uint8_t* __restrict__ bgr, float4* __restrict__ layers, int H, int W, Taps taps, float4 kos,
                  float4 zin) {
  constexpr int R = 2, GW = kTX + 2 * R + 2, GH = kTY + 2 * R + 2;     // gray tile with halo R + 1
  __shared__ float g_s[GH][GW];
  __shared__ float h_s[GH][kTX + 2];                                    // after the horizontal pass, halo 1 in x
  __shared__ float b_s[kTY + 2][kTX + 2];                               // blurred gray, halo 1
  const int tx = threadIdx.x % kTX, ty = threadIdx.x / kTX;
  const int x0 = blockIdx.x * kTX, y0 = blockIdx.y * kTY;
  for (int e = threadIdx.x; e < GW * GH; e += kTX * kTY) {
    const int ly = e / GW, lx = e - ly * GW;
    const int gy = min(max(y0 + ly - R - 1, 0), H - 1), gx = min(max(x0 + lx - R - 1, 0), W - 1);
    const uint8_t* p = bgr + 3 * ((size_t)gy * W + gx);
    const int y = (p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + 16384) >> 15;
    g_s[ly][lx] = __fdiv_rn((float)y, 255.0f);
  }
  __syncthreads();
  // horizontal pass at columns x0-1 .. x0+kTX.  The replicate border is applied to the IMAGE: column c of the
  // padded row is the clamped image column, which is what the tile already holds.
  for (int e = threadIdx.x; e < GH * (kTX + 2); e += kTX * kTY) {
    const int ly = e / (kTX + 2), lx = e - ly * (kTX + 2);
    const int gx = min(max(x0 + lx - 1, 0), W - 1);                    // image column this entry stands for
    h_s[ly][lx] = sym_taps<float, R>(taps, [&](int i) {
      const int c = min(max(gx + i, 0), W - 1);                        // clamped image column
      return g_s[ly][c - (x0 - R - 1) >= 0 ? min(c - (x0 - R - 1), GW - 1) : 0];
    });
  }
  __syncthreads();
  for (int e = threadIdx.x; e < (kTY + 2) * (kTX + 2); e += kTX * kTY) {
    const int ly = e / (kTX + 2), lx = e - ly * (kTX + 2);
    const int gy = min(max(y0 + ly - 1, 0), H - 1);
    b_s[ly][lx] = sym_taps<float, R>(taps, [&](int i) {
      const int r = min(max(gy + i, 0), H - 1);
      return h_s[min(max(r - (y0 - R - 1), 0), GH - 1)][lx];
    });
  }
  __syncthreads();
  const int x = x0 + tx, y = y0 + ty;
  if (x >= W || y >= H) return;
  // Sobel ksize=1, scale 0.5, replicate: b_s index of image column c is c - (x0 - 1), rows alike
  const int xl = max(x - 1, 0) - (x0 - 1), xr = min(x + 1, W - 1) - (x0 - 1);
  const int yu = max(y - 1, 0) - (y0 - 1), yd = min(y + 1, H - 1) - (y0 - 1);
  const float dx = __fmul_rn(0.5f, __fsub_rn(b_s[ty + 1][xr], b_s[ty + 1][xl]));
  const float dy = __fmul_rn(0.5f, __fsub_rn(b_s[yd][tx + 1], b_s[yu][tx + 1]));
  float4 o;
  o.x = fmaxf(__fadd_rn(__fmul_rn(kos.x, dx), __fmul_rn(zin.x, dy)), 0.f);
  o.y = fmaxf(__fadd_rn(__fmul_rn(kos.y, dx), __fmul_rn(zin.y, dy)), 0.f);
  o.z = fmaxf(__fadd_rn(__fmul_rn(kos.z, dx), __fmul_rn(zin.z, dy)), 0.f);
  o.w = fmaxf(__fadd_rn(__fmul_rn(kos.w, dx), __fmul_rn(zin.w, dy)), 0.f);
  layers[(size_t)y * W + x] = o;
}

// separable Gaussian on the float4 layers, horizontal then vertical pass in one kernel through shared memory
template <int R>
__global__ void __launch_bounds__(kTX* kTY)
blur_layers_kernel(const float4* __restrict__ in, float4* __restrict__ out, int H, int W, Taps taps) {
  __shared__ float4 h_s[kTY + 2 * R][kTX];       // horizontal pass results for the rows the vertical pass needs
  const int tx = threadIdx.x % kTX, ty = threadIdx.x / kTX;
  const int x0 = blockIdx.x * kTX, y0 = blockIdx.y * kTY;
  const int x = min(x0 + tx, W - 1);
  for (int ly = ty; ly < kTY + 2 * R; ly += kTY) {
    const int gy = min(max(y0 + ly - R, 0), H - 1);
    const float4* rowp = in + (size_t)gy * W;
    h_s[ly][tx] = sym_taps<float4, R>(taps, [&](int i) { return rowp[min(max(x + i, 0), W - 1)]; });
  }
  __syncthreads();
  const int y = y0 + ty;
  if (x0 + tx >= W || y >= H) return;
  out[(size_t)y * W + x] = sym_taps<float4, R>(taps, [&](int i) {
    const int r = min(max(y + i, 0), H - 1);      // clamped image row; rows y0-R .. y0+kTY-1+R are in the tile
    return h_s[r - (y0 - R) < 0 ? 0 : r - (y0 - R)][tx];
  });
}

// bilinear petal sampling (bi_get_histogram semantics).  One block per 32 x 8 pixel tile: the four cubes are staged
// in shared memory with a halo of 6, every thread produces (pixel, region) float4s such that consecutive threads
// write consecutive 16-byte pieces of the descriptor array (the 272 B/pixel write is the HBM-bound part).
constexpr int kSX = 32, kSY = 8, kSHalo = 6;
constexpr int kSW = kSX + 2 * kSHalo + 1, kSH = kSY + 2 * kSHalo + 1;

__global__ void __launch_bounds__(256)
sample_kernel(const float4* __restrict__ cubes, size_t cube_stride, float4* __restrict__ desc, int H, int W, GridPts gp) {
  extern __shared__ float4 cube_s[];             // [4][kSH][kSW]
  __shared__ double ax_s[kSX][17], ay_s[kSY][17];
  __shared__ int mx_s[kSX][17], my_s[kSY][17];
  __shared__ int gcube_s[17];
  if (threadIdx.x < 17) gcube_s[threadIdx.x] = gp.cube[threadIdx.x];
  const int x0 = blockIdx.x * kSX, y0 = blockIdx.y * kSY;
  for (int e = threadIdx.x; e < 4 * kSH * kSW; e += 256) {
    const int c = e / (kSH * kSW), r = e - c * (kSH * kSW);
    const int ly = r / kSW, lx = r - ly * kSW;
    const int gy = y0 + ly - kSHalo, gx = x0 + lx - kSHalo;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = cubes[(size_t)c * cube_stride + (size_t)gy * W + gx];
    cube_s[e] = v;
  }
  __syncthreads();
  // per-tile tables: the double-precision coordinate arithmetic of bi_get_histogram depends on (column, region)
  // and (row, region) separately, so it is done once per tile instead of once per (pixel, region)
  for (int e = threadIdx.x; e < (kSX + kSY) * 17; e += 256) {
    const int k = e / 17, reg = e - k * 17;
    if (k < kSX) {
      const double x = (double)(x0 + k) + gp.dx[reg];
      const int mnx = (int)x;
      ax_s[k][reg] = (double)(mnx + 1) - x;
      mx_s[k][reg] = ((reg == 0) || (x >= 0.0 && x < (double)(W - 1))) && mnx < W - 2 ? mnx - x0 + kSHalo : -1;
    } else {
      const double y = (double)(y0 + k - kSX) + gp.dy[reg];
      const int mny = (int)y;
      ay_s[k - kSX][reg] = (double)(mny + 1) - y;
      my_s[k - kSX][reg] = ((reg == 0) || (y >= 0.0 && y < (double)(H - 1))) && mny < H - 2 ? mny - y0 + kSHalo : -1;
    }
  }
  __syncthreads();
  const int ny = min(kSY, H - y0), nx = min(kSX, W - x0);
  const int items = ny * nx * 17;                 // (pixel, region) pairs of the tile, row-major like the output
  for (int e = threadIdx.x; e < items; e += 256) {
    const int pixel = e / 17, reg = e - pixel * 17;
    const int ly = pixel / nx, lx = pixel - ly * nx;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    const int cx = mx_s[lx][reg], cy = my_s[ly][reg];
    if (cx >= 0 && cy >= 0) {
      const double alpha = ax_s[lx][reg], beta = ay_s[ly][reg];
      const float w0 = (float)(alpha * beta);
      const double w0d = (double)w0;
      const float w1 = (float)(beta - w0d);
      const float w2 = (float)(alpha - w0d);
      const float w3 = (float)(1.0 + w0d - alpha - beta);
      const float4* c = cube_s + ((size_t)gcube_s[reg] * kSH + cy) * kSW + cx;
      const float4 A = c[0], C = c[1], B = c[kSW], D = c[kSW + 1];
      o = f4_add(f4_add(f4_add(f4_scale(w0, A), f4_scale(w1, C)), f4_scale(w2, B)), f4_scale(w3, D));
    }
    desc[((size_t)(y0 + ly) * W + x0 + lx) * 17 + reg] = o;
  }
}

static Taps taps_of(int which) {
  Taps t{};
  for (int i = 0; i < kDaisyBlurSize[which]; ++i) t.k[i] = kDaisyBlurTaps[which][i];
  return t;
}

static int blur_layers(const float4* in, float4* out, int H, int W, int which, cudaStream_t s) {
  const Taps t = taps_of(which);
  dim3 grid((W + kTX - 1) / kTX, (H + kTY - 1) / kTY);
  switch (kDaisyBlurSize[which] / 2) {
    case 1: blur_layers_kernel<1><<<grid, kTX * kTY, 0, s>>>(in, out, H, W, t); break;
    case 2: blur_layers_kernel<2><<<grid, kTX * kTY, 0, s>>>(in, out, H, W, t); break;
    case 3: blur_layers_kernel<3><<<grid, kTX * kTY, 0, s>>>(in, out, H, W, t); break;
    case 4: blur_layers_kernel<4><<<grid, kTX * kTY, 0, s>>>(in, out, H, W, t); break;
    default: return FLOWB200_EINVAL;
  }
  FB_LAUNCH_CHECK();
  return FLOWB200_OK;
}

}  // namespace flowb200

using namespace flowb200;

extern "C" size_t flowb200_daisy_workspace_bytes(int H, int W) {
  if (H <= 0 || W <= 0) return 0;
  size_t n = (size_t)H * W;
  return 6 * align_up(n * sizeof(float4));
}

extern "C" int flowb200_daisy(const uint8_t* bgr, int H, int W, float* desc, void* workspace, size_t workspace_bytes,
                              flowb200_stream_t stream) {
  if (!bgr || !desc || !workspace || H < 3 || W < 3) return FLOWB200_EINVAL;
  if (workspace_bytes < flowb200_daisy_workspace_bytes(H, W)) return FLOWB200_EWORKSPACE;
  const size_t n = (size_t)H * W;
  char* w = static_cast<char*>(workspace);
  float4* lay = reinterpret_cast<float4*>(w); w += align_up(n * sizeof(float4));
  float4* tmp = reinterpret_cast<float4*>(w); w += align_up(n * sizeof(float4));
  float4* cubes = reinterpret_cast<float4*>(w);
  const size_t cube_stride = align_up(n * sizeof(float4)) / sizeof(float4);

  // layered_gradient(): gray, GaussianBlur 5x5 sigma 0.5, central differences, rectified layers
  const float4 kos = make_float4(kDaisyCos[0], kDaisyCos[1], kDaisyCos[2], kDaisyCos[3]);
  const float4 zin = make_float4(kDaisySin[0], kDaisySin[1], kDaisySin[2], kDaisySin[3]);
  dim3 grid((W + kTX - 1) / kTX, (H + kTY - 1) / kTY);
  pre_layers_kernel<<<grid, kTX * kTY, 0, stream>>>(bgr, lay, H, W, taps_of(0), kos, zin);
  FB_LAUNCH_CHECK();
  // base smoothing sqrt(1.6^2 - 0.5^2), then the incremental blurs to cumulative sigma 0.625*(r+1)
  int rc = blur_layers(lay, tmp, H, W, 1, stream);
  if (rc) return rc;
  const float4* src = tmp;
  for (int r = 0; r < 4; ++r) {
    float4* dst = cubes + (size_t)r * cube_stride;
    rc = blur_layers(src, dst, H, W, 2 + r, stream);
    if (rc) return rc;
    src = dst;
  }
  GridPts gp;
  for (int i = 0; i < 17; ++i) {
    gp.dy[i] = kDaisyGridDy[i];
    gp.dx[i] = kDaisyGridDx[i];
    gp.cube[i] = kDaisyGridCube[i];
  }
  const size_t smem = (size_t)4 * kSH * kSW * sizeof(float4);
  FB_CUDA_CHECK(cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 sgrid((W + kSX - 1) / kSX, (H + kSY - 1) / kSY);
  sample_kernel<<<sgrid, 256, smem, stream>>>(cubes, cube_stride, reinterpret_cast<float4*>(desc), H, W, gp);
  FB_LAUNCH_CHECK();
  return FLOWB200_OK;
}
