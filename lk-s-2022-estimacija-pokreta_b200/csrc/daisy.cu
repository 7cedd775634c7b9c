// Dense DAISY(radius=5, q_radius=4, q_theta=4, q_hist=4, NRM_NONE, interpolation on) descriptors.
// Reference call sites: daisy i flann.py:66 (DAISY_create), :69-77 (izracunajDaisy: one keypoint per
// pixel, row-major).  The arithmetic lives in third-party opencv-contrib xfeatures2d (absent offline,
// unpinned); the algorithm is restated in oracle/daisy.py and this file computes
// the same float32 operations in the same order with explicit round-to-nearest intrinsics, so the
// result is bit-identical to the oracle, not merely within 1e-4.
//
// Layout: the four orientation layers of a pixel are one float4, so every blur tap and every petal
// tap is a single 16-byte access; the descriptor write (272 B/pixel) is the HBM-bound part and is
// issued as fully coalesced float4 stores (one thread per (pixel, region)).
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

__global__ void gray_kernel(const uint8_t* __restrict__ bgr, float* __restrict__ gray, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int b = bgr[3 * (size_t)i], g = bgr[3 * (size_t)i + 1], r = bgr[3 * (size_t)i + 2];
  int y = (b * 3735 + g * 19235 + r * 9798 + 16384) >> 15;     // cv2 BGR2GRAY, 15-bit fixed point
  gray[i] = __fdiv_rn((float)y, 255.0f);
}

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
  return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}
__device__ __forceinline__ float4 f4_scale(float k, float4 a) {
  return make_float4(__fmul_rn(k, a.x), __fmul_rn(k, a.y), __fmul_rn(k, a.z), __fmul_rn(k, a.w));
}
__device__ __forceinline__ float f4_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float f4_scale(float k, float a) { return __fmul_rn(k, a); }

// One separable pass, BORDER_REPLICATE: out = k[R]*c + sum_i k[R+i]*(p[-i] + p[+i])   (oracle blur_sep)
template <typename T, int R, bool VERT>
__global__ void blur_kernel(const T* __restrict__ in, T* __restrict__ out, int H, int W, Taps taps) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y;
  if (x >= W) return;
  const size_t row = (size_t)y * W;
  T acc = f4_scale(taps.k[R], in[row + x]);
#pragma unroll
  for (int i = 1; i <= R; ++i) {
    T a, b;
    if (VERT) {
      a = in[(size_t)max(y - i, 0) * W + x];
      b = in[(size_t)min(y + i, H - 1) * W + x];
    } else {
      a = in[row + max(x - i, 0)];
      b = in[row + min(x + i, W - 1)];
    }
    acc = f4_add(acc, f4_scale(taps.k[R + i], f4_add(a, b)));
  }
  out[row + x] = acc;
}

// central differences (Sobel ksize=1, scale 0.5, replicate) and the 4 rectified orientation layers
__global__ void grad_layers_kernel(const float* __restrict__ g, float4* __restrict__ layers, int H, int W,
                                   float4 kos, float4 zin) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y;
  if (x >= W) return;
  float dx = __fmul_rn(0.5f, __fsub_rn(g[(size_t)y * W + min(x + 1, W - 1)], g[(size_t)y * W + max(x - 1, 0)]));
  float dy = __fmul_rn(0.5f, __fsub_rn(g[(size_t)min(y + 1, H - 1) * W + x], g[(size_t)max(y - 1, 0) * W + x]));
  float4 o;
  o.x = fmaxf(__fadd_rn(__fmul_rn(kos.x, dx), __fmul_rn(zin.x, dy)), 0.f);
  o.y = fmaxf(__fadd_rn(__fmul_rn(kos.y, dx), __fmul_rn(zin.y, dy)), 0.f);
  o.z = fmaxf(__fadd_rn(__fmul_rn(kos.z, dx), __fmul_rn(zin.z, dy)), 0.f);
  o.w = fmaxf(__fadd_rn(__fmul_rn(kos.w, dx), __fmul_rn(zin.w, dy)), 0.f);
  layers[(size_t)y * W + x] = o;
}

// bilinear petal sampling (bi_get_histogram semantics): thread = (pixel, region)
__global__ void sample_kernel(const float4* __restrict__ cubes, size_t cube_stride, float4* __restrict__ desc, int H,
                              int W, GridPts gp) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)H * W * 17;
  if (t >= total) return;
  const int reg = (int)(t % 17);
  const int pix = (int)(t / 17);
  const int py = pix / W, px = pix - py * W;
  const double y = (double)py + gp.dy[reg], x = (double)px + gp.dx[reg];
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool inside = (reg == 0) || (x >= 0.0 && x < (double)(W - 1) && y >= 0.0 && y < (double)(H - 1));
  const int mnx = (int)x, mny = (int)y;
  if (inside && mnx < W - 2 && mny < H - 2) {
    const double alpha = (double)(mnx + 1) - x, beta = (double)(mny + 1) - y;
    const float w0 = (float)(alpha * beta);
    const float w1 = (float)(beta - (double)w0);
    const float w2 = (float)(alpha - (double)w0);
    const float w3 = (float)(1.0 + (double)w0 - alpha - beta);
    const float4* c = cubes + (size_t)gp.cube[reg] * cube_stride;
    const float4 A = c[(size_t)mny * W + mnx], C = c[(size_t)mny * W + mnx + 1];
    const float4 B = c[(size_t)(mny + 1) * W + mnx], D = c[(size_t)(mny + 1) * W + mnx + 1];
    o = f4_add(f4_add(f4_add(f4_scale(w0, A), f4_scale(w1, C)), f4_scale(w2, B)), f4_scale(w3, D));
  }
  desc[t] = o;
}

template <typename T>
static int blur2d(const T* in, T* tmp, T* out, int H, int W, int which, cudaStream_t s) {
  const int ksize = kDaisyBlurSize[which];
  Taps t{};
  for (int i = 0; i < ksize; ++i) t.k[i] = kDaisyBlurTaps[which][i];
  dim3 block(128), grid((W + 127) / 128, H);
  switch (ksize / 2) {
#define FB_BLUR_CASE(R)                                                  \
  case R:                                                                \
    blur_kernel<T, R, false><<<grid, block, 0, s>>>(in, tmp, H, W, t);   \
    blur_kernel<T, R, true><<<grid, block, 0, s>>>(tmp, out, H, W, t);   \
    break;
    FB_BLUR_CASE(1) FB_BLUR_CASE(2) FB_BLUR_CASE(3) FB_BLUR_CASE(4)
#undef FB_BLUR_CASE
    default: return FLOWB200_EINVAL;
  }
  FB_LAUNCH_CHECK_N(2);
  return FLOWB200_OK;
}

}  // namespace flowb200

using namespace flowb200;

extern "C" size_t flowb200_daisy_workspace_bytes(int H, int W) {
  if (H <= 0 || W <= 0) return 0;
  size_t n = (size_t)H * W;
  return 3 * align_up(n * sizeof(float)) + 6 * align_up(n * sizeof(float4));
}

extern "C" int flowb200_daisy(const uint8_t* bgr, int H, int W, float* desc, void* workspace, size_t workspace_bytes,
                              flowb200_stream_t stream) {
  if (!bgr || !desc || !workspace || H < 3 || W < 3) return FLOWB200_EINVAL;
  if (workspace_bytes < flowb200_daisy_workspace_bytes(H, W)) return FLOWB200_EWORKSPACE;
  const size_t n = (size_t)H * W;
  char* w = static_cast<char*>(workspace);
  float* g0 = reinterpret_cast<float*>(w); w += align_up(n * sizeof(float));
  float* g1 = reinterpret_cast<float*>(w); w += align_up(n * sizeof(float));
  float* g2 = reinterpret_cast<float*>(w); w += align_up(n * sizeof(float));
  float4* lay = reinterpret_cast<float4*>(w); w += align_up(n * sizeof(float4));
  float4* tmp = reinterpret_cast<float4*>(w); w += align_up(n * sizeof(float4));
  float4* cubes = reinterpret_cast<float4*>(w);
  const size_t cube_stride = align_up(n * sizeof(float4)) / sizeof(float4);

  gray_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(bgr, g0, (int)n);
  FB_LAUNCH_CHECK();
  int rc = blur2d<float>(g0, g1, g2, H, W, 0, stream);            // layered_gradient(): GaussianBlur 5x5, 0.5
  if (rc) return rc;
  const float4 kos = make_float4(kDaisyCos[0], kDaisyCos[1], kDaisyCos[2], kDaisyCos[3]);
  const float4 zin = make_float4(kDaisySin[0], kDaisySin[1], kDaisySin[2], kDaisySin[3]);
  dim3 block(128), grid((W + 127) / 128, H);
  grad_layers_kernel<<<grid, block, 0, stream>>>(g2, lay, H, W, kos, zin);
  FB_LAUNCH_CHECK();
  // base smoothing sqrt(1.6^2 - 0.5^2), then the incremental blurs to cumulative sigma 0.625*(r+1)
  rc = blur2d<float4>(lay, tmp, lay, H, W, 1, stream);
  if (rc) return rc;
  const float4* src = lay;
  for (int r = 0; r < 4; ++r) {
    float4* dst = cubes + (size_t)r * cube_stride;
    rc = blur2d<float4>(src, tmp, dst, H, W, 2 + r, stream);
    if (rc) return rc;
    src = dst;
  }
  GridPts gp;
  for (int i = 0; i < 17; ++i) {
    gp.dy[i] = kDaisyGridDy[i];
    gp.dx[i] = kDaisyGridDx[i];
    gp.cube[i] = kDaisyGridCube[i];
  }
  const size_t total = n * 17;
  sample_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(cubes, cube_stride, reinterpret_cast<float4*>(desc),
                                                                   H, W, gp);
  FB_LAUNCH_CHECK();
  return FLOWB200_OK;
}
