// Micro-benchmark: how fast can ONE CTA stream contiguous 20 KB chunks from L2 into shared memory with
// cp.async.bulk + mbarriers (producer thread / consumer thread, S stages)?  Build:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../lk-s-2022-estimacija-pokreta_b200/csrc tma_stream.cu -o tma_stream
#include <cstdio>
#include <cstdlib>
#include "sm100_ptx.cuh"
using namespace flowb200;

template <int S>
__global__ void stream_kernel(const char* __restrict__ src, size_t region_bytes, int chunk_bytes, int nchunks, int pieces) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full[S], empty[S];
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  const size_t nregion = region_bytes / chunk_bytes;
  if (threadIdx.x == 0) {
    for (int c = 0; c < nchunks; ++c) {
      const int st = c % S, ph = (c / S) & 1;
      ptx::mbar_wait(&empty[st], ph ^ 1);
      ptx::mbar_arrive_expect_tx(&full[st], chunk_bytes);
      const size_t idx = ((size_t)blockIdx.x * 7919 + (size_t)c * 104729) % nregion;
      const int pb = chunk_bytes / pieces;
      for (int q = 0; q < pieces; ++q)
        ptx::bulk_load(smem + st * chunk_bytes + q * pb, src + idx * chunk_bytes + q * pb, pb, &full[st]);
    }
  } else if (threadIdx.x == 32) {
    for (int c = 0; c < nchunks; ++c) {
      const int st = c % S, ph = (c / S) & 1;
      ptx::mbar_wait(&full[st], ph);
      ptx::mbar_arrive(&empty[st]);
    }
  }
}

template <int S>
void run(const char* d, size_t region, int chunk, int nchunks, int ctas_per_sm, int pieces) {
  const int grid = 148 * ctas_per_sm;
  const size_t smem = (size_t)S * chunk;
  cudaFuncSetAttribute(stream_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  stream_kernel<S><<<grid, 64, smem>>>(d, region, chunk, 64, pieces);
  cudaEventRecord(a);
  stream_kernel<S><<<grid, 64, smem>>>(d, region, chunk, nchunks, pieces);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double bytes = (double)grid * nchunks * chunk;
  printf("S=%d chunk=%d B pieces=%d ctas/SM=%d region=%zu MB: %.3f ms, %.1f GB/s total, %.2f GB/s per CTA, %.2f us/chunk  (%s)\n", S, chunk,
         pieces, ctas_per_sm, region >> 20, ms, bytes / ms / 1e6, bytes / ms / 1e6 / grid, ms * 1e3 / nchunks,
         cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv) {
  if (argc > 1) {   // record-stream mode: chunks of the K-set record size from a region far larger than L2
    const size_t total = 4096ull << 20;
    char* d; cudaMalloc(&d, total); cudaMemset(d, 1, total);
    const int chunk = atoi(argv[1]);
    for (int cps : {1, 2, 3, 4}) {
      run<1>(d, total, chunk, 2000, cps, 1);
      run<2>(d, total, chunk, 2000, cps, 1);
      run<4>(d, total, chunk, 2000, cps, 1);
      run<8>(d, total, chunk, 2000, cps, 1);
    }
    return 0;
  }
  const size_t total = 512ull << 20;
  char* d; cudaMalloc(&d, total); cudaMemset(d, 1, total);
  for (size_t region : {(size_t)64 << 20, (size_t)512 << 20}) {
    run<2>(d, region, 20480, 2000, 1, 1);
    run<3>(d, region, 20480, 2000, 1, 1);
    run<6>(d, region, 20480, 2000, 1, 1);
    run<3>(d, region, 20480, 2000, 1, 10);
    run<3>(d, region, 20480, 2000, 2, 1);
    run<3>(d, region, 20480, 2000, 4, 1);
    run<3>(d, region, 4096, 2000, 1, 1);
    run<8>(d, region, 4096, 4000, 4, 1);
  }
  return 0;
}
