// Lock-free union-find on an int32 parent array (device): roots are minimal indices, union = atomicMin on the
// larger root.  Used by segments.cu (flow segments) and edges.cu (Canny hysteresis).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace flowb200 {

__device__ __forceinline__ int uf_find(volatile int32_t* parent, int x) {
  int p;
  while ((p = parent[x]) != x) x = p;
  return x;
}

__device__ __forceinline__ void uf_merge(int32_t* parent, int x, int y) {
  bool done;
  do {
    x = uf_find(parent, x);
    y = uf_find(parent, y);
    if (x < y) {
      const int old = atomicMin(&parent[y], x);
      done = old == y;
      y = old;
    } else if (y < x) {
      const int old = atomicMin(&parent[x], y);
      done = old == x;
      x = old;
    } else {
      done = true;
    }
  } while (!done);
}

}  // namespace flowb200
