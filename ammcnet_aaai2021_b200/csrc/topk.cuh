// Sorted running top-K (smallest first) kept in registers; strict total order (distance, then index) so the result
// does not depend on the order in which partial lists are merged.
#pragma once
#include <math.h>

namespace ammc {

template <int K>
struct TopK {
  float v[K];
  int id[K];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < K; ++i) { v[i] = INFINITY; id[i] = 0x7fffffff; }
  }
  static __device__ __forceinline__ bool before(float a, int ia, float b, int ib) {
    return a < b || (a == b && ia < ib);
  }
  __device__ __forceinline__ void insert(float d, int j) {
    if (!before(d, j, v[K - 1], id[K - 1])) return;
    v[K - 1] = d; id[K - 1] = j;
#pragma unroll
    for (int i = K - 1; i > 0; --i) {
      if (before(v[i], id[i], v[i - 1], id[i - 1])) {
        float tv = v[i]; v[i] = v[i - 1]; v[i - 1] = tv;
        int ti = id[i]; id[i] = id[i - 1]; id[i - 1] = ti;
      }
    }
  }
};

}  // namespace ammc
