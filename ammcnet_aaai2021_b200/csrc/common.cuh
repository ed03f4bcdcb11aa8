// Shared host/device helpers for the ammc_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/ammc_b200.h"

namespace ammc {

// thread-local error message returned by ammc_last_error()
char* err_buf();
int fail(int code, const char* fmt, ...);

#define AMMC_CUDA_CHECK(expr)                                                                   \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return ::ammc::fail(AMMC_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                          __FILE__, __LINE__);                                                  \
  } while (0)

#define AMMC_LAUNCH_CHECK(name)                                                                 \
  do {                                                                                          \
    cudaError_t _e = cudaGetLastError();                                                        \
    if (_e != cudaSuccess)                                                                      \
      return ::ammc::fail(AMMC_ECUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
  } while (0)

#define AMMC_REQUIRE(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) return ::ammc::fail(AMMC_EINVAL, __VA_ARGS__);   \
  } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// bump allocator over the caller's workspace (256-byte aligned pieces)
struct Workspace {
  char* base;
  size_t size;
  size_t off;
  Workspace(void* p, size_t n) : base((char*)p), size(n), off(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T), 256);
    if (off + bytes > size) { off = size + 1; return nullptr; }
    T* r = (T*)(base + off);
    off += bytes;
    return r;
  }
  bool ok() const { return off <= size; }
};

int num_sms();

// largest power of two s with bound * s < 2^15 (s = 1 for bound == 0 or non-finite)
__host__ __device__ __forceinline__ float q_scale_for_bound(float bound) {
  if (!(bound > 0.f) || !(bound < 3.0e38f)) return 1.f;
  int e;
  frexpf(bound, &e);                 // bound = m * 2^e, m in [0.5, 1)
  int se = 15 - e;
  se = se > 100 ? 100 : (se < -100 ? -100 : se);
  return ldexpf(1.f, se);
}


__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32); result valid in every thread.
__device__ __forceinline__ float block_sum(float v, float* red /* >= 33 floats of smem */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    float t = lane < nw ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

}  // namespace ammc
