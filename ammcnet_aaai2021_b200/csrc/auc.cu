// ROC-AUC on the GPU ("GPU ROC-AUC (sort + scan)").
//
// sklearn.metrics.roc_curve + auc, as called at reference Code/main/eval_metric.py:428-429, equals the Mann-Whitney
// statistic with tied scores sharing their average rank:
//     AUC = (R_pos - n_pos (n_pos + 1) / 2) / (n_pos * n_neg),   R_pos = sum of 1-based average ranks of the positives.
// Steps: (1) order-preserving uint32 key of every fp32 score, (2) bitonic sort of (key, label) pairs (shared-memory stages
// for spans <= 2048, global stages above), (3) per element the tie group [lower_bound, upper_bound) by binary search,
// rank sum and positive count accumulated in integers (2*rank is an integer: exact), (4) the ratio in fp64.
// The result is the exact rational value sklearn's trapezoid rule produces, to fp64 rounding.
#include "common.cuh"

namespace ammc {

constexpr int AUC_SMEM_ELEMS = 2048;     // elements sorted per block in shared memory (1024 threads, 2 per thread)

__device__ __forceinline__ uint32_t float_key(float f) {
  uint32_t u = __float_as_uint(f + 0.0f);                   // -0.0 + 0.0 = +0.0: both zeros tie, as they compare equal
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);       // total order of IEEE floats as unsigned ints
}

__global__ void auc_keys_kernel(const float* __restrict__ scores, const int8_t* __restrict__ labels, int pos_label,
                                uint32_t* __restrict__ keys, uint8_t* __restrict__ flags, int64_t T, int64_t P) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  if (i < T) { keys[i] = float_key(scores[i]); flags[i] = labels[i] == pos_label ? 1 : 0; }
  else { keys[i] = 0xFFFFFFFFu; flags[i] = 2; }            // padding sorts to the end
}

__device__ __forceinline__ void cmp_swap(uint32_t& ka, uint8_t& fa, uint32_t& kb, uint8_t& fb, bool up) {
  if ((ka > kb) == up) { uint32_t t = ka; ka = kb; kb = t; uint8_t f = fa; fa = fb; fb = f; }
}

// all stages with span k <= AUC_SMEM_ELEMS for k from k_begin (each block owns AUC_SMEM_ELEMS consecutive elements)
__global__ void __launch_bounds__(1024) auc_bitonic_smem_kernel(uint32_t* __restrict__ keys, uint8_t* __restrict__ flags,
                                                                int k_begin, int k_end, int j_begin) {
  __shared__ uint32_t sk[AUC_SMEM_ELEMS];
  __shared__ uint8_t sf[AUC_SMEM_ELEMS];
  const int64_t base = (int64_t)blockIdx.x * AUC_SMEM_ELEMS;
  for (int i = threadIdx.x; i < AUC_SMEM_ELEMS; i += 1024) { sk[i] = keys[base + i]; sf[i] = flags[base + i]; }
  __syncthreads();
  for (int k = k_begin; k <= k_end; k <<= 1) {
    for (int j = (k == k_begin ? j_begin : k >> 1); j > 0; j >>= 1) {
      const int t = threadIdx.x;
      const int i = 2 * t - (t & (j - 1));                  // lower index of the pair
      const int l = i + j;
      const bool up = (((base + i) & k) == 0);
      cmp_swap(sk[i], sf[i], sk[l], sf[l], up);
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < AUC_SMEM_ELEMS; i += 1024) { keys[base + i] = sk[i]; flags[base + i] = sf[i]; }
}

__global__ void auc_bitonic_global_kernel(uint32_t* __restrict__ keys, uint8_t* __restrict__ flags, int64_t k, int64_t j,
                                          int64_t P) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P / 2) return;
  const int64_t i = 2 * t - (t & (j - 1));
  const int64_t l = i + j;
  const bool up = ((i & k) == 0);
  uint32_t ka = keys[i], kb = keys[l];
  uint8_t fa = flags[i], fb = flags[l];
  if ((ka > kb) == up) { keys[i] = kb; keys[l] = ka; flags[i] = fb; flags[l] = fa; }
}

// acc[0] += sum over positives of (lo + hi + 1) = 2 * average 1-based rank ; acc[1] += #positives
__global__ void auc_ranks_kernel(const uint32_t* __restrict__ keys, const uint8_t* __restrict__ flags,
                                 unsigned long long* __restrict__ acc, int64_t T) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long r2 = 0, np = 0;
  if (i < T && flags[i] == 1) {
    const uint32_t key = keys[i];
    int64_t lo = 0, hi = i;                                 // lower_bound in [0, i]
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if (keys[m] < key) lo = m + 1; else hi = m; }
    const int64_t first = lo;
    lo = i; hi = T;                                         // upper_bound in [i, T)
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if (keys[m] <= key) lo = m + 1; else hi = m; }
    r2 = (unsigned long long)(first + lo + 1);
    np = 1;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { r2 += __shfl_xor_sync(0xffffffffu, r2, o); np += __shfl_xor_sync(0xffffffffu, np, o); }
  if ((threadIdx.x & 31) == 0 && np) { atomicAdd(&acc[0], r2); atomicAdd(&acc[1], np); }
}

__global__ void auc_final_kernel(const unsigned long long* __restrict__ acc, double* __restrict__ auc, int64_t T) {
  const double n_pos = (double)acc[1], n_neg = (double)T - n_pos;
  if (n_pos == 0.0 || n_neg == 0.0) { auc[0] = nan(""); return; }
  const double r_pos = 0.5 * (double)acc[0];
  auc[0] = (r_pos - 0.5 * n_pos * (n_pos + 1.0)) / (n_pos * n_neg);
}

static int64_t pow2_at_least(int64_t n, int64_t floor_) {
  int64_t p = floor_;
  while (p < n) p <<= 1;
  return p;
}

}  // namespace ammc

using namespace ammc;

extern "C" size_t ammc_auc_workspace_bytes(int64_t T) {
  const int64_t P = pow2_at_least(T, AUC_SMEM_ELEMS);
  return align_up((size_t)P * 4, 256) + align_up((size_t)P, 256) + 256;
}

extern "C" int ammc_roc_auc(const float* scores, const int8_t* labels, int pos_label, double* auc, void* workspace,
                            size_t workspace_bytes, int64_t T, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AMMC_REQUIRE(scores && labels && auc && T > 0 && T <= (1LL << 26), "bad argument");
  const int64_t P = pow2_at_least(T, AUC_SMEM_ELEMS);
  Workspace ws(workspace, workspace_bytes);
  uint32_t* keys = ws.take<uint32_t>(P);
  uint8_t* flags = ws.take<uint8_t>(P);
  unsigned long long* acc = ws.take<unsigned long long>(2);
  if (!ws.ok()) return fail(AMMC_EWORKSPACE, "workspace too small");
  auc_keys_kernel<<<ceil_div(P, 256), 256, 0, st>>>(scores, labels, pos_label, keys, flags, T, P);
  AMMC_LAUNCH_CHECK("auc_keys_kernel");
  const int blocks = (int)(P / AUC_SMEM_ELEMS);
  auc_bitonic_smem_kernel<<<blocks, 1024, 0, st>>>(keys, flags, 2, AUC_SMEM_ELEMS, 1);
  AMMC_LAUNCH_CHECK("auc_bitonic_smem_kernel");
  for (int64_t k = 2 * AUC_SMEM_ELEMS; k <= P; k <<= 1) {
    for (int64_t j = k >> 1; j >= AUC_SMEM_ELEMS; j >>= 1) {
      auc_bitonic_global_kernel<<<ceil_div(P / 2, 256), 256, 0, st>>>(keys, flags, k, j, P);
      AMMC_LAUNCH_CHECK("auc_bitonic_global_kernel");
    }
    // remaining strides of this k fit one block's span
    auc_bitonic_smem_kernel<<<blocks, 1024, 0, st>>>(keys, flags, (int)k, (int)k, AUC_SMEM_ELEMS / 2);   // k <= 2^26
    AMMC_LAUNCH_CHECK("auc_bitonic_smem_kernel");
  }
  AMMC_CUDA_CHECK(cudaMemsetAsync(acc, 0, 16, st));
  auc_ranks_kernel<<<ceil_div(T, 256), 256, 0, st>>>(keys, flags, acc, T);
  AMMC_LAUNCH_CHECK("auc_ranks_kernel");
  auc_final_kernel<<<1, 1, 0, st>>>(acc, auc, T);
  AMMC_LAUNCH_CHECK("auc_final_kernel");
  return 0;
}
