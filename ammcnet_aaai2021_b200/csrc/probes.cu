// Hardware-behaviour probes behind the kernel designs (debug entry points; compiled ONLY into libammc_b200_debug.so --
// `python -m ammcnet_aaai2021_b200.build --debug` -- never into the product library).
//
// ammc_debug_fp8_probe: the mixed-precision chain proposed in DESIGN.md section 8 for the dominant conv kernel --
//   D  = A8 . B8^T                 kind::f8f6f4 (e4m3 x e4m3, K = 32 per MMA, 128-byte swizzled K-major rows of 128 elements)
//   D  = D * 2^-SCALE + A16 . B16^T   kind::f16 (fp16 x fp16) with the `scale-input-d` immediate on its first MMA
// checks on a 128 x 64 tile that (i) sm_100a executes the plain (non block-scaled) fp8 kind with the same shared-memory
// descriptors as bf16, (ii) a descriptor advance of +2 (32 bytes) is the K step for 8-bit operands as well, (iii)
// scale-input-d rescales the accumulator exactly as documented.
#ifdef AMMC_DEBUG_PROBES
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/ammc_b200_debug.h"

namespace ammc {

int make_map_generic(CUtensorMap* m, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                     const uint64_t* strides, const uint32_t* box, int swizzle128);   // amft_conv.cu

constexpr int FP8_PROBE_SCALE = 12;

__global__ void __launch_bounds__(128) fp8_probe_kernel(const __grid_constant__ CUtensorMap tmA8,
                                                        const __grid_constant__ CUtensorMap tmB8,
                                                        const __grid_constant__ CUtensorMap tmA16,
                                                        const __grid_constant__ CUtensorMap tmB16, float* out, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a8 = smem;                 // [128][128] e4m3   16 KB
  uint8_t* b8 = smem + 16384;         // [ 64][128] e4m3    8 KB
  uint8_t* a16 = smem + 24576;        // [128][ 64] fp16   16 KB
  uint8_t* b16 = smem + 40960;        // [ 64][ 64] fp16    8 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::mbar_init(done, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) { ptx::tmem_alloc(slot, 64); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    ptx::mbar_expect_tx(bar, 49152);
    ptx::tma_load_2d(a8, &tmA8, bar, 0, 0);
    ptx::tma_load_2d(b8, &tmB8, bar, 0, 0);
    ptx::tma_load_2d(a16, &tmA16, bar, 0, 0);
    ptx::tma_load_2d(b16, &tmB16, bar, 0, 0);
    ptx::mbar_wait(bar, 0, 92);
    ptx::tc_fence_after();
    // instruction descriptors: c_format F32 (1) at bit 4, a/b formats 0 (E4M3 resp. F16), N>>3 at bit 17, M>>4 at bit 24
    constexpr uint32_t idesc = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t zero = 0;
    const uint64_t da8 = ptx::umma_desc_k_sw128(ptx::smem_u32(a8)), db8 = ptx::umma_desc_k_sw128(ptx::smem_u32(b8));
    const uint64_t da16 = ptx::umma_desc_k_sw128(ptx::smem_u32(a16)), db16 = ptx::umma_desc_k_sw128(ptx::smem_u32(b16));
    if (mode != 2) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {          // K = 4 x 32 fp8 elements, 32 bytes per step
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}" ::"r"(tmem),
                     "l"(da8 + 2 * k), "l"(db8 + 2 * k), "r"(idesc), "r"(k != 0 ? 1u : 0u), "r"(zero), "r"(zero), "r"(zero),
                     "r"(zero)
                     : "memory");
      }
    }
    if (mode != 1) {
      // first fp16 MMA: D = D * 2^-SCALE + A.B (mode 0: after the fp8 part; mode 2: fp16 part alone, no accumulate)
      if (mode == 0) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p, %9;\n\t}" ::"r"(tmem),
                     "l"(da16), "l"(db16), "r"(idesc), "r"(1u), "r"(zero), "r"(zero), "r"(zero), "r"(zero),
                     "n"(FP8_PROBE_SCALE)
                     : "memory");
      } else {
        ptx::mma_f16_ss(tmem, da16, db16, idesc, 0u);
      }
#pragma unroll
      for (int k = 1; k < 4; ++k) ptx::mma_f16_ss(tmem, da16 + 2 * k, db16 + 2 * k, idesc, 1u);
    }
    ptx::mma_commit(done);
  }
  ptx::mbar_wait(done, 0, 93);
  ptx::tc_fence_after();
  const int r = warp * 32 + lane;
  for (int c32 = 0; c32 < 2; ++c32) {
    uint32_t v[32];
    ptx::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c32 * 32, v);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) out[r * 64 + c32 * 32 + j] = __uint_as_float(v[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 64); }
}

// MMA issue-to-retire rate of one SM: `iters` groups of 4 back-to-back MMAs (M = 128, N = n) on whatever bytes sit in
// shared memory, kind::f16 (K = 16) or kind::f8f6f4 (K = 32); cycles[0] = clock64 ticks from first issue to commit arrival.
__global__ void __launch_bounds__(128) mma_rate_kernel(long long* cycles, int fp8, int n, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + 49152);
  uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x * 16; i < 49152; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  ptx::fence_proxy_async();
  if (threadIdx.x == 0) { ptx::mbar_init(done, 1); ptx::fence_mbar_init(); }
  if (warp == 0) { ptx::tmem_alloc(slot, 256); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (fp8 ? 0u : ((1u << 7) | (1u << 10))) | (((uint32_t)n >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t da = ptx::umma_desc_k_sw128(ptx::smem_u32(smem)), db = ptx::umma_desc_k_sw128(ptx::smem_u32(smem + 16384));
    const uint32_t zero = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (fp8)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}" ::"r"(tmem),
                       "l"(da + 2 * k), "l"(db + 2 * k), "r"(idesc), "r"(1u), "r"(zero), "r"(zero), "r"(zero), "r"(zero)
                       : "memory");
        else
          ptx::mma_f16_ss(tmem, da + 2 * k, db + 2 * k, idesc, 1u);
      }
    }
    ptx::mma_commit(done);
    ptx::mbar_wait(done, 0, 94);
    cycles[0] = clock64() - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 256); }
}


}  // namespace ammc

using namespace ammc;

// a8 [128][128], b8 [64][128] e4m3 bytes; a16 [128][64], b16 [64][64] fp16; out [128][64] fp32.
// mode 0: out = (a8.b8^T) * 2^-12 + a16.b16^T ; mode 1: out = a8.b8^T ; mode 2: out = a16.b16^T
extern "C" int ammc_debug_fp8_probe(const void* a8, const void* b8, const void* a16, const void* b16, float* out, int mode,
                                    void* stream) {
  AMMC_REQUIRE(a8 && b8 && a16 && b16 && out && mode >= 0 && mode <= 2, "bad argument");
  CUtensorMap mA8, mB8, mA16, mB16;
  {
    uint64_t dims[2] = {128, 128}, strides[1] = {128};
    uint32_t box[2] = {128, 128};
    if (int rc = make_map_generic(&mA8, a8, 1, 2, dims, strides, box, 1)) return rc;
  }
  {
    uint64_t dims[2] = {128, 64}, strides[1] = {128};
    uint32_t box[2] = {128, 64};
    if (int rc = make_map_generic(&mB8, b8, 1, 2, dims, strides, box, 1)) return rc;
  }
  {
    uint64_t dims[2] = {64, 128}, strides[1] = {128};
    uint32_t box[2] = {64, 128};
    if (int rc = make_map_generic(&mA16, a16, 2, 2, dims, strides, box, 1)) return rc;
  }
  {
    uint64_t dims[2] = {64, 64}, strides[1] = {128};
    uint32_t box[2] = {64, 64};
    if (int rc = make_map_generic(&mB16, b16, 2, 2, dims, strides, box, 1)) return rc;
  }
  const int smem = 49152 + 256 + 1024;
  AMMC_CUDA_CHECK(cudaFuncSetAttribute(fp8_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  fp8_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(mA8, mB8, mA16, mB16, out, mode);
  AMMC_LAUNCH_CHECK("fp8_probe_kernel");
  return 0;
}

// cycles: device long long[1].  fp8 = 0: kind::f16 on bf16 (K = 16 per MMA); 1: kind::f8f6f4 on e4m3 (K = 32 per MMA).
extern "C" int ammc_debug_mma_rate(long long* cycles, int fp8, int n, int iters, void* stream) {
  AMMC_REQUIRE(cycles && (n == 64 || n == 128 || n == 256) && iters > 0 && iters <= 100000, "bad argument");
  const int smem = 49152 + 256 + 1024;
  AMMC_CUDA_CHECK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  mma_rate_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(cycles, fp8, n, iters);
  AMMC_LAUNCH_CHECK("mma_rate_kernel");
  return 0;
}
#endif  // AMMC_DEBUG_PROBES
