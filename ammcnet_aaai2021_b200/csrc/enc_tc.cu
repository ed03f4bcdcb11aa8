// `enc` 1x1 convolution of the memory module on the tensor cores, straight from the fp32 NCHW features.
//
//   z[n, :] = enc_w . x[n, :] + enc_b        (reference Code/models/unet.py:321,326)    C -> D = 64
//
// x is what the cuDNN encoder leaves in HBM: NCHW fp32, i.e. for a pixel tile the *pixels* are contiguous and the
// contraction axis (channels) is strided -- not a layout tcgen05 can consume as fp32.  Instead of a separate pack pass
// (read 134 MB + write 134 MB at b=64) the kernel converts on the fly:
//   warp 0      TMA: fp32 box [64 channels][128 pixels] (no swizzle) + the matching 64-channel slices of the bf16 hi/lo
//               weight planes [64 d][64 ch] (128B swizzle) per pipeline stage
//   warps 2-9   converters: thread = pixel x channel half; read its 32 staged fp32 values, split x = hi + lo (bf16), write both as rows
//               of the canonical K-major 128B-swizzled UMMA tile (16-byte chunk index XOR (row & 7)), then
//               fence.proxy.async + mbarrier arrive
//   warp 1      MMA: per stage  hi*W_hi + hi*W_lo + lo*W_hi  (M128 x N64 x K16, fp32 accumulate in TMEM)
//   warps 10-13 epilogue: + bias -> z fp32 [N, D], its row-scaled fp16 rounding zp (the addressing filter's operand), ||z||^2
// Error of the split-bf16 x3 product: ~2^-17 relative (measured on the conv kernel), i.e. z agrees with the fp32 FFMA
// kernel to ~1e-5; the exact fp32 refine stage then ranks the candidates with that z.
#include "common.cuh"
#include "ptx.cuh"
#include <cuda_bf16.h>

namespace ammc {

constexpr int ENC_THREADS = 448;                   // warp 0 TMA, 1 MMA, 2-9 converters, 10-13 epilogue
constexpr int ENC_D = 64;                          // output channels handled by this kernel
constexpr int ENC_BK = 64;                         // channels per stage
constexpr int ENC_STAGE_X = ENC_BK * 128 * 4;      // fp32 staging [64 ch][128 px]            32 KB
constexpr int ENC_STAGE_A = 128 * ENC_BK * 2;      // one bf16 A tile [128 px][64 ch]          16 KB
constexpr int ENC_STAGE_W = ENC_D * ENC_BK * 2;    // one bf16 weight slice [64 d][64 ch]       8 KB
// two rings: the fp32 staging buffers are released as soon as the converters have read them (3 deep, so TMA runs ahead
// of the conversion), the bf16 operand buffers (A hi/lo written by the converters + the weight slices) when their MMAs retire
constexpr int ENC_X_STAGES = 3;
constexpr int ENC_AW_STAGE = 2 * ENC_STAGE_A + 2 * ENC_STAGE_W;              // 48 KB
constexpr int ENC_AW_STAGES = 2;
constexpr int ENC_AW_OFFSET = ENC_X_STAGES * ENC_STAGE_X;                   // 96 KB
constexpr int ENC_BAR_OFFSET = ENC_AW_OFFSET + ENC_AW_STAGES * ENC_AW_STAGE;   // 192 KB
constexpr int ENC_SMEM = ENC_BAR_OFFSET + 256 + 1024;

struct EncParams {
  int N, HW, C, tiles;
  const float* bias;
  float* z;                 // [N][64]
  __nv_bfloat16* zp;        // [N][64] fp16 bits of z * s_n (s_n: per-row power of two), or null
  float* znorm2;            // [N][2] (||z_n||^2, 1 / s_n), or null
  unsigned* amax_bits;      // AMAX: max |x| over the whole input as the bits of a non-negative float (atomicMax; zeroed by the host)
};

template <bool AMAX>
__global__ void __launch_bounds__(ENC_THREADS, 1)
enc_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const EncParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* x_full = reinterpret_cast<uint64_t*>(smem + ENC_BAR_OFFSET);     // TMA landed an fp32 staging buffer
  uint64_t* x_empty = x_full + ENC_X_STAGES;                                 // the 128 converter threads have read it
  uint64_t* w_full = x_empty + ENC_X_STAGES;                                 // TMA landed the weight slices of an operand stage
  uint64_t* conv_bar = w_full + ENC_AW_STAGES;                               // converters wrote the A tiles
  uint64_t* aw_free = conv_bar + ENC_AW_STAGES;                              // MMAs of the operand stage retired
  uint64_t* tmem_full = aw_free + ENC_AW_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb_per_tile = p.C / ENC_BK;
  const int tiles_per_img = p.HW / 128;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmX);
    ptx::prefetch_tensormap(&tmW);
    for (int s = 0; s < ENC_X_STAGES; ++s) { ptx::mbar_init(&x_full[s], 1); ptx::mbar_init(&x_empty[s], 256); }
    for (int s = 0; s < ENC_AW_STAGES; ++s) {
      ptx::mbar_init(&w_full[s], 1);
      ptx::mbar_init(&conv_bar[s], 256);
      ptx::mbar_init(&aw_free[s], 1);
    }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tmem_full[a], 1); ptx::mbar_init(&tmem_empty[a], 128); }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 2 * ENC_D);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int sx = 0; uint32_t phx = 0;
      int sa = 0; uint32_t pha = 0;
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
        const int img = t / tiles_per_img, p0 = (t % tiles_per_img) * 128;
        for (int kb = 0; kb < kb_per_tile; ++kb) {
          ptx::mbar_wait(&x_empty[sx], phx ^ 1, 51);
          ptx::mbar_expect_tx(&x_full[sx], ENC_STAGE_X);
          ptx::tma_load_3d(smem + sx * ENC_STAGE_X, &tmX, &x_full[sx], p0, kb * ENC_BK, img);
          if (++sx == ENC_X_STAGES) { sx = 0; phx ^= 1; }
          ptx::mbar_wait(&aw_free[sa], pha ^ 1, 57);
          ptx::mbar_expect_tx(&w_full[sa], 2 * ENC_STAGE_W);
          uint8_t* wdst = smem + ENC_AW_OFFSET + sa * ENC_AW_STAGE + 2 * ENC_STAGE_A;
          ptx::tma_load_3d(wdst, &tmW, &w_full[sa], kb * ENC_BK, 0, 0);
          ptx::tma_load_3d(wdst + ENC_STAGE_W, &tmW, &w_full[sa], kb * ENC_BK, 0, 1);
          if (++sa == ENC_AW_STAGES) { sa = 0; pha ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    {
      // whole warp converged, one elected lane issues (ptx::mma_f16_ss_warp): N = 64 MMAs are short, the issue loop counts
      constexpr uint32_t idesc = ptx::umma_idesc(1, 128, ENC_D);
      int s = 0; uint32_t ph = 0;
      int it = 0;
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_ph = (it >> 1) & 1;
        ptx::mbar_wait(&tmem_empty[acc], acc_ph ^ 1, 52);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ENC_D;
        for (int kb = 0; kb < kb_per_tile; ++kb) {
          ptx::mbar_wait(&w_full[s], ph, 53);        // weight slices landed
          ptx::mbar_wait(&conv_bar[s], ph, 54);      // A tiles written by the converters
          ptx::tc_fence_after();
          const uint32_t base = ptx::smem_u32(smem + ENC_AW_OFFSET + s * ENC_AW_STAGE);
          const uint64_t a_hi = ptx::umma_desc_k_sw128(base);
          const uint64_t a_lo = ptx::umma_desc_k_sw128(base + ENC_STAGE_A);
          const uint64_t w_hi = ptx::umma_desc_k_sw128(base + 2 * ENC_STAGE_A);
          const uint64_t w_lo = ptx::umma_desc_k_sw128(base + 2 * ENC_STAGE_A + ENC_STAGE_W);
#pragma unroll
          for (int k4 = 0; k4 < ENC_BK / 16; ++k4) {
            ptx::mma_f16_ss_warp(d_tmem, a_hi + 2 * k4, w_hi + 2 * k4, idesc, (kb | k4) != 0 ? 1u : 0u);
            ptx::mma_f16_ss_warp(d_tmem, a_hi + 2 * k4, w_lo + 2 * k4, idesc, 1u);
            ptx::mma_f16_ss_warp(d_tmem, a_lo + 2 * k4, w_hi + 2 * k4, idesc, 1u);
          }
          ptx::mma_commit_warp(&aw_free[s]);
          if (kb == kb_per_tile - 1) ptx::mma_commit_warp(&tmem_full[acc]);
          if (++s == ENC_AW_STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp < 10) {
    // ---------------------------------------------------------------- converters: fp32 staging -> bf16 hi/lo UMMA tiles
    // eight warps: two per 32-pixel group, each converting 32 of the stage's 64 channels (one converter warp per
    // scheduler was the bound: ~450 dependent instructions per thread and k-block)
    const int row = ((warp - 2) & 3) * 32 + lane;             // pixel within the tile
    const int c8_0 = ((warp - 2) >> 2) * (ENC_BK / 16);       // first 8-channel chunk of this warp's half
    int sx = 0; uint32_t phx = 0;
    int sa = 0; uint32_t pha = 0;
    float amax = 0.f;                                         // AMAX: the values pass through registers here anyway
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
      for (int kb = 0; kb < kb_per_tile; ++kb) {
        ptx::mbar_wait(&x_full[sx], phx, 55);                 // fp32 staging landed
        ptx::mbar_wait(&aw_free[sa], pha ^ 1, 58);            // the MMAs that last read these A tiles have retired
        // explicit shared-space accesses (the generic pointer derived from the aligned base would compile to LD.E/ST.E)
        const uint32_t xs = ptx::smem_u32(smem + sx * ENC_STAGE_X) + row * 4;                     // [ch][128 px] fp32
        const uint32_t a_hi = ptx::smem_u32(smem + ENC_AW_OFFSET + sa * ENC_AW_STAGE) + row * 128;
        const uint32_t a_lo = a_hi + ENC_STAGE_A;
#pragma unroll
        for (int c8 = c8_0; c8 < c8_0 + ENC_BK / 16; ++c8) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = ptx::lds_f32(xs + (c8 * 8 + j) * 512);
          if (AMAX) {
#pragma unroll
            for (int j = 0; j < 8; j += 2) amax = fmaxf(amax, fmaxf(fabsf(v[j]), fabsf(v[j + 1])));
          }
          uint32_t hp[4], lp[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            // packed conversions: one F2FP per pair; float(hi) is a 16-bit shift of its bit pattern
            const uint32_t h = ptx::pack_bf16x2(v[2 * j], v[2 * j + 1]);
            const float r0 = v[2 * j] - __uint_as_float(h << 16);
            const float r1 = v[2 * j + 1] - __uint_as_float(h & 0xffff0000u);
            hp[j] = h;
            lp[j] = ptx::pack_bf16x2(r0, r1);
          }
          const uint32_t chunk = (uint32_t)((c8 ^ (row & 7)) * 16);   // 128B swizzle: 16-byte chunk index XOR (row mod 8)
          ptx::sts_v4(a_hi + chunk, hp[0], hp[1], hp[2], hp[3]);
          ptx::sts_v4(a_lo + chunk, lp[0], lp[1], lp[2], lp[3]);
        }
        ptx::mbar_arrive(&x_empty[sx]);                       // staging buffer read: TMA may refill it
        ptx::fence_proxy_async();                             // generic-proxy writes -> visible to the tensor core
        ptx::mbar_arrive(&conv_bar[sa]);
        if (++sx == ENC_X_STAGES) { sx = 0; phx ^= 1; }
        if (++sa == ENC_AW_STAGES) { sa = 0; pha ^= 1; }
      }
    }
    if (AMAX) {
      amax = warp_max(amax);
      if (lane == 0 && amax > 0.f) atomicMax(p.amax_bits, __float_as_uint(amax));
    }
  } else {
    // ---------------------------------------------------------------- epilogue: z, bf16(z), ||z||^2
    const int q = warp & 3;
    const int r = q * 32 + lane;
    int it = 0;
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      const size_t n = (size_t)t * 128 + r;
      ptx::mbar_wait(&tmem_full[acc], acc_ph, 56);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ENC_D;
      // pass 1: the row's largest component fixes its power-of-two fp16 scale (addressing filter operand, addr_tc.cu)
      float zmax = 0.f;
      if (p.zp) {
#pragma unroll
        for (int c32 = 0; c32 < ENC_D / 32; ++c32) {
          uint32_t v[32];
          ptx::tmem_ld_32x32(taddr + c32 * 32, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) zmax = fmaxf(zmax, fabsf(__uint_as_float(v[j]) + __ldg(p.bias + c32 * 32 + j)));
        }
      }
      const float zsc = q_scale_for_bound(zmax);
      float zn2 = 0.f;
#pragma unroll
      for (int c32 = 0; c32 < ENC_D / 32; ++c32) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(taddr + c32 * 32, v);
        ptx::tmem_ld_wait();
        // 32-byte stores (STG.256): one request per sector; a row of z is 256 B, of zp 128 B, both 32-byte aligned
        float* zr = p.z + n * ENC_D + c32 * 32;
        uint32_t zb[16];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float zv = __uint_as_float(v[8 * g + j]) + __ldg(p.bias + c32 * 32 + 8 * g + j);
            zn2 = fmaf(zv, zv, zn2);
            o[j] = __float_as_uint(zv);
          }
          ptx::stg_v8(zr + 8 * g, o);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            zb[4 * g + j] = ptx::pack_f16x2(__uint_as_float(o[2 * j]) * zsc, __uint_as_float(o[2 * j + 1]) * zsc);
        }
        if (p.zp) {
          __nv_bfloat16* zpr = p.zp + n * ENC_D + c32 * 32;
          const uint32_t (&z8)[2][8] = *reinterpret_cast<const uint32_t (*)[2][8]>(zb);
          ptx::stg_v8(zpr, z8[0]);
          ptx::stg_v8(zpr + 16, z8[1]);
        }
      }
      if (p.znorm2) reinterpret_cast<float2*>(p.znorm2)[n] = make_float2(zn2, 1.f / zsc);
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tmem_empty[acc]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 2 * ENC_D);
  }
}

int make_map_generic(CUtensorMap* m, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                     const uint64_t* strides, const uint32_t* box, int swizzle128);   // amft_conv.cu
int pack_weights_1x1(const float* w, void* wp, int Cout, int Cin, cudaStream_t st);

bool enc_tc_supported(int b, int HW, int C, int D) { return D == ENC_D && C % ENC_BK == 0 && HW % 128 == 0 && b > 0; }

size_t enc_tc_ws_bytes(int C) { return align_up((size_t)2 * ENC_D * C * 2, 256); }

// x [b][C][HW] fp32, enc_w [64][C] fp32 -> z [N][64] (+ zp, znorm2 when non-null).  wp_ws: enc_tc_ws_bytes(C).
int run_enc_tc(const float* x, const float* enc_w, const float* enc_b, float* z, __nv_bfloat16* zp, float* znorm2,
               void* wp_ws, bool wp_ready, unsigned* amax_bits, int b, int HW, int C, cudaStream_t st) {
  if (!wp_ready)
    if (int rc = pack_weights_1x1(enc_w, wp_ws, ENC_D, C, st)) return rc;
  CUtensorMap tmX, tmW;
  {
    uint64_t dims[3] = {(uint64_t)HW, (uint64_t)C, (uint64_t)b};
    uint64_t strides[2] = {(uint64_t)HW * 4, (uint64_t)C * HW * 4};
    uint32_t box[3] = {128, (uint32_t)ENC_BK, 1};
    if (int rc = make_map_generic(&tmX, x, 4, 3, dims, strides, box, 0)) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)C, (uint64_t)ENC_D, 2};
    uint64_t strides[2] = {(uint64_t)C * 2, (uint64_t)ENC_D * C * 2};
    uint32_t box[3] = {(uint32_t)ENC_BK, (uint32_t)ENC_D, 1};
    if (int rc = make_map_generic(&tmW, wp_ws, 2, 3, dims, strides, box, 1)) return rc;
  }
  EncParams p;
  p.N = b * HW; p.HW = HW; p.C = C; p.tiles = b * (HW / 128);
  p.bias = enc_b; p.z = z; p.zp = zp; p.znorm2 = znorm2; p.amax_bits = amax_bits;
  static bool configured[64] = {false};
  int dev = 0;
  AMMC_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    AMMC_CUDA_CHECK(cudaFuncSetAttribute(enc_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ENC_SMEM));
    AMMC_CUDA_CHECK(cudaFuncSetAttribute(enc_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ENC_SMEM));
    configured[dev] = true;
  }
  if (amax_bits) {
    AMMC_CUDA_CHECK(cudaMemsetAsync(amax_bits, 0, 4, st));
    enc_tc_kernel<true><<<min(num_sms(), p.tiles), ENC_THREADS, ENC_SMEM, st>>>(tmX, tmW, p);
  } else {
    enc_tc_kernel<false><<<min(num_sms(), p.tiles), ENC_THREADS, ENC_SMEM, st>>>(tmX, tmW, p);
  }
  AMMC_LAUNCH_CHECK("enc_tc_kernel");
  return 0;
}

AMMC_DEFINE_TIMEOUT_READER(timeout_reader_enc)

}  // namespace ammc
