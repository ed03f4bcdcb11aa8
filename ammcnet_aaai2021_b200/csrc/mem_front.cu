// Front half of the memory module in ONE persistent kernel (shipped shapes: D = 64, M <= 256):
//
//   z = enc(x)            1x1 conv C -> 64 on tcgen05, fp32 NCHW input converted on the fly      (reference unet.py:321,326)
//   dist, top-k           ||z||^2 - 2 z.e + ||e||^2, k nearest items                              (unet.py:283-293)
//   q1, read, commit      straight-through top-1 value, gathered top-k rows, per-pixel SSE        (unet.py:295-297,310-311)
//
// Before: enc_tc_kernel -> (z, bf16 z, ||z||^2 through HBM) -> addr_tc_kernel -> (candidate lists through HBM) ->
// refine_kernel.  Here the three stages share one CTA per SM and the intermediates never leave the chip:
//
//   warp 0        TMA: fp32 boxes [32 ch][128 px] of x into a 4-deep staging ring (two independent halves), the bf16 hi/lo
//                 slices of enc.weight per K block, and -- once -- the whole bf16 bank (32 KB, stays resident)
//   warps 2-9     converters: staged fp32 -> bf16 hi/lo rows of the K-major 128B-swizzled UMMA A tiles (as enc_tc.cu);
//                 max|x| for the q-plane scale falls out of the same pass
//   warp 1        MMA: enc (hi.Whi + hi.Wlo + lo.Whi, M128 x N64 x K16) into one of two 64-column accumulators, and for
//                 the PREVIOUS tile the similarity  S = fp16(z s_n) . fp16(bank t)^T  (M128 x N256 x K16 x 4) into a 256-column
//                 accumulator -- issued two K blocks into the next tile so the tensor core never waits for the epilogue
//   warps 10-13   epilogue, thread = pixel.  P1: z = acc + bias -> HBM (fp32, for backward / the rare exact re-scan),
//                 row-scaled fp16(z) -> the swizzled A tile of the similarity MMA, ||z||^2 with the summation tree of the fp32 kernel.
//                 P2: approximate scores from TMEM, k smallest + every column within a rigorous bf16 error margin
//                 (the filter of addr_tc.cu), then the EXACT fp32 distance of those few candidates with z re-read from
//                 its TMEM accumulator -- same fmaf chain as the generic kernel, so the indices are bit-identical --
//                 top-k, q1, bf16 hi/lo planes of the read (operand of the dec GEMM), per-pixel SSE.
//   Rows whose candidate list overflows (near-degenerate neighbourhoods) are queued for rescan_kernel (addr_tc.cu).
//
// Per 128-pixel tile the kernel reads 256 KB of x and writes 32 KB z + 32 KB q1 + 64 KB read planes + indices; the enc
// stage (HBM / converter bound) hides the addressing stages completely.
#include "common.cuh"
#include "ptx.cuh"
#include "topk.cuh"
#include "addr_tail.cuh"
#include <cuda_bf16.h>
#include <float.h>

namespace ammc {

constexpr int MF_THREADS = 448;                    // warp 0 TMA, 1 MMA, 2-9 converters, 10-13 epilogue
constexpr int MF_D = 64;
constexpr int MF_BK = 64;                          // channels per enc K block
constexpr int MF_MPAD = 256;                       // items, padded (zero rows, +inf norms)
constexpr int MF_CAP = 16;                         // exact-refine candidates per pixel
constexpr int MF_GS = 4;                           // candidates refined together (shared z reads, independent chains)
constexpr int MF_X_STAGE = 32 * 128 * 4;           // fp32 staging box [32 ch][128 px]          16 KB
constexpr int MF_X_STAGES = 4;                     // stages h, h + 2 belong to channel half h
constexpr int MF_STAGE_A = 128 * MF_BK * 2;        // one bf16 A tile [128 px][64 ch]           16 KB
constexpr int MF_STAGE_W = MF_D * MF_BK * 2;       // one bf16 weight slice [64 d][64 ch]        8 KB
constexpr int MF_AW_STAGE = 2 * MF_STAGE_A + 2 * MF_STAGE_W;   // 48 KB
constexpr int MF_AW_STAGES = 2;
constexpr int MF_AW_OFFSET = MF_X_STAGES * MF_X_STAGE;                      //  64 KB
constexpr int MF_BANK_OFFSET = MF_AW_OFFSET + MF_AW_STAGES * MF_AW_STAGE;   // 160 KB
constexpr int MF_BANK_BYTES = MF_MPAD * MF_D * 2;                           //  32 KB
constexpr int MF_ZP_OFFSET = MF_BANK_OFFSET + MF_BANK_BYTES;                // 192 KB
constexpr int MF_LIST_OFFSET = MF_ZP_OFFSET + MF_STAGE_A;                   // 208 KB: uint16 [128][MF_CAP]
constexpr int MF_TAB_OFFSET = MF_LIST_OFFSET + 128 * MF_CAP * 2;            // en2pad [256] fp32, bias [64] fp32
constexpr int MF_BAR_OFFSET = MF_TAB_OFFSET + MF_MPAD * 4 + MF_D * 4;
constexpr int MF_SMEM = MF_BAR_OFFSET + 256 + 1024;
static_assert(MF_SMEM <= 232448, "shared memory budget");

struct FrontParams {
  int N, HW, C, M, tiles;
  const float* bias;        // [64]
  const float* bank_t;      // [M][64] fp32
  const float* en2;         // [M]
  const float* en2pad;      // [256], +inf beyond M
  const float* emax;        // [2] max ||e||, 1 / t (inverse power-of-two fp16 scale of the bank)
  float* z;                 // [N][64]
  float* q1;                // [N][64]
  int64_t* idx;             // [N][K]
  float* sse_px;            // [N]
  __nv_bfloat16* read_planes;   // [2][N][K*64] or null
  long long read_plane_stride;
  int* stats;               // [0] rows queued for the exact re-scan, [2] length of the re-scan list
  int* rescan_list;         // [N]
  unsigned* amax_bits;      // AMAX: max |x| (atomicMax on the bit pattern; zeroed by the host)
};

// branch-free insertion of x into the ascending list m[0..KSEL)
template <int KSEL>
__device__ __forceinline__ void mf_sel_insert(float (&m)[KSEL], float x) {
#pragma unroll
  for (int i = 0; i < KSEL; ++i) {
    const float lo = fminf(m[i], x);
    x = fmaxf(m[i], x);
    m[i] = lo;
  }
}

// XBF16: x arrives as bf16 NCHW (the bf16 feature-I/O variant, BASELINE configs[2]).  The staging boxes are half as large,
// the converters copy the values into the hi tile unchanged (x = hi exactly, lo = 0) and the lo.Whi MMA is skipped -- it
// would add exact zeros, so z and everything after it is bit-identical to the fp32 kernel fed the widened tensor.
template <int K, bool AMAX, bool XBF16>
__global__ void __launch_bounds__(MF_THREADS, 1)
mem_front_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                 const __grid_constant__ CUtensorMap tmBank, const FrontParams p) {
  constexpr int KSEL = K <= 2 ? 2 : 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* x_full = reinterpret_cast<uint64_t*>(smem + MF_BAR_OFFSET);     // TMA landed a staging box
  uint64_t* x_empty = x_full + MF_X_STAGES;                                 // its 128 converter threads have read it
  uint64_t* w_full = x_empty + MF_X_STAGES;                                 // weight slices of an operand stage landed
  uint64_t* conv_bar = w_full + MF_AW_STAGES;                               // converters wrote the A tiles
  uint64_t* aw_free = conv_bar + MF_AW_STAGES;                              // MMAs of the operand stage retired
  uint64_t* tmem_full = aw_free + MF_AW_STAGES;                             // enc accumulator complete
  uint64_t* tmem_empty = tmem_full + 2;                                     // epilogue done with it (after P2)
  uint64_t* bank_full = tmem_empty + 2;
  uint64_t* zp_ready = bank_full + 1;                                       // bf16(z) tile written (128 threads)
  uint64_t* s_full = zp_ready + 1;                                          // similarity accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 1);
  float* en2pad_s = reinterpret_cast<float*>(smem + MF_TAB_OFFSET);
  float* bias_s = en2pad_s + MF_MPAD;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb_per_tile = p.C / MF_BK;
  const int tiles_per_img = p.HW / 128;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmX);
    ptx::prefetch_tensormap(&tmW);
    ptx::prefetch_tensormap(&tmBank);
    for (int s = 0; s < MF_X_STAGES; ++s) { ptx::mbar_init(&x_full[s], 1); ptx::mbar_init(&x_empty[s], 128); }
    for (int s = 0; s < MF_AW_STAGES; ++s) {
      ptx::mbar_init(&w_full[s], 1);
      ptx::mbar_init(&conv_bar[s], 256);
      ptx::mbar_init(&aw_free[s], 1);
    }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tmem_full[a], 1); ptx::mbar_init(&tmem_empty[a], 128); }
    ptx::mbar_init(bank_full, 1);
    ptx::mbar_init(zp_ready, 128);
    ptx::mbar_init(s_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);          // [0,128): two enc accumulators; [256,512): similarity
    ptx::tmem_relinquish();
  }
  for (int i = threadIdx.x; i < MF_MPAD + MF_D; i += MF_THREADS)
    en2pad_s[i] = i < MF_MPAD ? __ldg(p.en2pad + i) : __ldg(p.bias + i - MF_MPAD);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_expect_tx(bank_full, MF_BANK_BYTES);
      ptx::tma_load_2d(smem + MF_BANK_OFFSET, &tmBank, bank_full, 0, 0);
      int kx = 0;                        // staging boxes issued so far: stage = kx & 3, parity from kx >> 2
      int sa = 0; uint32_t pha = 0;
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
        const int img = t / tiles_per_img, p0 = (t % tiles_per_img) * 128;
        for (int kb = 0; kb < kb_per_tile; ++kb) {
#pragma unroll
          for (int h = 0; h < 2; ++h, ++kx) {
            const int sx = kx & 3;
            ptx::mbar_wait(&x_empty[sx], ((kx >> 2) & 1) ^ 1, 61);
            ptx::mbar_expect_tx(&x_full[sx], XBF16 ? MF_X_STAGE / 2 : MF_X_STAGE);
            ptx::tma_load_3d(smem + sx * MF_X_STAGE, &tmX, &x_full[sx], p0, kb * MF_BK + 32 * h, img);
          }
          ptx::mbar_wait(&aw_free[sa], pha ^ 1, 62);
          ptx::mbar_expect_tx(&w_full[sa], 2 * MF_STAGE_W);
          uint8_t* wdst = smem + MF_AW_OFFSET + sa * MF_AW_STAGE + 2 * MF_STAGE_A;
          ptx::tma_load_3d(wdst, &tmW, &w_full[sa], kb * MF_BK, 0, 0);
          ptx::tma_load_3d(wdst + MF_STAGE_W, &tmW, &w_full[sa], kb * MF_BK, 0, 1);
          if (++sa == MF_AW_STAGES) { sa = 0; pha ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp converged, one lane issues)
    constexpr uint32_t idesc_enc = ptx::umma_idesc(1, 128, MF_D);
    constexpr uint32_t idesc_sim = ptx::umma_idesc(0, 128, MF_MPAD);      // fp16 operands
    const uint32_t s_tmem = tmem_base + 256;
    const uint64_t zp_desc = ptx::umma_desc_k_sw128(ptx::smem_u32(smem + MF_ZP_OFFSET));
    const uint64_t bank_desc = ptx::umma_desc_k_sw128(ptx::smem_u32(smem + MF_BANK_OFFSET));
    const int kb_sim = kb_per_tile > 2 ? 2 : kb_per_tile - 1;    // where, inside the next tile, the similarity is issued
    auto issue_similarity = [&](int j) {
      ptx::mbar_wait(zp_ready, (uint32_t)(j & 1), 63);
      ptx::tc_fence_after();
#pragma unroll
      for (int k4 = 0; k4 < MF_D / 16; ++k4)
        ptx::mma_f16_ss_warp(s_tmem, zp_desc + 2 * k4, bank_desc + 2 * k4, idesc_sim, k4 != 0 ? 1u : 0u);
      ptx::mma_commit_warp(s_full);
    };
    ptx::mbar_wait(bank_full, 0, 64);
    int s = 0; uint32_t ph = 0;
    int it = 0;
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      ptx::mbar_wait(&tmem_empty[acc], acc_ph ^ 1, 65);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * MF_D;
      for (int kb = 0; kb < kb_per_tile; ++kb) {
        if (kb == kb_sim && it > 0) issue_similarity(it - 1);
        ptx::mbar_wait(&w_full[s], ph, 66);
        ptx::mbar_wait(&conv_bar[s], ph, 67);
        ptx::tc_fence_after();
        const uint32_t base = ptx::smem_u32(smem + MF_AW_OFFSET + s * MF_AW_STAGE);
        const uint64_t a_hi = ptx::umma_desc_k_sw128(base);
        const uint64_t a_lo = ptx::umma_desc_k_sw128(base + MF_STAGE_A);
        const uint64_t w_hi = ptx::umma_desc_k_sw128(base + 2 * MF_STAGE_A);
        const uint64_t w_lo = ptx::umma_desc_k_sw128(base + 2 * MF_STAGE_A + MF_STAGE_W);
#pragma unroll
        for (int k4 = 0; k4 < MF_BK / 16; ++k4) {
          ptx::mma_f16_ss_warp(d_tmem, a_hi + 2 * k4, w_hi + 2 * k4, idesc_enc, (kb | k4) != 0 ? 1u : 0u);
          ptx::mma_f16_ss_warp(d_tmem, a_hi + 2 * k4, w_lo + 2 * k4, idesc_enc, 1u);
          if (!XBF16) ptx::mma_f16_ss_warp(d_tmem, a_lo + 2 * k4, w_hi + 2 * k4, idesc_enc, 1u);
        }
        ptx::mma_commit_warp(&aw_free[s]);
        if (kb == kb_per_tile - 1) ptx::mma_commit_warp(&tmem_full[acc]);
        if (++s == MF_AW_STAGES) { s = 0; ph ^= 1; }
      }
    }
    if (it > 0) issue_similarity(it - 1);
  } else if (warp < 10) {
    // ---------------------------------------------------------------- converters: fp32 staging -> bf16 hi/lo UMMA tiles
    const int row = ((warp - 2) & 3) * 32 + lane;             // pixel within the tile
    const int half = (warp - 2) >> 2;                         // which 32 of the K block's 64 channels
    int kx = 0;                                               // boxes of this half consumed: stage half + 2 * (kx & 1)
    int sa = 0; uint32_t pha = 0;
    float amax = 0.f;
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
      for (int kb = 0; kb < kb_per_tile; ++kb, ++kx) {
        const int sx = half + 2 * (kx & 1);
        ptx::mbar_wait(&x_full[sx], (uint32_t)((kx >> 1) & 1), 68);
        ptx::mbar_wait(&aw_free[sa], pha ^ 1, 69);
        const uint32_t xs = ptx::smem_u32(smem + sx * MF_X_STAGE) + row * (XBF16 ? 2 : 4);   // [32 ch][128 px]
        const uint32_t a_hi = ptx::smem_u32(smem + MF_AW_OFFSET + sa * MF_AW_STAGE) + row * 128;
        const uint32_t a_lo = a_hi + MF_STAGE_A;
#pragma unroll
        for (int c8l = 0; c8l < 4; ++c8l) {
          const int c8 = half * 4 + c8l;                               // 16-byte chunk of the 128-byte row
          const uint32_t chunk = (uint32_t)((c8 ^ (row & 7)) * 16);   // 128B swizzle
          if (XBF16) {
            uint32_t b[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) b[j] = ptx::lds_u16(xs + (c8l * 8 + j) * 256);
            if (AMAX) {
#pragma unroll
              for (int j = 0; j < 8; ++j) amax = fmaxf(amax, fabsf(__uint_as_float(b[j] << 16)));
            }
            ptx::sts_v4(a_hi + chunk, b[0] | (b[1] << 16), b[2] | (b[3] << 16), b[4] | (b[5] << 16), b[6] | (b[7] << 16));
          } else {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = ptx::lds_f32(xs + (c8l * 8 + j) * 512);
            if (AMAX) {
#pragma unroll
              for (int j = 0; j < 8; j += 2) amax = fmaxf(amax, fmaxf(fabsf(v[j]), fabsf(v[j + 1])));
            }
            uint32_t hp[4], lp[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) ptx::split_pack_bf16x2(v[2 * j], v[2 * j + 1], hp[j], lp[j]);
            ptx::sts_v4(a_hi + chunk, hp[0], hp[1], hp[2], hp[3]);
            ptx::sts_v4(a_lo + chunk, lp[0], lp[1], lp[2], lp[3]);
          }
        }
        ptx::mbar_arrive(&x_empty[sx]);
        ptx::fence_proxy_async();
        ptx::mbar_arrive(&conv_bar[sa]);
        if (++sa == MF_AW_STAGES) { sa = 0; pha ^= 1; }
      }
    }
    if (AMAX) {
      amax = warp_max(amax);
      if (lane == 0 && amax > 0.f) atomicMax(p.amax_bits, __float_as_uint(amax));
    }
  } else {
    // ---------------------------------------------------------------- epilogue: thread = pixel
    const int q = warp & 3;
    const int r = q * 32 + lane;
    uint16_t* list = reinterpret_cast<uint16_t*>(smem + MF_LIST_OFFSET) + r * MF_CAP;
    const uint32_t zp_row = ptx::smem_u32(smem + MF_ZP_OFFSET) + r * 128;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const float emax = __ldg(p.emax), tinv = __ldg(p.emax + 1);
    int it = 0;
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      const size_t n = (size_t)t * 128 + r;
      const uint32_t z_tmem = lane_base + acc * MF_D;
      // ---- P1: z, bf16(z), ||z||^2
      ptx::mbar_wait(&tmem_full[acc], acc_ph, 70);
      ptx::tc_fence_after();
      float zmax = 0.f;                            // pass 1: the row's power-of-two fp16 scale (filter operand)
#pragma unroll
      for (int c32 = 0; c32 < 2; ++c32) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(z_tmem + c32 * 32, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) zmax = fmaxf(zmax, fabsf(__uint_as_float(v[j]) + bias_s[c32 * 32 + j]));
      }
      const float zsc = q_scale_for_bound(zmax);
      float zs[4] = {0.f, 0.f, 0.f, 0.f};          // partial sums over d = part (mod 4): the tree of team_zn2
#pragma unroll
      for (int c32 = 0; c32 < 2; ++c32) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(z_tmem + c32 * 32, v);
        ptx::tmem_ld_wait();
        float* zr = p.z + n * MF_D + c32 * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float zv = __uint_as_float(v[8 * g + j]) + bias_s[c32 * 32 + 8 * g + j];
            zs[j & 3] = fmaf(zv, zv, zs[j & 3]);
            o[j] = __float_as_uint(zv);
          }
          ptx::stg_v8(zr + 8 * g, o);
          const int c8 = c32 * 4 + g;
          ptx::sts_v4(zp_row + (uint32_t)((c8 ^ (r & 7)) * 16),
                      ptx::pack_f16x2(__uint_as_float(o[0]) * zsc, __uint_as_float(o[1]) * zsc),
                      ptx::pack_f16x2(__uint_as_float(o[2]) * zsc, __uint_as_float(o[3]) * zsc),
                      ptx::pack_f16x2(__uint_as_float(o[4]) * zsc, __uint_as_float(o[5]) * zsc),
                      ptx::pack_f16x2(__uint_as_float(o[6]) * zsc, __uint_as_float(o[7]) * zsc));
        }
      }
      const float zn2 = (zs[0] + zs[1]) + (zs[2] + zs[3]);
      ptx::fence_proxy_async();
      ptx::mbar_arrive(zp_ready);
      // ---- P2: filter on the approximate scores, exact refine of the survivors
      ptx::mbar_wait(s_full, (uint32_t)(it & 1), 71);
      ptx::tc_fence_after();
      const uint32_t s_tmem = lane_base + 256;
      // |a~ - a| <= 4u(1+u) ||z|| ||e||, u = 2^-11 (fp16 rounding of both scaled operands, Cauchy-Schwarz); everything
      // within twice that of the KSEL-th smallest approximate score is a superset of the exact top-KSEL (addr_tc.cu)
      const float margin = 8.f * 0.00048828125f * 1.01f * sqrtf(zn2) * emax + 1e-5f * (zn2 + emax * emax) + 1e-30f;
      const float cs = -2.f * tinv / zsc;          // a~ = ||e||^2 + cs * (z~.e~): undoes the two power-of-two scales
      float m[KSEL];
#pragma unroll
      for (int i = 0; i < KSEL; ++i) m[i] = INFINITY;
#pragma unroll 1
      for (int c = 0; c < MF_MPAD / 32; ++c) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(s_tmem + c * 32, v);
        ptx::tmem_ld_wait();
        float ma[4][KSEL];
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int i = 0; i < KSEL; ++i) ma[g][i] = INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          mf_sel_insert<KSEL>(ma[j & 3], fmaf(cs, __uint_as_float(v[j]), en2pad_s[c * 32 + j]));
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int i = 0; i < KSEL; ++i) mf_sel_insert<KSEL>(m, ma[g][i]);
      }
      const float thr = fminf(m[KSEL - 1] + margin, FLT_MAX);   // padded columns score +inf and never hit
      int cnt = 0;
#pragma unroll 1
      for (int c = 0; c < MF_MPAD / 32; ++c) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(s_tmem + c * 32, v);
        ptx::tmem_ld_wait();
        uint32_t hits = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          hits |= (fmaf(cs, __uint_as_float(v[j]), en2pad_s[c * 32 + j]) <= thr ? 1u : 0u) << j;
        while (hits) {
          const int j = __ffs(hits) - 1;
          hits &= hits - 1;
          if (cnt < MF_CAP) list[cnt] = (uint16_t)(c * 32 + j);
          ++cnt;
        }
      }
      const bool rescan = cnt > MF_CAP || cnt < K;
      const int ncand = rescan ? 0 : cnt;
      TopK<K> top;
      top.init();
      const int ncand_max = __reduce_max_sync(0xffffffffu, ncand);
      // exact distances, MF_GS candidates at a time: one TMEM read of a z chunk serves all of them, their bank rows are
      // 4 x MF_GS independent 16-byte loads in flight and MF_GS independent fmaf chains (each still ascending in d)
      for (int base = 0; base < ncand_max; base += MF_GS) {
        const float4* e4[MF_GS];
        int col[MF_GS];
        float dot[MF_GS];
#pragma unroll
        for (int u = 0; u < MF_GS; ++u) {
          col[u] = base + u < ncand ? (int)list[base + u] : 0;
          e4[u] = reinterpret_cast<const float4*>(p.bank_t + (size_t)col[u] * MF_D);
          dot[u] = 0.f;
        }
#pragma unroll
        for (int c16 = 0; c16 < 4; ++c16) {
          float4 e[MF_GS][4];
#pragma unroll
          for (int u = 0; u < MF_GS; ++u)
#pragma unroll
            for (int g = 0; g < 4; ++g) e[u][g] = __ldg(e4[u] + c16 * 4 + g);
          uint32_t zc[16];
          ptx::tmem_ld_32x16(z_tmem + c16 * 16, zc);
          ptx::tmem_ld_wait();
          float zb[16];
#pragma unroll
          for (int d = 0; d < 16; ++d) zb[d] = __uint_as_float(zc[d]) + bias_s[c16 * 16 + d];
#pragma unroll
          for (int u = 0; u < MF_GS; ++u)
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              dot[u] = fmaf(zb[g * 4 + 0], e[u][g].x, dot[u]);
              dot[u] = fmaf(zb[g * 4 + 1], e[u][g].y, dot[u]);
              dot[u] = fmaf(zb[g * 4 + 2], e[u][g].z, dot[u]);
              dot[u] = fmaf(zb[g * 4 + 3], e[u][g].w, dot[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < MF_GS; ++u)
          if (base + u < ncand) top.insert(exact_dist(zn2, dot[u], __ldg(p.en2 + col[u])), col[u]);
      }
#pragma unroll
      for (int i = 0; i < K; ++i) top.id[i] = min(top.id[i], p.M - 1);
      if (rescan) {                                  // rescan_kernel produces every output of this row from z in HBM
        atomicAdd(&p.stats[0], 1);
        p.rescan_list[atomicAdd(&p.stats[2], 1)] = (int)n;
      } else {
        if (K == 2) {
          *reinterpret_cast<longlong2*>(p.idx + n * 2) = make_longlong2((long long)top.id[0], (long long)top.id[1]);
        } else {
#pragma unroll
          for (int i = 0; i < K; ++i) p.idx[n * K + i] = (int64_t)top.id[i];
        }
      }
      // outputs of the row (arithmetic of team_emit_row, one thread instead of a 4-lane team), one pass over d: per 16
      // components one TMEM read of z, the K winning rows (L1-resident: they were candidates a moment ago), q1 + SSE from
      // the nearest, bf16 hi/lo planes of all K (A operand of the dec GEMM).  TMEM loads are warp-collective: re-scan rows
      // walk along with their stores masked off.
      {
        const float4* er[K];
#pragma unroll
        for (int j = 0; j < K; ++j) er[j] = reinterpret_cast<const float4*>(p.bank_t + (size_t)top.id[j] * MF_D);
        float sp[4] = {0.f, 0.f, 0.f, 0.f};        // team_emit_row: lane `part` owns the float4 chunks i = part (mod 4)
        float* q1r = p.q1 + n * MF_D;
        const bool planes = !rescan && p.read_planes != nullptr;
#pragma unroll
        for (int c16 = 0; c16 < 4; ++c16) {
          float4 ev[K][4];
#pragma unroll
          for (int j = 0; j < K; ++j)
#pragma unroll
            for (int g = 0; g < 4; ++g) ev[j][g] = __ldg(er[j] + c16 * 4 + g);
          uint32_t zc[16];
          ptx::tmem_ld_32x16(z_tmem + c16 * 16, zc);
          ptx::tmem_ld_wait();
          uint32_t o[16];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int d = c16 * 16 + g * 4;
            const float z0 = __uint_as_float(zc[g * 4 + 0]) + bias_s[d + 0], z1 = __uint_as_float(zc[g * 4 + 1]) + bias_s[d + 1];
            const float z2 = __uint_as_float(zc[g * 4 + 2]) + bias_s[d + 2], z3 = __uint_as_float(zc[g * 4 + 3]) + bias_s[d + 3];
            const float d0 = ev[0][g].x - z0, d1 = ev[0][g].y - z1, d2 = ev[0][g].z - z2, d3 = ev[0][g].w - z3;
            o[g * 4 + 0] = __float_as_uint(z0 + d0); o[g * 4 + 1] = __float_as_uint(z1 + d1);
            o[g * 4 + 2] = __float_as_uint(z2 + d2); o[g * 4 + 3] = __float_as_uint(z3 + d3);
            float s = sp[g];                       // chunk index i = c16 * 4 + g  ->  part = g
            s = fmaf(d0, d0, s); s = fmaf(d1, d1, s); s = fmaf(d2, d2, s); s = fmaf(d3, d3, s);
            sp[g] = s;
          }
          if (!rescan) {
            const uint32_t (&o8)[2][8] = *reinterpret_cast<const uint32_t (*)[2][8]>(o);
            ptx::stg_v8(q1r + c16 * 16, o8[0]);
            ptx::stg_v8(q1r + c16 * 16 + 8, o8[1]);
          }
          if (planes) {
#pragma unroll
            for (int j = 0; j < K; ++j) {
              uint32_t h[8], l[8];
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                ptx::split_pack_bf16x2(ev[j][g].x, ev[j][g].y, h[2 * g], l[2 * g]);
                ptx::split_pack_bf16x2(ev[j][g].z, ev[j][g].w, h[2 * g + 1], l[2 * g + 1]);
              }
              __nv_bfloat16* hp = p.read_planes + (n * K + j) * MF_D + c16 * 16;
              ptx::stg_v8(hp, h);
              ptx::stg_v8(hp + p.read_plane_stride, l);
            }
          }
        }
        if (!rescan) p.sse_px[n] = (sp[0] + sp[1]) + (sp[2] + sp[3]);
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tmem_empty[acc]);           // z accumulator free (P2 re-read it)
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

int make_map_generic(CUtensorMap* m, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                     const uint64_t* strides, const uint32_t* box, int swizzle128);   // amft_conv.cu

bool mem_front_supported(int b, int HW, int C, int D, int M, int k) {
  return D == MF_D && C % MF_BK == 0 && C >= MF_BK && HW % 128 == 0 && b > 0 && M >= 1 && M <= MF_MPAD && k >= 1 && k <= 4 &&
         k <= M;
}

// x [b][C][HW] fp32 (x_bf16 = 0) or bf16 (1); enc_wp [2][64][C] bf16 planes; bank_hi [256][64] bf16 (zero rows beyond M); en2pad [256]; others as
// FrontParams.  stats[0], stats[2] must be zero on entry.
int run_mem_front(const void* x, int x_bf16, const void* enc_wp, const float* enc_b, const void* bank_hi, const float* bank_t,
                  const float* en2, const float* en2pad, const float* emax, float* z, float* q1, int64_t* idx, float* sse_px,
                  __nv_bfloat16* read_planes, int* stats, int* rescan_list, unsigned* amax_bits, int b, int HW, int C, int M,
                  int k, cudaStream_t st) {
  CUtensorMap tmX, tmW, tmBank;
  {
    uint64_t dims[3] = {(uint64_t)HW, (uint64_t)C, (uint64_t)b};
    const uint64_t eb = x_bf16 ? 2 : 4;
    uint64_t strides[2] = {(uint64_t)HW * eb, (uint64_t)C * HW * eb};
    uint32_t box[3] = {128, 32, 1};
    if (int rc = make_map_generic(&tmX, x, (int)eb, 3, dims, strides, box, 0)) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)C, (uint64_t)MF_D, 2};
    uint64_t strides[2] = {(uint64_t)C * 2, (uint64_t)MF_D * C * 2};
    uint32_t box[3] = {(uint32_t)MF_BK, (uint32_t)MF_D, 1};
    if (int rc = make_map_generic(&tmW, enc_wp, 2, 3, dims, strides, box, 1)) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)MF_D, (uint64_t)MF_MPAD};
    uint64_t strides[1] = {(uint64_t)MF_D * 2};
    uint32_t box[2] = {(uint32_t)MF_D, (uint32_t)MF_MPAD};
    if (int rc = make_map_generic(&tmBank, bank_hi, 2, 2, dims, strides, box, 1)) return rc;
  }
  FrontParams p;
  p.N = b * HW; p.HW = HW; p.C = C; p.M = M; p.tiles = b * (HW / 128);
  p.bias = enc_b; p.bank_t = bank_t; p.en2 = en2; p.en2pad = en2pad; p.emax = emax;
  p.z = z; p.q1 = q1; p.idx = idx; p.sse_px = sse_px;
  p.read_planes = read_planes; p.read_plane_stride = (long long)p.N * k * MF_D;
  p.stats = stats; p.rescan_list = rescan_list; p.amax_bits = amax_bits;
  if (amax_bits) AMMC_CUDA_CHECK(cudaMemsetAsync(amax_bits, 0, 4, st));
  static bool configured[64] = {false};
  int dev = 0;
  AMMC_CUDA_CHECK(cudaGetDevice(&dev));
  const int grid = min(num_sms(), p.tiles);
#define AMMC_MF_LAUNCH(KK, AM, XB) mem_front_kernel<KK, AM, XB><<<grid, MF_THREADS, MF_SMEM, st>>>(tmX, tmW, tmBank, p)
#define AMMC_MF_ATTR(KK, AM, XB)                                                                                          \
  AMMC_CUDA_CHECK(cudaFuncSetAttribute(mem_front_kernel<KK, AM, XB>, cudaFuncAttributeMaxDynamicSharedMemorySize, MF_SMEM))
  // the attribute must be set for every instantiation that may run on this device: set all of them once
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    AMMC_MF_ATTR(1, false, false); AMMC_MF_ATTR(1, true, false); AMMC_MF_ATTR(2, false, false); AMMC_MF_ATTR(2, true, false);
    AMMC_MF_ATTR(3, false, false); AMMC_MF_ATTR(3, true, false); AMMC_MF_ATTR(4, false, false); AMMC_MF_ATTR(4, true, false);
    AMMC_MF_ATTR(1, false, true); AMMC_MF_ATTR(1, true, true); AMMC_MF_ATTR(2, false, true); AMMC_MF_ATTR(2, true, true);
    AMMC_MF_ATTR(3, false, true); AMMC_MF_ATTR(3, true, true); AMMC_MF_ATTR(4, false, true); AMMC_MF_ATTR(4, true, true);
    configured[dev] = true;
  }
  const bool am = amax_bits != nullptr;
#define AMMC_MF_CASE(KK)                                                                        \
  case KK:                                                                                      \
    if (x_bf16) { if (am) AMMC_MF_LAUNCH(KK, true, true); else AMMC_MF_LAUNCH(KK, false, true); } \
    else { if (am) AMMC_MF_LAUNCH(KK, true, false); else AMMC_MF_LAUNCH(KK, false, false); }     \
    break;
  switch (k) {
    AMMC_MF_CASE(1) AMMC_MF_CASE(2) AMMC_MF_CASE(3) AMMC_MF_CASE(4)
    default: return fail(AMMC_EUNSUPPORTED, "fused memory front supports k <= 4");
  }
#undef AMMC_MF_CASE
#undef AMMC_MF_ATTR
#undef AMMC_MF_LAUNCH
  AMMC_LAUNCH_CHECK("mem_front_kernel");
  return 0;
}

AMMC_DEFINE_TIMEOUT_READER(timeout_reader_front)

}  // namespace ammc
