// 3x3 convolution for the wide, shallow U-Net layers (Cout <= 128 at 128^2 / 256^2 pixels; reference
// Code/models/unet.py:23-59): the implicit GEMM of amft_conv.cu re-fetches every activation tile once per tap and per
// split-bf16 pass (27x) and is L2->smem bound there.  This kernel loads a halo tile once and forms the nine tap
// operands as shifted shared-memory descriptors over it.
#include "common.cuh"
#include "ptx.cuh"
#include <cuda_bf16.h>

namespace ammc {

int make_map_2d_bf16(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint32_t box_inner,
                     uint32_t box_outer);

#ifdef AMMC_DEBUG_PROBES
// ------------------------------------------------------------------------------------------------
// probe: does a K-major SWIZZLE_128B operand descriptor work when it starts at an arbitrary 128-byte row of a tile that
// TMA wrote 1024-byte aligned?  out[m][n] = sum_k A[row_off + m][k] * B[n][k],  m < 128, n < 64, k < 64.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) desc_probe_kernel(const __grid_constant__ CUtensorMap tmA,
                                                         const __grid_constant__ CUtensorMap tmB, float* out, int rows,
                                                         int row_off, int base_off) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int a_bytes = (rows * 128 + 1023) & ~1023;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + a_bytes + 8192);
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
    ptx::mbar_expect_tx(bar, rows * 128 + 8192);
    ptx::tma_load_2d(smem, &tmA, bar, 0, 0);
    ptx::tma_load_2d(smem + a_bytes, &tmB, bar, 0, 0);
    ptx::mbar_wait(bar, 0, 90);
    ptx::tc_fence_after();
    constexpr uint32_t idesc = ptx::umma_idesc(1, 128, 64);
    const uint32_t a_addr = ptx::smem_u32(smem) + row_off * 128;
    const uint64_t adesc = ptx::umma_desc_k_sw128(a_addr) | ((uint64_t)(base_off & 7) << 49);
    const uint64_t bdesc = ptx::umma_desc_k_sw128(ptx::smem_u32(smem + a_bytes));
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) ptx::mma_f16_ss(tmem, adesc + 2 * k4, bdesc + 2 * k4, idesc, k4 != 0 ? 1u : 0u);
    ptx::mma_commit(done);
  }
  ptx::mbar_wait(done, 0, 91);
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
#endif  // AMMC_DEBUG_PROBES


// ------------------------------------------------------------------------------------------------
// halo kernel
//   tile        T_H x T_W = 4 x 30 output pixels, computed as 4 rows of WP = 32 "padded-linear" pixels (M = 128; the two
//               extra columns per row are garbage and never stored).  The (T_H+2) x WP halo box of one 64-channel block
//               lands in shared memory as 192 rows of 128 B (TMA zero-fills out-of-image pixels = the conv padding);
//               output row m of tap (dy,dx) reads halo row m + (1+dy)*WP + (1+dx), so the tap's A operand is the SAME
//               tile with its descriptor start address advanced by that many 128-byte rows (the 128B swizzle is a
//               function of the absolute shared-memory address, so any row offset is legal -- tools/desc_probe.py).
//   A traffic   192 rows per 120 outputs per plane instead of 9 x 128 rows x 3 passes.
//   B operand   [BLOCK_N x 64] single-plane weight tiles of one tap.  Cin = 64 and Cout = 64: all 18 tiles (9 taps x
//               hi/lo) stay resident in shared memory for the whole persistent CTA; otherwise they stream through a ring.
//   schedule    per (tile, 64-channel block): A_hi stage -> taps x {B_hi, B_lo};  A_lo stage -> taps x {B_hi}.
//   roles       warp 0: A producer, warp 1: MMA issuer (+TMEM alloc), warp 2: B producer, warps 3-6: epilogue.
// ------------------------------------------------------------------------------------------------
constexpr int HALO_WP = 32, HALO_TH = 4, HALO_TW = 30;
constexpr int HALO_ROWS = (HALO_TH + 2) * HALO_WP;            // 192 rows written by TMA
constexpr int HALO_A_STAGE = 25600;                            // 194 rows are read (2 stale ones feed garbage columns)
constexpr int HALO_A_STAGES = 3;
constexpr int HALO_B_BYTES = 18 * 64 * 128;                    // 144 KB: 18 tiles at N = 64, 9 at N = 128
constexpr int HALO_BAR_OFFSET = HALO_A_STAGES * HALO_A_STAGE + HALO_B_BYTES;
constexpr int HALO_SMEM = HALO_BAR_OFFSET + 512 + 1024;
constexpr int HALO_THREADS = 7 * 32;

struct HaloParams {
  int b, H, W, Cin, Cout;
  int tiles_x, tiles_y, num_tiles;
  int n_pass, act, resident;
  const float* scale;
  const float* shift;
  __nv_bfloat16* out_planes;
  long long out_plane_stride;
  int out_cs, out_c_off;
  float* out_nchw;
  int cout_valid;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const HaloParams p) {
  constexpr int B_TILE = BLOCK_N * 128;
  constexpr int NB = HALO_B_BYTES / B_TILE;       // resident mode: one slot per (plane, tap)
  constexpr int NBS = NB / 3;                     // streaming mode: stages of three tiles
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_b = smem + HALO_A_STAGES * HALO_A_STAGE;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + HALO_BAR_OFFSET);
  uint64_t* a_empty = a_full + HALO_A_STAGES;
  uint64_t* b_full = a_empty + HALO_A_STAGES;
  uint64_t* b_empty = b_full + NB;
  uint64_t* tmem_full = b_empty + NB;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncc = p.Cin / 64;
  const int n_planes = p.n_pass == 3 ? 2 : 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    for (int s = 0; s < HALO_A_STAGES; ++s) { ptx::mbar_init(&a_full[s], 1); ptx::mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < NB; ++s) { ptx::mbar_init(&b_full[s], 1); ptx::mbar_init(&b_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tmem_full[a], 1); ptx::mbar_init(&tmem_empty[a], 128); }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 2 * BLOCK_N);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------------ A producer: one halo plane per stage
      int s = 0; uint32_t ph = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int tx = t % p.tiles_x, t2 = t / p.tiles_x;
        const int ty = t2 % p.tiles_y, img = t2 / p.tiles_y;
        for (int cc = 0; cc < ncc; ++cc)
          for (int pl = 0; pl < n_planes; ++pl) {
            ptx::mbar_wait(&a_empty[s], ph ^ 1, 61);
            ptx::mbar_expect_tx(&a_full[s], HALO_ROWS * 128);
            ptx::tma_load_5d(smem + s * HALO_A_STAGE, &tmA, &a_full[s], cc * 64, tx * HALO_TW - 1, ty * HALO_TH - 1, img, pl);
            if (++s == HALO_A_STAGES) { s = 0; ph ^= 1; }
          }
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {
      // ------------------------------------------------------------------ B producer
      if (p.resident) {
        for (int pl = 0; pl < n_planes; ++pl)
          for (int tap = 0; tap < 9; ++tap) {
            const int slot = pl * 9 + tap;
            ptx::mbar_expect_tx(&b_full[slot], B_TILE);
            ptx::tma_load_3d(smem_b + slot * B_TILE, &tmB, &b_full[slot], tap * p.Cin, 0, pl);
          }
      } else {
        // streamed weights: one stage = the three taps of one kernel row of one plane (3 tiles, one barrier round trip)
        int s = 0; uint32_t ph = 0;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x)
          for (int cc = 0; cc < ncc; ++cc)
            for (int pl = 0; pl < n_planes; ++pl)
              for (int dyg = 0; dyg < 3; ++dyg) {
                const int nb = (pl == 0 && n_planes == 2) ? 2 : 1;
                for (int bp = 0; bp < nb; ++bp) {
                  ptx::mbar_wait(&b_empty[s], ph ^ 1, 62);
                  ptx::mbar_expect_tx(&b_full[s], 3 * B_TILE);
#pragma unroll
                  for (int dx = 0; dx < 3; ++dx)
                    ptx::tma_load_3d(smem_b + (s * 3 + dx) * B_TILE, &tmB, &b_full[s], (dyg * 3 + dx) * p.Cin + cc * 64, 0, bp);
                  if (++s == NBS) { s = 0; ph ^= 1; }
                }
              }
      }
    }
  } else if (warp == 1) {
    {
      // ------------------------------------------------------------------ MMA issuer
      // The whole warp runs this loop converged (uniform values); one elected lane issues each MMA / commit.
      // One thread issues every MMA; at N = 64 an MMA occupies the tensor core for only ~32-48 cycles, so this loop
      // must stay at a handful of instructions per MMA: taps fully unrolled, descriptors advanced by constants.
      constexpr uint32_t idesc = ptx::umma_idesc(1, 128, BLOCK_N);
      int sa = 0; uint32_t pha = 0;
      int sb = 0; uint32_t phb = 0;
      int it = 0;
      const uint64_t bdesc_ring = ptx::umma_desc_k_sw128(ptx::smem_u32(smem_b));
      if (p.resident) {
        for (int slot = 0; slot < 9 * n_planes; ++slot) ptx::mbar_wait(&b_full[slot], 0, 65);   // loaded once
        ptx::tc_fence_after();
      }
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_ph = (it >> 1) & 1;
        ptx::mbar_wait(&tmem_empty[acc], acc_ph ^ 1, 63);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        uint32_t accum = 0;
        for (int cc = 0; cc < ncc; ++cc)
          for (int pl = 0; pl < n_planes; ++pl) {
            ptx::mbar_wait(&a_full[sa], pha, 64);
            ptx::tc_fence_after();
            const uint64_t adesc0 = ptx::umma_desc_k_sw128(ptx::smem_u32(smem + sa * HALO_A_STAGE));
            const bool two = (pl == 0 && n_planes == 2);        // the hi activation plane meets both weight planes
            if (p.resident) {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const uint64_t ad = adesc0 + ((((tap / 3) * HALO_WP + tap % 3) * 128) >> 4);
                const uint64_t bd = bdesc_ring + ((tap * B_TILE) >> 4);
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                  ptx::mma_f16_ss_warp(d_tmem, ad + 2 * k4, bd + 2 * k4, idesc, accum);
                  accum = 1;
                }
                if (two) {
#pragma unroll
                  for (int k4 = 0; k4 < 4; ++k4)
                    ptx::mma_f16_ss_warp(d_tmem, ad + 2 * k4, bd + ((9 * B_TILE) >> 4) + 2 * k4, idesc, 1u);
                }
              }
            } else {
              for (int dyg = 0; dyg < 3; ++dyg) {
                const uint64_t ad0 = adesc0 + (uint64_t)((dyg * HALO_WP * 128) >> 4);
                for (int bp = 0; bp < (two ? 2 : 1); ++bp) {
                  ptx::mbar_wait(&b_full[sb], phb, 66);
                  ptx::tc_fence_after();
                  const uint64_t bd0 = bdesc_ring + (uint64_t)((sb * 3 * B_TILE) >> 4);
#pragma unroll
                  for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                      ptx::mma_f16_ss_warp(d_tmem, ad0 + ((dx * 128) >> 4) + 2 * k4, bd0 + ((dx * B_TILE) >> 4) + 2 * k4, idesc,
                                           accum);
                      accum = 1;
                    }
                  }
                  ptx::mma_commit_warp(&b_empty[sb]);
                  if (++sb == NBS) { sb = 0; phb ^= 1; }
                }
              }
            }
            ptx::mma_commit_warp(&a_empty[sa]);
            if (++sa == HALO_A_STAGES) { sa = 0; pha ^= 1; }
          }
        ptx::mma_commit_warp(&tmem_full[acc]);
      }
    }
  } else if (warp >= 3) {
    // -------------------------------------------------------------------- epilogue warps (TMEM lane quarter = warp % 4)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int ly = r / HALO_WP, lx = r % HALO_WP;
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      const int tx = t % p.tiles_x, t2 = t / p.tiles_x;
      const int ty = t2 % p.tiles_y, img = t2 / p.tiles_y;
      const int hh = ty * HALO_TH + ly, ww = tx * HALO_TW + lx;
      const bool valid = lx < HALO_TW && hh < p.H && ww < p.W;
      ptx::mbar_wait(&tmem_full[acc], acc_ph, 67);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (int c32 = 0; c32 < BLOCK_N / 32; ++c32) {
        const int cbase = c32 * 32;
        if (!p.out_planes && cbase >= p.cout_valid) break;      // zero-padded output channels (outc): nothing to store
        uint32_t v[32];
        ptx::tmem_ld_32x32(taddr + c32 * 32, v);
        ptx::tmem_ld_wait();
        float y[32];
#pragma unroll
        for (int g4 = 0; g4 < 8; ++g4) {
          const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + cbase) + g4);
          const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + cbase) + g4);
          y[4 * g4 + 0] = fmaf(__uint_as_float(v[4 * g4 + 0]), sc.x, sh.x);
          y[4 * g4 + 1] = fmaf(__uint_as_float(v[4 * g4 + 1]), sc.y, sh.y);
          y[4 * g4 + 2] = fmaf(__uint_as_float(v[4 * g4 + 2]), sc.z, sh.z);
          y[4 * g4 + 3] = fmaf(__uint_as_float(v[4 * g4 + 3]), sc.w, sh.w);
        }
        if (p.act == 2) {                 // tanh is ~25 instructions per element: keep it off the common path
          const int nt = p.out_planes ? 32 : p.cout_valid - cbase;     // and off the zero-padded channels
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nt) y[j] = tanhf(y[j]);
        } else if (p.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] = fmaxf(y[j], 0.f);
        }
        if (valid) {
          if (p.out_nchw) {
            const size_t hw = (size_t)p.H * p.W;
            size_t o = ((size_t)img * p.cout_valid + cbase) * hw + (size_t)hh * p.W + ww;
            const int nv = p.cout_valid - cbase;
#pragma unroll
            for (int j = 0; j < 32; ++j) { if (j < nv) p.out_nchw[o] = y[j]; o += hw; }
          }
          if (p.out_planes) {
            const size_t pix = ((size_t)img * p.H + hh) * p.W + ww;
            __nv_bfloat16* hi = p.out_planes + pix * p.out_cs + p.out_c_off + cbase;
            __nv_bfloat16* lo = hi + p.out_plane_stride;
            uint32_t hp[16], lp[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              ptx::split_pack_bf16x2(y[2 * j], y[2 * j + 1], hp[j], lp[j]);
            }
            if (((reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 31) == 0) {
              // 32 channels = two 32-byte sectors per plane: STG.256, one request per sector
              const uint32_t (&h8)[2][8] = *reinterpret_cast<const uint32_t (*)[2][8]>(hp);
              const uint32_t (&l8)[2][8] = *reinterpret_cast<const uint32_t (*)[2][8]>(lp);
              ptx::stg_v8(hi, h8[0]);
              ptx::stg_v8(hi + 16, h8[1]);
              ptx::stg_v8(lo, l8[0]);
              ptx::stg_v8(lo + 16, l8[1]);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                reinterpret_cast<uint4*>(hi)[j] = make_uint4(hp[4 * j], hp[4 * j + 1], hp[4 * j + 2], hp[4 * j + 3]);
                reinterpret_cast<uint4*>(lo)[j] = make_uint4(lp[4 * j], lp[4 * j + 1], lp[4 * j + 2], lp[4 * j + 3]);
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tmem_empty[acc]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 2 * BLOCK_N);
  }
}

int make_map_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                  const uint32_t* box);

static int g_halo_mode = 1;      // 1: use the halo kernel where it applies; 0: never (A/B measurements)

// Shapes the halo kernel takes: 3x3, Cout 64 or 128 (one N tile), no residual / transposed scatter.
bool halo_conv_applies(const ammc_conv_layer& L) {
  if (!g_halo_mode || L.taps != 9 || L.up2x || L.res_nchw) return false;
  if (L.Cout != 64 && L.Cout != 128) return false;
  // below 64 columns the 30-pixel tile quantisation wastes more than the tap reuse saves
  return L.w >= 64 && L.h >= 4;
}

int halo_conv_run(const ammc_conv_layer& L, cudaStream_t st) {
  const int in_cs = L.in_cs > 0 ? L.in_cs : L.Cin;
  const int cout_valid = L.cout_valid > 0 ? L.cout_valid : L.Cout;
  HaloParams p;
  p.b = L.b; p.H = L.h; p.W = L.w; p.Cin = L.Cin; p.Cout = L.Cout;
  p.tiles_x = ceil_div(L.w, HALO_TW);
  p.tiles_y = ceil_div(L.h, HALO_TH);
  p.num_tiles = p.tiles_x * p.tiles_y * L.b;
  p.n_pass = L.precision;
  p.act = L.act;
  p.resident = (L.Cin == 64 && L.Cout == 64) ? 1 : 0;
  p.scale = L.scale; p.shift = L.shift;
  p.out_planes = (__nv_bfloat16*)L.out_planes;
  p.out_cs = L.out_cs > 0 ? L.out_cs : L.Cout;
  p.out_c_off = L.out_c_off;
  p.out_plane_stride = (long long)L.b * L.h * L.w * p.out_cs;
  p.out_nchw = L.out_nchw;
  p.cout_valid = cout_valid;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[5] = {(uint64_t)L.Cin, (uint64_t)L.w, (uint64_t)L.h, (uint64_t)L.b, 2};
    uint64_t strides[4] = {(uint64_t)in_cs * 2, (uint64_t)L.w * in_cs * 2, (uint64_t)L.h * L.w * in_cs * 2,
                           (uint64_t)L.b * L.h * L.w * in_cs * 2};
    uint32_t box[5] = {64, HALO_WP, HALO_TH + 2, 1, 1};
    if (int rc = make_map_bf16(&tmA, (const __nv_bfloat16*)L.in_planes + L.in_c_off, 5, dims, strides, box)) return rc;
  }
  {
    const uint64_t K = 9ull * L.Cin;
    uint64_t dims[3] = {K, (uint64_t)L.Cout, 2};
    uint64_t strides[2] = {K * 2, (uint64_t)L.Cout * K * 2};
    uint32_t box[3] = {64, (uint32_t)L.Cout, 1};
    if (int rc = make_map_bf16(&tmB, L.wp, 3, dims, strides, box)) return rc;
  }
  static bool configured[64] = {false};
  int dev = 0;
  AMMC_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    AMMC_CUDA_CHECK(cudaFuncSetAttribute(conv_halo_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, HALO_SMEM));
    AMMC_CUDA_CHECK(cudaFuncSetAttribute(conv_halo_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, HALO_SMEM));
    configured[dev] = true;
  }
  const int grid = min(num_sms(), p.num_tiles);
  if (L.Cout == 64) conv_halo_kernel<64><<<grid, HALO_THREADS, HALO_SMEM, st>>>(tmA, tmB, p);
  else conv_halo_kernel<128><<<grid, HALO_THREADS, HALO_SMEM, st>>>(tmA, tmB, p);
  AMMC_LAUNCH_CHECK("conv_halo_kernel");
  return 0;
}

AMMC_DEFINE_TIMEOUT_READER(timeout_reader_halo)

}  // namespace ammc

using namespace ammc;

#ifdef AMMC_DEBUG_PROBES
extern "C" int ammc_debug_desc_probe(const void* a, const void* b, float* out, int rows, int row_off, int base_off,
                                     void* stream) {
  AMMC_REQUIRE(a && b && out && rows >= 128 && rows <= 256 && row_off >= 0 && row_off + 128 <= rows, "bad argument");
  CUtensorMap tmA, tmB;
  if (int rc = make_map_2d_bf16(&tmA, a, 64, (uint64_t)rows, 64, (uint32_t)rows)) return rc;
  if (int rc = make_map_2d_bf16(&tmB, b, 64, 64, 64, 64)) return rc;
  const int smem = ((rows * 128 + 1023) & ~1023) + 8192 + 256 + 1024;
  AMMC_CUDA_CHECK(cudaFuncSetAttribute(desc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  desc_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(tmA, tmB, out, rows, row_off, base_off);
  AMMC_LAUNCH_CHECK("desc_probe_kernel");
  return 0;
}
#endif  // AMMC_DEBUG_PROBES

extern "C" int ammc_set_conv_halo_mode(int on) {
  g_halo_mode = on ? 1 : 0;
  return 0;
}
