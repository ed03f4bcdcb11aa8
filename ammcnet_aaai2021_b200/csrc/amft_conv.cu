// AMFT (`bridge`, reference Code/models/unet.py:956-965): 3x3 convolution + folded BatchNorm + ReLU (+ residual)
// as an implicit GEMM on the 5th-generation tensor cores.
//
//   GEMM view      M = pixels (128 per tile: W_box x H_box x B_box), N = output channels (BLOCK_N per tile),
//                  K = 9 taps x Cin, walked as 64-channel blocks of one tap at a time.
//   A operand      NHWC bf16 activation planes.  The tile of tap (dy,dx) is the SAME 5-D TMA box shifted by
//                  (dx,dy): out-of-bounds coordinates are zero-filled by TMA, which is exactly the conv padding,
//                  so no im2col buffer exists.  The box lands in shared memory as 128 rows x 128 B with the
//                  128-byte swizzle = the canonical K-major UMMA layout.
//   B operand      weights repacked once to [Cout][tap*Cin + cin] bf16 (K-major), TMA box 64 x BLOCK_N.
//   accumulate     fp32 in TMEM (tcgen05.mma kind::f16, M=128, N=BLOCK_N, K=16), two accumulator buffers so the
//                  epilogue of tile i overlaps the MMAs of tile i+1.
//   precision      operands are split x = hi + lo (two bf16 planes).  precision=1 multiplies hi*hi only;
//                  precision=3 runs the K loop three times over (hi,hi), (hi,lo), (lo,hi) into the same
//                  accumulator -> ~2^-17 relative error, the fp32-parity mode.
//   epilogue       TMEM -> registers (tcgen05.ld 32x32b), y = relu(acc*scale[c] + shift[c]) and either
//                  (a) NHWC bf16 hi/lo planes (the next conv's A operand) or (b) + residual -> NCHW fp32.
//   roles          warp 0 lane 0: TMA producer; warp 1: MMA issuer (the warp runs converged, one elected lane issues;
//                  it also owns TMEM alloc/dealloc); warps 2-5: epilogue (one TMEM lane quarter each, bf16 planes stored
//                  with 256-bit st.global.v8).  Persistent over tiles, one CTA per SM.
//   general form   ammc_conv_layer (include/ammc_b200.h): channel windows on both sides, rows wider than 128 pixels,
//                  transposed 2x2 conv as a scattering epilogue, zero-padded channel counts, tanh; wide shallow layers are
//                  routed to halo_conv.cu.
#include "common.cuh"
#include "ptx.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace ammc {

constexpr int CONV_THREADS = 192;
constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                       // bf16 elements = 128 bytes = one swizzle row
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;    // 16 KB

struct ConvParams {
  int b, H, W, Cin, Cout;
  int W_box, H_box, B_box, groups_h, groups_w;
  int rows_valid;                // pixels actually covered by one TMA box (<= 128; < 128 when w does not divide 128)
  int tiles_m, tiles_n, ntaps;
  int n_pass;               // 1: hi*hi only; 3: (hi,hi) + (hi,lo) + (lo,hi)
  int act;                  // 0: none, 1: ReLU, 2: tanh (single-CTA kernel only)
  const float* scale;
  const float* shift;
  __nv_bfloat16* out_planes;     // [2][b][Ho][Wo][out_cs] (channels out_c_off .. of a possibly wider buffer) or null
  long long out_plane_stride;    // elements between the hi and lo plane
  int out_cs, out_c_off;
  int up2x, up_cout;             // transposed 2x2/stride-2 conv: column n = (dy*2+dx)*up_cout + co lands on pixel (2h+dy, 2w+dx)
  float* out_nchw;               // [b][cout_valid][H][W] or null
  const float* res_nchw;         // same shape or null
  int cout_valid;                // channels of out_nchw / res_nchw (< Cout when the weights were zero-padded)
  // mixed fp16 + e4m3 operand format "q" (pair kernel only; ptx.cuh split_pack_q)
  int out_fmt;                   // 0: out_planes are bf16 hi/lo planes; 1: q planes (out_q16 / out_q8)
  __half* out_q16;               // [b][H][W][Cout] fp16
  uint8_t* out_q8;               // [2][b][H][W][Cout] e4m3 (h8, l8)
  long long out_q8_stride;       // bytes between the h8 and l8 plane
  const float* out_qs;           // device scalar: power-of-two scale of the q output
  const float* in_qs;            // device scalars: scales of the q input and the q weights (PAIR_Q mode), else null
  const float* w_qs;
  int io_bf16;                   // pair kernel: out_nchw / res_nchw point to bf16 NCHW tensors (bf16 feature-I/O variant)
  int a_reuse;                   // PAIR_Q, 3x3: one (H_box+2)-row A tile per (dx, channel block) serves the three dy taps
  const float* out_l1;           // PAIR_Q with q output: per-output-channel sum_k |w| -- the kernel derives the output scale itself
  float* out_qs_store;           // ... and CTA 0 stores it for the consumer
};

// m-tile index -> first image / row / column of its TMA box
__device__ __forceinline__ void tile_origin(int mt, int groups_w, int groups_h, int W_box, int H_box, int B_box,
                                            int& img0, int& h0, int& w0) {
  const int wg = mt % groups_w, t2 = mt / groups_w;
  w0 = wg * W_box;
  h0 = (t2 % groups_h) * H_box;
  img0 = (t2 / groups_h) * B_box;
}

template <int BLOCK_N>
struct ConvSmem {
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;  // barriers + tmem slot + alignment slack
};

template <int BLOCK_N>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvParams p) {
  using S = ConvSmem<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + S::STAGES;
  uint64_t* tmem_full = empty_bar + S::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int kb_per_pass = p.ntaps * (p.Cin / BLOCK_K);
  const int kb_total = kb_per_pass * p.n_pass;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    for (int s = 0; s < S::STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
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
      // ------------------------------------------------------------------ TMA producer
      int s = 0; uint32_t ph = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int mt = t / p.tiles_n, nt = t % p.tiles_n;
        int img0, h0, w0;
        tile_origin(mt, p.groups_w, p.groups_h, p.W_box, p.H_box, p.B_box, img0, h0, w0);
        const int n0 = nt * BLOCK_N;
        for (int ps = 0; ps < p.n_pass; ++ps) {
          const int pa = ps == 2 ? 1 : 0, pb = ps == 1 ? 1 : 0;   // (hi,hi) (hi,lo) (lo,hi)
          for (int tap = 0; tap < p.ntaps; ++tap) {
            const int dy = p.ntaps == 9 ? tap / 3 - 1 : 0, dx = p.ntaps == 9 ? tap % 3 - 1 : 0;
            for (int cc = 0; cc < p.Cin / BLOCK_K; ++cc) {
              ptx::mbar_wait(&empty_bar[s], ph ^ 1, 1);
              ptx::mbar_expect_tx(&full_bar[s], p.rows_valid * 128 + S::B_BYTES);
              uint8_t* a_dst = smem + s * S::STAGE_BYTES;
              ptx::tma_load_5d(a_dst, &tmA, &full_bar[s], cc * BLOCK_K, w0 + dx, h0 + dy, img0, pa);
              ptx::tma_load_3d(a_dst + A_BYTES, &tmB, &full_bar[s], tap * p.Cin + cc * BLOCK_K, n0, pb);
              if (++s == S::STAGES) { s = 0; ph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      // ------------------------------------------------------------------ MMA issuer: the whole warp runs the loop
      // converged on warp-uniform values and one elected lane issues (see ptx::mma_f16_ss_warp); with short MMAs
      // (N <= 128) a single-lane issue loop, not the tensor core, was the bound
      constexpr uint32_t idesc = ptx::umma_idesc(1, BLOCK_M, BLOCK_N);
      int s = 0; uint32_t ph = 0;
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_ph = (it >> 1) & 1;
        ptx::mbar_wait(&tmem_empty[acc], acc_ph ^ 1, 2);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < kb_total; ++kb) {
          ptx::mbar_wait(&full_bar[s], ph, 3);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem + s * S::STAGE_BYTES);
          const uint64_t adesc = ptx::umma_desc_k_sw128(a_addr);
          const uint64_t bdesc = ptx::umma_desc_k_sw128(a_addr + A_BYTES);
#pragma unroll
          for (int k4 = 0; k4 < BLOCK_K / 16; ++k4) {
            // advance 16 elements (32 B) along K inside the swizzled row: +2 in the (>>4) address field
            ptx::mma_f16_ss_warp(d_tmem, adesc + 2 * k4, bdesc + 2 * k4, idesc, (kb | k4) != 0 ? 1u : 0u);
          }
          ptx::mma_commit_warp(&empty_bar[s]);
          if (kb == kb_total - 1) ptx::mma_commit_warp(&tmem_full[acc]);
          if (++s == S::STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // -------------------------------------------------------------------- epilogue warps
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;            // accumulator row = pixel within the tile
    const int per_img = p.W_box * p.H_box;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      const int mt = t / p.tiles_n, nt = t % p.tiles_n;
      int img0, h0, w0;
      tile_origin(mt, p.groups_w, p.groups_h, p.W_box, p.H_box, p.B_box, img0, h0, w0);
      const int img = img0 + r / per_img;
      const int rem = r % per_img;
      const int hh = h0 + rem / p.W_box, ww = w0 + rem % p.W_box;
      // rows past the box hold stale smem, columns past the row are TMA zero fill: never stored
      const bool valid = img < p.b && hh < p.H && ww < p.W && r < p.rows_valid;
      const int n0 = nt * BLOCK_N;
      ptx::mbar_wait(&tmem_full[acc], acc_ph, 4);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (int c32 = 0; c32 < BLOCK_N / 32; ++c32) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(taddr + c32 * 32, v);
        ptx::tmem_ld_wait();
        const int cbase = n0 + c32 * 32;
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
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] = tanhf(y[j]);
        } else if (p.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] = fmaxf(y[j], 0.f);
        }
        if (valid) {
          const size_t hw = (size_t)p.H * p.W;
          const size_t o0 = ((size_t)img * p.cout_valid + cbase) * hw + (size_t)hh * p.W + ww;
          const int nv = p.cout_valid - cbase;       // channels of this chunk that exist in the NCHW tensors
          if (p.res_nchw) {                      // residual first: both output forms carry it
            size_t o = o0;
#pragma unroll
            for (int j = 0; j < 32; ++j) { if (j < nv) y[j] += __ldg(p.res_nchw + o); o += hw; }
          }
          if (p.out_nchw) {
            size_t o = o0;
#pragma unroll
            for (int j = 0; j < 32; ++j) { if (j < nv) p.out_nchw[o] = y[j]; o += hw; }
          }
          if (p.out_planes) {
            size_t pix;
            int oc = cbase;
            if (p.up2x) {
              const int q4 = cbase / p.up_cout;
              oc = cbase - q4 * p.up_cout;
              pix = ((size_t)img * (2 * p.H) + 2 * hh + (q4 >> 1)) * (2 * p.W) + 2 * ww + (q4 & 1);
            } else {
              pix = ((size_t)img * p.H + hh) * p.W + ww;
            }
            __nv_bfloat16* hi = p.out_planes + pix * p.out_cs + p.out_c_off + oc;
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

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): one 256-pixel x 256-channel tile per pair of SMs.  Each CTA stages its own 128 pixel
// rows of A and its own 128-channel half of B (32 KB per stage instead of 48 KB -> 6 stages, one third less L2->smem
// traffic); the leader issues M=256 MMAs that read both CTAs' shared memory and writes rows 0-127 / 128-255 of the
// accumulator into the leader's / the peer's TMEM.  Barrier protocol: both CTAs' TMA bytes are credited to the
// leader's `full` barrier; `empty` and `tmem_full` are released in both CTAs by a multicast tcgen05.commit; the
// peer's epilogue threads arrive remotely on the leader's `tmem_empty`.
// ------------------------------------------------------------------------------------------------
constexpr int PAIR_N = 256;
constexpr int PAIR_B_BYTES = (PAIR_N / 2) * BLOCK_K * 2;      // this CTA's half of the B tile: 16 KB
constexpr int PAIR_TILE_BYTES = A_BYTES + PAIR_B_BYTES;        // one (A, B-half) pair of one plane: 32 KB
constexpr int PAIR_RING_BYTES = 6 * PAIR_TILE_BYTES;           // 192 KB: 6 stages of 32 KB, or 3 fused stages of 64 KB
constexpr int PAIR_BAR_OFFSET = PAIR_RING_BYTES;
constexpr int PAIR_SMEM = PAIR_BAR_OFFSET + 256 + 1024;
// PAIR_FUSED3 (precision 3): a stage holds the hi AND lo planes of both operands, loaded once and used by the three MMA
// groups hi*hi, hi*lo, lo*hi -> one third less L2->smem traffic than streaming the K loop three times.
// PAIR_Q (precision 2): mixed fp16 + e4m3 operands, two pass-equivalents instead of three.  The K loop runs twice per tile
// over 128-channel blocks of 64 KB stages: phase 0 streams the e4m3 planes [A_h8 | A_l8 | B_h8 | B_l8] and issues the
// cross terms h8.l8 + l8.h8 as kind::f8f6f4 MMAs (K = 32 each: twice the MACs per tensor-core cycle), phase 1 streams
// the fp16 planes [A_0 | A_1 | B_0 | B_1] (two 64-channel halves) and issues h16.h16; the first fp16 MMA rescales the
// cross-term sum by 2^-4 through scale-input-d, so one TMEM accumulator holds the whole product.  Each plane is read
// from L2 exactly once.
enum { PAIR_STREAM = 0, PAIR_FUSED3 = 1, PAIR_Q = 2 };
// PAIR_Q shared memory: an A ring (2 stages of two 24 KB planes: up to 192 pixel rows of 128 bytes) and a B ring (3 stages
// of two 16 KB planes).  With a_reuse the A tile of a (dx, channel block) carries H_box + 2 image rows and the three dy
// taps are the same tile with the UMMA descriptor advanced by W_box rows -- A traffic from L2 drops from 9 to 3 tiles per
// channel block (the kernel is bound by L2 -> shared-memory bandwidth: 64 KB per 1024 tensor-core cycles and SM otherwise).
constexpr int Q_A_PLANE = 192 * 128;                 // 24 KB
constexpr int Q_A_STAGE = 2 * Q_A_PLANE;             // 48 KB
constexpr int Q_A_STAGES = 2;
constexpr int Q_B_STAGE = 2 * PAIR_B_BYTES;          // 32 KB
constexpr int Q_B_STAGES = 4;
constexpr int Q_B_RING = Q_A_STAGES * Q_A_STAGE;     // 96 KB
constexpr int Q_RING_BYTES = Q_B_RING + Q_B_STAGES * Q_B_STAGE;      // 224 KB
constexpr int Q_SMEM = Q_RING_BYTES + 256 + 1024;
static_assert(Q_SMEM <= 232448, "q rings exceed the 227 KB of shared memory a CTA can have");
template <int MODE> struct PairCfg {
  static constexpr int PLANES = MODE == PAIR_STREAM ? 1 : 2;
  static constexpr int STAGE_BYTES = PLANES * PAIR_TILE_BYTES;
  static constexpr int STAGES = PAIR_RING_BYTES / STAGE_BYTES;
};
constexpr int PAIR_EPI_WARPS = 8;
constexpr int PAIR_THREADS = 64 + 32 * PAIR_EPI_WARPS;    // warp 0: TMA, warp 1: MMA, warps 2-9: epilogue

// IO16: out_nchw / res_nchw are bf16 NCHW tensors (bf16 feature-I/O variant) -- a compile-time variant, so that the fp32
// kernels carry none of its branches
template <int MODE, bool IO16 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PAIR_THREADS, 1)
conv_igemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmA8, const __grid_constant__ CUtensorMap tmB8, const ConvParams p) {
  constexpr bool FUSED3 = MODE == PAIR_FUSED3;
  constexpr bool QMODE = MODE == PAIR_Q;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  using Cfg = PairCfg<MODE>;
  constexpr int PAIR_STAGES = Cfg::STAGES;
  constexpr int PAIR_STAGE_BYTES = Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (QMODE ? Q_RING_BYTES : PAIR_BAR_OFFSET));
  uint64_t* empty_bar = full_bar + PAIR_STAGES;
  // PAIR_Q: [a_full 2][a_empty 2][b_full 3][b_empty 3] instead of [full][empty]
  uint64_t* a_full = full_bar;
  uint64_t* a_empty = a_full + Q_A_STAGES;
  uint64_t* b_full = a_empty + Q_A_STAGES;
  uint64_t* b_empty = b_full + Q_B_STAGES;
  uint64_t* tmem_full = QMODE ? b_empty + Q_B_STAGES : empty_bar + PAIR_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();            // 0 = leader
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int pair_tiles_m = (p.tiles_m + 1) >> 1;
  const int num_tiles = pair_tiles_m * p.tiles_n;
  const int kb_per_pass = p.ntaps * (p.Cin / BLOCK_K);
  const int kb_total = FUSED3 ? kb_per_pass : kb_per_pass * p.n_pass;      // PAIR_Q walks its own loops

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    if (QMODE) { ptx::prefetch_tensormap(&tmA8); ptx::prefetch_tensormap(&tmB8); }
    if (QMODE) {
      for (int s = 0; s < 2 * Q_A_STAGES + 2 * Q_B_STAGES; ++s) ptx::mbar_init(&a_full[s], 1);
    } else {
      for (int s = 0; s < PAIR_STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tmem_full[a], 1); ptx::mbar_init(&tmem_empty[a], 2 * 32 * PAIR_EPI_WARPS); }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_2sm(tmem_slot, 2 * PAIR_N);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------------ TMA producer (both CTAs)
      int s = 0; uint32_t ph = 0;
      int sa = 0, sb = 0; uint32_t pha = 0, phb = 0;       // PAIR_Q rings
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        const int mt = (t / p.tiles_n) * 2 + (int)rank, nt = t % p.tiles_n;
        int img0, h0, w0;
        tile_origin(mt, p.groups_w, p.groups_h, p.W_box, p.H_box, p.B_box, img0, h0, w0);
        const int n0 = nt * PAIR_N + (int)rank * (PAIR_N / 2);
        if (QMODE) {
          // A groups: with a_reuse one tile per (dx, channel block) feeding the three dy taps, else one tile per tap
          const int ngroups = p.a_reuse ? 3 : p.ntaps, nsub = p.a_reuse ? 3 : 1;
          const int rows_a = p.a_reuse ? p.W_box * (p.H_box + 2) * p.B_box : p.rows_valid;
          for (int phase = 0; phase < 2; ++phase) {
            for (int g = 0; g < ngroups; ++g) {
              const int dx = p.a_reuse ? g - 1 : (p.ntaps == 9 ? g % 3 - 1 : 0);
              const int dy0 = p.a_reuse ? -1 : (p.ntaps == 9 ? g / 3 - 1 : 0);
              for (int cc = 0; cc < p.Cin / 128; ++cc) {
                const int c0 = cc * 128;
                ptx::mbar_wait(&a_empty[sa], pha ^ 1, 41);
                if (rank == 0) ptx::mbar_expect_tx(&a_full[sa], 2 * 2 * rows_a * 128);
                uint8_t* a_dst = smem + sa * Q_A_STAGE;
                if (phase == 0) {                       // [A_h8 | A_l8]: rows of 128 e4m3
                  ptx::tma_load_5d_2sm(a_dst, &tmA8, &a_full[sa], c0, w0 + dx, h0 + dy0, img0, 0);
                  ptx::tma_load_5d_2sm(a_dst + Q_A_PLANE, &tmA8, &a_full[sa], c0, w0 + dx, h0 + dy0, img0, 1);
                } else {                                // [A_0 | A_1]: two 64-channel fp16 halves
                  ptx::tma_load_5d_2sm(a_dst, &tmA, &a_full[sa], c0, w0 + dx, h0 + dy0, img0, 0);
                  ptx::tma_load_5d_2sm(a_dst + Q_A_PLANE, &tmA, &a_full[sa], c0 + 64, w0 + dx, h0 + dy0, img0, 0);
                }
                if (++sa == Q_A_STAGES) { sa = 0; pha ^= 1; }
                for (int sub = 0; sub < nsub; ++sub) {
                  const int tap = p.a_reuse ? sub * 3 + g : g;
                  const int k0 = tap * p.Cin + c0;
                  ptx::mbar_wait(&b_empty[sb], phb ^ 1, 45);
                  if (rank == 0) ptx::mbar_expect_tx(&b_full[sb], 2 * Q_B_STAGE);
                  uint8_t* b_dst = smem + Q_B_RING + sb * Q_B_STAGE;
                  if (phase == 0) {                     // [B_h8 | B_l8]
                    ptx::tma_load_3d_2sm(b_dst, &tmB8, &b_full[sb], k0, n0, 0);
                    ptx::tma_load_3d_2sm(b_dst + PAIR_B_BYTES, &tmB8, &b_full[sb], k0, n0, 1);
                  } else {                              // [B_0 | B_1]
                    ptx::tma_load_3d_2sm(b_dst, &tmB, &b_full[sb], k0, n0, 0);
                    ptx::tma_load_3d_2sm(b_dst + PAIR_B_BYTES, &tmB, &b_full[sb], k0 + 64, n0, 0);
                  }
                  if (++sb == Q_B_STAGES) { sb = 0; phb ^= 1; }
                }
              }
            }
          }
          continue;
        }
        for (int ps = 0; ps < (FUSED3 ? 1 : p.n_pass); ++ps) {
          const int pa = ps == 2 ? 1 : 0, pb = ps == 1 ? 1 : 0;   // (hi,hi) (hi,lo) (lo,hi)
          for (int tap = 0; tap < p.ntaps; ++tap) {
            const int dy = p.ntaps == 9 ? tap / 3 - 1 : 0, dx = p.ntaps == 9 ? tap % 3 - 1 : 0;
            for (int cc = 0; cc < p.Cin / BLOCK_K; ++cc) {
              ptx::mbar_wait(&empty_bar[s], ph ^ 1, 41);
              if (rank == 0) ptx::mbar_expect_tx(&full_bar[s], 2 * Cfg::PLANES * (p.rows_valid * 128 + PAIR_B_BYTES));
              uint8_t* a_dst = smem + s * PAIR_STAGE_BYTES;
              if (FUSED3) {
#pragma unroll
                for (int pl = 0; pl < 2; ++pl) {       // [A_hi | B_hi | A_lo | B_lo]
                  ptx::tma_load_5d_2sm(a_dst + pl * PAIR_TILE_BYTES, &tmA, &full_bar[s], cc * BLOCK_K, w0 + dx, h0 + dy, img0, pl);
                  ptx::tma_load_3d_2sm(a_dst + pl * PAIR_TILE_BYTES + A_BYTES, &tmB, &full_bar[s],
                                       tap * p.Cin + cc * BLOCK_K, n0, pl);
                }
              } else {
                ptx::tma_load_5d_2sm(a_dst, &tmA, &full_bar[s], cc * BLOCK_K, w0 + dx, h0 + dy, img0, pa);
                ptx::tma_load_3d_2sm(a_dst + A_BYTES, &tmB, &full_bar[s], tap * p.Cin + cc * BLOCK_K, n0, pb);
              }
              if (++s == PAIR_STAGES) { s = 0; ph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ------------------------------------------------------------------ MMA issuer (leader only)
      constexpr uint32_t idesc = ptx::umma_idesc(1, 2 * BLOCK_M, PAIR_N);
      int s = 0; uint32_t ph = 0;
      int sa = 0, sb = 0; uint32_t pha = 0, phb = 0;       // PAIR_Q rings
      int it = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_ph = (it >> 1) & 1;
        ptx::mbar_wait(&tmem_empty[acc], acc_ph ^ 1, 42);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * PAIR_N;
        if (QMODE) {
          constexpr uint32_t idesc_q = ptx::umma_idesc(0, 2 * BLOCK_M, PAIR_N);    // format 0 = F16 resp. E4M3
          const int ngroups = p.a_reuse ? 3 : p.ntaps, nsub = p.a_reuse ? 3 : 1;
          const int ncc = p.Cin / 128;
          for (int phase = 0; phase < 2; ++phase) {
            for (int gc = 0; gc < ngroups * ncc; ++gc) {
              ptx::mbar_wait(&a_full[sa], pha, 43);
              const uint32_t a_base = ptx::smem_u32(smem + sa * Q_A_STAGE);
              for (int sub = 0; sub < nsub; ++sub) {
                ptx::mbar_wait(&b_full[sb], phb, 46);
                ptx::tc_fence_after();
                const uint32_t a_addr = a_base + (p.a_reuse ? sub * p.W_box * 128 : 0);    // dy tap = W_box rows further
                const uint32_t b_addr = ptx::smem_u32(smem + Q_B_RING + sb * Q_B_STAGE);
                const uint64_t a0 = ptx::umma_desc_k_sw128(a_addr), a1 = ptx::umma_desc_k_sw128(a_addr + Q_A_PLANE);
                const uint64_t b0 = ptx::umma_desc_k_sw128(b_addr), b1 = ptx::umma_desc_k_sw128(b_addr + PAIR_B_BYTES);
                if (phase == 0) {                       // cross terms: h8.l8 + l8.h8, four K = 32 steps per 128-byte row
                  const uint32_t first = (gc | sub) == 0 ? 0u : 1u;
#pragma unroll
                  for (int k4 = 0; k4 < 4; ++k4) {
                    ptx::mma_f8_ss_2sm(d_tmem, a0 + 2 * k4, b1 + 2 * k4, idesc_q, k4 == 0 ? first : 1u);
                    ptx::mma_f8_ss_2sm(d_tmem, a1 + 2 * k4, b0 + 2 * k4, idesc_q, 1u);
                  }
                } else {                                // main product; its first MMA folds the cross terms in (x 2^-4)
                  if ((gc | sub) == 0) ptx::mma_f16_ss_2sm_scaled<ptx::Q_SHIFT>(d_tmem, a0, b0, idesc_q);
                  else ptx::mma_f16_ss_2sm(d_tmem, a0, b0, idesc_q, 1u);
#pragma unroll
                  for (int k4 = 1; k4 < 4; ++k4) ptx::mma_f16_ss_2sm(d_tmem, a0 + 2 * k4, b0 + 2 * k4, idesc_q, 1u);
#pragma unroll
                  for (int k4 = 0; k4 < 4; ++k4) ptx::mma_f16_ss_2sm(d_tmem, a1 + 2 * k4, b1 + 2 * k4, idesc_q, 1u);
                }
                ptx::mma_commit_2sm(&b_empty[sb], 3);
                if (++sb == Q_B_STAGES) { sb = 0; phb ^= 1; }
              }
              ptx::mma_commit_2sm(&a_empty[sa], 3);
              if (++sa == Q_A_STAGES) { sa = 0; pha ^= 1; }
            }
          }
          ptx::mma_commit_2sm(&tmem_full[acc], 3);
          continue;
        }
        for (int kb = 0; kb < kb_total; ++kb) {
          ptx::mbar_wait(&full_bar[s], ph, 43);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem + s * PAIR_STAGE_BYTES);
          const uint64_t adesc = ptx::umma_desc_k_sw128(a_addr);
          const uint64_t bdesc = ptx::umma_desc_k_sw128(a_addr + A_BYTES);
          if (FUSED3) {
            const uint64_t adesc_lo = ptx::umma_desc_k_sw128(a_addr + PAIR_TILE_BYTES);
            const uint64_t bdesc_lo = ptx::umma_desc_k_sw128(a_addr + PAIR_TILE_BYTES + A_BYTES);
#pragma unroll
            for (int k4 = 0; k4 < BLOCK_K / 16; ++k4) {
              ptx::mma_f16_ss_2sm(d_tmem, adesc + 2 * k4, bdesc + 2 * k4, idesc, (kb | k4) != 0 ? 1u : 0u);
              ptx::mma_f16_ss_2sm(d_tmem, adesc + 2 * k4, bdesc_lo + 2 * k4, idesc, 1u);
              ptx::mma_f16_ss_2sm(d_tmem, adesc_lo + 2 * k4, bdesc + 2 * k4, idesc, 1u);
            }
          } else {
#pragma unroll
            for (int k4 = 0; k4 < BLOCK_K / 16; ++k4)
              ptx::mma_f16_ss_2sm(d_tmem, adesc + 2 * k4, bdesc + 2 * k4, idesc, (kb | k4) != 0 ? 1u : 0u);
          }
          ptx::mma_commit_2sm(&empty_bar[s], 3);
          if (kb == kb_total - 1) ptx::mma_commit_2sm(&tmem_full[acc], 3);
          if (++s == PAIR_STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // -------------------------------------------------------------------- epilogue: 8 warps per CTA (own 128 rows);
    // two warps share a TMEM lane quarter and split the 256 columns -> twice the loads/stores in flight, which is what
    // bounds the kernel when K is short (the 1x1 `dec` GEMM: 2 K-blocks per tile, epilogue exposed)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int per_img = p.W_box * p.H_box;
    constexpr int HALF_N = PAIR_N / 2;
    const size_t hw = (size_t)p.H * p.W;
    // PAIR_Q writing q planes: the power-of-two scale of the output follows from a rigorous bound,
    //   |y_c| <= |scale_c| * sum_k |w_ck| * max|x| + |shift_c|,   max|x| <= 2^15 / s16(input),
    // evaluated here by the (otherwise idle) epilogue warps of every CTA while the first tile's MMAs run -- no extra launch
    float oq_self = 1.f;
    if (QMODE && p.out_l1 != nullptr) {
      float* red = reinterpret_cast<float*>(tmem_slot + 4);          // 8 floats of the barrier block's slack
      const float X = 32768.f / __ldg(p.in_qs);
      float m = 0.f;
      for (int c = threadIdx.x - 64; c < p.Cout; c += 32 * PAIR_EPI_WARPS)
        m = fmaxf(m, fabsf(__ldg(p.scale + c)) * __ldg(p.out_l1 + c) * X * 1.01f + fabsf(__ldg(p.shift + c)));
      m = warp_max(m);
      if (lane == 0) red[warp - 2] = m;
      asm volatile("bar.sync 1, %0;" ::"n"(32 * PAIR_EPI_WARPS) : "memory");
#pragma unroll
      for (int i = 0; i < PAIR_EPI_WARPS; ++i) m = fmaxf(m, red[i]);
      oq_self = q_scale_for_bound(m);
      if (blockIdx.x == 0 && threadIdx.x == 64) *p.out_qs_store = oq_self;
    }
    int it = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      const int mt = (t / p.tiles_n) * 2 + (int)rank, nt = t % p.tiles_n;
      int img0, h0, w0;
      tile_origin(mt, p.groups_w, p.groups_h, p.W_box, p.H_box, p.B_box, img0, h0, w0);
      const int img = img0 + r / per_img;
      const int rem = r % per_img;
      const int hh = h0 + rem / p.W_box, ww = w0 + rem % p.W_box;
      const bool valid = img < p.b && mt < p.tiles_m && hh < p.H && ww < p.W && r < p.rows_valid;
      const int n0 = nt * PAIR_N + half * HALF_N;
      // NHWC destination of this thread's pixel; a transposed conv scatters the (dy,dx) group of this 128-column half
      const int q4 = p.up2x ? n0 / p.up_cout : 0;
      const size_t opix = p.up2x ? ((size_t)img * (2 * p.H) + 2 * hh + (q4 >> 1)) * (2 * p.W) + 2 * ww + (q4 & 1)
                                 : ((size_t)img * p.H + hh) * p.W + ww;
      __nv_bfloat16* const oplane = p.out_planes + opix * p.out_cs + p.out_c_off - q4 * p.up_cout;
      // 32-byte aligned plane rows (every shipped layout): STG.256
      const bool planes32 = ((reinterpret_cast<uintptr_t>(oplane) | (uintptr_t)(p.out_plane_stride * 2)) & 31) == 0;
      const size_t obase = ((size_t)img * p.Cout + n0) * hw + (size_t)hh * p.W + ww;
      const bool has_res = valid && p.res_nchw != nullptr;
      // q operands carry power-of-two scales: the accumulator holds (x.w) * s16x * s16w
      const float acc_inv = QMODE ? 1.f / (__ldg(p.in_qs) * __ldg(p.w_qs)) : 1.f;
      const float oq = p.out_fmt == 1 ? ((QMODE && p.out_l1 != nullptr) ? oq_self : __ldg(p.out_qs)) : 1.f;
      __half* const oq16 = p.out_q16 + opix * p.out_cs + p.out_c_off;
      uint8_t* const oq8 = p.out_q8 + opix * p.out_cs + p.out_c_off;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * PAIR_N + half * HALF_N;
      // 16 columns at a time, software-pipelined two chunks deep: the residual loads of chunks c+1 and c+2 are in
      // flight while chunk c is converted and stored (the loads are the only latency on the epilogue's critical path)
      float rv0[16], rv1[16];
      constexpr bool io16 = IO16;
      const unsigned short* const res16 = reinterpret_cast<const unsigned short*>(p.res_nchw);
      auto load_res = [&](size_t i) -> float {
        return io16 ? __uint_as_float((uint32_t)__ldg(res16 + i) << 16) : __ldg(p.res_nchw + i);
      };
      if (has_res) {
#pragma unroll
        for (int j = 0; j < 16; ++j) rv0[j] = load_res(obase + (size_t)j * hw);
#pragma unroll
        for (int j = 0; j < 16; ++j) rv1[j] = load_res(obase + (size_t)(16 + j) * hw);
      }
      ptx::mbar_wait(&tmem_full[acc], acc_ph, 44);        // the first residual loads overlap the tail of the MMAs
      ptx::tc_fence_after();
#pragma unroll 1
      for (int c16 = 0; c16 < HALF_N / 16; c16 += 2) {
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          float (&rv)[16] = sub == 0 ? rv0 : rv1;
          const int cc = c16 + sub;
          const size_t oc = obase + (size_t)cc * 16 * hw;
          uint32_t v[16];
          ptx::tmem_ld_32x16(taddr + cc * 16, v);
          ptx::tmem_ld_wait();
          const int cbase = n0 + cc * 16;
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + cbase + 4 * g4));
            const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + cbase + 4 * g4));
            const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float a = fmaf(QMODE ? __uint_as_float(v[4 * g4 + j]) * acc_inv : __uint_as_float(v[4 * g4 + j]), scv[j], shv[j]);
              a = p.act ? fmaxf(a, 0.f) : a;
              v[4 * g4 + j] = __float_as_uint(has_res ? a + rv[4 * g4 + j] : a);
            }
          }
          if (has_res && cc + 2 < HALF_N / 16) {          // refill this buffer with the chunk two steps ahead
            const size_t on = obase + (size_t)(cc + 2) * 16 * hw;
#pragma unroll
            for (int j = 0; j < 16; ++j) rv[j] = load_res(on + (size_t)j * hw);
          }
          if (valid) {
            if (p.out_nchw) {
              size_t o = oc;
              if (io16) {                                       // the operand planes below keep the unrounded fp32 value
                __nv_bfloat16* const o16 = reinterpret_cast<__nv_bfloat16*>(p.out_nchw);
#pragma unroll
                for (int j = 0; j < 16; ++j) { o16[o] = __float2bfloat16_rn(__uint_as_float(v[j])); o += hw; }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) { p.out_nchw[o] = __uint_as_float(v[j]); o += hw; }
              }
            }
            if (p.out_fmt == 1) {
              // q planes: 16 channels = 32 B of fp16 + 16 B of h8 + 16 B of l8
              uint32_t h16[8], h8[8], l8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j)
                ptx::split_pack_q(__uint_as_float(v[2 * j]) * oq, __uint_as_float(v[2 * j + 1]) * oq, h16[j], h8[j], l8[j]);
              ptx::stg_v8(oq16 + cbase, h16);
              *reinterpret_cast<uint4*>(oq8 + cbase) =
                  make_uint4(h8[0] | (h8[1] << 16), h8[2] | (h8[3] << 16), h8[4] | (h8[5] << 16), h8[6] | (h8[7] << 16));
              *reinterpret_cast<uint4*>(oq8 + p.out_q8_stride + cbase) =
                  make_uint4(l8[0] | (l8[1] << 16), l8[2] | (l8[3] << 16), l8[4] | (l8[5] << 16), l8[6] | (l8[7] << 16));
            } else if (p.out_planes) {
              __nv_bfloat16* hi = oplane + cbase;
              __nv_bfloat16* lo = hi + p.out_plane_stride;
              uint32_t hp[8], lp[8];
#pragma unroll
              for (int j = 0; j < 8; ++j)
                ptx::split_pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]), hp[j], lp[j]);
              if (planes32) {                                  // 16 channels = one 32-byte sector per plane: one request
                ptx::stg_v8(hi, hp);
                ptx::stg_v8(lo, lp);
              } else {
#pragma unroll
                for (int g8 = 0; g8 < 2; ++g8) {
                  reinterpret_cast<uint4*>(hi)[g8] = make_uint4(hp[4 * g8], hp[4 * g8 + 1], hp[4 * g8 + 2], hp[4 * g8 + 3]);
                  reinterpret_cast<uint4*>(lo)[g8] = make_uint4(lp[4 * g8], lp[4 * g8 + 1], lp[4 * g8 + 2], lp[4 * g8 + 3]);
                }
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive_cluster(&tmem_empty[acc], 0);      // the leader's MMA thread waits for both CTAs' epilogues
    }
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm(tmem_base, 2 * PAIR_N);
  }
}

// ------------------------------------------------------------------------------------------------
// operand packing
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// x [b][C][HW] fp32  ->  xp [2][b][HW][C] bf16.  32x32 (c, p) tiles transposed through shared memory.
__global__ void __launch_bounds__(256) pack_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xp,
                                                         int C, int HW, long long plane_stride) {
  __shared__ float tile[32][33];
  const int img = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    int c = c0 + r, pp = p0 + tx;
    tile[r][tx] = (c < C && pp < HW) ? x[((size_t)img * C + c) * HW + pp] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    int pp = p0 + r, c = c0 + tx;
    if (pp < HW && c < C) {
      __nv_bfloat16 hi, lo;
      split_bf16(tile[tx][r], hi, lo);
      size_t o = ((size_t)img * HW + pp) * C + c;
      xp[o] = hi;
      xp[o + plane_stride] = lo;
    }
  }
}

// w [Cout][Cin][taps] fp32 -> wp [2][Cout][taps*Cin] bf16 with k = tap*Cin + cin   (taps = 9 for 3x3, 1 for 1x1)
__global__ void pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int Cout, int Cin,
                                    int taps) {
  const long long total = (long long)Cout * Cin * taps;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int cin = (int)(e % Cin);
  const int tap = (int)((e / Cin) % taps);
  const int co = (int)(e / ((long long)Cin * taps));
  __nv_bfloat16 hi, lo;
  split_bf16(w[((size_t)co * Cin + cin) * taps + tap], hi, lo);
  wp[e] = hi;
  wp[e + total] = lo;
}

__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var, float eps,
                               float* __restrict__ scale, float* __restrict__ shift, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = gamma[c] / sqrtf(var[c] + eps);
  scale[c] = s;
  shift[c] = beta[c] - mean[c] * s;
}

// ------------------------------------------------------------------------------------------------
// q-format (fp16 + e4m3) packers and scales.  Buffer layouts (n = element count):
//   activations  [n fp16][n e4m3 h8][n e4m3 l8][float s16 @ 4n]                     NHWC, 4n + 16 bytes
//   weights      [n fp16][n h8][n l8][Cout floats: sum_k |w_ck| @ 4n][float s16 @ 4n + 4 Cout]   4n + 4 Cout + 16 bytes
// ------------------------------------------------------------------------------------------------
// max |x| as the bit pattern of a non-negative float (monotone under unsigned compare); *out must be zeroed first
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, long long n, unsigned* __restrict__ out) {
  float m = 0.f;
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(x[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}
__global__ void qscale_from_absmax_kernel(const unsigned* __restrict__ amax_bits, float* __restrict__ qs) {
  if (threadIdx.x == 0) qs[0] = q_scale_for_bound(__uint_as_float(amax_bits[0]));
}
// weights: one block per output channel; w [Cout][Cin][taps] -> q planes with k = tap*Cin + cin, L1 row sums.
// dgrad = 1: the data-gradient weights of the same tensor, w'[ci][co][tap] = w[co][ci][taps - 1 - tap] (taps flipped, channel
// roles swapped): Cout / Cin are then the ROWS / COLUMNS of the packed matrix, i.e. the original Cin / Cout.
__global__ void __launch_bounds__(256) pack_weights_q_kernel(const float* __restrict__ w, uint8_t* __restrict__ wq, int Cout,
                                                             int Cin, int taps, const unsigned* __restrict__ amax_bits,
                                                             int dgrad = 0) {
  __shared__ float red[33];
  const int co = blockIdx.x;
  const long long K = (long long)taps * Cin, n = (long long)Cout * K;
  // weights sit with their maximum in (2^7, 2^8]: fp16 and e4m3 planes share one scale (no 2^-7 between them)
  const float s = q_scale_for_bound(__uint_as_float(amax_bits[0])) * 0.0078125f;
  __half* w16 = reinterpret_cast<__half*>(wq);
  uint8_t* h8 = wq + 2 * n;
  uint8_t* l8 = wq + 3 * n;
  float l1 = 0.f;
  for (long long e = threadIdx.x; e < K; e += blockDim.x) {
    const int cin = (int)(e % Cin), tap = (int)(e / Cin);
    const float v = dgrad ? w[((size_t)cin * Cout + co) * taps + (taps - 1 - tap)] : w[((size_t)co * Cin + cin) * taps + tap];
    l1 += fabsf(v);
    const float t = v * s;
    const __half hh = __float2half_rn(t);
    w16[co * K + e] = hh;
    h8[co * K + e] = (uint8_t)(ptx::pack_e4m3x2(t, 0.f) & 0xff);
    l8[co * K + e] = (uint8_t)(ptx::pack_e4m3x2((t - __half2float(hh)) * 2048.f, 0.f) & 0xff);
  }
  l1 = block_sum(l1, red);
  if (threadIdx.x == 0) {
    reinterpret_cast<float*>(wq + 4 * n)[co] = l1;
    if (co == 0) reinterpret_cast<float*>(wq + 4 * n + 4 * (long long)Cout)[0] = s;
  }
}
// activations: x [b][C][HW] fp32 -> q planes [b][HW][C] (+ optionally the bf16 hi/lo planes of the same tensor, which the
// training path needs beside them for the weight gradient).  64 (channel) x 32 (pixel) tiles through shared memory: 128-byte
// coalesced reads along the pixels, then one (pixel, 8-channel group) per thread with 16 / 8-byte stores.  C % 8 == 0.
__global__ void __launch_bounds__(256) pack_nhwc_q_kernel(const float* __restrict__ x, uint8_t* __restrict__ xq,
                                                          __nv_bfloat16* __restrict__ xp, int C, int HW, long long n,
                                                          const float* __restrict__ qs) {
  __shared__ float tile[64][33];
  const int img = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float s = qs[0];
#pragma unroll
  for (int r = ty; r < 64; r += 8) {
    const int c = c0 + r, pp = p0 + tx;
    tile[r][tx] = (c < C && pp < HW) ? __ldg(x + ((size_t)img * C + c) * HW + pp) : 0.f;
  }
  __syncthreads();
  const int px = threadIdx.x >> 3, cg = threadIdx.x & 7;         // pixel within the tile, 8-channel group
  const int pp = p0 + px, c = c0 + cg * 8;
  if (pp < HW && c < C) {
    const size_t o = ((size_t)img * HW + pp) * C + c;
    uint32_t h16[4], h8[4], l8[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      ptx::split_pack_q(tile[cg * 8 + 2 * j][px] * s, tile[cg * 8 + 2 * j + 1][px] * s, h16[j], h8[j], l8[j]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(xq) + o) = make_uint4(h16[0], h16[1], h16[2], h16[3]);
    *reinterpret_cast<uint2*>(xq + 2 * n + o) = make_uint2(h8[0] | (h8[1] << 16), h8[2] | (h8[3] << 16));
    *reinterpret_cast<uint2*>(xq + 3 * n + o) = make_uint2(l8[0] | (l8[1] << 16), l8[2] | (l8[3] << 16));
    if (xp) {
      uint32_t hp[4], lp[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) ptx::split_pack_bf16x2(tile[cg * 8 + 2 * j][px], tile[cg * 8 + 2 * j + 1][px], hp[j], lp[j]);
      *reinterpret_cast<uint4*>(xp + o) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
      *reinterpret_cast<uint4*>(xp + n + o) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
    }
  }
}

// same, scalar stores: channel counts that are not a multiple of 8
__global__ void __launch_bounds__(256) pack_nhwc_q_scalar_kernel(const float* __restrict__ x, uint8_t* __restrict__ xq, int C,
                                                                 int HW, long long n, const float* __restrict__ qs) {
  __shared__ float tile[32][33];
  const int img = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float s = qs[0];
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, pp = p0 + tx;
    tile[r][tx] = (c < C && pp < HW) ? x[((size_t)img * C + c) * HW + pp] : 0.f;
  }
  __syncthreads();
  __half* x16 = reinterpret_cast<__half*>(xq);
  for (int r = ty; r < 32; r += 8) {
    const int pp = p0 + r, c = c0 + tx;
    if (pp < HW && c < C) {
      const float t = tile[tx][r] * s;
      uint32_t h16, h8, l8;
      ptx::split_pack_q(t, 0.f, h16, h8, l8);
      const size_t o = ((size_t)img * HW + pp) * C + c;
      x16[o] = __ushort_as_half((unsigned short)(h16 & 0xffff));
      xq[2 * n + o] = (uint8_t)(h8 & 0xff);
      xq[3 * n + o] = (uint8_t)(l8 & 0xff);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host: tensor maps + launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

static int make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                    const cuuint32_t* box) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(AMMC_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(AMMC_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

// Any element size (1 = bytes / fp8, 2 = bf16 / fp16, 4 = fp32), with or without the 128B swizzle.
int make_map_generic(CUtensorMap* m, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                     const uint64_t* strides, const uint32_t* box, int swizzle128) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(AMMC_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t d[5], s[4];
  cuuint32_t bx[5], estr[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides[i];
  const CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : (elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  CUresult r = fn(m, dt, (cuuint32_t)rank,
                  const_cast<void*>(base), d, s, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(AMMC_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

int make_map_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                  const uint32_t* box) {
  cuuint64_t d[5], s[4];
  cuuint32_t bx[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides[i];
  return make_map(m, base, rank, d, s, bx);
}

// [outer][inner] bf16 row-major matrix, 128B-swizzled boxes of box_outer rows x box_inner (=64) elements
int make_map_2d_bf16(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint32_t box_inner,
                     uint32_t box_outer) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {inner * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  return make_map(m, base, 2, dims, strides, box);
}

template <int BLOCK_N>
static int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p, cudaStream_t st) {
  using S = ConvSmem<BLOCK_N>;
  static bool configured[64] = {false};
  int dev = 0;
  AMMC_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    AMMC_CUDA_CHECK(cudaFuncSetAttribute(conv_igemm_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         S::TOTAL));
    configured[dev] = true;
  }
  int grid = min(num_sms(), p.tiles_m * p.tiles_n);
  conv_igemm_kernel<BLOCK_N><<<grid, CONV_THREADS, S::TOTAL, st>>>(tmA, tmB, p);
  AMMC_LAUNCH_CHECK("conv_igemm_kernel");
  return 0;
}

static int g_conv_fused3 = 1;      // pair kernel, precision 3: load hi+lo planes once per K block (1) or stream K 3x (0)
static int g_conv_pair_mode = 1;   // 1: CTA-pair kernel when Cout % 256 == 0; 0: single-CTA kernel everywhere
static int g_conv_q_reuse = 1;     // precision 2: A tiles with dy halo rows shared by three taps (0: one A tile per tap)

bool conv_shape_supported(int b, int Cin, int Cout, int h, int w) {
  return b > 0 && h > 0 && w > 0 && Cin % 64 == 0 && Cout % 64 == 0;
}

// w [Cout][Cin] fp32 -> [2][Cout][Cin] bf16 planes (1x1 conv / plain GEMM weights)
int pack_weights_1x1(const float* w, void* wp, int Cout, int Cin, cudaStream_t st) {
  const long long total = (long long)Cout * Cin;
  pack_weights_kernel<<<ceil_div(total, 256), 256, 0, st>>>(w, (__nv_bfloat16*)wp, Cout, Cin, 1);
  AMMC_LAUNCH_CHECK("pack_weights_kernel");
  return 0;
}

bool halo_conv_applies(const ammc_conv_layer& L);
int halo_conv_run(const ammc_conv_layer& L, cudaStream_t st);

// The one launch path of the engine: every public conv entry point fills an ammc_conv_layer and lands here.
int conv_run(const ammc_conv_layer& L, cudaStream_t st) {
  const int b = L.b, h = L.h, w = L.w, Cin = L.Cin, Cout = L.Cout, ntaps = L.taps;
  AMMC_REQUIRE(L.in_planes && L.wp && L.scale && L.shift && (L.out_planes || L.out_nchw), "null pointer argument");
  AMMC_REQUIRE(b > 0 && h > 0 && w > 0, "bad shape b=%d h=%d w=%d", b, h, w);
  AMMC_REQUIRE(ntaps == 9 || ntaps == 1, "taps must be 9 (3x3) or 1 (1x1), got %d", ntaps);
  if (L.precision < 1 || L.precision > 3) return fail(AMMC_EINVAL, "precision must be 1, 2 or 3 (got %d)", L.precision);
  AMMC_REQUIRE(L.in_fmt == (L.precision == 2 ? 1 : 0), "precision 2 takes q-format operands (in_fmt = 1), 1 and 3 take "
               "bf16 hi/lo planes (in_fmt = 0); got precision %d, in_fmt %d", L.precision, L.in_fmt);
  AMMC_REQUIRE(L.out_fmt == 0 || L.out_fmt == 1, "out_fmt must be 0 (bf16 hi/lo planes) or 1 (q planes)");
  if (Cin % 64 != 0 || Cout % 64 != 0)
    return fail(AMMC_EUNSUPPORTED, "tcgen05 conv needs Cin and Cout multiples of 64 (got %d, %d)", Cin, Cout);
  const int in_cs = L.in_cs > 0 ? L.in_cs : Cin;
  const int cout_valid = L.cout_valid > 0 ? L.cout_valid : Cout;
  AMMC_REQUIRE(in_cs % 8 == 0 && L.in_c_off % 8 == 0 && L.in_c_off >= 0 && L.in_c_off + Cin <= in_cs,
               "input channel window [%d, %d) does not fit a %d-channel buffer (8-channel alignment)", L.in_c_off,
               L.in_c_off + Cin, in_cs);
  AMMC_REQUIRE(cout_valid <= Cout, "cout_valid %d > Cout %d", cout_valid, Cout);
  AMMC_REQUIRE(L.act >= 0 && L.act <= 2, "act must be 0 (none), 1 (ReLU) or 2 (tanh)");
  const int up_cout = L.up2x ? Cout / 4 : Cout;
  if (L.up2x) {
    AMMC_REQUIRE(ntaps == 1 && !L.out_nchw && !L.res_nchw && L.out_planes && cout_valid == Cout,
                 "a transposed 2x2 conv is a 1x1 GEMM with NHWC plane output only");
    AMMC_REQUIRE(Cout % 4 == 0 && up_cout % 32 == 0, "transposed conv needs Cout (=4*channels) with channels %% 32 == 0");
  }
  const int out_cs = L.out_cs > 0 ? L.out_cs : up_cout;
  AMMC_REQUIRE(!L.out_planes || (out_cs % 8 == 0 && L.out_c_off % 8 == 0 && L.out_c_off >= 0 &&
                                 L.out_c_off + (L.up2x ? up_cout : cout_valid) <= out_cs),
               "output channel window does not fit a %d-channel buffer (8-channel alignment)", out_cs);
  AMMC_REQUIRE(!L.out_planes || cout_valid == Cout, "NHWC plane output needs cout_valid == Cout");
  if (L.precision != 2 && L.out_fmt == 0 && !L.io_bf16 && halo_conv_applies(L)) return halo_conv_run(L, st);   // wide, shallow layers
  ConvParams p;
  p.b = b; p.H = h; p.W = w; p.Cin = Cin; p.Cout = Cout;
  // one TMA box = W_box x H_box x B_box pixels <= 128: whole rows, as many as fit; whole images when a full image
  // fits; rows wider than 128 pixels are cut into 128-pixel segments (the last one zero-filled past the row end)
  p.W_box = min(w, 128);
  p.groups_w = ceil_div(w, p.W_box);
  p.H_box = p.groups_w == 1 ? min(h, 128 / w) : 1;
  p.B_box = (p.groups_w == 1 && p.H_box == h) ? max(1, 128 / (w * h)) : 1;
  p.groups_h = ceil_div(h, p.H_box);
  p.rows_valid = p.W_box * p.H_box * p.B_box;
  p.tiles_m = ceil_div(b, p.B_box) * p.groups_h * p.groups_w;
  const int block_n = (Cout % 256 == 0) ? 256 : (Cout % 128 == 0 ? 128 : 64);
  p.tiles_n = Cout / block_n;
  p.ntaps = ntaps;
  p.n_pass = L.precision;
  p.act = L.act;
  p.scale = L.scale; p.shift = L.shift;
  p.out_planes = (__nv_bfloat16*)L.out_planes;
  p.out_cs = out_cs; p.out_c_off = L.out_c_off;
  p.up2x = L.up2x ? 1 : 0; p.up_cout = up_cout;
  p.out_plane_stride = (long long)b * h * w * (L.up2x ? 4 : 1) * out_cs;
  p.out_nchw = L.out_nchw;
  p.res_nchw = L.res_nchw;
  p.io_bf16 = L.io_bf16 ? 1 : 0;
  p.cout_valid = cout_valid;
  p.out_fmt = 0; p.out_q16 = nullptr; p.out_q8 = nullptr; p.out_q8_stride = 0; p.out_qs = nullptr;
  p.in_qs = nullptr; p.w_qs = nullptr; p.out_l1 = nullptr; p.out_qs_store = nullptr;
  const bool qmode = L.precision == 2;
  p.a_reuse = (qmode && ntaps == 9 && g_conv_q_reuse && p.B_box == 1 && p.groups_w == 1 && p.W_box % 8 == 0 &&
               p.W_box * (p.H_box + 2) <= 192) ? 1 : 0;
  const long long n_out = (long long)b * h * w * out_cs, n_in = (long long)b * h * w * in_cs, n_w = (long long)Cout * ntaps * Cin;
  if (L.out_fmt == 1) {
    AMMC_REQUIRE(L.out_planes && !L.up2x && out_cs % 16 == 0 && L.out_c_off % 16 == 0 && cout_valid == Cout,
                 "q-format output needs a plain conv writing 16-channel aligned windows");
    p.out_fmt = 1;
    p.out_planes = nullptr;
    p.out_q16 = (__half*)L.out_planes;
    p.out_q8 = (uint8_t*)L.out_planes + 2 * n_out;
    p.out_q8_stride = n_out;
    p.out_qs = (const float*)((const uint8_t*)L.out_planes + 4 * n_out);
  }
  if (qmode) {
    AMMC_REQUIRE(Cin % 128 == 0 && Cout % 256 == 0 && !L.up2x && L.act != 2 && cout_valid == Cout && in_cs == Cin,
                 "precision 2 (fp16 + e4m3) needs Cin %% 128 == 0, Cout %% 256 == 0, dense input planes, act 0/1 (got Cin=%d Cout=%d)",
                 Cin, Cout);
    p.in_qs = (const float*)((const uint8_t*)L.in_planes + 4 * n_in);
    p.w_qs = (const float*)((const uint8_t*)L.wp + 4 * n_w + 4 * (long long)Cout);
    if (L.out_fmt == 1) {       // the kernel's epilogue warps derive the output scale from the weights' L1 row sums
      p.out_l1 = (const float*)((const uint8_t*)L.wp + 4 * n_w);
      p.out_qs_store = const_cast<float*>(p.out_qs);
    }
  }

  CUtensorMap tmA, tmB, tmA8, tmB8;
  if (qmode) {
    // fp16 plane (one plane, 64-channel boxes) and the two e4m3 planes (128-channel boxes = 128-byte rows)
    uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)w, (uint64_t)h, (uint64_t)b, 1};
    uint64_t st16[4] = {(uint64_t)Cin * 2, (uint64_t)w * Cin * 2, (uint64_t)h * w * Cin * 2, (uint64_t)b * h * w * Cin * 2};
    const uint32_t hb = (uint32_t)(p.a_reuse ? p.H_box + 2 : p.H_box);       // a_reuse: the tile carries its dy halo rows
    uint32_t box16[5] = {64, (uint32_t)p.W_box, hb, (uint32_t)p.B_box, 1};
    if (int rc = make_map_generic(&tmA, L.in_planes, 2, 5, dims, st16, box16, 1)) return rc;
    dims[4] = 2;
    uint64_t st8[4] = {(uint64_t)Cin, (uint64_t)w * Cin, (uint64_t)h * w * Cin, (uint64_t)n_in};
    uint32_t box8[5] = {128, (uint32_t)p.W_box, hb, (uint32_t)p.B_box, 1};
    if (int rc = make_map_generic(&tmA8, (const uint8_t*)L.in_planes + 2 * n_in, 1, 5, dims, st8, box8, 1)) return rc;
    const uint64_t K = (uint64_t)ntaps * Cin;
    uint64_t wd[3] = {K, (uint64_t)Cout, 1}, ws16[2] = {K * 2, (uint64_t)Cout * K * 2};
    uint32_t wb16[3] = {64, (uint32_t)(PAIR_N / 2), 1};
    if (int rc = make_map_generic(&tmB, L.wp, 2, 3, wd, ws16, wb16, 1)) return rc;
    wd[2] = 2;
    uint64_t ws8[2] = {K, (uint64_t)Cout * K};
    uint32_t wb8[3] = {128, (uint32_t)(PAIR_N / 2), 1};
    if (int rc = make_map_generic(&tmB8, (const uint8_t*)L.wp + 2 * n_w, 1, 3, wd, ws8, wb8, 1)) return rc;
  }
  if (!qmode) {
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)b, 2};
    cuuint64_t strides[4] = {(cuuint64_t)in_cs * 2, (cuuint64_t)w * in_cs * 2, (cuuint64_t)h * w * in_cs * 2,
                             (cuuint64_t)b * h * w * in_cs * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)p.W_box, (cuuint32_t)p.H_box, (cuuint32_t)p.B_box, 1};
    if (int rc = make_map(&tmA, (const __nv_bfloat16*)L.in_planes + L.in_c_off, 5, dims, strides, box)) return rc;
  }
  // the pair epilogue maps each 128-column half to ONE (dy,dx) group of a transposed conv
  const bool pair = block_n == 256 && (g_conv_pair_mode || qmode || L.out_fmt == 1) && L.act != 2 && cout_valid == Cout &&
                    (!L.up2x || up_cout % 128 == 0);
  AMMC_REQUIRE(pair || (!qmode && L.out_fmt == 0), "q-format operands need the CTA-pair kernel (Cout %% 256 == 0)");
  if (L.io_bf16 && !pair)
    return fail(AMMC_EUNSUPPORTED, "bf16 NCHW output / residual needs the CTA-pair kernel (Cout %% 256 == 0, act 0/1)");
  if (!qmode) {
    // B boxes are per-CTA halves (128 channels) in the pair kernel
    const cuuint64_t K = (cuuint64_t)ntaps * Cin;
    cuuint64_t dims[3] = {K, (cuuint64_t)Cout, 2};
    cuuint64_t strides[2] = {K * 2, (cuuint64_t)Cout * K * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)(pair ? PAIR_N / 2 : block_n), 1};
    if (int rc = make_map(&tmB, L.wp, 3, dims, strides, box)) return rc;
  }
  if (pair) {
    static bool configured[64] = {false};
    int dev = 0;
    AMMC_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !configured[dev]) {
      AMMC_CUDA_CHECK(cudaFuncSetAttribute(conv_igemm_pair_kernel<PAIR_STREAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM));
      AMMC_CUDA_CHECK(cudaFuncSetAttribute(conv_igemm_pair_kernel<PAIR_FUSED3>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM));
      AMMC_CUDA_CHECK(cudaFuncSetAttribute(conv_igemm_pair_kernel<PAIR_Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, Q_SMEM));
      AMMC_CUDA_CHECK(cudaFuncSetAttribute(conv_igemm_pair_kernel<PAIR_STREAM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM));
      AMMC_CUDA_CHECK(cudaFuncSetAttribute(conv_igemm_pair_kernel<PAIR_FUSED3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM));
      AMMC_CUDA_CHECK(cudaFuncSetAttribute(conv_igemm_pair_kernel<PAIR_Q, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Q_SMEM));
      configured[dev] = true;
    }
    const int pair_tiles = ((p.tiles_m + 1) / 2) * p.tiles_n;
    const int clusters = min(num_sms() / 2, pair_tiles);
    const bool io16 = p.io_bf16 != 0;
    if (qmode) {
      if (io16) conv_igemm_pair_kernel<PAIR_Q, true><<<2 * clusters, PAIR_THREADS, Q_SMEM, st>>>(tmA, tmB, tmA8, tmB8, p);
      else conv_igemm_pair_kernel<PAIR_Q><<<2 * clusters, PAIR_THREADS, Q_SMEM, st>>>(tmA, tmB, tmA8, tmB8, p);
    } else if (L.precision == 3 && g_conv_fused3) {
      if (io16) conv_igemm_pair_kernel<PAIR_FUSED3, true><<<2 * clusters, PAIR_THREADS, PAIR_SMEM, st>>>(tmA, tmB, tmA, tmB, p);
      else conv_igemm_pair_kernel<PAIR_FUSED3><<<2 * clusters, PAIR_THREADS, PAIR_SMEM, st>>>(tmA, tmB, tmA, tmB, p);
    } else {
      if (io16) conv_igemm_pair_kernel<PAIR_STREAM, true><<<2 * clusters, PAIR_THREADS, PAIR_SMEM, st>>>(tmA, tmB, tmA, tmB, p);
      else conv_igemm_pair_kernel<PAIR_STREAM><<<2 * clusters, PAIR_THREADS, PAIR_SMEM, st>>>(tmA, tmB, tmA, tmB, p);
    }
    AMMC_LAUNCH_CHECK("conv_igemm_pair_kernel");
    return 0;
  }
  switch (block_n) {
    case 256: return launch_conv<256>(tmA, tmB, p, st);
    case 128: return launch_conv<128>(tmA, tmB, p, st);
    default: return launch_conv<64>(tmA, tmB, p, st);
  }
}

int absmax_f32(const float* x, long long n, unsigned* out_bits, cudaStream_t st) {
  AMMC_CUDA_CHECK(cudaMemsetAsync(out_bits, 0, 4, st));
  absmax_kernel<<<min(ceil_div(n, 1024), 1184), 256, 0, st>>>(x, n, out_bits);
  AMMC_LAUNCH_CHECK("absmax_kernel");
  return 0;
}

// Shared by the plain 3x3 (ntaps = 9) and 1x1 (ntaps = 1) entry points.
int conv_igemm(const void* xp, const void* wp, const float* scale, const float* shift, void* out_planes,
               float* out_nchw, const float* res_nchw, int b, int Cin, int Cout, int h, int w, int ntaps,
               int precision, int relu, cudaStream_t st) {
  ammc_conv_layer L = {};
  L.in_planes = xp; L.wp = wp; L.taps = ntaps; L.scale = scale; L.shift = shift; L.act = relu ? 1 : 0;
  L.out_planes = out_planes; L.out_nchw = out_nchw; L.res_nchw = res_nchw;
  L.b = b; L.h = h; L.w = w; L.Cin = Cin; L.Cout = Cout; L.precision = precision;
  L.in_fmt = precision == 2 ? 1 : 0;          // precision 2: q operands in, q planes out (the next conv's operand)
  L.out_fmt = (precision == 2 && out_planes) ? 1 : 0;
  return conv_run(L, st);
}

}  // namespace ammc

namespace ammc { AMMC_DEFINE_TIMEOUT_READER(timeout_reader_conv) }

using namespace ammc;

extern "C" int ammc_conv_layer_run(const ammc_conv_layer* layer, void* stream) {
  AMMC_REQUIRE(layer != nullptr, "null layer descriptor");
  return conv_run(*layer, (cudaStream_t)stream);
}

extern "C" size_t ammc_q_act_bytes(int64_t n) { return (size_t)(4 * n + 16); }
extern "C" size_t ammc_q_weight_bytes(int Cout, int K) { return (size_t)(4 * (int64_t)Cout * K + 4 * (int64_t)Cout + 16); }

extern "C" int ammc_pack_conv_weights_q(const float* w, void* wq, int Cout, int Cin, int taps, void* stream) {
  AMMC_REQUIRE(w && wq && Cout > 0 && Cin > 0 && (taps == 9 || taps == 1), "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)Cout * Cin * taps;
  unsigned* amax = reinterpret_cast<unsigned*>((uint8_t*)wq + 4 * n + 4 * (long long)Cout + 4);   // scratch word of the tail
  AMMC_CUDA_CHECK(cudaMemsetAsync(amax, 0, 4, st));
  absmax_kernel<<<min(ceil_div(n, 1024), 1184), 256, 0, st>>>(w, n, amax);
  AMMC_LAUNCH_CHECK("absmax_kernel");
  pack_weights_q_kernel<<<Cout, 256, 0, st>>>(w, (uint8_t*)wq, Cout, Cin, taps, amax);
  AMMC_LAUNCH_CHECK("pack_weights_q_kernel");
  return 0;
}

static int pack_nhwc_q_impl(const float* x, void* xq, void* xp, int b, int C, int h, int w, void* stream);

// forward AND data-gradient q weights of one conv weight tensor, one max|w| reduction for both
extern "C" int ammc_pack_conv_weights_q_pair(const float* w, void* wq, void* wq_dgrad, int Cout, int Cin, int taps,
                                             void* stream) {
  AMMC_REQUIRE(w && wq && wq_dgrad && Cout > 0 && Cin > 0 && (taps == 9 || taps == 1), "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)Cout * Cin * taps;
  unsigned* amax = reinterpret_cast<unsigned*>((uint8_t*)wq + 4 * n + 4 * (long long)Cout + 4);   // scratch word of the tail
  AMMC_CUDA_CHECK(cudaMemsetAsync(amax, 0, 4, st));
  absmax_kernel<<<min(ceil_div(n, 1024), 1184), 256, 0, st>>>(w, n, amax);
  AMMC_LAUNCH_CHECK("absmax_kernel");
  pack_weights_q_kernel<<<Cout, 256, 0, st>>>(w, (uint8_t*)wq, Cout, Cin, taps, amax, 0);
  AMMC_LAUNCH_CHECK("pack_weights_q_kernel");
  pack_weights_q_kernel<<<Cin, 256, 0, st>>>(w, (uint8_t*)wq_dgrad, Cin, Cout, taps, amax, 1);
  AMMC_LAUNCH_CHECK("pack_weights_q_kernel (dgrad)");
  return 0;
}

extern "C" int ammc_pack_nhwc_q(const float* x, void* xq, int b, int C, int h, int w, void* stream) {
  return pack_nhwc_q_impl(x, xq, nullptr, b, C, h, w, stream);
}

// the q buffer AND the bf16 hi/lo planes [2][b,h,w,C] of the same tensor in one pass over x (training: the forward conv
// takes the q operand, the weight gradient the bf16 planes); C % 8 == 0
extern "C" int ammc_pack_nhwc_q_planes(const float* x, void* xq, void* xp, int b, int C, int h, int w, void* stream) {
  AMMC_REQUIRE(xp != nullptr && C % 8 == 0, "bf16 planes output needs a buffer and C %% 8 == 0");
  return pack_nhwc_q_impl(x, xq, xp, b, C, h, w, stream);
}

static int pack_nhwc_q_impl(const float* x, void* xq, void* xp, int b, int C, int h, int w, void* stream) {
  AMMC_REQUIRE(x && xq && b > 0 && C > 0 && h > 0 && w > 0 && b <= 65535, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)b * C * h * w;
  float* qs = reinterpret_cast<float*>((uint8_t*)xq + 4 * n);
  unsigned* amax = reinterpret_cast<unsigned*>(qs + 1);
  AMMC_CUDA_CHECK(cudaMemsetAsync(amax, 0, 4, st));
  absmax_kernel<<<min(ceil_div(n, 1024), 1184), 256, 0, st>>>(x, n, amax);
  AMMC_LAUNCH_CHECK("absmax_kernel");
  qscale_from_absmax_kernel<<<1, 32, 0, st>>>(amax, qs);
  AMMC_LAUNCH_CHECK("qscale_from_absmax_kernel");
  if (C % 8 == 0)
    pack_nhwc_q_kernel<<<dim3(ceil_div(h * w, 32), ceil_div(C, 64), b), 256, 0, st>>>(x, (uint8_t*)xq, (__nv_bfloat16*)xp, C,
                                                                                      h * w, n, qs);
  else
    pack_nhwc_q_scalar_kernel<<<dim3(ceil_div(h * w, 32), ceil_div(C, 32), b), 256, 0, st>>>(x, (uint8_t*)xq, C, h * w, n, qs);
  AMMC_LAUNCH_CHECK("pack_nhwc_q_kernel");
  return 0;
}

extern "C" int ammc_set_conv_pair_mode(int on) {
  g_conv_pair_mode = on ? 1 : 0;       // bit 1 (value 2/3): keep the pair kernel but stream the K loop three times
  g_conv_fused3 = (on & 2) ? 0 : 1;
  g_conv_q_reuse = (on & 4) ? 0 : 1;   // bit 2 (value 5): precision 2 without the dy-halo A tiles (A/B measurement)
  return 0;
}

extern "C" int ammc_pack_conv_weights(const float* w, void* wp, int Cout, int Cin, void* stream) {
  AMMC_REQUIRE(w && wp && Cout > 0 && Cin > 0, "bad argument");
  long long total = (long long)Cout * Cin * 9;
  pack_weights_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)wp, Cout, Cin, 9);
  AMMC_LAUNCH_CHECK("pack_weights_kernel");
  return 0;
}

extern "C" int ammc_pack_conv_weights_1x1(const float* w, void* wp, int Cout, int Cin, void* stream) {
  AMMC_REQUIRE(w && wp && Cout > 0 && Cin > 0, "bad argument");
  long long total = (long long)Cout * Cin;
  pack_weights_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)wp, Cout, Cin, 1);
  AMMC_LAUNCH_CHECK("pack_weights_kernel");
  return 0;
}

extern "C" int ammc_conv1x1_bn_relu(const void* xp, const void* wp, const float* scale, const float* shift,
                                    void* out_planes, float* out_nchw, const float* res_nchw, int b, int Cin, int Cout,
                                    int h, int w, int precision, int relu, void* stream) {
  return conv_igemm(xp, wp, scale, shift, out_planes, out_nchw, res_nchw, b, Cin, Cout, h, w, 1, precision, relu,
                    (cudaStream_t)stream);
}

namespace ammc { int pack_nhwc64(const float* x, void* xp, int b, int C, int HW, cudaStream_t st); }   // mem_simt.cu

extern "C" int ammc_pack_nhwc(const float* x, void* xp, int b, int C, int h, int w, void* stream) {
  AMMC_REQUIRE(x && xp && b > 0 && C > 0 && h > 0 && w > 0, "bad argument");
  AMMC_REQUIRE(b <= 65535, "batch %d too large for one pack launch", b);
  const int HW = h * w;
  if (C % 8 == 0) return pack_nhwc64(x, xp, b, C, HW, (cudaStream_t)stream);   // 16-byte stores, 2.8x faster
  pack_nhwc_kernel<<<dim3(ceil_div(HW, 32), ceil_div(C, 32), b), 256, 0, (cudaStream_t)stream>>>(
      x, (__nv_bfloat16*)xp, C, HW, (long long)b * HW * C);
  AMMC_LAUNCH_CHECK("pack_nhwc_kernel");
  return 0;
}

extern "C" int ammc_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                            float* scale, float* shift, int C, void* stream) {
  AMMC_REQUIRE(gamma && beta && mean && var && scale && shift && C > 0, "bad argument");
  bn_fold_kernel<<<ceil_div(C, 256), 256, 0, (cudaStream_t)stream>>>(gamma, beta, mean, var, eps, scale, shift, C);
  AMMC_LAUNCH_CHECK("bn_fold_kernel");
  return 0;
}

extern "C" int ammc_conv3x3_bn_relu(const void* xp, const void* wp, const float* scale, const float* shift,
                                    void* out_planes, float* out_nchw, const float* res_nchw, int b, int Cin, int Cout,
                                    int h, int w, int precision, int relu, void* stream) {
  return conv_igemm(xp, wp, scale, shift, out_planes, out_nchw, res_nchw, b, Cin, Cout, h, w, 9, precision, relu,
                    (cudaStream_t)stream);
}
