// Memory addressing on the tensor cores (north-star kernel (a)): similarity contraction -> top-k, with the
// similarity matrix living only in TMEM.
//
// Exactness scheme ("filter + refine"): the reference ranks fp32 distances (Code/models/unet.py:283-293).  tcgen05 has
// no fp32 MMA, so the N x M x D contraction runs ONCE in fp16 as a *filter*: queries are scaled per row and the bank per
// tensor by powers of two (max component -> [2^14, 2^15), so neither overflow nor lost small rows can occur), rounded to
// fp16 (u = 2^-11) and multiplied on the tensor cores; the epilogue forms the approximate score
//   a~_j = ||e_j||^2 - 2 (z~.e~_j) / (s_n t)          (||z||^2 is rank-irrelevant)
// and keeps every item within a margin of the k-th smallest.  |a~_j - a_j| <= 2 (2u + u^2) ||z|| ||e_j|| (rounding of both
// operands, Cauchy-Schwarz), so with margin = 8 u (1 + 1%) ||z|| max||e|| (+ fp32 accumulation slack) the kept set is a
// superset of the exact top-k.  fp16 instead of bf16 shrinks that margin -- and with it the candidate lists -- eightfold:
// at D = 1024, M = 8192 the bf16 margin admitted ~11 items per column half (lists of 12 overflowed on 6 % of the rows,
// each paying an exact scan of the whole bank), the fp16 margin ~2-3.  Rows whose k best approximate scores (and the next
// survivor) lie more than the margin apart are ranked by the filter itself; for the others the tail recomputes the
// candidates' exact fp32 distances with the SAME arithmetic as the generic fp32 path (mem_simt.cu) and ranks them; rows whose
// list still overflowed (near-degenerate neighbourhoods) are re-scanned exactly over all M items, so the indices are
// bit-identical to the fp32 path for every input; the number of such rows is reported in the stats block.
//
//   addr_tc_kernel<BLOCK_N, KSEL, SWEEP>   persistent; per 128-query tile loops over item tiles; TMA (128B swizzle) -> smem
//                             -> tcgen05.mma (M128 x N BLOCK_N x K16, fp32 accum in TMEM, 2 accumulator buffers); 8 epilogue
//                             warps: tcgen05.ld -> a~ -> the KSEL best and every column within the margin (two passes per
//                             tile, or one sweep with a running threshold for long banks); at the end of a query tile the
//                             two column halves are merged and the row is DECIDED when the approximate ranking is proven
//   addr_tail_kernel<K>       D >= 128, warp per query: decided rows are a pure gather; the others get the exact fp32
//                             distances of their candidates (same fmaf chains as the fp32 kernel), then the same gather
//   refine_kernel<K>          D = 64, 4-lane team per query: exact distances of the candidates, top-k, gathers (read, q1),
//                             per-pixel SSE, EMA statistics
//   rescan_kernel<K>          exact scan over all items for rows whose candidate list overflowed (rare)
#include "common.cuh"
#include "ptx.cuh"
#include "topk.cuh"
#include "addr_tail.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <float.h>

namespace ammc {

constexpr int ADDR_EPI_WARPS = 8;
constexpr int ADDR_THREADS = 64 + 32 * ADDR_EPI_WARPS;   // warp 0: TMA, warp 1: MMA, warps 2-9: epilogue
constexpr int ADDR_CAPH = 12;                           // candidate slots per (query, column half)
constexpr int ADDR_CAND = 2 * ADDR_CAPH;                // candidate slots per query
constexpr int ADDR_BLOCK_K = 64;
constexpr int ADDR_A_BYTES = 128 * ADDR_BLOCK_K * 2;

struct AddrParams {
  int N, D, M, Mpad;
  int tiles_q, tiles_i;
  const float* en2pad;  // [Mpad], +inf beyond M
  const float2* zmeta;  // [N]  (||z_n||^2, 1 / s_n): fp32 norm and the inverse of the row's power-of-two fp16 scale
  const float* emax;    // [2]  max_j ||e_j||, 1 / t (inverse of the bank's power-of-two fp16 scale)
  int k;                // items wanted per query (<= KSEL of the instantiation)
  int* cand;            // [N][ADDR_CAND]  merged candidate list of the query (both column halves)
  int* cand_cnt;        // [N][2]  [0] entries used (> ADDR_CAND: a list overflowed -> exact re-scan), [1] 1 when the
                        //         filter alone decided the row: cand[0..k) are the final indices in rank order
};

template <int BLOCK_N>
struct AddrSmem {
  static constexpr int B_BYTES = BLOCK_N * ADDR_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = ADDR_A_BYTES + B_BYTES;
  static constexpr int STAGES = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr int LIST_OFFSET = STAGES * STAGE_BYTES;         // float [256 threads][CAPH] + uint16 [256][CAPH]
  static constexpr int XCH_OFFSET = LIST_OFFSET + 256 * ADDR_CAPH * 6;   // float [128 rows][4] + int [128]: half 1 -> half 0
  static constexpr int BAR_OFFSET = XCH_OFFSET + 128 * 4 * 4 + 128 * 4;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
};

// branch-free insertion of x into the ascending list m[0..KSEL)
template <int KSEL>
__device__ __forceinline__ void sel_insert(float (&m)[KSEL], float x) {
#pragma unroll
  for (int i = 0; i < KSEL; ++i) {
    const float lo = fminf(m[i], x);
    x = fmaxf(m[i], x);
    m[i] = lo;
  }
}

// v[j] for a run-time j without spilling the register array to local memory (5-level select tree)
__device__ __forceinline__ uint32_t select32(const uint32_t (&v)[32], int j) {
  uint32_t a[16], b[8], c[4], d[2];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (j & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = (j & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[2 * i + 1] : b[2 * i];
#pragma unroll
  for (int i = 0; i < 2; ++i) d[i] = (j & 8) ? c[2 * i + 1] : c[2 * i];
  return (j & 16) ? d[1] : d[0];
}

// SWEEP = true: the epilogue visits every accumulator column ONCE with a running threshold (bank of several item tiles);
// false: two branch-free passes per tile (one or two item tiles, where the threshold never gets the chance to tighten).
template <int BLOCK_N, int KSEL, bool SWEEP>
__global__ void __launch_bounds__(ADDR_THREADS, 1)
addr_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const AddrParams p) {
  using S = AddrSmem<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* list_val = reinterpret_cast<float*>(smem + S::LIST_OFFSET);
  uint16_t* list_idx = reinterpret_cast<uint16_t*>(smem + S::LIST_OFFSET + 256 * ADDR_CAPH * 4);
  float* xch_m = reinterpret_cast<float*>(smem + S::XCH_OFFSET);             // [128][4] best scores of column half 1
  int* xch_cnt = reinterpret_cast<int*>(smem + S::XCH_OFFSET + 128 * 4 * 4);  // [128] its list length, -1 = overflowed
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + S::STAGES;
  uint64_t* tmem_full = empty_bar + S::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb_per_tile = p.D / ADDR_BLOCK_K;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    for (int s = 0; s < S::STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tmem_full[a], 1); ptx::mbar_init(&tmem_empty[a], 32 * ADDR_EPI_WARPS); }
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
      int s = 0; uint32_t ph = 0;
      for (int tq = blockIdx.x; tq < p.tiles_q; tq += gridDim.x) {
        for (int ti = 0; ti < p.tiles_i; ++ti) {
          for (int kb = 0; kb < kb_per_tile; ++kb) {
            ptx::mbar_wait(&empty_bar[s], ph ^ 1, 11);
            ptx::mbar_expect_tx(&full_bar[s], S::STAGE_BYTES);
            uint8_t* a_dst = smem + s * S::STAGE_BYTES;
            ptx::tma_load_2d(a_dst, &tmA, &full_bar[s], kb * ADDR_BLOCK_K, tq * 128);
            ptx::tma_load_2d(a_dst + ADDR_A_BYTES, &tmB, &full_bar[s], kb * ADDR_BLOCK_K, ti * BLOCK_N);
            if (++s == S::STAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc(0, 128, BLOCK_N);      // fp16 operands
      int s = 0; uint32_t ph = 0;
      int it = 0;
      for (int tq = blockIdx.x; tq < p.tiles_q; tq += gridDim.x) {
        for (int ti = 0; ti < p.tiles_i; ++ti, ++it) {
          const int acc = it & 1;
          const uint32_t acc_ph = (it >> 1) & 1;
          ptx::mbar_wait(&tmem_empty[acc], acc_ph ^ 1, 12);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
          for (int kb = 0; kb < kb_per_tile; ++kb) {
            ptx::mbar_wait(&full_bar[s], ph, 13);
            ptx::tc_fence_after();
            const uint32_t a_addr = ptx::smem_u32(smem + s * S::STAGE_BYTES);
            const uint64_t adesc = ptx::umma_desc_k_sw128(a_addr);
            const uint64_t bdesc = ptx::umma_desc_k_sw128(a_addr + ADDR_A_BYTES);
#pragma unroll
            for (int k4 = 0; k4 < ADDR_BLOCK_K / 16; ++k4)
              ptx::mma_f16_ss(d_tmem, adesc + 2 * k4, bdesc + 2 * k4, idesc, (kb | k4) != 0 ? 1u : 0u);
            ptx::mma_commit(&empty_bar[s]);
            if (kb == kb_per_tile - 1) ptx::mma_commit(&tmem_full[acc]);
            if (++s == S::STAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue: 2 warps per TMEM lane quarter, each
    // owning half of the tile's columns.  Pass A keeps the KSEL smallest approximate scores (values only, branch-free);
    // pass B re-reads the tile and records every column whose score is within the safety margin of the KSEL-th best.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int HALF_N = BLOCK_N / 2;
    constexpr int CHUNKS = HALF_N / 32;
    float* lv = list_val + (size_t)((warp - 2) * 32 + lane) * ADDR_CAPH;
    uint16_t* li = list_idx + (size_t)((warp - 2) * 32 + lane) * ADDR_CAPH;
    const float emax = __ldg(p.emax), tinv = __ldg(p.emax + 1);
    int it = 0;
    for (int tq = blockIdx.x; tq < p.tiles_q; tq += gridDim.x) {
      const int n = tq * 128 + q * 32 + lane;
      const float2 zm = n < p.N ? __ldg(p.zmeta + n) : make_float2(0.f, 0.f);
      const float zn2 = zm.x;
      const float cs = -2.f * zm.y * tinv;           // a~ = ||e||^2 + cs * (z~.e~): undoes the two power-of-two scales
      // |a~ - a| <= 4u(1+u) ||z|| ||e||, u = 2^-11 (fp16 rounding of both operands, Cauchy-Schwarz); candidates within
      // twice that of the KSEL-th smallest approximate score are a superset of the exact top-KSEL (see file header).
      // The second term bounds, again worst case, what fp32 evaluation adds on both sides (u32 = 2^-24): the D-term
      // accumulations of the MMA and of the exact kernel's dot product (2 * 2 * D u32 ||z|| ||e||), of ||e||^2
      // (D u32 ||e||^2) and the roundings of the three-term distance (3 u32 (||z|| + ||e||)^2) -- doubled like the first.
      const float zn = sqrtf(zn2);
      const float margin = 8.f * 0.00048828125f * 1.01f * zn * emax +
                           1.1920929e-7f * (4.f * (float)p.D * zn * emax + (float)p.D * emax * emax +
                                            3.f * (zn + emax) * (zn + emax)) + 1e-30f;
      float m[KSEL];
#pragma unroll
      for (int i = 0; i < KSEL; ++i) m[i] = INFINITY;
      int cnt = 0;
      bool overflow = false;
      float thr = FLT_MAX;
      for (int ti = 0; ti < p.tiles_i; ++ti, ++it) {
        const int acc = it & 1;
        const uint32_t acc_ph = (it >> 1) & 1;
        ptx::mbar_wait(&tmem_full[acc], acc_ph, 14);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N + half * HALF_N;
        const int col0 = ti * BLOCK_N + half * HALF_N;
        // accumulator chunks (32 columns) and their item norms are fetched ONE CHUNK AHEAD into a second register buffer:
        // the TMEM load and the global loads of chunk c + 1 are in flight while chunk c is scanned
        uint32_t vbuf[2][32];
        float4 ebuf[2][8];
        const float4* e4base = reinterpret_cast<const float4*>(p.en2pad + col0);
        auto fetch = [&](int c, uint32_t (&vv)[32], float4 (&ee)[8]) {
          ptx::tmem_ld_32x32(taddr + c * 32, vv);
#pragma unroll
          for (int j = 0; j < 8; ++j) ee[j] = __ldg(e4base + c * 8 + j);
        };
        if (SWEEP) {
        // ---- one sweep over the tile with a RUNNING threshold thr = (KSEL-th best score so far) + margin.  thr only
        // falls, so every column within the row's FINAL threshold is a hit when it is visited: the list stays a superset
        // of the final candidates (pruned at the end).  After the first columns of a query tile almost nothing passes the
        // threshold, so a column costs one FMA and one compare; the insertions run only for the few hits.  With D <= 256
        // the two-pass epilogue (~12 instructions per column), not the MMAs, set the pace: (65 536, 8192, 256) went from
        // 0.44 to 0.36 ms.  (For D >= 512 the L2 -> shared-memory operand stream bounds the kernel instead.)
        auto push = [&](float av, int col) {                 // record a candidate; the caller has tested av <= thr
          if (cnt == ADDR_CAPH) {                            // full: drop entries the tightened threshold no longer admits
            int w = 0;
            for (int i = 0; i < ADDR_CAPH; ++i) {
              const float x = lv[i];
              const uint16_t ci = li[i];
              if (x <= thr) { lv[w] = x; li[w] = ci; ++w; }
            }
            cnt = w;
          }
          if (cnt < ADDR_CAPH) { lv[cnt] = av; li[cnt] = (uint16_t)col; ++cnt; }
          else overflow = true;
        };
        fetch(0, vbuf[0], ebuf[0]);
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
          ptx::tmem_ld_wait();
          if (c + 1 < CHUNKS) fetch(c + 1, vbuf[(c + 1) & 1], ebuf[(c + 1) & 1]);
          uint32_t (&v)[32] = vbuf[c & 1];
          float4 (&e4)[8] = ebuf[c & 1];
          uint32_t hits = 0;
          const bool first = ti == 0 && c == 0;              // first 32 columns of the row: no threshold yet ->
          if (first) {                                       // branch-free selection of the KSEL best, then the hits
            float ma[4][KSEL];
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
              for (int i = 0; i < KSEL; ++i) ma[g][i] = INFINITY;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 e = e4[j];
              sel_insert<KSEL>(ma[0], fmaf(cs, __uint_as_float(v[4 * j + 0]), e.x));
              sel_insert<KSEL>(ma[1], fmaf(cs, __uint_as_float(v[4 * j + 1]), e.y));
              sel_insert<KSEL>(ma[2], fmaf(cs, __uint_as_float(v[4 * j + 2]), e.z));
              sel_insert<KSEL>(ma[3], fmaf(cs, __uint_as_float(v[4 * j + 3]), e.w));
            }
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
              for (int i = 0; i < KSEL; ++i) sel_insert<KSEL>(m, ma[g][i]);
            thr = fminf(m[KSEL - 1] + margin, FLT_MAX);      // padded columns score +inf and never hit
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 e = e4[j];
            hits |= (fmaf(cs, __uint_as_float(v[4 * j + 0]), e.x) <= thr ? 1u : 0u) << (4 * j + 0);
            hits |= (fmaf(cs, __uint_as_float(v[4 * j + 1]), e.y) <= thr ? 1u : 0u) << (4 * j + 1);
            hits |= (fmaf(cs, __uint_as_float(v[4 * j + 2]), e.z) <= thr ? 1u : 0u) << (4 * j + 2);
            hits |= (fmaf(cs, __uint_as_float(v[4 * j + 3]), e.w) <= thr ? 1u : 0u) << (4 * j + 3);
          }
          while (hits) {
            const int j = __ffs(hits) - 1;
            hits &= hits - 1;
            const int col = col0 + c * 32 + j;
            const float av = fmaf(cs, __uint_as_float(select32(v, j)), __ldg(p.en2pad + col));
            if (!first) {                                    // the first chunk's best are in m already
              sel_insert<KSEL>(m, av);
              thr = fminf(m[KSEL - 1] + margin, FLT_MAX);
            }
            if (av <= thr) push(av, col);
          }
        }
        } else {
        // ---- pass A
        fetch(0, vbuf[0], ebuf[0]);
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
          ptx::tmem_ld_wait();
          if (c + 1 < CHUNKS) fetch(c + 1, vbuf[(c + 1) & 1], ebuf[(c + 1) & 1]);
          else fetch(0, vbuf[(c + 1) & 1], ebuf[(c + 1) & 1]);       // first chunk of pass B
          uint32_t (&v)[32] = vbuf[c & 1];
          float4 (&e4)[8] = ebuf[c & 1];
          float ma[4][KSEL];
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int i = 0; i < KSEL; ++i) ma[g][i] = INFINITY;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 e = e4[j];
            sel_insert<KSEL>(ma[0], fmaf(cs, __uint_as_float(v[4 * j + 0]), e.x));
            sel_insert<KSEL>(ma[1], fmaf(cs, __uint_as_float(v[4 * j + 1]), e.y));
            sel_insert<KSEL>(ma[2], fmaf(cs, __uint_as_float(v[4 * j + 2]), e.z));
            sel_insert<KSEL>(ma[3], fmaf(cs, __uint_as_float(v[4 * j + 3]), e.w));
          }
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int i = 0; i < KSEL; ++i) sel_insert<KSEL>(m, ma[g][i]);
        }
        // ---- pass B
        thr = fminf(m[KSEL - 1] + margin, FLT_MAX);   // padded columns score +inf and never hit
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
          ptx::tmem_ld_wait();
          // pass A left chunk 0 of this pass in buffer CHUNKS & 1; the buffers keep alternating from there
          if (c + 1 < CHUNKS) fetch(c + 1, vbuf[(CHUNKS + c + 1) & 1], ebuf[(CHUNKS + c + 1) & 1]);
          uint32_t (&v)[32] = vbuf[(CHUNKS + c) & 1];
          float4 (&e4)[8] = ebuf[(CHUNKS + c) & 1];
          uint32_t hits = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 e = e4[j];
            hits |= (fmaf(cs, __uint_as_float(v[4 * j + 0]), e.x) <= thr ? 1u : 0u) << (4 * j + 0);
            hits |= (fmaf(cs, __uint_as_float(v[4 * j + 1]), e.y) <= thr ? 1u : 0u) << (4 * j + 1);
            hits |= (fmaf(cs, __uint_as_float(v[4 * j + 2]), e.z) <= thr ? 1u : 0u) << (4 * j + 2);
            hits |= (fmaf(cs, __uint_as_float(v[4 * j + 3]), e.w) <= thr ? 1u : 0u) << (4 * j + 3);
          }
          while (hits) {
            const int j = __ffs(hits) - 1;
            hits &= hits - 1;
            const int col = col0 + c * 32 + j;
            const float av = fmaf(cs, __uint_as_float(select32(v, j)), __ldg(p.en2pad + col));
            if (cnt == ADDR_CAPH) {            // full: drop entries that the tightened threshold no longer admits
              int w = 0;
              for (int i = 0; i < ADDR_CAPH; ++i) {
                const float x = lv[i];
                const uint16_t ci = li[i];
                if (x <= thr) { lv[w] = x; li[w] = ci; ++w; }
              }
              cnt = w;
            }
            if (cnt < ADDR_CAPH) { lv[cnt] = av; li[cnt] = (uint16_t)col; ++cnt; }
            else overflow = true;
          }
        }
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&tmem_empty[acc]);
      }
      // ---- merge the two column halves of the row and decide.  Half 1 hands its best scores and list over in shared
      // memory; the half-0 thread prunes both lists against the row's final threshold, and -- new in round 2 -- checks
      // whether the approximate ranking is already PROVEN: every candidate's approximate score is within margin / 2 of
      // the fp32 distance the exact kernel would compute, so k candidates whose consecutive scores (and the next one, if
      // any survived) lie more than `margin` apart are in their final order and need no exact distance at all.
      const int row = q * 32 + lane;
      if (half == 1) {
#pragma unroll
        for (int i = 0; i < KSEL; ++i) xch_m[row * 4 + i] = m[i];
        xch_cnt[row] = overflow ? -1 : cnt;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (half == 0 && n < p.N) {
#pragma unroll
        for (int i = 0; i < KSEL; ++i) sel_insert<KSEL>(m, xch_m[row * 4 + i]);
        const float thr_final = fminf(m[KSEL - 1] + margin, FLT_MAX);
        const int ocnt = xch_cnt[row];
        const bool ovf = overflow || ocnt < 0;
        const float* olv = lv + (size_t)4 * 32 * ADDR_CAPH;      // the partner warp (warp + 4) holds the other half
        const uint16_t* oli = li + (size_t)4 * 32 * ADDR_CAPH;
        int* dst = p.cand + (size_t)n * ADDR_CAND;
        float sv[KSEL + 1];
        int si[KSEL + 1];
#pragma unroll
        for (int i = 0; i <= KSEL; ++i) { sv[i] = INFINITY; si[i] = 0; }
        int w = 0;
        auto take = [&](float x, int c) {
          if (x <= thr_final) {
            dst[w++] = c;
#pragma unroll
            for (int i = 0; i <= KSEL; ++i) {
              const bool lt = x < sv[i];
              const float tx = lt ? sv[i] : x; const int tc = lt ? si[i] : c;
              sv[i] = lt ? x : sv[i]; si[i] = lt ? c : si[i];
              x = tx; c = tc;
            }
          }
        };
        for (int i = 0; i < cnt; ++i) take(lv[i], (int)li[i]);
        for (int i = 0; i < ocnt; ++i) take(olv[i], (int)oli[i]);
        bool decided = !ovf;
#pragma unroll
        for (int i = 0; i < KSEL; ++i)
          if (i < p.k) decided = decided && (sv[i + 1] - sv[i] > margin);      // inf - inf = NaN -> undecided
        if (decided) {
#pragma unroll
          for (int i = 0; i < KSEL; ++i)
            if (i < p.k) dst[i] = si[i];
          w = p.k;
        }
        p.cand_cnt[(size_t)n * 2] = ovf ? ADDR_CAND + 1 : w;
        p.cand_cnt[(size_t)n * 2 + 1] = decided ? 1 : 0;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");   // the lists are free for the next query tile
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
// operand prep
// ------------------------------------------------------------------------------------------------
// z [N][D] fp32 -> zp fp16 of the row scaled by a power of two s_n (max component -> [2^14, 2^15)) and
// zmeta[n] = (||z_n||^2, 1 / s_n); one warp per query row
__global__ void __launch_bounds__(256) pack_rows_f16_kernel(const float* __restrict__ z, __half* __restrict__ zp,
                                                             float2* __restrict__ zmeta, long long N, int D) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* zr = z + row * D;
  __half* zo = zp + row * D;
  float s = 0.f, mx = 0.f;
  for (int d = lane * 2; d < D; d += 64) {      // D is a multiple of 64
    const float2 v = __ldg(reinterpret_cast<const float2*>(zr + d));
    s = fmaf(v.x, v.x, fmaf(v.y, v.y, s));
    mx = fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y)));
  }
  s = warp_sum(s);
  const float sc = q_scale_for_bound(warp_max(mx));
  for (int d = lane * 2; d < D; d += 64) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(zr + d));
    *reinterpret_cast<uint32_t*>(zo + d) = ptx::pack_f16x2(v.x * sc, v.y * sc);
  }
  if (lane == 0) zmeta[row] = make_float2(s, 1.f / sc);
}

// max |bank| as the bits of a non-negative float (atomicMax); *out zeroed by the host
__global__ void __launch_bounds__(256) bank_absmax_kernel(const float* __restrict__ bank_t, long long n, unsigned* __restrict__ out) {
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(bank_t[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}

// bank_t [M][D] fp32 -> bank_hi [Mpad][D] fp16 of the bank scaled by one power of two t (zero rows beyond M);
// en2pad [Mpad] (+inf beyond M); emax[0] = max ||e||, emax[1] = 1 / t.  amax_bits = max |bank| from bank_absmax_kernel.
__global__ void bank_pack_kernel(const float* __restrict__ bank_t, const float* __restrict__ en2,
                                 __half* __restrict__ bank_hi, float* __restrict__ en2pad,
                                 float* __restrict__ emax, const unsigned* __restrict__ amax_bits, int D, int M, int Mpad) {
  const int m = blockIdx.x;
  const float t = q_scale_for_bound(__uint_as_float(amax_bits[0]));
  for (int d = threadIdx.x; d < D; d += blockDim.x)
    bank_hi[(size_t)m * D + d] = __float2half_rn(m < M ? bank_t[(size_t)m * D + d] * t : 0.f);
  if (threadIdx.x == 0) {
    en2pad[m] = m < M ? en2[m] : INFINITY;
    if (m < M) atomicMax(reinterpret_cast<int*>(emax), __float_as_int(sqrtf(en2[m])));  // non-negative floats order as ints
    if (m == 0) emax[1] = 1.f / t;
  }
}

// ------------------------------------------------------------------------------------------------
// refine: exact fp32 ranking of the candidates (or of all items when the list overflowed), gathers, commit partials,
// EMA statistics.  Arithmetic identical to address_kernel in mem_simt.cu -> identical indices.
// ------------------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void team_merge(TopK<K>& top) {
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
    float ov[K]; int oi[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
      ov[i] = __shfl_xor_sync(0xffffffffu, top.v[i], o);
      oi[i] = __shfl_xor_sync(0xffffffffu, top.id[i], o);
    }
#pragma unroll
    for (int i = 0; i < K; ++i) top.insert(ov[i], oi[i]);
  }
}

// stats: [0] rows needing the exact re-scan (reported), [1] path id, [2] length of the re-scan work list
template <int K>
__global__ void __launch_bounds__(256) refine_kernel(
    const float* __restrict__ z, const int* __restrict__ cand, const int* __restrict__ cand_cnt,
    const float* __restrict__ bank_t, const float* __restrict__ en2,
    float* __restrict__ read, float* __restrict__ q1, int64_t* __restrict__ idx, float* __restrict__ sse_px,
    float* __restrict__ counts, float* __restrict__ embed_sum, int* __restrict__ stats, int* __restrict__ rescan_list,
    __nv_bfloat16* __restrict__ read_planes, long long read_plane_stride, int N, int D, int M) {
  const int t = threadIdx.x;
  const int px = t >> 2, part = t & 3;
  const int n = blockIdx.x * 64 + px;
  bool valid = n < N;
  const float* zr = z + (size_t)(valid ? n : 0) * D;
  const float zn2 = team_zn2(zr, D, part, valid);
  TopK<K> top;
  top.init();
  if (valid) {
    const int c0 = cand_cnt[(size_t)n * 2];
    const bool rescan = c0 > ADDR_CAND || c0 < K;                                 // uniform within the 4-lane team
    if (rescan) {
      if (part == 0) {
        atomicAdd(&stats[0], 1);
        rescan_list[atomicAdd(&stats[2], 1)] = n;
      }
      valid = false;                                                              // rescan_kernel owns this row
    } else {
      const int* cr = cand + (size_t)n * ADDR_CAND;
      for (int c = part; c < c0; c += 4) {
        const int j = cr[c];
        top.insert(exact_dist(zn2, exact_dot(zr, bank_t + (size_t)j * D, D), en2[j]), j);
      }
    }
  }
  team_merge<K>(top);
#pragma unroll
  for (int i = 0; i < K; ++i) top.id[i] = min(top.id[i], M - 1);
  team_emit_row<K>(zr, bank_t, top.id, (int64_t)n, D, M, part, valid, read, q1, idx, sse_px, counts, embed_sum,
                   read_planes, read_plane_stride);
}

// ------------------------------------------------------------------------------------------------
// Tail for D >= 128: ONE WARP PER QUERY, every global access a coalesced 16-byte-per-lane row segment.
//   * rows the filter decided (cand_cnt[n][1] == 1): no distance is computed; the warp streams z and the k item rows
//     and writes q1 / read / idx / the commit partial -- a pure gather at HBM speed.
//   * ambiguous rows: z and up to TAIL_CM candidate rows pass through shared memory in 128-element segments (coalesced
//     loads, the next segment in flight while this one is used); lane c then walks candidate c's fmaf chain in
//     ascending d and every lane the ||z||^2 chain of its component (lane & 3), i.e. exactly the arithmetic of
//     exact_dot / team_zn2 above, so the ranking is bit-identical to refine_kernel and to the fp32 kernel.
// The commit partial keeps the summation tree of team_emit_row (four interleaved chains over the 16-byte chunks, then
// (s0 + s1) + (s2 + s3)): a chain crosses the lanes in order, eight hops per 512 elements.
// ------------------------------------------------------------------------------------------------
constexpr int TAIL_CM = 4;      // candidate rows chained per pass (one lane each)
constexpr int TAIL_ROW4 = 33;   // staged segment pitch in 16-byte units: 32 + 1 shifts the banks row to row

__device__ __forceinline__ float chain4(float acc, const float4& a, const float4& b) {
  acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
  return acc;
}

template <int K>
__global__ void __launch_bounds__(512, 2) addr_tail_kernel(
    const float* __restrict__ z, const int* __restrict__ cand, const int* __restrict__ cand_cnt,
    const float* __restrict__ bank_t, const float* __restrict__ en2,
    float* __restrict__ read, float* __restrict__ q1, int64_t* __restrict__ idx, float* __restrict__ sse_px,
    float* __restrict__ counts, float* __restrict__ embed_sum, int* __restrict__ stats, int* __restrict__ rescan_list,
    __nv_bfloat16* __restrict__ read_planes, long long read_plane_stride, int N, int D, int M) {
  __shared__ float4 tail_smem[16 * (1 + TAIL_CM) * TAIL_ROW4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  float4* sz4 = tail_smem + warp * (1 + TAIL_CM) * TAIL_ROW4;   // this warp's staged 128-element segment of z ...
  float4* se4 = sz4 + TAIL_ROW4;                                 // ... and of up to TAIL_CM candidate rows
  const int chunks = D >> 2;
  const int part = lane & 3;
  for (int n = blockIdx.x * wpb + warp; n < N; n += gridDim.x * wpb) {
    const int cnt = __ldg(cand_cnt + (size_t)n * 2);
    const int decided = __ldg(cand_cnt + (size_t)n * 2 + 1);
    if (cnt > ADDR_CAND || cnt < K) {                       // a list overflowed (or M < K items scored): exact re-scan
      if (lane == 0) {
        atomicAdd(&stats[0], 1);
        rescan_list[atomicAdd(&stats[2], 1)] = n;
      }
      continue;
    }
    const int* cr = cand + (size_t)n * ADDR_CAND;
    const float4* z4 = reinterpret_cast<const float4*>(z + (size_t)n * D);
    int ids[K];
    if (decided) {
#pragma unroll
      for (int i = 0; i < K; ++i) ids[i] = __ldg(cr + i);
    } else {
      TopK<K> top;
      top.init();
      float zn2 = 0.f;
      for (int c0 = 0; c0 < cnt; c0 += TAIL_CM) {
        const int nc = min(TAIL_CM, cnt - c0);
        const bool dotl = lane < nc;
        const int jme = dotl ? __ldg(cr + c0 + lane) : 0;
        const float4* e4[TAIL_CM];
#pragma unroll
        for (int c = 0; c < TAIL_CM; ++c)
          e4[c] = reinterpret_cast<const float4*>(bank_t + (size_t)__shfl_sync(0xffffffffu, jme, c) * D);
        const float4* er4 = se4 + (dotl ? lane : 0) * TAIL_ROW4;   // lane c < nc chains candidate c (the others idle on row 0)
        float acc = 0.f, zacc = 0.f;
        // segments of 128 elements: the next one is fetched (coalesced, 512 bytes per row) while this one is chained
        float4 pz, pe[TAIL_CM];
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        pz = lane < chunks ? __ldg(z4 + lane) : zero4;
#pragma unroll
        for (int c = 0; c < TAIL_CM; ++c) pe[c] = (c < nc && lane < chunks) ? __ldg(e4[c] + lane) : zero4;
        for (int s0 = 0; s0 < chunks; s0 += 32) {
          sz4[lane] = pz;
#pragma unroll
          for (int c = 0; c < TAIL_CM; ++c) se4[c * TAIL_ROW4 + lane] = pe[c];
          __syncwarp();
          const int nx = s0 + 32 + lane;
          pz = nx < chunks ? __ldg(z4 + nx) : zero4;
#pragma unroll
          for (int c = 0; c < TAIL_CM; ++c) pe[c] = (c < nc && nx < chunks) ? __ldg(e4[c] + nx) : zero4;
          const int lim = min(32, chunks - s0);
          // candidate chain + (same pass, same shared-memory read of z) the ||z||^2 chain of component `part`; shared
          // memory wavefronts, not instruction issue, bound this loop (ncu), hence selects instead of a second read
#pragma unroll 4
          for (int i = 0; i < lim; ++i) {
            const float4 a = sz4[i];
            acc = chain4(acc, a, er4[i]);
            const float c = (part & 1) ? ((part & 2) ? a.w : a.y) : ((part & 2) ? a.z : a.x);
            zacc = fmaf(c, c, zacc);
          }
          __syncwarp();
        }
        if (c0 == 0) {
          const float s0 = __shfl_sync(0xffffffffu, zacc, 0), s1 = __shfl_sync(0xffffffffu, zacc, 1);
          const float s2 = __shfl_sync(0xffffffffu, zacc, 2), s3 = __shfl_sync(0xffffffffu, zacc, 3);
          zn2 = (s0 + s1) + (s2 + s3);
        }
        const float dme = dotl ? exact_dist(zn2, acc, __ldg(en2 + jme)) : INFINITY;
        for (int c = 0; c < nc; ++c) top.insert(__shfl_sync(0xffffffffu, dme, c), __shfl_sync(0xffffffffu, jme, c));
      }
#pragma unroll
      for (int i = 0; i < K; ++i) ids[i] = top.id[i];
    }
#pragma unroll
    for (int i = 0; i < K; ++i) ids[i] = min(ids[i], M - 1);
    // ---- outputs
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < K; ++i) idx[(size_t)n * K + i] = (int64_t)ids[i];
      if (counts) atomicAdd(&counts[ids[0]], 1.f);
    }
    const float4* e1 = reinterpret_cast<const float4*>(bank_t + (size_t)ids[0] * D);
    float4* q4 = reinterpret_cast<float4*>(q1 + (size_t)n * D);
    float carry = 0.f;                                      // chain `part` of the commit partial, replicated per part
    for (int g = 0; g < chunks; g += 128) {                 // 4 chunks per lane and pass
      float4 zv[4], ev[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int i = g + 32 * jj + lane;
        const bool ok = i < chunks;
        zv[jj] = ok ? __ldg(z4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        ev[jj] = ok ? __ldg(e1 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int i = g + 32 * jj + lane;
        const bool ok = i < chunks;
        const float4 df = make_float4(ev[jj].x - zv[jj].x, ev[jj].y - zv[jj].y, ev[jj].z - zv[jj].z, ev[jj].w - zv[jj].w);
        if (ok) {
          q4[i] = make_float4(zv[jj].x + df.x, zv[jj].y + df.y, zv[jj].z + df.z, zv[jj].w + df.w);
          if (read) reinterpret_cast<float4*>(read + ((size_t)n * K) * D)[i] = ev[jj];
          if (embed_sum) {
            atomicAdd(&embed_sum[(size_t)(4 * i + 0) * M + ids[0]], zv[jj].x);
            atomicAdd(&embed_sum[(size_t)(4 * i + 1) * M + ids[0]], zv[jj].y);
            atomicAdd(&embed_sum[(size_t)(4 * i + 2) * M + ids[0]], zv[jj].z);
            atomicAdd(&embed_sum[(size_t)(4 * i + 3) * M + ids[0]], zv[jj].w);
          }
        }
        if (g + 32 * jj < chunks) {                          // uniform: this group of 32 chunks exists
#pragma unroll
          for (int l = 0; l < 8; ++l) {
            const float nv = chain4(carry, df, df);
            const float pass = ((lane >> 2) == l && ok) ? nv : carry;
            carry = __shfl_sync(0xffffffffu, pass, 4 * l + part);
          }
        }
      }
    }
    {
      float s = carry + __shfl_xor_sync(0xffffffffu, carry, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (lane == 0) sse_px[n] = s;
    }
    if (read) {
#pragma unroll
      for (int j = 1; j < K; ++j) {
        const float4* er = reinterpret_cast<const float4*>(bank_t + (size_t)ids[j] * D);
        float4* rr = reinterpret_cast<float4*>(read + ((size_t)n * K + j) * D);
        for (int i = lane; i < chunks; i += 32) rr[i] = __ldg(er + i);
      }
    }
    if (read_planes) {                                      // bf16 hi/lo split of the read: A operand of the `dec` GEMM
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const float4* er = reinterpret_cast<const float4*>(bank_t + (size_t)ids[j] * D);
        __nv_bfloat16* hp = read_planes + ((size_t)n * K + j) * D;
        for (int i = lane; i < chunks; i += 32) {
          const float4 v = __ldg(er + i);
          uint2 ph, pl;
          ptx::split_pack_bf16x2(v.x, v.y, ph.x, pl.x);
          ptx::split_pack_bf16x2(v.z, v.w, ph.y, pl.y);
          *reinterpret_cast<uint2*>(hp + 4 * i) = ph;
          *reinterpret_cast<uint2*>(hp + read_plane_stride + 4 * i) = pl;
        }
      }
    }
  }
}

// Exact scan over all M items for the (rare) rows whose candidate list overflowed: one warp per row.
template <int K>
__global__ void __launch_bounds__(256) rescan_kernel(
    const float* __restrict__ z, const float* __restrict__ bank_t, const float* __restrict__ en2,
    float* __restrict__ read, float* __restrict__ q1, int64_t* __restrict__ idx, float* __restrict__ sse_px,
    float* __restrict__ counts, float* __restrict__ embed_sum, const int* __restrict__ stats,
    const int* __restrict__ rescan_list, __nv_bfloat16* __restrict__ read_planes, long long read_plane_stride, int D,
    int M) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int total = stats[2];
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total; w += warps) {
    const int n = rescan_list[w];
    const float* zr = z + (size_t)n * D;
    const float zn2 = __shfl_sync(0xffffffffu, team_zn2(zr, D, lane & 3, true), 0);
    TopK<K> top;
    top.init();
    for (int j = lane; j < M; j += 32)
      top.insert(exact_dist(zn2, exact_dot(zr, bank_t + (size_t)j * D, D), en2[j]), j);
#pragma unroll
    for (int o = 1; o <= 16; o <<= 1) {
      float ov[K]; int oi[K];
#pragma unroll
      for (int i = 0; i < K; ++i) {
        ov[i] = __shfl_xor_sync(0xffffffffu, top.v[i], o);
        oi[i] = __shfl_xor_sync(0xffffffffu, top.id[i], o);
      }
#pragma unroll
      for (int i = 0; i < K; ++i) top.insert(ov[i], oi[i]);
    }
#pragma unroll
    for (int i = 0; i < K; ++i) top.id[i] = min(top.id[i], M - 1);
    team_emit_row<K>(zr, bank_t, top.id, (int64_t)n, D, M, lane & 3, lane < 4, read, q1, idx, sse_px, counts, embed_sum,
                     read_planes, read_plane_stride);
  }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
int make_map_2d_bf16(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint32_t box_inner,
                     uint32_t box_outer);   // amft_conv.cu

int pack_bank_padded(const float* bank_t, const float* en2, void* bank_hi, float* en2pad, float* emax4, int D, int M, int Mpad,
                     cudaStream_t st);

int addr_block_n(int M) { return M >= 192 ? 256 : (M >= 96 ? 128 : 64); }

size_t addr_tc_ws_bytes(int64_t N, int D, int M) {
  const int bn = addr_block_n(M);
  const int Mpad = (int)align_up(M, bn);
  return align_up((size_t)N * D * 2, 256) + align_up((size_t)Mpad * D * 2, 256) + align_up((size_t)Mpad * 4, 256) +
         align_up((size_t)N * ADDR_CAND * 4, 256) + align_up((size_t)N * 2 * 4, 256) + align_up((size_t)N * 8, 256) +
         align_up((size_t)N * 4, 256) + 512;
}

bool addr_tc_supported(int64_t N, int D, int M, int k) {
  return D % 64 == 0 && D >= 64 && M >= 16 && M <= 65536 && k <= 4 && N >= 1;
}

template <int BLOCK_N, int KSEL, bool SWEEP>
static int launch_addr3(const CUtensorMap& tmA, const CUtensorMap& tmB, const AddrParams& p, cudaStream_t st) {
  using S = AddrSmem<BLOCK_N>;
  static bool configured[64] = {false};
  int dev = 0;
  AMMC_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    AMMC_CUDA_CHECK(cudaFuncSetAttribute(addr_tc_kernel<BLOCK_N, KSEL, SWEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         S::TOTAL));
    configured[dev] = true;
  }
  int grid = min(num_sms(), p.tiles_q);
  addr_tc_kernel<BLOCK_N, KSEL, SWEEP><<<grid, ADDR_THREADS, S::TOTAL, st>>>(tmA, tmB, p);
  AMMC_LAUNCH_CHECK("addr_tc_kernel");
  return 0;
}

template <int BLOCK_N, int KSEL>
static int launch_addr2(const CUtensorMap& tmA, const CUtensorMap& tmB, const AddrParams& p, cudaStream_t st) {
  // long banks: the running threshold is tight for most of the sweep.  Measured break-even: 8 item tiles when the epilogue
  // sets the pace (D <= 256), 16 otherwise (at 4 tiles the sweep is 5 % slower than the two passes)
  const bool sweep = p.tiles_i >= 16 || (p.tiles_i >= 8 && p.D <= 256);
  return sweep ? launch_addr3<BLOCK_N, KSEL, true>(tmA, tmB, p, st) : launch_addr3<BLOCK_N, KSEL, false>(tmA, tmB, p, st);
}

template <int BLOCK_N>
static int launch_addr(const CUtensorMap& tmA, const CUtensorMap& tmB, const AddrParams& p, int k, cudaStream_t st) {
  return k <= 2 ? launch_addr2<BLOCK_N, 2>(tmA, tmB, p, st) : launch_addr2<BLOCK_N, 4>(tmA, tmB, p, st);
}

static int launch_filter(const void* zp, const void* bank_hi, AddrParams& p, int bn, int k, cudaStream_t st) {
  CUtensorMap tmA, tmB;
  if (int rc = make_map_2d_bf16(&tmA, zp, (uint64_t)p.D, (uint64_t)p.N, 64, 128)) return rc;
  if (int rc = make_map_2d_bf16(&tmB, bank_hi, (uint64_t)p.D, (uint64_t)p.Mpad, 64, (uint32_t)bn)) return rc;
  return bn == 256 ? launch_addr<256>(tmA, tmB, p, k, st)
                   : (bn == 128 ? launch_addr<128>(tmA, tmB, p, k, st) : launch_addr<64>(tmA, tmB, p, k, st));
}

// z fp32 [N][D] (+ optional pre-packed bf16 copy zp), bank_t fp32 [M][D], en2 [M]  ->  outputs as address_kernel.
// `ws` must provide addr_tc_ws_bytes(); stats[0] += rows that needed the exact fallback.
int run_address_tc(const float* z, const void* zp_in, const float* zmeta_in, __nv_bfloat16* read_planes,
                   const float* bank_t, const float* en2, float* read, float* q1, int64_t* idx, float* sse_px,
                   float* counts, float* embed_sum, int* stats, Workspace& ws, int64_t N, int D, int M, int k,
                   cudaStream_t st) {
  const long long rps = (long long)N * k * D;
  const int bn = addr_block_n(M);
  const int Mpad = (int)align_up(M, bn);
  __half* zp = ws.take<__half>((size_t)N * D);
  __half* bank_hi = ws.take<__half>((size_t)Mpad * D);
  float* en2pad = ws.take<float>(Mpad);
  int* cand = ws.take<int>((size_t)N * ADDR_CAND);
  int* cand_cnt = ws.take<int>((size_t)N * 2);
  float2* zmeta = ws.take<float2>(N);
  int* rescan_list = ws.take<int>(N);
  float* emax = ws.take<float>(4);
  if (!ws.ok()) return fail(AMMC_EWORKSPACE, "workspace too small");
  if (zp_in && zmeta_in) {                  // the tensor-core enc epilogue already produced the scaled fp16 rows and norms
    zp = reinterpret_cast<__half*>(const_cast<void*>(zp_in));
    zmeta = reinterpret_cast<float2*>(const_cast<float*>(zmeta_in));
  } else {
    pack_rows_f16_kernel<<<ceil_div(N, 8), 256, 0, st>>>(z, zp, zmeta, N, D);
    AMMC_LAUNCH_CHECK("pack_rows_f16_kernel");
  }
  if (int rc = pack_bank_padded(bank_t, en2, bank_hi, en2pad, emax, D, M, Mpad, st)) return rc;

  AddrParams p;
  p.N = (int)N; p.D = D; p.M = M; p.Mpad = Mpad;
  p.tiles_q = ceil_div(N, 128);
  p.tiles_i = Mpad / bn;
  p.en2pad = en2pad; p.zmeta = zmeta; p.emax = emax; p.cand = cand; p.cand_cnt = cand_cnt; p.k = k;
  if (int rc = launch_filter(zp, bank_hi, p, bn, k, st)) return rc;
  const int blocks = ceil_div(N, 64);
  const bool tail = D >= 128;                      // warp-per-query tail; below that a 4-lane team per query is denser
  const int tail_blocks = min(ceil_div(N, 16), 2 * num_sms());
  switch (k) {
#define AMMC_RF_CASE(KK)                                                                                      \
  case KK:                                                                                                    \
    if (tail) {                                                                                               \
      addr_tail_kernel<KK><<<tail_blocks, 512, 0, st>>>(                                                      \
          z, cand, cand_cnt, bank_t, en2, read, q1, idx, sse_px, counts, embed_sum, stats, rescan_list,       \
          read_planes, rps, (int)N, D, M);                                                                    \
    } else {                                                                                                  \
      refine_kernel<KK><<<blocks, 256, 0, st>>>(z, cand, cand_cnt, bank_t, en2, read, q1, idx, sse_px, counts, \
                                                embed_sum, stats, rescan_list, read_planes, rps, (int)N, D, M); \
    }                                                                                                         \
    AMMC_LAUNCH_CHECK("refine_kernel / addr_tail_kernel");                                                    \
    rescan_kernel<KK><<<2 * num_sms(), 256, 0, st>>>(z, bank_t, en2, read, q1, idx, sse_px, counts, embed_sum, \
                                                     stats, rescan_list, read_planes, rps, D, M);             \
    break;
    AMMC_RF_CASE(1) AMMC_RF_CASE(2) AMMC_RF_CASE(3) AMMC_RF_CASE(4)
#undef AMMC_RF_CASE
    default: return fail(AMMC_EUNSUPPORTED, "tensor-core addressing supports k <= 4");
  }
  AMMC_LAUNCH_CHECK("rescan_kernel");
  return 0;
}

// exact re-scan of the rows the fused front kernel (mem_front.cu) queued: stats[2] rows listed in rescan_list
int run_rescan(const float* z, const float* bank_t, const float* en2, float* q1, int64_t* idx, float* sse_px, int* stats,
               int* rescan_list, __nv_bfloat16* read_planes, long long read_plane_stride, int D, int M, int k,
               cudaStream_t st) {
  switch (k) {
#define AMMC_RS_CASE(KK)                                                                                                  \
  case KK:                                                                                                                \
    rescan_kernel<KK><<<2 * num_sms(), 256, 0, st>>>(z, bank_t, en2, nullptr, q1, idx, sse_px, nullptr, nullptr, stats,   \
                                                     rescan_list, read_planes, read_plane_stride, D, M);                  \
    break;
    AMMC_RS_CASE(1) AMMC_RS_CASE(2) AMMC_RS_CASE(3) AMMC_RS_CASE(4)
#undef AMMC_RS_CASE
    default: return fail(AMMC_EUNSUPPORTED, "re-scan supports k <= 4");
  }
  AMMC_LAUNCH_CHECK("rescan_kernel");
  return 0;
}

// bank_t [M][D], en2 [M] -> bank_hi [Mpad][D] fp16 (power-of-two scaled), en2pad [Mpad], emax4 = {max ||e||, 1 / t, scratch}
int pack_bank_padded(const float* bank_t, const float* en2, void* bank_hi, float* en2pad, float* emax4, int D, int M, int Mpad,
                     cudaStream_t st) {
  AMMC_CUDA_CHECK(cudaMemsetAsync(emax4, 0, 16, st));
  unsigned* amax = reinterpret_cast<unsigned*>(emax4 + 2);
  bank_absmax_kernel<<<min(ceil_div((long long)M * D, 256), 592), 256, 0, st>>>(bank_t, (long long)M * D, amax);
  AMMC_LAUNCH_CHECK("bank_absmax_kernel");
  bank_pack_kernel<<<Mpad, 128, 0, st>>>(bank_t, en2, (__half*)bank_hi, en2pad, emax4, amax, D, M, Mpad);
  AMMC_LAUNCH_CHECK("bank_pack_kernel");
  return 0;
}

// mem_simt.cu
__global__ void bank_transpose_kernel(const float* __restrict__ embed, float* __restrict__ bank_t, int D, int M);
__global__ void bank_norms_kernel(const float* __restrict__ embed, float* __restrict__ en2, int D, int M);

}  // namespace ammc

namespace ammc { AMMC_DEFINE_TIMEOUT_READER(timeout_reader_addr) }

using namespace ammc;

// ---- staged entry points (what ammc_quantize_fwd / ammc_mem_fwd compose internally; used by the cfg5 microbench) ----
extern "C" int ammc_addr_padded_items(int M) { return (int)align_up(M, addr_block_n(M)); }

extern "C" int ammc_addr_pack_queries(const float* z, void* zp, float* zmeta, int64_t N, int D, void* stream) {
  AMMC_REQUIRE(z && zp && zmeta && N > 0 && D > 0 && D % 64 == 0, "bad argument (D must be a multiple of 64)");
  pack_rows_f16_kernel<<<ceil_div(N, 8), 256, 0, (cudaStream_t)stream>>>(z, (__half*)zp, (float2*)zmeta, N, D);
  AMMC_LAUNCH_CHECK("pack_rows_f16_kernel");
  return 0;
}

extern "C" int ammc_addr_pack_bank(const float* embed, float* bank_t, float* en2, void* bank_hi, float* en2pad,
                                   float* emax, int D, int M, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AMMC_REQUIRE(embed && bank_t && en2 && bank_hi && en2pad && emax && D > 0 && M > 0, "bad argument");
  const int Mpad = ammc_addr_padded_items(M);
  bank_transpose_kernel<<<dim3(ceil_div(M, 32), ceil_div(D, 32)), dim3(32, 8), 0, st>>>(embed, bank_t, D, M);
  AMMC_LAUNCH_CHECK("bank_transpose_kernel");
  bank_norms_kernel<<<ceil_div(M, 32), dim3(32, 8), 0, st>>>(embed, en2, D, M);
  AMMC_LAUNCH_CHECK("bank_norms_kernel");
  return pack_bank_padded(bank_t, en2, bank_hi, en2pad, emax, D, M, Mpad, st);
}

extern "C" int ammc_addr_filter(const void* zp, const float* zmeta, const void* bank_hi, const float* en2pad,
                                const float* emax, int* cand, int* cand_cnt, int64_t N, int D, int M, int k,
                                void* stream) {
  AMMC_REQUIRE(zp && zmeta && bank_hi && en2pad && emax && cand && cand_cnt, "null pointer argument");
  if (!addr_tc_supported(N, D, M, k))
    return fail(AMMC_EUNSUPPORTED, "tensor-core addressing needs D %% 64 == 0, 16 <= M <= 65536, k <= 4 (got D=%d M=%d k=%d)", D, M, k);
  AMMC_REQUIRE(N < (1LL << 31), "N too large");
  const int bn = addr_block_n(M);
  AddrParams p;
  p.N = (int)N; p.D = D; p.M = M; p.Mpad = ammc_addr_padded_items(M);
  p.tiles_q = ceil_div(N, 128);
  p.tiles_i = p.Mpad / bn;
  p.en2pad = en2pad; p.zmeta = reinterpret_cast<const float2*>(zmeta); p.emax = emax; p.cand = cand; p.cand_cnt = cand_cnt;
  p.k = k;
  return launch_filter(zp, bank_hi, p, bn, k, (cudaStream_t)stream);
}
