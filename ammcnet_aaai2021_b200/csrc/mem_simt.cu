// Memory module (VQ codebook with top-k read), generic-shape fp32 path.
//
// Replaces the ATen op sequence of reference Code/models/unet.py:282-331,384-387 (forward), its autograd
// backward, and the EMA bank update (unet.py:298-309).  Everything is exact fp32 FMA arithmetic, so the
// top-k ranking follows the reference's fp32 ranking; `dist`, the one-hot matrix and the permuted copies
// the reference materialises never exist in HBM.
//
// Kernels (N = b*h*w pixels, C channels, D embed_dim, M items, k):
//   bank_transpose / bank_norms / dec_table   per-call operand prep (tiny):  E^T [M,D], ||e||^2 [M],
//                                             T_j = (dec_w[:, jD:(j+1)D] . E)^T  [k][M][C]
//   enc1x1_kernel     z[n,:] = enc_w . x[n,:] + enc_b          NCHW in, [N,D] out         (unet.py:326)
//   address_kernel    dist -> top-k -> idx, read, q1, per-pixel SSE, EMA statistics      (unet.py:283-311)
//   dec_gather_kernel out[n,:] = sum_j T_j[idx_j(n)] + dec_b (+ x[n,:])  NCHW out         (unet.py:328-330,386)
//        (dec is linear and the read is a concatenation of bank rows, so dec(read_n) is a sum of k rows of
//         the precomputed tables: the N x kD x C contraction becomes a gather)
//   sse_frame_kernel                          deterministic per-frame and global commit reductions (unet.py:310), q-plane scale
//   backward: gz_kernel, then on tcgen05 when D, C % 64 == 0 (1x1 conv engine for the input gradient, weight-gradient
//             kernel of amft_train.cu for both weight gradients; pack_nhwc64 / read_planes / channel_sum prepare operands),
//             else gx_kernel, genc_w_kernel, gdec_scatter_priv_kernel (gdec_scatter_kernel for huge banks), gdec_w_kernel
//   ema_*                                     unet.py:298-309
#include "common.cuh"
#include "topk.cuh"
#include "addr_tail.cuh"
#include <cuda_bf16.h>
#include <float.h>

namespace ammc {

// ------------------------------------------------------------------------------------------------
// prep
// ------------------------------------------------------------------------------------------------
__global__ void bank_transpose_kernel(const float* __restrict__ embed, float* __restrict__ bank_t, int D, int M) {
  __shared__ float tile[32][33];
  const int m0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int d = d0 + r, m = m0 + threadIdx.x;
    tile[r][threadIdx.x] = (d < D && m < M) ? embed[(size_t)d * M + m] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int m = m0 + r, d = d0 + threadIdx.x;
    if (m < M && d < D) bank_t[(size_t)m * D + d] = tile[threadIdx.x][r];
  }
}

// ||e_m||^2 for every item.  Block = 32 items x 8 interleaved partial sums (d = g, g + 8, ...: independent coalesced
// loads, 16 in flight per thread), combined in a fixed order -- deterministic, and the ONE source of the norms every
// addressing path (fp32, tensor-core, fused front) ranks with.  Launch with dim3(32, 8) threads.
__global__ void __launch_bounds__(256) bank_norms_kernel(const float* __restrict__ embed, float* __restrict__ en2, int D, int M) {
  __shared__ float part[8][33];
  const int m = blockIdx.x * 32 + threadIdx.x, g = threadIdx.y;
  float acc = 0.f;
  if (m < M) {
    int d = g;
    for (; d + 8 * 15 < D; d += 8 * 16) {
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = embed[(size_t)(d + 8 * j) * M + m];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc = fmaf(v[j], v[j], acc);
    }
    for (; d < D; d += 8) {
      const float v = embed[(size_t)d * M + m];
      acc = fmaf(v, v, acc);
    }
  }
  part[g][threadIdx.x] = acc;
  __syncthreads();
  if (g == 0 && m < M) {
    float s = part[0][threadIdx.x];
#pragma unroll
    for (int i = 1; i < 8; ++i) s += part[i][threadIdx.x];
    en2[m] = s;
  }
}

// T[j][m][c] = sum_d dec_w[c][j*D + d] * embed[d][m]
__global__ void dec_table_kernel(const float* __restrict__ dec_w, const float* __restrict__ embed,
                                 float* __restrict__ T, int C, int D, int M, int k) {
  __shared__ float Ws[32][33];  // [c][d]
  __shared__ float Es[32][33];  // [d][m]
  const int c0 = blockIdx.x * 32, m0 = blockIdx.y * 32, j = blockIdx.z;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int d0 = 0; d0 < D; d0 += 32) {
    for (int r = ty; r < 32; r += 8) {
      int c = c0 + r, d = d0 + tx;
      Ws[r][tx] = (c < C && d < D) ? dec_w[(size_t)c * (k * D) + (size_t)j * D + d] : 0.f;
      int dd = d0 + r, m = m0 + tx;
      Es[r][tx] = (dd < D && m < M) ? embed[(size_t)dd * M + m] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int d = 0; d < 32; ++d) {
      float wv = Ws[tx][d];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(wv, Es[d][ty * 4 + i], acc[i]);
    }
    __syncthreads();
  }
  int c = c0 + tx;
  if (c < C) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int m = m0 + ty * 4 + i;
      if (m < M) T[((size_t)j * M + m) * C + c] = acc[i];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// enc 1x1:  z[n, d] = sum_c w[d, c] * x[img, c, p] + bias[d]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) enc1x1_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, float* __restrict__ z,
                                                      __nv_bfloat16* __restrict__ zp, int N, int HW, int C, int D) {
  __shared__ __align__(16) float Xs[16][64];
  __shared__ __align__(16) float Ws[16][68];
  const int t = threadIdx.x;
  const int n0 = blockIdx.x * 64, d0 = blockIdx.y * 64;
  const int tx = t & 15, ty = t >> 4;  // tx -> d group, ty -> px group
  // fixed pixel for this thread's X loads
  const int lp = t & 63, lc = t >> 6;
  const int ln = n0 + lp;
  const bool lvalid = ln < N;
  const int limg = lvalid ? ln / HW : 0, lpp = lvalid ? ln % HW : 0;
  const float* xbase = x + ((size_t)limg * C) * HW + lpp;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int c0 = 0; c0 < C; c0 += 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int c = c0 + lc + 4 * i;
      Xs[lc + 4 * i][lp] = (lvalid && c < C) ? xbase[(size_t)c * HW] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = t + i * 256;
      int d = e >> 4, c = e & 15;
      Ws[c][d] = (d0 + d < D && c0 + c < C) ? w[(size_t)(d0 + d) * C + c0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&Xs[kk][ty * 4]);
      float4 bq = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int n = n0 + ty * 4 + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int d = d0 + tx * 4 + j;
      if (d < D) {
        const float v = acc[i][j] + bias[d];
        z[(size_t)n * D + d] = v;
        if (zp) zp[(size_t)n * D + d] = __float2bfloat16_rn(v);   // operand of the tensor-core addressing filter
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// addressing: distances -> top-k -> gathers -> commit partials (+ EMA statistics)
// ------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(256) address_kernel(
    const float* __restrict__ z, const float* __restrict__ embed, const float* __restrict__ en2,
    const float* __restrict__ bank_t, float* __restrict__ read, float* __restrict__ q1,
    int64_t* __restrict__ idx, float* __restrict__ sse_px, float* __restrict__ counts,
    float* __restrict__ embed_sum, __nv_bfloat16* __restrict__ read_planes, int N, int D, int M) {
  __shared__ __align__(16) float Zs[16][68];
  __shared__ __align__(16) float Es[16][64];
  __shared__ float Ds[64][65];
  __shared__ float zn2_s[64];
  const int t = threadIdx.x;
  const int n0 = blockIdx.x * 64;
  const int tx = t & 15, ty = t >> 4;     // tx -> item group, ty -> px group (GEMM phase)
  const int px = t >> 2, part = t & 3;    // team-of-4 mapping (scan / gather phases)
  const int n_team = n0 + px;
  const bool team_valid = n_team < N;

  // ||z||^2 per pixel
  {
    const float s = team_zn2(z + (size_t)(team_valid ? n_team : 0) * D, D, part, team_valid);
    if (part == 0) zn2_s[px] = s;
  }
  TopK<K> top;
  top.init();
  __syncthreads();

  for (int m0 = 0; m0 < M; m0 += 64) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int d0 = 0; d0 < D; d0 += 16) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int e = t + i * 256;
        int p = e >> 4, d = e & 15;
        Zs[d][p] = (n0 + p < N && d0 + d < D) ? z[(size_t)(n0 + p) * D + d0 + d] : 0.f;
        int dd = e >> 6, m = e & 63;
        Es[dd][m] = (d0 + dd < D && m0 + m < M) ? embed[(size_t)(d0 + dd) * M + m0 + m] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        float4 a = *reinterpret_cast<const float4*>(&Zs[kk][ty * 4]);
        float4 bq = *reinterpret_cast<const float4*>(&Es[kk][tx * 4]);
        float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
    // dist = (||z||^2 - 2 z.e) + ||e||^2, same association as unet.py:283-288
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int m = m0 + tx * 4 + j;
      float e2 = m < M ? en2[m] : 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float dv = exact_dist(zn2_s[ty * 4 + i], acc[i][j], e2);
        Ds[ty * 4 + i][tx * 4 + j] = m < M ? dv : FLT_MAX;
      }
    }
    __syncthreads();
    for (int s = 0; s < 16; ++s) {
      int j = part + 4 * s;
      if (m0 + j < M) top.insert(Ds[px][j], m0 + j);
    }
    __syncthreads();
  }
  // merge the four partial lists of a team (butterfly; order-independent thanks to the total order)
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
  // NaN distances (a diverged bank) never enter the list; keep the gathers in bounds regardless
#pragma unroll
  for (int i = 0; i < K; ++i) top.id[i] = min(top.id[i], M - 1);
  team_emit_row<K>(z + (size_t)(team_valid ? n_team : 0) * D, bank_t, top.id, (int64_t)n_team, D, M, part, team_valid,
                   read, q1, idx, sse_px, counts, embed_sum, read_planes, (long long)N * K * D);
}

// One block per frame: per-frame commit partial (unet.py:310 summed over the frame's pixels).  The block that finishes
// last (ticket counter) also produces what used to be two more launches: `diff`, the batch mean over all N*D elements
// (frames summed in index order by one warp-strided loop: deterministic), and -- when asked -- the power-of-two scale of
// the module's q-plane output from max|x| and the prepared dec bound (see dec_bound_kernel).
__global__ void sse_frame_kernel(const float* __restrict__ sse_px, float* __restrict__ sse_frame, int64_t rows,
                                 float* __restrict__ diff, double inv_count, unsigned* __restrict__ ticket,
                                 const unsigned* __restrict__ amax_bits, const unsigned* __restrict__ bound_bits,
                                 int residual, float* __restrict__ qs) {
  __shared__ float red[33];
  __shared__ bool last;
  const float* p = sse_px + (size_t)blockIdx.x * rows;
  float s = 0.f;
  if ((rows & 3) == 0 && (((uintptr_t)p) & 15) == 0) {       // 16-byte loads, four in flight (one block per frame)
    const float4* p4 = reinterpret_cast<const float4*>(p);
    const int64_t n4 = rows >> 2;
#pragma unroll 4
    for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 v = __ldg(p4 + i);
      s += (v.x + v.y) + (v.z + v.w);
    }
  } else {
    for (int64_t i = threadIdx.x; i < rows; i += blockDim.x) s += p[i];
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    sse_frame[blockIdx.x] = s;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float t = 0.f;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += __ldcg(sse_frame + i);
  t = block_sum(t, red);
  if (threadIdx.x == 0) {
    diff[0] = (float)((double)t * inv_count);
    *ticket = 0;                                             // ready for the next call on this workspace
    if (qs) qs[0] = q_scale_for_bound((residual ? __uint_as_float(amax_bits[0]) : 0.f) + __uint_as_float(bound_bits[0]));
  }
}

// ------------------------------------------------------------------------------------------------
// dec 1x1 as a table gather + bias + residual, NCHW output
// ------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(256) dec_gather_kernel(
    const float* __restrict__ T, const int64_t* __restrict__ idx, const float* __restrict__ dec_b,
    const float* __restrict__ x, float* __restrict__ out, int N, int HW, int C, int M, int chunks_per_block) {
  __shared__ float tile[32][33];
  __shared__ int idx_s[32][K];
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  const int n0 = blockIdx.x * 32;
  for (int e = threadIdx.x; e < 32 * K; e += 256) {
    int p = e / K, j = e % K;
    idx_s[p][j] = (n0 + p < N) ? (int)idx[(size_t)(n0 + p) * K + j] : 0;
  }
  const int n = n0 + lane;
  const bool nvalid = n < N;
  const int img = nvalid ? n / HW : 0, pp = nvalid ? n % HW : 0;
  __syncthreads();
  const int cbeg = blockIdx.y * chunks_per_block * 32;
  for (int cc = 0; cc < chunks_per_block; ++cc) {
    const int c0 = cbeg + cc * 32;
    if (c0 >= C) break;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int p = wv * 4 + r;
      float v = 0.f;
      if (n0 + p < N && c0 + lane < C) {
#pragma unroll
        for (int j = 0; j < K; ++j) v += T[((size_t)j * M + idx_s[p][j]) * C + c0 + lane];
      }
      tile[p][lane] = v;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int ch = wv * 4 + r, c = c0 + ch;
      if (nvalid && c < C) {
        size_t o = ((size_t)img * C + c) * HW + pp;
        float v = tile[lane][ch] + dec_b[c];
        if (x) v += x[o];
        out[o] = v;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// embed_code
// ------------------------------------------------------------------------------------------------
__global__ void embed_code_kernel(const int64_t* __restrict__ ids, const float* __restrict__ embed,
                                  float* __restrict__ out, int64_t n_ids, int D, int M) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_ids * D) return;
  int64_t i = e / D;
  int d = (int)(e % D);
  int64_t id = ids[i];
  out[e] = (id >= 0 && id < M) ? embed[(size_t)d * M + id] : 0.f;
}

// ------------------------------------------------------------------------------------------------
// EMA bank update (unet.py:298-309)
// ------------------------------------------------------------------------------------------------
__global__ void ema_cluster_kernel(float* __restrict__ cluster_size, const float* __restrict__ counts, int M,
                                   float decay) {
  const float om = 1.f - decay;
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x)
    cluster_size[m] = __fadd_rn(__fmul_rn(cluster_size[m], decay), __fmul_rn(om, counts[m]));
}

// Every block recomputes n = sum(cluster_size) in the same order, so all blocks use the identical value.
__global__ void __launch_bounds__(256) ema_embed_kernel(float* __restrict__ embed, float* __restrict__ embed_avg,
                                                         const float* __restrict__ embed_sum,
                                                         const float* __restrict__ cluster_size, int D, int M,
                                                         float decay, float eps, int elems_per_block) {
  __shared__ float red[33];
  float s = 0.f;
  for (int m = threadIdx.x; m < M; m += blockDim.x) s += cluster_size[m];
  const float n = block_sum(s, red);
  const float om = 1.f - decay;
  const float denom = __fadd_rn(n, (float)((double)M * (double)eps));
  const int64_t total = (int64_t)D * M;
  const int64_t beg = (int64_t)blockIdx.x * elems_per_block;
  const int64_t end = min(total, beg + elems_per_block);
  for (int64_t e = beg + threadIdx.x; e < end; e += blockDim.x) {
    int m = (int)(e % M);
    float cs = __fmul_rn(__fdiv_rn(__fadd_rn(cluster_size[m], eps), denom), n);
    float avg = __fadd_rn(__fmul_rn(embed_avg[e], decay), __fmul_rn(om, embed_sum[e]));
    embed_avg[e] = avg;
    embed[e] = __fdiv_rn(avg, cs);
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// gz[n,d] = g_diff * 2 (z - e1) / (N D) + g_q1 ;  g_enc_b[d] += sum_n gz
__global__ void __launch_bounds__(256) gz_kernel(const float* __restrict__ z, const float* __restrict__ bank_t,
                                                  const int64_t* __restrict__ idx, const float* __restrict__ g_diff,
                                                  const float* __restrict__ g_q1, float* __restrict__ gz,
                                                  float* __restrict__ g_enc_b, int N, int D, int K, double inv_count) {
  extern __shared__ float colsum[];  // D
  for (int d = threadIdx.x; d < D; d += blockDim.x) colsum[d] = 0.f;
  __syncthreads();
  const float coef = (float)((double)g_diff[0] * 2.0 * inv_count);
  const int n0 = blockIdx.x * 64;
  const int rows = min(64, N - n0);
  for (int e = threadIdx.x; e < rows * D; e += blockDim.x) {
    int p = e / D, d = e % D;
    size_t n = (size_t)(n0 + p);
    int i1 = (int)idx[n * K];
    float v = coef * (z[n * D + d] - bank_t[(size_t)i1 * D + d]);
    if (g_q1) v += g_q1[n * D + d];
    gz[n * D + d] = v;
    atomicAdd(&colsum[d], v);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) atomicAdd(&g_enc_b[d], colsum[d]);
}

// Same values with 16-byte accesses for D = 4 * (power of two <= 256): a thread owns four embedding coordinates of every R-th
// row of the block's 64 rows, keeps their column sums in registers (no per-element shared atomics), and -- when gzp is given --
// also writes g_z as bf16 hi/lo planes [2][N*D], the operand format of the tensor-core input / weight gradients (what
// pack_planes_f32 would produce from gz in a second pass).
__global__ void __launch_bounds__(256) gz4_kernel(const float* __restrict__ z, const float* __restrict__ bank_t,
                                                   const int64_t* __restrict__ idx, const float* __restrict__ g_diff,
                                                   const float* __restrict__ g_q1, float* __restrict__ gz,
                                                   __nv_bfloat16* __restrict__ gzp, float* __restrict__ g_enc_b, int N, int D,
                                                   int K, double inv_count) {
  __shared__ float4 part[256];
  const int L = D >> 2, R = 256 / L;
  const int lane = threadIdx.x & (L - 1), rl = threadIdx.x / L;
  const float coef = (float)((double)g_diff[0] * 2.0 * inv_count);
  const int n0 = blockIdx.x * 64;
  const size_t plane = (size_t)N * D;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p = rl; p < 64 && n0 + p < N; p += R) {
    const size_t n = (size_t)(n0 + p), o = n * D + 4 * lane;
    const int i1 = (int)idx[n * K];
    const float4 zz = *reinterpret_cast<const float4*>(z + o);
    const float4 ee = __ldg(reinterpret_cast<const float4*>(bank_t + (size_t)i1 * D + 4 * lane));
    float4 v = make_float4(coef * (zz.x - ee.x), coef * (zz.y - ee.y), coef * (zz.z - ee.z), coef * (zz.w - ee.w));
    if (g_q1) {
      const float4 q = *reinterpret_cast<const float4*>(g_q1 + o);
      v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    }
    *reinterpret_cast<float4*>(gz + o) = v;
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    if (gzp) {
      const float vv[4] = {v.x, v.y, v.z, v.w};
      uint32_t hp[2], lp[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(vv[2 * j]), h1 = __float2bfloat16_rn(vv[2 * j + 1]);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(vv[2 * j] - __bfloat162float(h0));
        const __nv_bfloat16 l1 = __float2bfloat16_rn(vv[2 * j + 1] - __bfloat162float(h1));
        hp[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        lp[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
      }
      *reinterpret_cast<uint2*>(gzp + o) = make_uint2(hp[0], hp[1]);
      *reinterpret_cast<uint2*>(gzp + plane + o) = make_uint2(lp[0], lp[1]);
    }
  }
  part[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < L) {
    float4 t = part[threadIdx.x];
    for (int r = 1; r < R; ++r) {
      const float4 u = part[threadIdx.x + r * L];
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    atomicAdd(&g_enc_b[4 * threadIdx.x + 0], t.x);
    atomicAdd(&g_enc_b[4 * threadIdx.x + 1], t.y);
    atomicAdd(&g_enc_b[4 * threadIdx.x + 2], t.z);
    atomicAdd(&g_enc_b[4 * threadIdx.x + 3], t.w);
  }
}

static inline bool gz4_ok(int D, const void* a, const void* b, const void* c, const void* d) {
  const int L = D >> 2;
  return D % 4 == 0 && L >= 1 && L <= 256 && (L & (L - 1)) == 0 &&
         (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d) & 15) == 0;
}

// gx[img,c,p] = sum_d enc_w[d,c] gz[n,d] (+ g_out[img,c,p])
__global__ void __launch_bounds__(256) gx_kernel(const float* __restrict__ gz, const float* __restrict__ enc_w,
                                                  const float* __restrict__ g_out, float* __restrict__ gx,
                                                  int N, int HW, int C, int D, int residual) {
  __shared__ __align__(16) float As[16][68];  // [d][px]
  __shared__ __align__(16) float Bs[16][64];  // [d][c]
  const int t = threadIdx.x;
  const int n0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const int tx = t & 15, ty = t >> 4;  // tx -> px group, ty -> c group
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int d0 = 0; d0 < D; d0 += 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = t + i * 256;
      int p = e >> 4, d = e & 15;
      As[d][p] = (n0 + p < N && d0 + d < D) ? gz[(size_t)(n0 + p) * D + d0 + d] : 0.f;
      int dd = e >> 6, c = e & 63;
      Bs[dd][c] = (d0 + dd < D && c0 + c < C) ? enc_w[(size_t)(d0 + dd) * C + c0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[kk][tx * 4]);
      float4 bq = *reinterpret_cast<const float4*>(&Bs[kk][ty * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int n = n0 + tx * 4 + i;
    if (n >= N) continue;
    int img = n / HW, pp = n % HW;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = c0 + ty * 4 + j;
      if (c >= C) continue;
      size_t o = ((size_t)img * C + c) * HW + pp;
      float v = acc[i][j];
      if (residual) v += g_out[o];
      gx[o] = v;
    }
  }
}

// g_enc_w[d,c] += sum_{n in split} gz[n,d] x[img,c,p]   (split-K over pixels, atomics)
__global__ void __launch_bounds__(256) genc_w_kernel(const float* __restrict__ gz, const float* __restrict__ x,
                                                      float* __restrict__ g_enc_w, int N, int HW, int C, int D,
                                                      int px_per_split) {
  __shared__ __align__(16) float As[16][64];  // [n][d]
  __shared__ __align__(16) float Bs[16][68];  // [n][c]
  const int t = threadIdx.x;
  const int d0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const int nbeg = blockIdx.z * px_per_split, nend = min(N, nbeg + px_per_split);
  const int tx = t & 15, ty = t >> 4;  // tx -> c group, ty -> d group
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int nn = nbeg; nn < nend; nn += 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = t + i * 256;
      int r = e >> 6, d = e & 63;
      int n = nn + r;
      As[r][d] = (n < nend && d0 + d < D) ? gz[(size_t)n * D + d0 + d] : 0.f;
      int r2 = e & 15, c = e >> 4;
      int n2 = nn + r2;
      float v = 0.f;
      if (n2 < nend && c0 + c < C) {
        int img = n2 / HW, pp = n2 % HW;
        v = x[((size_t)img * C + c0 + c) * HW + pp];
      }
      Bs[r2][c] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 bq = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int d = d0 + ty * 4 + i;
    if (d >= D) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = c0 + tx * 4 + j;
      if (c < C) atomicAdd(&g_enc_w[(size_t)d * C + c], acc[i][j]);
    }
  }
}

// Same reduction with a per-block private table in shared memory: a block owns 32 channels and PRIV_PX pixels, adds every
// (pixel, j) row into table[j*M + idx][32] with shared-memory atomics (lane = channel = bank: conflict-free) and flushes the
// table to G once -- 67 M contended global atomics become N*k*C / PRIV_PX * ... ~ 8 M spread ones (523 -> ~60 us at b=64).
constexpr int PRIV_PX = 2048;
template <int K>
__global__ void __launch_bounds__(256) gdec_scatter_priv_kernel(const float* __restrict__ g_out,
                                                                 const int64_t* __restrict__ idx, float* __restrict__ G,
                                                                 float* __restrict__ g_dec_b, int N, int HW, int C, int M) {
  extern __shared__ float table[];             // [K*M][32]
  __shared__ float tile[32][33];               // [ch][px]
  __shared__ int idx_s[32][K];
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  const int c0 = blockIdx.y * 32;
  const int n_beg = blockIdx.x * PRIV_PX, n_end = min(N, n_beg + PRIV_PX);
  for (int e = threadIdx.x; e < K * M * 32; e += 256) table[e] = 0.f;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  __syncthreads();
  for (int n0 = n_beg; n0 < n_end; n0 += 32) {
    for (int e = threadIdx.x; e < 32 * K; e += 256) {
      const int p = e / K, j = e % K;
      idx_s[p][j] = (n0 + p < n_end) ? (int)idx[(size_t)(n0 + p) * K + j] : 0;
    }
    const int n = n0 + lane;
    const bool nvalid = n < n_end;
    const int img = nvalid ? n / HW : 0, pp = nvalid ? n % HW : 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int ch = wv * 4 + r, c = c0 + ch;
      const float v = (nvalid && c < C) ? g_out[((size_t)img * C + c) * HW + pp] : 0.f;
      tile[ch][lane] = v;
      bsum[r] += v;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int p = wv * 4 + r;
      if (n0 + p < n_end) {
        const float v = tile[lane][p];
#pragma unroll
        for (int j = 0; j < K; ++j) atomicAdd(&table[(j * M + idx_s[p][j]) * 32 + lane], v);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float s = warp_sum(bsum[r]);
    const int c = c0 + wv * 4 + r;
    if (lane == 0 && c < C) atomicAdd(&g_dec_b[c], s);
  }
  if (c0 + lane < C)
    for (int row = wv; row < K * M; row += 8) {
      const float v = table[row * 32 + lane];
      if (v != 0.f) atomicAdd(&G[(size_t)row * C + c0 + lane], v);
    }
}

// G[j][idx_j(n)][c] += g_out[img,c,p] ;  g_dec_b[c] += sum_n g_out
template <int K>
__global__ void __launch_bounds__(256) gdec_scatter_kernel(const float* __restrict__ g_out,
                                                            const int64_t* __restrict__ idx, float* __restrict__ G,
                                                            float* __restrict__ g_dec_b, int N, int HW, int C, int M,
                                                            int chunks_per_block) {
  __shared__ float tile[32][33];  // [ch][px]
  __shared__ int idx_s[32][K];
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  const int n0 = blockIdx.x * 32;
  for (int e = threadIdx.x; e < 32 * K; e += 256) {
    int p = e / K, j = e % K;
    idx_s[p][j] = (n0 + p < N) ? (int)idx[(size_t)(n0 + p) * K + j] : 0;
  }
  const int n = n0 + lane;
  const bool nvalid = n < N;
  const int img = nvalid ? n / HW : 0, pp = nvalid ? n % HW : 0;
  __syncthreads();
  const int cbeg = blockIdx.y * chunks_per_block * 32;
  for (int cc = 0; cc < chunks_per_block; ++cc) {
    const int c0 = cbeg + cc * 32;
    if (c0 >= C) break;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int ch = wv * 4 + r, c = c0 + ch;
      float v = (nvalid && c < C) ? g_out[((size_t)img * C + c) * HW + pp] : 0.f;
      tile[ch][lane] = v;
      float s = warp_sum(v);
      if (lane == 0 && c < C) atomicAdd(&g_dec_b[c], s);
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int p = wv * 4 + r;
      if (n0 + p < N && c0 + lane < C) {
        float v = tile[lane][p];
#pragma unroll
        for (int j = 0; j < K; ++j) atomicAdd(&G[((size_t)j * M + idx_s[p][j]) * C + c0 + lane], v);
      }
    }
    __syncthreads();
  }
}

// ---- operands of the tensor-core weight-gradient GEMMs (K = pixels) -------------------------------------------------
// g [b][C][HW] fp32 -> NHWC bf16 hi/lo planes [2][b][HW][C].  A block transposes a [64 ch][32 px] tile through shared
// memory: coalesced 128-byte row loads, then every thread emits 8 channels of one pixel as one 16-byte store per plane.
__global__ void __launch_bounds__(256) pack_nhwc64_kernel(const float* __restrict__ g, __nv_bfloat16* __restrict__ gp, int C,
                                                           int HW, long long plane_stride) {
  __shared__ float tile[64][33];
  const int img = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = ty; r < 64; r += 8) {
    const int c = c0 + r, pp = p0 + tx;
    tile[r][tx] = (c < C && pp < HW) ? g[((size_t)img * C + c) * HW + pp] : 0.f;
  }
  __syncthreads();
  const int px = threadIdx.x >> 3, cg = threadIdx.x & 7;         // pixel within the tile, 8-channel group
  const int pp = p0 + px, c = c0 + cg * 8;
  if (pp < HW && c < C) {                                        // C is a multiple of 8 on this path
    uint32_t hp[4], lp[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float v0 = tile[cg * 8 + 2 * j][px], v1 = tile[cg * 8 + 2 * j + 1][px];
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
      const __nv_bfloat16 l0 = __float2bfloat16_rn(v0 - __bfloat162float(h0));
      const __nv_bfloat16 l1 = __float2bfloat16_rn(v1 - __bfloat162float(h1));
      hp[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      lp[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    __nv_bfloat16* dst = gp + ((size_t)img * HW + pp) * C + c;
    *reinterpret_cast<uint4*>(dst) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
    *reinterpret_cast<uint4*>(dst + plane_stride) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
  }
}

// HW % 64 == 0, C % 64 == 0, 16-byte aligned x: 64 x 64 tiles, 16-byte accesses on both sides (256 contiguous bytes per
// channel row and pass on the NCHW side instead of 128)
// SUMS: the block also writes the sums of its 64 channels over its 64 pixels to part[img * gridDim.x + blockIdx.x][C]
// (16-lane shuffle reduction of the values already in registers) -- the channel sums of g (g_dec_b) without a second read of g.
template <bool SUMS>
__global__ void __launch_bounds__(256, 4) pack_nhwc64x64_kernel(const float* __restrict__ g, __nv_bfloat16* __restrict__ gp, int C,
                                                                 int HW, long long plane_stride, float* __restrict__ part) {
  __shared__ float tile[64][65];
  const int img = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float4 v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    v[i] = __ldg(reinterpret_cast<const float4*>(g + ((size_t)img * C + c0 + ty + 16 * i) * HW + p0 + 4 * tx));
  if (SUMS) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a = (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);      // the 16 lanes that share ty
      if (tx == 0) part[((size_t)img * gridDim.x + blockIdx.x) * C + c0 + ty + 16 * i] = a;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty + 16 * i;
    tile[r][4 * tx + 0] = v[i].x; tile[r][4 * tx + 1] = v[i].y; tile[r][4 * tx + 2] = v[i].z; tile[r][4 * tx + 3] = v[i].w;
  }
  __syncthreads();
  const int cg = threadIdx.x & 7;
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int px = (threadIdx.x >> 3) + 32 * hh;
    uint32_t hp[4], lp[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float v0 = tile[cg * 8 + 2 * j][px], v1 = tile[cg * 8 + 2 * j + 1][px];
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
      const __nv_bfloat16 l0 = __float2bfloat16_rn(v0 - __bfloat162float(h0));
      const __nv_bfloat16 l1 = __float2bfloat16_rn(v1 - __bfloat162float(h1));
      hp[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      lp[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    __nv_bfloat16* dst = gp + ((size_t)img * HW + p0 + px) * C + c0 + cg * 8;
    *reinterpret_cast<uint4*>(dst) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
    *reinterpret_cast<uint4*>(dst + plane_stride) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
  }
}

static inline bool pack64x64_ok(const float* x, int C, int HW) { return HW % 64 == 0 && C % 64 == 0 && ((uintptr_t)x & 15) == 0; }

// colsum[c] += sum over the rows of part [rows][C]; grid (C / 32, 8), 256 threads = 8 row lanes x 32 channels
__global__ void __launch_bounds__(256) colsum_rows_kernel(const float* __restrict__ part, float* __restrict__ colsum, int rows, int C) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), lane_r = threadIdx.x >> 5;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const int step = 8 * gridDim.y;
  int r = blockIdx.y * 8 + lane_r;
  for (; r + 3 * step < rows; r += 4 * step) {
    a0 += part[(size_t)r * C + c]; a1 += part[(size_t)(r + step) * C + c];
    a2 += part[(size_t)(r + 2 * step) * C + c]; a3 += part[(size_t)(r + 3 * step) * C + c];
  }
  for (; r < rows; r += step) a0 += part[(size_t)r * C + c];
  red[lane_r][threadIdx.x & 31] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (lane_r == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    atomicAdd(&colsum[c], t);
  }
}

// pack + channel sums of g in one read of g (64 x 64-tile shapes only: the caller checks pack64x64_ok); colsum must be zeroed
static int pack_nhwc64_colsum(const float* g, void* gp, float* part, float* colsum, int b, int C, int HW, cudaStream_t st) {
  pack_nhwc64x64_kernel<true><<<dim3(HW / 64, C / 64, b), 256, 0, st>>>(g, (__nv_bfloat16*)gp, C, HW, (long long)b * HW * C, part);
  AMMC_LAUNCH_CHECK("pack_nhwc64x64_kernel");
  colsum_rows_kernel<<<dim3(C / 32, 8), 256, 0, st>>>(part, colsum, b * (HW / 64), C);
  AMMC_LAUNCH_CHECK("colsum_rows_kernel");
  return 0;
}

int pack_nhwc64(const float* x, void* xp, int b, int C, int HW, cudaStream_t st) {     // C % 8 == 0
  if (pack64x64_ok(x, C, HW))
    pack_nhwc64x64_kernel<false><<<dim3(HW / 64, C / 64, b), 256, 0, st>>>(x, (__nv_bfloat16*)xp, C, HW, (long long)b * HW * C,
                                                                          nullptr);
  else
    pack_nhwc64_kernel<<<dim3(ceil_div(HW, 32), ceil_div(C, 64), b), 256, 0, st>>>(x, (__nv_bfloat16*)xp, C, HW,
                                                                                    (long long)b * HW * C);
  AMMC_LAUNCH_CHECK("pack_nhwc64_kernel");
  return 0;
}

// colsum[c] += sum over images and pixels of g [b][C][HW]  (g_dec_b): one block per (channel, image group)
__global__ void __launch_bounds__(256) channel_sum_kernel(const float* __restrict__ g, float* __restrict__ colsum, int b,
                                                           int C, int HW, int imgs_per_block) {
  __shared__ float red[33];
  const int c = blockIdx.x;
  const int i0 = blockIdx.y * imgs_per_block, i1 = min(b, i0 + imgs_per_block);
  float acc = 0.f;
  for (int img = i0; img < i1; ++img) {
    const float* row = g + ((size_t)img * C + c) * HW;
    for (int i = threadIdx.x; i < HW; i += 256) acc += row[i];
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(&colsum[c], acc);
}

// read[n][j*D + d] = bank_t[idx[n][j]][d] as bf16 hi/lo planes [2][N][k*D] (what the forward's refine kernel wrote)
__global__ void __launch_bounds__(256) read_planes_kernel(const int64_t* __restrict__ idx, const float* __restrict__ bank_t,
                                                           __nv_bfloat16* __restrict__ rp, long long N, int D, int k,
                                                           long long plane_stride) {
  const long long total = N * k * (D / 4);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int d4 = (int)(e % (D / 4));
    const long long nj = e / (D / 4);
    const float4 v = __ldg(reinterpret_cast<const float4*>(bank_t + (size_t)idx[nj] * D) + d4);
    const float vv[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      h[i] = __float2bfloat16_rn(vv[i]);
      l[i] = __float2bfloat16_rn(vv[i] - __bfloat162float(h[i]));
    }
    uint2 ph, pl;
    ph.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
    ph.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
    pl.x = (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16);
    pl.y = (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16);
    __nv_bfloat16* dst = rp + (size_t)nj * D + 4 * d4;
    *reinterpret_cast<uint2*>(dst) = ph;
    *reinterpret_cast<uint2*>(dst + plane_stride) = pl;
  }
}

// g_dec_w[c][j*D + d] = sum_m G[j][m][c] * embed[d][m]
__global__ void gdec_w_kernel(const float* __restrict__ G, const float* __restrict__ embed,
                              float* __restrict__ g_dec_w, int C, int D, int M, int k) {
  __shared__ float Gs[32][33];  // [m][c]
  __shared__ float Es[32][33];  // [d][m]
  const int d0 = blockIdx.x * 32, c0 = blockIdx.y * 32, j = blockIdx.z;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int m0 = 0; m0 < M; m0 += 32) {
    for (int r = ty; r < 32; r += 8) {
      int m = m0 + r, c = c0 + tx;
      Gs[r][tx] = (m < M && c < C) ? G[((size_t)j * M + m) * C + c] : 0.f;
      int d = d0 + r, mm = m0 + tx;
      Es[r][tx] = (d < D && mm < M) ? embed[(size_t)d * M + mm] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int m = 0; m < 32; ++m) {
      float ev = Es[tx][m];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(Gs[m][ty * 4 + i], ev, acc[i]);
    }
    __syncthreads();
  }
  int d = d0 + tx;
  if (d < D) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int c = c0 + ty * 4 + i;
      if (c < C) g_dec_w[(size_t)c * (k * D) + (size_t)j * D + d] = acc[i];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct MemWs {
  int* stats; float* bank_t; float* en2; float* T; float* sse_px; __nv_bfloat16* read_planes;
  bool own_bank;                                    // bank_t / en2 live in the workspace and are (re)derived per call
  const __nv_bfloat16* zp; const float* znorm2;     // produced by the tensor-core enc epilogue, else null
};


// addr_tc.cu
size_t addr_tc_ws_bytes(int64_t N, int D, int M);
bool addr_tc_supported(int64_t N, int D, int M, int k);
int run_address_tc(const float* z, const void* zp_in, const float* zmeta_in, __nv_bfloat16* read_planes,
                   const float* bank_t, const float* en2, float* read, float* q1, int64_t* idx, float* sse_px,
                   float* counts, float* embed_sum, int* stats, Workspace& ws, int64_t N, int D, int M, int k,
                   cudaStream_t st);
int run_rescan(const float* z, const float* bank_t, const float* en2, float* q1, int64_t* idx, float* sse_px, int* stats,
               int* rescan_list, __nv_bfloat16* read_planes, long long read_plane_stride, int D, int M, int k, cudaStream_t st);
int pack_bank_padded(const float* bank_t, const float* en2, void* bank_hi, float* en2pad, float* emax, int D, int M, int Mpad,
                     cudaStream_t st);
// mem_front.cu
bool mem_front_supported(int b, int HW, int C, int D, int M, int k);
int run_mem_front(const void* x, int x_bf16, const void* enc_wp, const float* enc_b, const void* bank_hi, const float* bank_t,
                  const float* en2, const float* en2pad, const float* emax, float* z, float* q1, int64_t* idx, float* sse_px,
                  __nv_bfloat16* read_planes, int* stats, int* rescan_list, unsigned* amax_bits, int b, int HW, int C, int M,
                  int k, cudaStream_t st);
// enc_tc.cu
bool enc_tc_supported(int b, int HW, int C, int D);
size_t enc_tc_ws_bytes(int C);
int run_enc_tc(const float* x, const float* enc_w, const float* enc_b, float* z, __nv_bfloat16* zp, float* znorm2,
               void* wp_ws, bool wp_ready, unsigned* amax_bits, int b, int HW, int C, cudaStream_t st);

// amft_conv.cu
bool conv_shape_supported(int b, int Cin, int Cout, int h, int w);
int pack_planes_f32(const float* x, void* xp, long long n, cudaStream_t st);
int conv_wgrad_run(const void* gy_nhwc_planes, const void* x_nhwc_planes, float* gw, int b, int Cin, int Cout, int h, int w,
                   int ntaps, int precision, cudaStream_t st);
int pack_weights_1x1(const float* w, void* wp, int Cout, int Cin, cudaStream_t st);
int conv_igemm(const void* xp, const void* wp, const float* scale, const float* shift, void* out_planes,
               float* out_nchw, const float* res_nchw, int b, int Cin, int Cout, int h, int w, int ntaps, int precision,
               int relu, cudaStream_t st);
int conv_run(const ammc_conv_layer& L, cudaStream_t st);
int absmax_f32(const float* x, long long n, unsigned* out_bits, cudaStream_t st);

// Power-of-two scale of the module output written as q planes (precision-2 operand of the AMFT block), from a rigorous
// bound:  |out_c| <= max|x| (residual) + sum_j ||dec_w[c, jD:(j+1)D]||_2 * max_m ||e_m||_2 + |dec_b[c]|   (Cauchy-Schwarz)
// One warp per output channel (a single block walking all C channels was latency-bound: 150 us); the per-channel bounds
// meet in an atomicMax on the bit pattern, a one-thread kernel turns the sum into the scale.
__global__ void __launch_bounds__(256) dec_bound_kernel(const float* __restrict__ dec_w, const float* __restrict__ dec_b,
                                                        const float* __restrict__ en2, int C, int D, int M, int k,
                                                        unsigned* __restrict__ bound_bits) {
  __shared__ float red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float e2 = 0.f;
  for (int m = threadIdx.x; m < M; m += blockDim.x) e2 = fmaxf(e2, en2[m]);
  e2 = warp_max(e2);
  if (lane == 0) red[warp] = e2;
  __syncthreads();
  e2 = red[0];
  for (int i = 1; i < 8; ++i) e2 = fmaxf(e2, red[i]);
  const float emax = sqrtf(e2);
  const int c = blockIdx.x * 8 + warp;
  if (c >= C) return;
  float bound = 0.f;
  for (int j = 0; j < k; ++j) {
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) { const float wv = dec_w[(size_t)c * k * D + j * D + d]; ss = fmaf(wv, wv, ss); }
    bound += sqrtf(warp_sum(ss));
  }
  if (lane == 0) atomicMax(bound_bits, __float_as_uint(bound * emax * 1.01f + fabsf(dec_b[c])));
}

// Everything the forward derives from the parameters alone (constant between optimizer / EMA steps): packed once by
// ammc_mem_prepare into a caller-owned buffer and reused by every ammc_mem_fwd until a parameter changes.
struct MemPrep {
  __nv_bfloat16* enc_wp;     // [2][D][C]   bf16 hi/lo planes of enc.weight
  __nv_bfloat16* dec_wp;     // [2][C][kD]  bf16 hi/lo planes of dec.weight
  float* bank_t;             // [M][D]      items as rows
  float* en2;                // [M]         ||e||^2
  __nv_bfloat16* bank_hi;    // [256][D]    fp16 bits of the scaled items, zero rows beyond M (fused front kernel)
  float* en2pad;             // [256]       ||e||^2, +inf beyond M
  float* emax;               // [4]         max ||e||, 1 / t (inverse power-of-two fp16 scale of the bank), scratch
  float* ones;               // [C]
  unsigned* dec_bound;       // [1]         bits of max_c (sum_j ||dec_w[c,j]|| max||e|| + |dec_b[c]|): q-plane scale bound
};
static size_t prep_bytes(int C, int D, int M, int k) {
  return align_up((size_t)2 * D * C * 2, 256) + align_up((size_t)2 * C * k * D * 2, 256) + align_up((size_t)M * D * 4, 256) +
         align_up((size_t)M * 4, 256) + align_up((size_t)256 * D * 2, 256) + align_up((size_t)256 * 4, 256) + 256 +
         align_up((size_t)C * 4, 256) + 256;
}
static int carve_prep(Workspace& ws, MemPrep& pr, int C, int D, int M, int k) {
  pr.enc_wp = ws.take<__nv_bfloat16>((size_t)2 * D * C);
  pr.dec_wp = ws.take<__nv_bfloat16>((size_t)2 * C * k * D);
  pr.bank_t = ws.take<float>((size_t)M * D);
  pr.en2 = ws.take<float>(M);
  pr.bank_hi = ws.take<__nv_bfloat16>((size_t)256 * D);
  pr.en2pad = ws.take<float>(256);
  pr.emax = ws.take<float>(4);          // max ||e||, 1 / fp16 scale of the bank, scratch
  pr.ones = ws.take<float>(C);
  pr.dec_bound = ws.take<unsigned>(1);
  if (!ws.ok()) return fail(AMMC_EWORKSPACE, "prepared-parameter buffer too small");
  return 0;
}

static int g_enc_mode = 0;
static int g_front_mode = 1;  // 1: fused enc + addressing kernel (mem_front.cu) when the shape allows; 0: staged kernels    // 0 auto, 1 fp32 FFMA (CUDA cores), 2 tensor-core GEMM (split-bf16 x3, converts NCHW on the fly)
static int g_dec_mode = 0;    // 0 auto, 1 fp32 table gather (CUDA cores), 2 tensor-core GEMM (split-bf16 x3)
static int g_addr_mode = 0;   // 0 auto, 1 generic fp32 (CUDA cores), 2 tensor-core filter + exact refine

static bool use_tc(int64_t N, int D, int M, int k) {
  if (g_addr_mode == 1) return false;
  return addr_tc_supported(N, D, M, k);
}

__global__ void fill_kernel(float* p, float v, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void fill2_kernel(float* p0, float v0, float* p1, float v1, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { p0[i] = v0; p1[i] = v1; }
}

static bool use_tc_dec(int b, int h, int w, int C, int D, int k) {
  if (g_dec_mode == 1) return false;
  return ((k * D) % 64 == 0) && (D % 4 == 0) && conv_shape_supported(b, k * D, C, h, w);
}

static size_t mem_ws_bytes(int64_t N, int C, int D, int M, int k, bool with_table) {
  size_t s = 256;                                   // stats block: [0] exact-fallback rows, [1] path (1 fp32, 2 tensor)
  s += addr_tc_ws_bytes(N, D, M);
  s += align_up((size_t)M * D * 4, 256);
  s += align_up((size_t)M * 4, 256);
  if (with_table) {
    s += prep_bytes(C, D, M, k) + align_up((size_t)N * 4, 256);      // prepared parameters (when not cached), re-scan list
    s += align_up((size_t)k * M * C * 4, 256);                       // dec tables (fp32 gather path)
    s += align_up((size_t)N * k * D * 2 * 2, 256);                    // read planes (tensor-core dec)
    s += align_up((size_t)C * k * D * 2 * 2, 256) + align_up((size_t)C * 4, 256);
    s += enc_tc_ws_bytes(C) + align_up((size_t)N * D * 2, 256) + align_up((size_t)N * 8, 256);   // tensor-core enc
  }
  s += align_up((size_t)N * 4, 256);
  return s;
}

static int carve(Workspace& ws, MemWs& m, int64_t N, int C, int D, int M, int k, bool with_table) {
  m.stats = ws.take<int>(64);
  m.read_planes = nullptr;
  m.zp = nullptr;
  m.znorm2 = nullptr;
  m.own_bank = !with_table;                         // the module path takes both from the prepared-parameter buffer
  m.bank_t = with_table ? nullptr : ws.take<float>((size_t)M * D);
  m.en2 = with_table ? nullptr : ws.take<float>(M);
  m.T = with_table ? ws.take<float>((size_t)k * M * C) : nullptr;
  m.sse_px = ws.take<float>(N);
  if (!ws.ok()) return fail(AMMC_EWORKSPACE, "workspace too small");
  return 0;
}

template <int K>
static void launch_address(const float* z, const float* embed, const MemWs& m, float* read, float* q1, int64_t* idx,
                           float* counts, float* embed_sum, int N, int D, int M, cudaStream_t st) {
  address_kernel<K><<<ceil_div(N, 64), 256, 0, st>>>(z, embed, m.en2, m.bank_t, read, q1, idx, m.sse_px, counts,
                                                      embed_sum, m.read_planes, N, D, M);
}

static int run_address(const float* z, const float* embed, const MemWs& m, float* read, float* q1, int64_t* idx,
                       float* counts, float* embed_sum, int64_t N, int D, int M, int k, cudaStream_t st, Workspace& ws) {
  AMMC_CUDA_CHECK(cudaMemsetAsync(m.stats, 0, 32, st));   // words 0-7; word 8 = max|x| bits of the enc pass
  if (m.own_bank) {                                       // Quantize_topk on its own: items as rows + norms derived here
    bank_transpose_kernel<<<dim3(ceil_div(M, 32), ceil_div(D, 32)), dim3(32, 8), 0, st>>>(embed, m.bank_t, D, M);
    AMMC_LAUNCH_CHECK("bank_transpose_kernel");
    bank_norms_kernel<<<ceil_div(M, 32), dim3(32, 8), 0, st>>>(embed, m.en2, D, M);
    AMMC_LAUNCH_CHECK("bank_norms_kernel");
  }
  if (counts) AMMC_CUDA_CHECK(cudaMemsetAsync(counts, 0, (size_t)M * 4, st));
  if (embed_sum) AMMC_CUDA_CHECK(cudaMemsetAsync(embed_sum, 0, (size_t)D * M * 4, st));
  if (g_addr_mode == 2 && !addr_tc_supported(N, D, M, k))
    return fail(AMMC_EUNSUPPORTED, "tensor-core addressing needs D %% 64 == 0, M >= 16, k <= 4 (got D=%d M=%d k=%d)", D, M, k);
  {
    const int path = use_tc(N, D, M, k) ? 2 : 1;
    AMMC_CUDA_CHECK(cudaMemsetAsync(m.stats + 1, path, 1, st));   // low byte of stats[1] (rest zeroed above)
  }
  if (use_tc(N, D, M, k))
    return run_address_tc(z, m.zp, m.znorm2, m.read_planes, m.bank_t, m.en2, read, q1, idx, m.sse_px, counts, embed_sum,
                          m.stats, ws, N, D, M, k, st);
  switch (k) {
    case 1: launch_address<1>(z, embed, m, read, q1, idx, counts, embed_sum, (int)N, D, M, st); break;
    case 2: launch_address<2>(z, embed, m, read, q1, idx, counts, embed_sum, (int)N, D, M, st); break;
    case 3: launch_address<3>(z, embed, m, read, q1, idx, counts, embed_sum, (int)N, D, M, st); break;
    case 4: launch_address<4>(z, embed, m, read, q1, idx, counts, embed_sum, (int)N, D, M, st); break;
    case 5: launch_address<5>(z, embed, m, read, q1, idx, counts, embed_sum, (int)N, D, M, st); break;
    case 6: launch_address<6>(z, embed, m, read, q1, idx, counts, embed_sum, (int)N, D, M, st); break;
    case 7: launch_address<7>(z, embed, m, read, q1, idx, counts, embed_sum, (int)N, D, M, st); break;
    case 8: launch_address<8>(z, embed, m, read, q1, idx, counts, embed_sum, (int)N, D, M, st); break;
    default: return fail(AMMC_EUNSUPPORTED, "k=%d outside 1..%d", k, AMMC_MAX_K);
  }
  AMMC_LAUNCH_CHECK("address_kernel");
  return 0;
}

static int run_commit(const MemWs& m, float* sse_frame, float* diff, int64_t N, int64_t rows_per_frame, int D,
                      cudaStream_t st, const unsigned* amax_bits = nullptr, const unsigned* bound_bits = nullptr,
                      int residual = 0, float* qs = nullptr) {
  int frames = (int)(N / rows_per_frame);
  unsigned* ticket = reinterpret_cast<unsigned*>(m.stats + 3);       // word 3 of the stats block (zeroed with it)
  sse_frame_kernel<<<frames, 256, 0, st>>>(m.sse_px, sse_frame, rows_per_frame, diff, 1.0 / ((double)N * (double)D), ticket,
                                           amax_bits, bound_bits, residual, qs);
  AMMC_LAUNCH_CHECK("sse_frame_kernel");
  return 0;
}

}  // namespace ammc

using namespace ammc;

extern "C" int ammc_set_addressing_mode(int mode) {
  AMMC_REQUIRE(mode >= 0 && mode <= 2, "addressing mode must be 0 (auto), 1 (fp32 CUDA-core) or 2 (tensor-core)");
  g_addr_mode = mode;
  return 0;
}

extern "C" size_t ammc_mem_workspace_bytes(int b, int h, int w, int C, int D, int M, int k) {
  return mem_ws_bytes((int64_t)b * h * w, C, D, M, k, true);
}

extern "C" size_t ammc_quantize_workspace_bytes(int64_t N, int D, int M, int k) {
  return mem_ws_bytes(N, 0, D, M, k, false);
}

static int check_dims(int64_t N, int C, int D, int M, int k) {
  AMMC_REQUIRE(N > 0 && N < (1LL << 31), "N=%lld out of range", (long long)N);
  AMMC_REQUIRE(D > 0 && M > 0 && C >= 0, "bad dims C=%d D=%d M=%d", C, D, M);
  if (k < 1 || k > AMMC_MAX_K) return fail(AMMC_EUNSUPPORTED, "k=%d outside 1..%d", k, AMMC_MAX_K);
  AMMC_REQUIRE(k <= M, "k=%d exceeds n_embed=%d", k, M);  // torch.topk raises for k > M as well
  return 0;
}

// Derive everything that depends on the parameters only (see MemPrep).
static int run_prepare(const MemPrep& pr, const float* enc_w, const float* embed, const float* dec_w, const float* dec_b,
                       int C, int D, int M, int k, bool tc_enc, bool tc_dec, cudaStream_t st) {
  bank_transpose_kernel<<<dim3(ceil_div(M, 32), ceil_div(D, 32)), dim3(32, 8), 0, st>>>(embed, pr.bank_t, D, M);
  AMMC_LAUNCH_CHECK("bank_transpose_kernel");
  bank_norms_kernel<<<ceil_div(M, 32), dim3(32, 8), 0, st>>>(embed, pr.en2, D, M);
  AMMC_LAUNCH_CHECK("bank_norms_kernel");
  if (M <= 256)
    if (int rc = pack_bank_padded(pr.bank_t, pr.en2, pr.bank_hi, pr.en2pad, pr.emax, D, M, 256, st)) return rc;
  if (tc_enc)
    if (int rc = pack_weights_1x1(enc_w, pr.enc_wp, D, C, st)) return rc;
  if (tc_dec) {
    if (int rc = pack_weights_1x1(dec_w, pr.dec_wp, C, k * D, st)) return rc;
    fill_kernel<<<ceil_div(C, 256), 256, 0, st>>>(pr.ones, 1.f, C);
    AMMC_LAUNCH_CHECK("fill_kernel");
  }
  AMMC_CUDA_CHECK(cudaMemsetAsync(pr.dec_bound, 0, 4, st));
  dec_bound_kernel<<<ceil_div(C, 8), 256, 0, st>>>(dec_w, dec_b, pr.en2, C, D, M, k, pr.dec_bound);
  AMMC_LAUNCH_CHECK("dec_bound_kernel");
  return 0;
}

extern "C" size_t ammc_mem_prep_bytes(int C, int D, int M, int k) { return prep_bytes(C, D, M, k); }

extern "C" int ammc_mem_prepare(const float* enc_w, const float* embed, const float* dec_w, const float* dec_b, void* prep,
                                size_t prep_bytes_given, int b, int h, int w, int C, int D, int M, int k, void* stream) {
  AMMC_REQUIRE(enc_w && embed && dec_w && dec_b && prep, "null pointer argument");
  if (int rc = check_dims((int64_t)b * h * w, C, D, M, k)) return rc;
  Workspace ws(prep, prep_bytes_given);
  MemPrep pr;
  if (int rc = carve_prep(ws, pr, C, D, M, k)) return rc;
  const bool tc_dec = use_tc_dec(b, h, w, C, D, k);
  const bool tc_enc = g_enc_mode != 1 && enc_tc_supported(b, h * w, C, D);
  return run_prepare(pr, enc_w, embed, dec_w, dec_b, C, D, M, k, tc_enc, tc_dec, (cudaStream_t)stream);
}

extern "C" int ammc_set_front_mode(int on) {
  g_front_mode = on ? 1 : 0;
  return 0;
}

// io16: x_any / out_any are bf16 NCHW (ammc_mem_fwd_io16), else fp32
static int mem_fwd_impl(const void* x_any, const float* enc_w, const float* enc_b, const float* embed,
                        const float* dec_w, const float* dec_b, void* out_any, float* q1, int64_t* idx, float* z,
                        float* sse_frame, float* diff, float* counts, float* embed_sum, void* out_planes,
                        int planes_fmt, const void* prep, void* workspace, size_t workspace_bytes, int b, int h, int w,
                        int C, int D, int M, int k, int residual, int io16, void* stream) {
  const float* x = reinterpret_cast<const float*>(x_any);     // only dereferenced as fp32 when io16 == 0
  float* out = reinterpret_cast<float*>(out_any);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t N = (int64_t)b * h * w;
  const int HW = h * w;
  if (int rc = check_dims(N, C, D, M, k)) return rc;
  AMMC_REQUIRE(x && enc_w && enc_b && embed && dec_w && dec_b && out && q1 && idx && z && sse_frame && diff,
               "null pointer argument");
  Workspace ws(workspace, workspace_bytes);
  MemWs m;
  if (int rc = carve(ws, m, N, C, D, M, k, true)) return rc;
  const bool tc_dec = use_tc_dec(b, h, w, C, D, k);
  if (g_dec_mode == 2 && !tc_dec)
    return fail(AMMC_EUNSUPPORTED, "tensor-core dec needs k*D %% 64 == 0, C %% 64 == 0 and a feature map the conv engine tiles");
  AMMC_REQUIRE(!out_planes || tc_dec, "out_planes requested but the tensor-core dec path is not in use "
                                      "(query ammc_mem_dec_uses_tensor first)");
  AMMC_REQUIRE(planes_fmt == 0 || planes_fmt == 1, "planes_fmt must be 0 (bf16 hi/lo) or 1 (q)");
  const bool q_planes = out_planes && planes_fmt == 1;
  AMMC_REQUIRE(!q_planes || C % 256 == 0, "q-format planes need C %% 256 == 0");
  unsigned* amax_bits = reinterpret_cast<unsigned*>(m.stats + 8);     // word 8 of the stats block
  const bool tc_enc = g_enc_mode != 1 && enc_tc_supported(b, HW, C, D);
  if (g_enc_mode == 2 && !tc_enc)
    return fail(AMMC_EUNSUPPORTED, "tensor-core enc needs embed_dim == 64, C %% 64 == 0 and h*w %% 128 == 0");
  // parameters-only quantities: from the caller's prepared buffer, else derived here into the workspace
  MemPrep pr;
  if (prep) {
    Workspace pws(const_cast<void*>(prep), prep_bytes(C, D, M, k));
    if (int rc = carve_prep(pws, pr, C, D, M, k)) return rc;
  } else {
    if (int rc = carve_prep(ws, pr, C, D, M, k)) return rc;
    if (int rc = run_prepare(pr, enc_w, embed, dec_w, dec_b, C, D, M, k, tc_enc, tc_dec, st)) return rc;
  }
  m.bank_t = pr.bank_t;
  m.en2 = pr.en2;
  if (tc_dec) {
    m.read_planes = ws.take<__nv_bfloat16>((size_t)N * k * D * 2);
    if (!ws.ok()) return fail(AMMC_EWORKSPACE, "workspace too small");
  } else {
    dec_table_kernel<<<dim3(ceil_div(C, 32), ceil_div(M, 32), k), dim3(32, 8), 0, st>>>(dec_w, embed, m.T, C, D, M, k);
    AMMC_LAUNCH_CHECK("dec_table_kernel");
  }
  // fused enc + addressing (eval): z, candidates and scores never leave the SM
  const bool front = g_front_mode && tc_enc && tc_dec && !counts && !embed_sum && g_addr_mode != 1 &&
                     mem_front_supported(b, HW, C, D, M, k);
  if (io16 && !front)
    return fail(AMMC_EUNSUPPORTED, "bf16 feature I/O runs on the fused front kernel + tensor-core dec only "
                                   "(embed_dim 64, n_embed <= 256, h*w %% 128 == 0, eval); widen the input instead");
  if (front) {
    int* rescan_list = ws.take<int>(N);
    if (!ws.ok()) return fail(AMMC_EWORKSPACE, "workspace too small");
    AMMC_CUDA_CHECK(cudaMemsetAsync(m.stats, 0, 32, st));
    AMMC_CUDA_CHECK(cudaMemsetAsync(m.stats + 1, 3, 1, st));            // path id 3: fused front kernel
    if (int rc = run_mem_front(x_any, io16, pr.enc_wp, enc_b, pr.bank_hi, pr.bank_t, pr.en2, pr.en2pad, pr.emax, z, q1, idx, m.sse_px,
                               m.read_planes, m.stats, rescan_list, (q_planes && residual) ? amax_bits : nullptr, b, HW, C,
                               M, k, st))
      return rc;
    if (int rc = run_rescan(z, pr.bank_t, pr.en2, q1, idx, m.sse_px, m.stats, rescan_list, m.read_planes,
                            (long long)N * k * D, D, M, k, st))
      return rc;
  } else {
    if (tc_enc) {
      __nv_bfloat16* zp = ws.take<__nv_bfloat16>((size_t)N * D);
      float* zn2 = ws.take<float>(2 * N);                // (||z||^2, 1 / row scale) pairs
      if (!ws.ok()) return fail(AMMC_EWORKSPACE, "workspace too small");
      // the converters see every input value: max|x| (for the q planes' scale) comes out of the same pass
      if (int rc = run_enc_tc(x, enc_w, enc_b, z, zp, zn2, pr.enc_wp, true, (q_planes && residual) ? amax_bits : nullptr, b,
                              HW, C, st))
        return rc;
      m.zp = zp;
      m.znorm2 = zn2;
    } else {
      enc1x1_kernel<<<dim3(ceil_div(N, 64), ceil_div(D, 64)), 256, 0, st>>>(x, enc_w, enc_b, z, nullptr, (int)N, HW, C, D);
      AMMC_LAUNCH_CHECK("enc1x1_kernel");
      if (q_planes && residual)
        if (int rc = absmax_f32(x, (long long)N * C, amax_bits, st)) return rc;
    }
    if (int rc = run_address(z, embed, m, nullptr, q1, idx, counts, embed_sum, N, D, M, k, st, ws)) return rc;
  }
  {
    float* qs = q_planes ? reinterpret_cast<float*>((uint8_t*)out_planes + 4 * N * C) : nullptr;   // scale slot of the q buffer
    if (int rc = run_commit(m, sse_frame, diff, N, HW, D, st, amax_bits, pr.dec_bound, residual, qs)) return rc;
  }
  const float* res = residual ? x : nullptr;
  if (tc_dec) {
    // dec(read) as a [N, kD] x [kD, C] GEMM on tcgen05 (split-bf16 x3); bias, residual and -- when asked -- the NHWC
    // operand planes of `out` (the AMFT block's input) all come out of the same epilogue
    ammc_conv_layer L = {};
    L.in_planes = m.read_planes; L.wp = pr.dec_wp; L.taps = 1; L.scale = pr.ones; L.shift = dec_b; L.act = 0;
    L.out_planes = out_planes; L.out_nchw = out; L.res_nchw = res;
    L.b = b; L.h = h; L.w = w; L.Cin = k * D; L.Cout = C; L.precision = 3;
    L.io_bf16 = io16;                   // residual x and `out` as bf16 NCHW
    if (q_planes) L.out_fmt = 1;        // its scale was written by the commit kernel above
    return conv_run(L, st);
  }
  const int chunks = 4;
  dim3 grid(ceil_div(N, 32), ceil_div(ceil_div(C, 32), chunks));
  switch (k) {
#define AMMC_DEC_CASE(KK) \
  case KK: dec_gather_kernel<KK><<<grid, 256, 0, st>>>(m.T, idx, dec_b, res, out, (int)N, HW, C, M, chunks); break;
    AMMC_DEC_CASE(1) AMMC_DEC_CASE(2) AMMC_DEC_CASE(3) AMMC_DEC_CASE(4)
    AMMC_DEC_CASE(5) AMMC_DEC_CASE(6) AMMC_DEC_CASE(7) AMMC_DEC_CASE(8)
#undef AMMC_DEC_CASE
  }
  AMMC_LAUNCH_CHECK("dec_gather_kernel");
  return 0;
}

extern "C" int ammc_mem_fwd(const float* x, const float* enc_w, const float* enc_b, const float* embed,
                            const float* dec_w, const float* dec_b, float* out, float* q1, int64_t* idx, float* z,
                            float* sse_frame, float* diff, float* counts, float* embed_sum, void* out_planes,
                            int planes_fmt, const void* prep, void* workspace, size_t workspace_bytes, int b, int h, int w,
                            int C, int D, int M, int k, int residual, void* stream) {
  return mem_fwd_impl(x, enc_w, enc_b, embed, dec_w, dec_b, out, q1, idx, z, sse_frame, diff, counts, embed_sum, out_planes,
                      planes_fmt, prep, workspace, workspace_bytes, b, h, w, C, D, M, k, residual, 0, stream);
}

extern "C" int ammc_mem_io16_supported(int b, int h, int w, int C, int D, int M, int k) {
  if (b <= 0 || h <= 0 || w <= 0 || C <= 0 || D <= 0 || M <= 0 || k <= 0) return 0;
  const bool tc_enc = g_enc_mode != 1 && enc_tc_supported(b, h * w, C, D);
  return (g_front_mode && tc_enc && use_tc_dec(b, h, w, C, D, k) && g_addr_mode != 1 && C % 256 == 0 &&
          mem_front_supported(b, h * w, C, D, M, k)) ? 1 : 0;
}

extern "C" int ammc_mem_fwd_io16(const void* x_bf16, const float* enc_w, const float* enc_b, const float* embed,
                                 const float* dec_w, const float* dec_b, void* out_bf16, float* q1, int64_t* idx, float* z,
                                 float* sse_frame, float* diff, void* out_planes, int planes_fmt, const void* prep,
                                 void* workspace, size_t workspace_bytes, int b, int h, int w, int C, int D, int M, int k,
                                 int residual, void* stream) {
  if (!ammc_mem_io16_supported(b, h, w, C, D, M, k))
    return fail(AMMC_EUNSUPPORTED, "bf16 feature I/O needs embed_dim 64, n_embed <= 256, C %% 256 == 0, h*w %% 128 == 0 "
                                   "(got C=%d D=%d M=%d h*w=%d); widen the input and call ammc_mem_fwd", C, D, M, h * w);
  return mem_fwd_impl(x_bf16, enc_w, enc_b, embed, dec_w, dec_b, out_bf16, q1, idx, z, sse_frame, diff, nullptr, nullptr,
                      out_planes, planes_fmt, prep, workspace, workspace_bytes, b, h, w, C, D, M, k, residual, 1, stream);
}

extern "C" int ammc_mem_dec_uses_tensor(int b, int h, int w, int C, int D, int M, int k) {
  (void)M;
  return use_tc_dec(b, h, w, C, D, k) ? 1 : 0;
}

extern "C" int ammc_set_enc_mode(int mode) {
  AMMC_REQUIRE(mode >= 0 && mode <= 2, "enc mode must be 0 (auto), 1 (fp32 FFMA) or 2 (tensor-core)");
  g_enc_mode = mode;
  return 0;
}

extern "C" int ammc_set_dec_mode(int mode) {
  AMMC_REQUIRE(mode >= 0 && mode <= 2, "dec mode must be 0 (auto), 1 (fp32 table gather) or 2 (tensor-core)");
  g_dec_mode = mode;
  return 0;
}

extern "C" int ammc_quantize_fwd(const float* z, const float* embed, float* read, float* q1, int64_t* idx,
                                 float* sse_frame, float* diff, float* counts, float* embed_sum, void* workspace,
                                 size_t workspace_bytes, int64_t N, int64_t rows_per_frame, int D, int M, int k,
                                 void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = check_dims(N, 0, D, M, k)) return rc;
  AMMC_REQUIRE(z && embed && read && q1 && idx && sse_frame && diff, "null pointer argument");
  AMMC_REQUIRE(rows_per_frame > 0 && N % rows_per_frame == 0, "N=%lld not a multiple of rows_per_frame=%lld",
               (long long)N, (long long)rows_per_frame);
  Workspace ws(workspace, workspace_bytes);
  MemWs m;
  if (int rc = carve(ws, m, N, 0, D, M, k, false)) return rc;
  if (int rc = run_address(z, embed, m, read, q1, idx, counts, embed_sum, N, D, M, k, st, ws)) return rc;
  return run_commit(m, sse_frame, diff, N, rows_per_frame, D, st);
}

extern "C" size_t ammc_quantize_bwd_workspace_bytes(int64_t N, int D, int M) {
  (void)N;
  return align_up((size_t)M * D * 4, 256) + align_up((size_t)D * 4, 256);
}

extern "C" int ammc_quantize_bwd(const float* z, const float* embed, const int64_t* idx, const float* g_diff,
                                 const float* g_q1, float* gz, void* workspace, size_t workspace_bytes, int64_t N,
                                 int D, int M, int k, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = check_dims(N, 0, D, M, k)) return rc;
  AMMC_REQUIRE(z && embed && idx && g_diff && gz, "null pointer argument");
  AMMC_REQUIRE(D <= 8192, "embed_dim %d too large for the backward column-sum buffer", D);
  Workspace ws(workspace, workspace_bytes);
  float* bank_t = ws.take<float>((size_t)M * D);
  float* colsum = ws.take<float>(D);
  if (!ws.ok()) return fail(AMMC_EWORKSPACE, "workspace too small");
  bank_transpose_kernel<<<dim3(ceil_div(M, 32), ceil_div(D, 32)), dim3(32, 8), 0, st>>>(embed, bank_t, D, M);
  AMMC_LAUNCH_CHECK("bank_transpose_kernel");
  AMMC_CUDA_CHECK(cudaMemsetAsync(colsum, 0, (size_t)D * 4, st));
  gz_kernel<<<ceil_div(N, 64), 256, (size_t)D * 4, st>>>(z, bank_t, idx, g_diff, g_q1, gz, colsum, (int)N, D, k,
                                                        1.0 / ((double)N * (double)D));
  AMMC_LAUNCH_CHECK("gz_kernel");
  return 0;
}

extern "C" int ammc_embed_code(const int64_t* ids, const float* embed, float* out, int64_t n_ids, int D, int M,
                               void* stream) {
  AMMC_REQUIRE(ids && embed && out && n_ids >= 0 && D > 0 && M > 0, "bad argument");
  if (n_ids == 0) return 0;
  int64_t total = n_ids * D;
  embed_code_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(ids, embed, out, n_ids, D, M);
  AMMC_LAUNCH_CHECK("embed_code_kernel");
  return 0;
}

extern "C" int ammc_ema_update(float* embed, float* cluster_size, float* embed_avg, const float* counts,
                               const float* embed_sum, int D, int M, float decay, float eps, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AMMC_REQUIRE(embed && cluster_size && embed_avg && counts && embed_sum && D > 0 && M > 0, "bad argument");
  ema_cluster_kernel<<<ceil_div(M, 256), 256, 0, st>>>(cluster_size, counts, M, decay);
  AMMC_LAUNCH_CHECK("ema_cluster_kernel");
  const int64_t total = (int64_t)D * M;
  const int per_block = 4096;
  ema_embed_kernel<<<ceil_div(total, per_block), 256, 0, st>>>(embed, embed_avg, embed_sum, cluster_size, D, M, decay,
                                                               eps, per_block);
  AMMC_LAUNCH_CHECK("ema_embed_kernel");
  return 0;
}

extern "C" size_t ammc_mem_bwd_workspace_bytes(int b, int h, int w, int C, int D, int M, int k) {
  int64_t N = (int64_t)b * h * w;
  // + tensor-core gx: bf16 hi/lo planes of g_z, packed enc_w^T, unit scale / zero shift vectors
  return align_up((size_t)M * D * 4, 256) + align_up((size_t)N * D * 4, 256) + align_up((size_t)k * M * C * 4, 256) +
         align_up((size_t)2 * N * D * 2, 256) + align_up((size_t)2 * C * D * 2, 256) + 2 * align_up((size_t)C * 4, 256) +
         align_up((size_t)C * D * 4, 256) +
         // tensor-core weight gradients: NHWC planes of x and g_out, planes of the gathered read, g_dec_w^T
         2 * align_up((size_t)2 * N * C * 2, 256) + align_up((size_t)2 * N * k * D * 2, 256) +
         align_up((size_t)k * D * C * 4, 256) +
         // per-block channel sums of g_out from its pack pass (one row of C values per 64 pixels)
         align_up((size_t)((N + 63) / 64) * C * 4, 256);
}

extern "C" int ammc_mem_bwd(const float* x, const float* enc_w, const float* embed, const int64_t* idx,
                            const float* z, const float* g_out, const float* g_diff, const float* g_q1, float* gx,
                            float* g_enc_w, float* g_enc_b, float* g_dec_w, float* g_dec_b, void* workspace,
                            size_t workspace_bytes, int b, int h, int w, int C, int D, int M, int k, int residual,
                            void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t N = (int64_t)b * h * w;
  const int HW = h * w;
  if (int rc = check_dims(N, C, D, M, k)) return rc;
  AMMC_REQUIRE(x && enc_w && embed && idx && z && g_out && g_diff && gx && g_enc_w && g_enc_b && g_dec_w && g_dec_b,
               "null pointer argument");
  AMMC_REQUIRE(D <= 8192, "embed_dim %d too large for the backward column-sum buffer", D);
  Workspace ws(workspace, workspace_bytes);
  float* bank_t = ws.take<float>((size_t)M * D);
  float* gz = ws.take<float>((size_t)N * D);
  float* G = ws.take<float>((size_t)k * M * C);
  if (!ws.ok()) return fail(AMMC_EWORKSPACE, "workspace too small");
  bank_transpose_kernel<<<dim3(ceil_div(M, 32), ceil_div(D, 32)), dim3(32, 8), 0, st>>>(embed, bank_t, D, M);
  AMMC_LAUNCH_CHECK("bank_transpose_kernel");
  const bool tc_gx = g_dec_mode != 1 && D % 64 == 0 && C % 64 == 0;     // ammc_set_dec_mode(1) keeps every 1x1 GEMM on CUDA cores
  const bool tc_wgrad = tc_gx && w <= 128 && b <= 65535;
  AMMC_CUDA_CHECK(cudaMemsetAsync(g_enc_b, 0, (size_t)D * 4, st));
  AMMC_CUDA_CHECK(cudaMemsetAsync(g_dec_b, 0, (size_t)C * 4, st));
  if (!tc_wgrad) {                 // the CUDA-core weight gradients accumulate; the tensor-core ones zero their own outputs
    AMMC_CUDA_CHECK(cudaMemsetAsync(g_enc_w, 0, (size_t)D * C * 4, st));
    AMMC_CUDA_CHECK(cudaMemsetAsync(G, 0, (size_t)k * M * C * 4, st));
  }
  __nv_bfloat16* gzp = nullptr;                    // bf16 hi/lo planes of g_z (tensor-core route)
  if (tc_gx) {
    gzp = ws.take<__nv_bfloat16>((size_t)2 * N * D);
    if (!ws.ok()) return fail(AMMC_EWORKSPACE, "workspace too small");
  }
  const bool gz4 = gz4_ok(D, z, bank_t, g_q1, gz) && (!gzp || ((uintptr_t)gzp & 7) == 0);
  if (gz4)
    gz4_kernel<<<ceil_div(N, 64), 256, 0, st>>>(z, bank_t, idx, g_diff, g_q1, gz, gzp, g_enc_b, (int)N, D, k,
                                                1.0 / ((double)N * (double)D));
  else
    gz_kernel<<<ceil_div(N, 64), 256, (size_t)D * 4, st>>>(z, bank_t, idx, g_diff, g_q1, gz, g_enc_b, (int)N, D, k,
                                                          1.0 / ((double)N * (double)D));
  AMMC_LAUNCH_CHECK("gz_kernel");
  if (tc_gx) {
    // gx = (residual ? g_out : 0) + g_z . enc_w  as a split-bf16 x3 1x1 GEMM on tcgen05 (the conv engine with K = D):
    // operands = NHWC planes of g_z and enc_w^T [C][D]; the residual add and the NCHW store are its epilogue
    __nv_bfloat16* wtp = ws.take<__nv_bfloat16>((size_t)2 * C * D);
    float* ones = ws.take<float>(C);
    float* zeros = ws.take<float>(C);
    float* wt = ws.take<float>((size_t)C * D);
    if (!ws.ok()) return fail(AMMC_EWORKSPACE, "workspace too small");
    if (!gz4)
      if (int rc = pack_planes_f32(gz, gzp, (long long)N * D, st)) return rc;
    bank_transpose_kernel<<<dim3(ceil_div(C, 32), ceil_div(D, 32)), dim3(32, 8), 0, st>>>(enc_w, wt, D, C);   // [D][C] -> [C][D]
    AMMC_LAUNCH_CHECK("bank_transpose_kernel");
    if (int rc = pack_weights_1x1(wt, wtp, C, D, st)) return rc;
    fill2_kernel<<<ceil_div(C, 256), 256, 0, st>>>(ones, 1.f, zeros, 0.f, C);
    AMMC_LAUNCH_CHECK("fill2_kernel");
    if (int rc = conv_igemm(gzp, wtp, ones, zeros, nullptr, gx, residual ? g_out : nullptr, b, D, C, h, w, 1, 3, 0, st))
      return rc;
    if (tc_wgrad) {
      // both weight gradients are GEMMs with K = pixels: the tcgen05 weight-gradient kernel (amft_train.cu) on NHWC bf16
      // hi/lo planes.   g_enc_w [D][C] = g_z^T . x ;   g_dec_w^T [kD][C] = read^T . g_out ;   g_dec_b from the pack pass
      __nv_bfloat16* xpl = ws.take<__nv_bfloat16>((size_t)2 * N * C);
      __nv_bfloat16* gopl = ws.take<__nv_bfloat16>((size_t)2 * N * C);
      __nv_bfloat16* rdpl = ws.take<__nv_bfloat16>((size_t)2 * N * k * D);
      float* gdwT = ws.take<float>((size_t)k * D * C);
      if (!ws.ok()) return fail(AMMC_EWORKSPACE, "workspace too small");
      if (int rc = pack_nhwc64(x, xpl, b, C, HW, st)) return rc;
      if (pack64x64_ok(g_out, C, HW)) {             // g_dec_b from the pack pass: g_out is read once
        float* part = ws.take<float>((size_t)(N / 64) * C);
        if (!ws.ok()) return fail(AMMC_EWORKSPACE, "workspace too small");
        if (int rc = pack_nhwc64_colsum(g_out, gopl, part, g_dec_b, b, C, HW, st)) return rc;
      } else {
        if (int rc = pack_nhwc64(g_out, gopl, b, C, HW, st)) return rc;
        const int per = max(1, b / 8);
        channel_sum_kernel<<<dim3(C, ceil_div(b, per)), 256, 0, st>>>(g_out, g_dec_b, b, C, HW, per);
        AMMC_LAUNCH_CHECK("channel_sum_kernel");
      }
      read_planes_kernel<<<num_sms() * 8, 256, 0, st>>>(idx, bank_t, rdpl, (long long)N, D, k, (long long)N * k * D);
      AMMC_LAUNCH_CHECK("read_planes_kernel");
      if (int rc = conv_wgrad_run(gzp, xpl, g_enc_w, b, C, D, h, w, 1, 3, st)) return rc;
      if (int rc = conv_wgrad_run(rdpl, gopl, gdwT, b, C, k * D, h, w, 1, 3, st)) return rc;
      bank_transpose_kernel<<<dim3(ceil_div(C, 32), ceil_div(k * D, 32)), dim3(32, 8), 0, st>>>(gdwT, g_dec_w, k * D, C);
      AMMC_LAUNCH_CHECK("bank_transpose_kernel");
      return 0;
    }
  } else {
    gx_kernel<<<dim3(ceil_div(N, 64), ceil_div(C, 64)), 256, 0, st>>>(gz, enc_w, g_out, gx, (int)N, HW, C, D, residual);
    AMMC_LAUNCH_CHECK("gx_kernel");
  }
  {
    int tiles = ceil_div(D, 64) * ceil_div(C, 64);
    int splits = max(1, min((int)ceil_div(N, 256), ceil_div(4 * num_sms(), tiles)));
    int per = (int)align_up((size_t)ceil_div(N, splits), 16);
    splits = ceil_div(N, per);
    genc_w_kernel<<<dim3(ceil_div(D, 64), ceil_div(C, 64), splits), 256, 0, st>>>(gz, x, g_enc_w, (int)N, HW, C, D, per);
    AMMC_LAUNCH_CHECK("genc_w_kernel");
  }
  const size_t priv_bytes = (size_t)k * M * 32 * sizeof(float);
  if (priv_bytes <= 160 * 1024) {
    // private shared-memory tables (shipped shapes: k*M = 512 rows -> 64 KB)
    dim3 grid(ceil_div(N, PRIV_PX), ceil_div(C, 32));
    switch (k) {
#define AMMC_SP_CASE(KK)                                                                                              \
  case KK:                                                                                                            \
    AMMC_CUDA_CHECK(cudaFuncSetAttribute(gdec_scatter_priv_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                         160 * 1024));                                                                \
    gdec_scatter_priv_kernel<KK><<<grid, 256, priv_bytes, st>>>(g_out, idx, G, g_dec_b, (int)N, HW, C, M);            \
    break;
      AMMC_SP_CASE(1) AMMC_SP_CASE(2) AMMC_SP_CASE(3) AMMC_SP_CASE(4)
      AMMC_SP_CASE(5) AMMC_SP_CASE(6) AMMC_SP_CASE(7) AMMC_SP_CASE(8)
#undef AMMC_SP_CASE
    }
    AMMC_LAUNCH_CHECK("gdec_scatter_priv_kernel");
  } else {
    const int chunks = 4;
    dim3 grid(ceil_div(N, 32), ceil_div(ceil_div(C, 32), chunks));
    switch (k) {
#define AMMC_SC_CASE(KK) \
  case KK: gdec_scatter_kernel<KK><<<grid, 256, 0, st>>>(g_out, idx, G, g_dec_b, (int)N, HW, C, M, chunks); break;
      AMMC_SC_CASE(1) AMMC_SC_CASE(2) AMMC_SC_CASE(3) AMMC_SC_CASE(4)
      AMMC_SC_CASE(5) AMMC_SC_CASE(6) AMMC_SC_CASE(7) AMMC_SC_CASE(8)
#undef AMMC_SC_CASE
    }
    AMMC_LAUNCH_CHECK("gdec_scatter_kernel");
  }
  gdec_w_kernel<<<dim3(ceil_div(D, 32), ceil_div(C, 32), k), dim3(32, 8), 0, st>>>(G, embed, g_dec_w, C, D, M, k);
  AMMC_LAUNCH_CHECK("gdec_w_kernel");
  return 0;
}
