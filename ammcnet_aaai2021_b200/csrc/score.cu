// Scoring: batched per-frame PSNR and the regularity-score reduction.
//
//   psnr_*          replaces the per-frame psnr_error calls (reference Code/utils/utils.py:130-148) issued one
//                   frame at a time by the scoring loop (Code/run_helper/test_helper.py:445-452): one coalesced
//                   128-bit pass over gen/gt for the whole batch, no host sync.  HBM-bound: 2*elems*4 B per frame.
//   score_*         replaces norm_score + mixing + 2-tap smoothing (Code/main/eval_metric.py:405-427).  Pure
//                   IEEE fp32 elementwise arithmetic with explicit _rn intrinsics (no FMA contraction), so the
//                   result is bit-identical to the numpy float32 reference.
#include "common.cuh"
#include <float.h>

namespace ammc {

constexpr int PSNR_THREADS = 256;
constexpr int PSNR_ELEMS_PER_BLOCK = PSNR_THREADS * 4 * 8;  // 8 float4 per thread

__device__ __forceinline__ float sqdiff01(float g, float t) {
  // ((gt + 1)/2 - (gen + 1)/2)^2, same association as utils.py:143-145
  float a = __fmul_rn(__fadd_rn(t, 1.0f), 0.5f);
  float b = __fmul_rn(__fadd_rn(g, 1.0f), 0.5f);
  float d = __fsub_rn(a, b);
  return d * d;
}

__global__ void __launch_bounds__(PSNR_THREADS) psnr_partial_kernel(const float* __restrict__ gen,
                                                                    const float* __restrict__ gt,
                                                                    float* __restrict__ partial, int64_t elems,
                                                                    int blocks_per_frame, int vec_ok) {
  __shared__ float red[33];
  const int frame = blockIdx.y;
  const float* g = gen + (size_t)frame * elems;
  const float* t = gt + (size_t)frame * elems;
  const int64_t beg = (int64_t)blockIdx.x * PSNR_ELEMS_PER_BLOCK;
  const int64_t end = min(elems, beg + PSNR_ELEMS_PER_BLOCK);
  float s = 0.f;
  if (vec_ok) {
    const float4* g4 = reinterpret_cast<const float4*>(g + beg);
    const float4* t4 = reinterpret_cast<const float4*>(t + beg);
    const int n4 = (int)((end - beg) >> 2);
    float4 gv[8], tv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int j = threadIdx.x + i * PSNR_THREADS;
      if (j < n4) { gv[i] = __ldg(g4 + j); tv[i] = __ldg(t4 + j); }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int j = threadIdx.x + i * PSNR_THREADS;
      if (j < n4) {
        s += sqdiff01(gv[i].x, tv[i].x) + sqdiff01(gv[i].y, tv[i].y) + sqdiff01(gv[i].z, tv[i].z) +
             sqdiff01(gv[i].w, tv[i].w);
      }
    }
    for (int64_t e = beg + ((int64_t)n4 << 2) + threadIdx.x; e < end; e += PSNR_THREADS) s += sqdiff01(g[e], t[e]);
  } else {
    for (int64_t e = beg + threadIdx.x; e < end; e += PSNR_THREADS) s += sqdiff01(g[e], t[e]);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[(size_t)frame * blocks_per_frame + blockIdx.x] = s;
}

__global__ void psnr_final_kernel(const float* __restrict__ partial, float* __restrict__ psnr, int n,
                                  int blocks_per_frame, float inv_elems) {
  // one warp per frame; fixed summation order -> deterministic
  const int frame = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (frame >= n) return;
  float s = 0.f;
  for (int i = lane; i < blocks_per_frame; i += 32) s += partial[(size_t)frame * blocks_per_frame + i];
  s = warp_sum(s);
  if (lane == 0) psnr[frame] = 10.f * log10f(__fdiv_rn(1.0f, __fmul_rn(inv_elems, s)));  // utils.py:147
}

// ------------------------------------------------------------------------------------------------
// score reduction
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_minmax(float& mn, float& mx, float* red /* 66 floats */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  mn = warp_min(mn);
  mx = warp_max(mx);
  __syncthreads();
  if (lane == 0) { red[wid] = mn; red[33 + wid] = mx; }
  __syncthreads();
  if (wid == 0) {
    float a = lane < nw ? red[lane] : FLT_MAX, b = lane < nw ? red[33 + lane] : -FLT_MAX;
    a = warp_min(a);
    b = warp_max(b);
    if (lane == 0) { red[32] = a; red[65] = b; }
  }
  __syncthreads();
  mn = red[32];
  mx = red[65];
}

// stage 1: per (record, video) min-max normalisation, first DECIDABLE_IDX=4 frames dropped (eval_metric.py:405-413).
// grid (V, 2); norm[rec][offsets[v] - 4 v + i - 4] = (d[i] - min) / max(d - min)
__global__ void score_video_kernel(const float* __restrict__ img, const float* __restrict__ fea,
                                   const int64_t* __restrict__ offsets, float* __restrict__ norm, int64_t t_out) {
  __shared__ float red[66];
  const int v = blockIdx.x, rec = blockIdx.y;
  const float* src = (rec == 0 ? img : fea) + offsets[v];
  const int64_t len = offsets[v + 1] - offsets[v];
  float mn = FLT_MAX, mx = -FLT_MAX;
  for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
    float d = src[i];
    mn = fminf(mn, d);
    mx = fmaxf(mx, d);
  }
  block_minmax(mn, mx, red);
  const float range = __fsub_rn(mx, mn);
  float* dst = norm + (size_t)rec * t_out + (offsets[v] - 4 * (int64_t)v);
  for (int64_t i = 4 + threadIdx.x; i < len; i += blockDim.x)
    dst[i - 4] = __fdiv_rn(__fsub_rn(src[i], mn), range);
}

// stage 2 (single block): global min-max of both records (eval_metric.py:414-416), mix (:425), smooth (:426)
__global__ void score_final_kernel(const float* __restrict__ norm, float* __restrict__ mixed,
                                   float* __restrict__ scores, int64_t t_out, float oml1, float l1, float oml2,
                                   float l2) {
  __shared__ float red[66];
  float mn[2], rg[2];
  for (int rec = 0; rec < 2; ++rec) {
    const float* p = norm + (size_t)rec * t_out;
    float a = FLT_MAX, b = -FLT_MAX;
    for (int64_t i = threadIdx.x; i < t_out; i += blockDim.x) {
      float d = p[i];
      a = fminf(a, d);
      b = fmaxf(b, d);
    }
    block_minmax(a, b, red);
    mn[rec] = a;
    rg[rec] = __fsub_rn(b, a);
    __syncthreads();
  }
  for (int64_t i = threadIdx.x; i < t_out; i += blockDim.x) {
    float pi = __fdiv_rn(__fsub_rn(norm[i], mn[0]), rg[0]);
    float fi = __fdiv_rn(__fsub_rn(norm[t_out + i], mn[1]), rg[1]);
    mixed[i] = __fadd_rn(__fmul_rn(oml1, pi), __fmul_rn(l1, __fsub_rn(1.0f, fi)));
  }
  __syncthreads();
  for (int64_t i = threadIdx.x; i < t_out; i += blockDim.x)
    scores[i] = i > 0 ? __fadd_rn(__fmul_rn(oml2, mixed[i - 1]), __fmul_rn(l2, mixed[i])) : mixed[0];
}

}  // namespace ammc

using namespace ammc;

extern "C" size_t ammc_psnr_workspace_bytes(int n, int64_t elems) {
  int bpf = ceil_div(elems, PSNR_ELEMS_PER_BLOCK);
  return align_up((size_t)n * bpf * 4, 256);
}

extern "C" int ammc_psnr_batch(const float* gen, const float* gt, float* psnr, void* workspace,
                               size_t workspace_bytes, int n, int64_t elems, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AMMC_REQUIRE(gen && gt && psnr && n >= 0 && elems > 0, "bad argument");
  if (n == 0) return 0;
  AMMC_REQUIRE(n <= 65535, "at most 65535 frames per call (got %d)", n);
  const int bpf = ceil_div(elems, PSNR_ELEMS_PER_BLOCK);
  if (workspace_bytes < ammc_psnr_workspace_bytes(n, elems) || !workspace)
    return fail(AMMC_EWORKSPACE, "workspace too small");
  float* partial = (float*)workspace;
  const int vec_ok = (elems % 4 == 0) && ((uintptr_t)gen % 16 == 0) && ((uintptr_t)gt % 16 == 0);
  psnr_partial_kernel<<<dim3(bpf, n), PSNR_THREADS, 0, st>>>(gen, gt, partial, elems, bpf, vec_ok);
  AMMC_LAUNCH_CHECK("psnr_partial_kernel");
  psnr_final_kernel<<<ceil_div(n, 8), 256, 0, st>>>(partial, psnr, n, bpf, (float)(1.0 / (double)elems));
  AMMC_LAUNCH_CHECK("psnr_final_kernel");
  return 0;
}

extern "C" size_t ammc_score_workspace_bytes(int64_t t_total, int n_videos) {
  (void)n_videos;
  return align_up((size_t)t_total * 2 * 4, 256) + align_up((size_t)t_total * 4, 256);
}

extern "C" int ammc_score_reduce(const float* img, const float* fea, const int64_t* offsets, int n_videos,
                                 float one_minus_lam1, float lam1, float one_minus_lam2, float lam2, float* scores,
                                 void* workspace, size_t workspace_bytes, int64_t t_total, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AMMC_REQUIRE(img && fea && offsets && scores && n_videos > 0 && t_total > 4LL * n_videos, "bad argument");
  const int64_t t_out = t_total - 4LL * n_videos;
  Workspace ws(workspace, workspace_bytes);
  float* norm = ws.take<float>((size_t)t_total * 2);
  float* mixed = ws.take<float>((size_t)t_total);
  if (!ws.ok()) return fail(AMMC_EWORKSPACE, "workspace too small");
  score_video_kernel<<<dim3(n_videos, 2), 256, 0, st>>>(img, fea, offsets, norm, t_out);
  AMMC_LAUNCH_CHECK("score_video_kernel");
  score_final_kernel<<<1, 1024, 0, st>>>(norm, mixed, scores, t_out, one_minus_lam1, lam1, one_minus_lam2, lam2);
  AMMC_LAUNCH_CHECK("score_final_kernel");
  return 0;
}
