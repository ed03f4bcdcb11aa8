// Image-space generator losses (reference Code/models/losses/losses_utils.py:17-59,124-129):
//   intensity  = mean over (b,h,w) of || gen - gt ||_2 over the channel axis                     (Intensity_Loss -> L2)
//   gradient   = mean over (b,h,w) of |dx| + |dy|, dx = s(h,w) - s(h,w-1), dy = s(h,w) - s(h-1,w),
//                s = sum over channels of (gt - gen), zero outside the image                     (Gradient_Loss, alpha = 1)
// The reference runs ~14 ATen kernels (pads, 4 convolutions, abs, pow, means) and their autograd twins per call; here the
// forward is one pass over gen/gt (+ a tiny deterministic final sum), the backward one pass writing d loss / d gen.
// HBM-bound: 2*C*H*W*4 bytes per frame each way (the left / upper neighbours are L1/L2 hits).
#include "common.cuh"

namespace ammc {

constexpr int LOSS_MAX_C = 8;

template <int C>
__device__ __forceinline__ float chan_sum_diff(const float* __restrict__ gen, const float* __restrict__ gt, size_t base,
                                               size_t plane) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) s += gt[base + c * plane] - gen[base + c * plane];
  return s;
}

// partial[block] = (sum of channel norms, sum of |dx|+|dy|) over the block's pixels; grid (ceil(W*H/256), n)
template <int C>
__global__ void __launch_bounds__(256) frame_loss_partial_kernel(const float* __restrict__ gen, const float* __restrict__ gt,
                                                                  float* __restrict__ partial, int H, int W) {
  __shared__ float red[33];
  const size_t plane = (size_t)H * W;
  const int p = blockIdx.x * 256 + threadIdx.x;
  const size_t img = (size_t)blockIdx.y * C * plane;
  float vi = 0.f, vg = 0.f;
  if (p < H * W) {
    const int h = p / W, w = p % W;
    float n2 = 0.f, s = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float d = gt[img + c * plane + p] - gen[img + c * plane + p];
      n2 = fmaf(d, d, n2);
      s += d;
    }
    vi = sqrtf(n2);
    const float sl = w > 0 ? chan_sum_diff<C>(gen, gt, img + p - 1, plane) : 0.f;
    const float su = h > 0 ? chan_sum_diff<C>(gen, gt, img + p - W, plane) : 0.f;
    vg = fabsf(s - sl) + fabsf(s - su);
  }
  const float a = block_sum(vi, red);
  __syncthreads();
  const float b = block_sum(vg, red);
  if (threadIdx.x == 0) {
    const size_t o = ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2;
    partial[o] = a;
    partial[o + 1] = b;
  }
}

// out[0] = intensity, out[1] = gradient: fixed-order sum of the partials (deterministic), divided by the pixel count
__global__ void __launch_bounds__(256) frame_loss_final_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                                int n_partials, double inv_count) {
  __shared__ double red[2][8];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < n_partials; i += 256) { a += (double)partial[2 * i]; b += (double)partial[2 * i + 1]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sa = 0.0, sb = 0.0;
    for (int i = 0; i < 8; ++i) { sa += red[0][i]; sb += red[1][i]; }
    out[0] = (float)(sa * inv_count);
    out[1] = (float)(sb * inv_count);
  }
}

__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

// grad_gen[c,p] = g_int * (-(d_c / ||d||) / P)  +  g_gd * (-(sign dx(p) - sign dx(right) + sign dy(p) - sign dy(below)) / P)
template <int C>
__global__ void __launch_bounds__(256) frame_loss_bwd_kernel(const float* __restrict__ gen, const float* __restrict__ gt,
                                                              const float* __restrict__ g_int, const float* __restrict__ g_gd,
                                                              float* __restrict__ grad_gen, int H, int W, float inv_count) {
  const size_t plane = (size_t)H * W;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= H * W) return;
  const size_t img = (size_t)blockIdx.y * C * plane;
  const int h = p / W, w = p % W;
  const float gi = g_int ? g_int[0] : 0.f, gg = g_gd ? g_gd[0] : 0.f;
  float d[C];
  float n2 = 0.f, s = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    d[c] = gt[img + c * plane + p] - gen[img + c * plane + p];
    n2 = fmaf(d[c], d[c], n2);
    s += d[c];
  }
  const float sl = w > 0 ? chan_sum_diff<C>(gen, gt, img + p - 1, plane) : 0.f;
  const float su = h > 0 ? chan_sum_diff<C>(gen, gt, img + p - W, plane) : 0.f;
  float t = sgn(s - sl) + sgn(s - su);
  if (w + 1 < W) t -= sgn(chan_sum_diff<C>(gen, gt, img + p + 1, plane) - s);
  if (h + 1 < H) t -= sgn(chan_sum_diff<C>(gen, gt, img + p + W, plane) - s);
  const float norm = sqrtf(n2);
  const float ki = norm > 0.f ? gi * inv_count / norm : 0.f;      // torch: the 2-norm has a zero (sub)gradient at 0
  const float kg = gg * inv_count * t;
#pragma unroll
  for (int c = 0; c < C; ++c) grad_gen[img + c * plane + p] = -(ki * d[c]) - kg;
}

}  // namespace ammc

using namespace ammc;

extern "C" size_t ammc_frame_losses_workspace_bytes(int n, int H, int W) {
  return align_up((size_t)n * ceil_div((int64_t)H * W, 256) * 2 * sizeof(float), 256);
}

extern "C" int ammc_frame_losses_fwd(const float* gen, const float* gt, float* out2, void* workspace, size_t workspace_bytes,
                                     int n, int C, int H, int W, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AMMC_REQUIRE(gen && gt && out2 && n > 0 && H > 0 && W > 0, "bad argument");
  AMMC_REQUIRE(C >= 1 && C <= LOSS_MAX_C, "frame losses support 1..%d channels (got %d)", LOSS_MAX_C, C);
  AMMC_REQUIRE(n <= 65535, "batch %d too large for one launch", n);
  if (!workspace || workspace_bytes < ammc_frame_losses_workspace_bytes(n, H, W)) return fail(AMMC_EWORKSPACE, "workspace too small");
  const int bx = ceil_div((int64_t)H * W, 256);
  float* partial = (float*)workspace;
  switch (C) {
#define AMMC_FL_CASE(CC) case CC: frame_loss_partial_kernel<CC><<<dim3(bx, n), 256, 0, st>>>(gen, gt, partial, H, W); break;
    AMMC_FL_CASE(1) AMMC_FL_CASE(2) AMMC_FL_CASE(3) AMMC_FL_CASE(4) AMMC_FL_CASE(5) AMMC_FL_CASE(6) AMMC_FL_CASE(7) AMMC_FL_CASE(8)
#undef AMMC_FL_CASE
  }
  AMMC_LAUNCH_CHECK("frame_loss_partial_kernel");
  frame_loss_final_kernel<<<1, 256, 0, st>>>(partial, out2, bx * n, 1.0 / ((double)n * H * W));
  AMMC_LAUNCH_CHECK("frame_loss_final_kernel");
  return 0;
}

extern "C" int ammc_frame_losses_bwd(const float* gen, const float* gt, const float* g_int, const float* g_gd, float* grad_gen,
                                     int n, int C, int H, int W, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AMMC_REQUIRE(gen && gt && grad_gen && (g_int || g_gd) && n > 0 && H > 0 && W > 0, "bad argument");
  AMMC_REQUIRE(C >= 1 && C <= LOSS_MAX_C, "frame losses support 1..%d channels (got %d)", LOSS_MAX_C, C);
  AMMC_REQUIRE(n <= 65535, "batch %d too large for one launch", n);
  const int bx = ceil_div((int64_t)H * W, 256);
  const float inv = (float)(1.0 / ((double)n * H * W));
  switch (C) {
#define AMMC_FL_CASE(CC) \
  case CC: frame_loss_bwd_kernel<CC><<<dim3(bx, n), 256, 0, st>>>(gen, gt, g_int, g_gd, grad_gen, H, W, inv); break;
    AMMC_FL_CASE(1) AMMC_FL_CASE(2) AMMC_FL_CASE(3) AMMC_FL_CASE(4) AMMC_FL_CASE(5) AMMC_FL_CASE(6) AMMC_FL_CASE(7) AMMC_FL_CASE(8)
#undef AMMC_FL_CASE
  }
  AMMC_LAUNCH_CHECK("frame_loss_bwd_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------------------------------------
// Element-wise mean objectives of the training step (reference Code/models/losses/losses_utils.py:10-15,103-113; consumer
// Code/models/losses/loss_zoo.py:331-336 and Code/run_helper/train_helper.py:318-326):
//   ELEM_L1        Flow_Loss          mean |a - b|
//   ELEM_LSGAN_G   Adversarial_Loss   mean (a - 1)^2 / 2                       (a = discriminator map of the prediction)
//   ELEM_LSGAN_D   Discriminate_Loss  mean (a - 1)^2 / 2 + mean b^2 / 2        (a = map of the real frame, b = of the fake)
// The reference spends 3-6 ATen kernels + their autograd twins per objective; here one grid-stride pass (16-byte loads when
// both pointers allow) + the fixed-order final sum, and one pass for the gradient.  HBM-bound: 4 bytes per element read.
// ------------------------------------------------------------------------------------------------------------------------
namespace ammc {

enum { ELEM_L1 = 0, ELEM_LSGAN_G = 1, ELEM_LSGAN_D = 2 };

template <int MODE>
__device__ __forceinline__ void elem_terms(float a, float b, float& sa, float& sb) {
  if (MODE == ELEM_L1) sa += fabsf(a - b);
  else if (MODE == ELEM_LSGAN_G) sa = fmaf(0.5f * (a - 1.f), a - 1.f, sa);
  else { sa = fmaf(0.5f * (a - 1.f), a - 1.f, sa); sb = fmaf(0.5f * b, b, sb); }
}

// partial[2*block] = sum of the a-terms, partial[2*block+1] = sum of the b-terms (ELEM_LSGAN_D only) over the block's share;
// n = element count shared by a and b (ELEM_LSGAN_D with maps of different sizes runs as two ELEM_LSGAN_G-style launches)
template <int MODE, bool VEC>
__global__ void __launch_bounds__(256) elem_loss_partial_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                 float* __restrict__ partial, int64_t n) {
  __shared__ float red[33];
  float sa = 0.f, sb = 0.f;
  const int64_t stride = (int64_t)gridDim.x * 256, t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (VEC) {
    const int64_t n4 = n >> 2;
    for (int64_t i = t; i < n4; i += stride) {
      const float4 va = reinterpret_cast<const float4*>(a)[i];
      float4 vb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE != ELEM_LSGAN_G) vb = reinterpret_cast<const float4*>(b)[i];
      elem_terms<MODE>(va.x, vb.x, sa, sb); elem_terms<MODE>(va.y, vb.y, sa, sb);
      elem_terms<MODE>(va.z, vb.z, sa, sb); elem_terms<MODE>(va.w, vb.w, sa, sb);
    }
    for (int64_t i = (n4 << 2) + t; i < n; i += stride) elem_terms<MODE>(a[i], MODE != ELEM_LSGAN_G ? b[i] : 0.f, sa, sb);
  } else {
    for (int64_t i = t; i < n; i += stride) elem_terms<MODE>(a[i], MODE != ELEM_LSGAN_G ? b[i] : 0.f, sa, sb);
  }
  const float ra = block_sum(sa, red);
  __syncthreads();
  const float rb = MODE == ELEM_LSGAN_D ? block_sum(sb, red) : 0.f;
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = ra; partial[2 * blockIdx.x + 1] = rb; }
}

// out[0] = (sum of a-partials + sum of b-partials) / n, fixed summation order (deterministic)
__global__ void __launch_bounds__(256) elem_loss_final_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                               int n_partials, double inv_count) {
  __shared__ double red[8];
  double s = 0.0;
  for (int i = threadIdx.x; i < 2 * n_partials; i += 256) s += (double)partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += red[i];
    out[0] = (float)(t * inv_count);
  }
}

// grad_a = g * d/da, grad_b = g * d/db (either may be NULL; ELEM_LSGAN_G has no b)
template <int MODE>
__global__ void __launch_bounds__(256) elem_loss_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                             const float* __restrict__ g, float* __restrict__ grad_a,
                                                             float* __restrict__ grad_b, int64_t n, float inv_count) {
  const float k = g[0] * inv_count;
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
    if (MODE == ELEM_L1) {
      const float d = a[i] - b[i];
      const float s = k * sgn(d);
      if (grad_a) grad_a[i] = s;
      if (grad_b) grad_b[i] = -s;
    } else {
      if (grad_a) grad_a[i] = k * (a[i] - 1.f);
      if (MODE == ELEM_LSGAN_D && grad_b) grad_b[i] = k * b[i];
    }
  }
}

static inline int elem_blocks(int64_t n) {
  const int64_t want = (n + 256 * 8 - 1) / (256 * 8);            // ~8 elements per thread before the grid stops growing
  const int64_t cap = (int64_t)num_sms() * 8;
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace ammc

extern "C" size_t ammc_elem_loss_workspace_bytes(int64_t n) {
  return align_up((size_t)elem_blocks(n > 0 ? n : 1) * 2 * sizeof(float), 256);
}

extern "C" int ammc_elem_loss_fwd(const float* a, const float* b, float* out1, int mode, int64_t n, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AMMC_REQUIRE(a && out1 && n > 0, "bad argument");
  AMMC_REQUIRE(mode >= ELEM_L1 && mode <= ELEM_LSGAN_D, "unknown objective %d", mode);
  AMMC_REQUIRE(mode == ELEM_LSGAN_G || b, "objective %d needs two tensors", mode);
  if (!workspace || workspace_bytes < ammc_elem_loss_workspace_bytes(n)) return fail(AMMC_EWORKSPACE, "workspace too small");
  const int blocks = elem_blocks(n);
  float* partial = (float*)workspace;
  const bool vec = ((((uintptr_t)a) | ((uintptr_t)(mode == ELEM_LSGAN_G ? a : b))) & 15) == 0;
#define AMMC_EL_LAUNCH(M)                                                                             \
  if (vec) elem_loss_partial_kernel<M, true><<<blocks, 256, 0, st>>>(a, b, partial, n);               \
  else elem_loss_partial_kernel<M, false><<<blocks, 256, 0, st>>>(a, b, partial, n);
  if (mode == ELEM_L1) { AMMC_EL_LAUNCH(ELEM_L1) }
  else if (mode == ELEM_LSGAN_G) { AMMC_EL_LAUNCH(ELEM_LSGAN_G) }
  else { AMMC_EL_LAUNCH(ELEM_LSGAN_D) }
#undef AMMC_EL_LAUNCH
  AMMC_LAUNCH_CHECK("elem_loss_partial_kernel");
  elem_loss_final_kernel<<<1, 256, 0, st>>>(partial, out1, blocks, 1.0 / (double)n);
  AMMC_LAUNCH_CHECK("elem_loss_final_kernel");
  return 0;
}

extern "C" int ammc_elem_loss_bwd(const float* a, const float* b, const float* g, float* grad_a, float* grad_b, int mode,
                                  int64_t n, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AMMC_REQUIRE(a && g && (grad_a || grad_b) && n > 0, "bad argument");
  AMMC_REQUIRE(mode >= ELEM_L1 && mode <= ELEM_LSGAN_D, "unknown objective %d", mode);
  AMMC_REQUIRE(mode == ELEM_LSGAN_G || b, "objective %d needs two tensors", mode);
  AMMC_REQUIRE(mode != ELEM_LSGAN_G || !grad_b, "objective %d has one tensor", mode);
  const int blocks = elem_blocks(n);
  const float inv = (float)(1.0 / (double)n);
  if (mode == ELEM_L1) elem_loss_bwd_kernel<ELEM_L1><<<blocks, 256, 0, st>>>(a, b, g, grad_a, grad_b, n, inv);
  else if (mode == ELEM_LSGAN_G) elem_loss_bwd_kernel<ELEM_LSGAN_G><<<blocks, 256, 0, st>>>(a, b, g, grad_a, grad_b, n, inv);
  else elem_loss_bwd_kernel<ELEM_LSGAN_D><<<blocks, 256, 0, st>>>(a, b, g, grad_a, grad_b, n, inv);
  AMMC_LAUNCH_CHECK("elem_loss_bwd_kernel");
  return 0;
}
