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
__global__ void __launch_bounds__(1024) frame_loss_final_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                                 int n_partials, double inv_count) {
  __shared__ double red[2][32];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < n_partials; i += 1024) {
    const float2 v = reinterpret_cast<const float2*>(partial)[i];
    a += (double)v.x;
    b += (double)v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sa = 0.0, sb = 0.0;
    for (int i = 0; i < 32; ++i) { sa += red[0][i]; sb += red[1][i]; }
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


// ---- 4 pixels per thread (W % 4 == 0, 16-byte aligned tensors): 16-byte loads / stores, the channel sums of the left / right /
// upper / lower neighbours come from the block's shared table where the neighbour quad belongs to the block, from L1/L2-resident
// global memory otherwise; a quarter of the partials for the final sum.
template <int C>
__device__ __forceinline__ float4 chan_sum_diff4(const float* __restrict__ gen, const float* __restrict__ gt, size_t base,
                                                 size_t plane) {
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float4 a = *reinterpret_cast<const float4*>(gt + base + c * plane);
    const float4 b = *reinterpret_cast<const float4*>(gen + base + c * plane);
    s.x += a.x - b.x; s.y += a.y - b.y; s.z += a.z - b.z; s.w += a.w - b.w;
  }
  return s;
}

template <int C>
__global__ void __launch_bounds__(256) frame_loss_partial4_kernel(const float* __restrict__ gen, const float* __restrict__ gt,
                                                                   float* __restrict__ partial, int H, int W) {
  __shared__ float red[33];
  __shared__ float4 sS[256];
  const size_t plane = (size_t)H * W;
  const int nq = (H * W) >> 2, wq = W >> 2, tid = threadIdx.x;
  const int q = blockIdx.x * 256 + tid;
  const size_t img = (size_t)blockIdx.y * C * plane;
  const bool live = q < nq;
  const int p = q << 2, h = live ? p / W : 0, w = live ? p - h * W : 0;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  float vi = 0.f, vg = 0.f;
  if (live) {
    float4 n2 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float4 a = *reinterpret_cast<const float4*>(gt + img + c * plane + p);
      const float4 b = *reinterpret_cast<const float4*>(gen + img + c * plane + p);
      const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z, dw = a.w - b.w;
      n2.x = fmaf(dx, dx, n2.x); n2.y = fmaf(dy, dy, n2.y); n2.z = fmaf(dz, dz, n2.z); n2.w = fmaf(dw, dw, n2.w);
      s.x += dx; s.y += dy; s.z += dz; s.w += dw;
    }
    vi = (sqrtf(n2.x) + sqrtf(n2.y)) + (sqrtf(n2.z) + sqrtf(n2.w));
  }
  sS[tid] = s;
  __syncthreads();
  if (live) {
    float sl0 = 0.f;
    if (w > 0) sl0 = tid > 0 ? sS[tid - 1].w : chan_sum_diff<C>(gen, gt, img + p - 1, plane);
    float4 su = make_float4(0.f, 0.f, 0.f, 0.f);
    if (h > 0) su = tid >= wq ? sS[tid - wq] : chan_sum_diff4<C>(gen, gt, img + p - W, plane);
    vg = (fabsf(s.x - sl0) + fabsf(s.x - su.x)) + (fabsf(s.y - s.x) + fabsf(s.y - su.y)) +
         (fabsf(s.z - s.y) + fabsf(s.z - su.z)) + (fabsf(s.w - s.z) + fabsf(s.w - su.w));
  }
  const float a = block_sum(vi, red);
  __syncthreads();
  const float b = block_sum(vg, red);
  if (tid == 0) {
    const size_t o = ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2;
    partial[o] = a;
    partial[o + 1] = b;
  }
}

template <int C>
__global__ void __launch_bounds__(256) frame_loss_bwd4_kernel(const float* __restrict__ gen, const float* __restrict__ gt,
                                                               const float* __restrict__ g_int, const float* __restrict__ g_gd,
                                                               float* __restrict__ grad_gen, int H, int W, float inv_count) {
  __shared__ float4 sS[256];
  const size_t plane = (size_t)H * W;
  const int nq = (H * W) >> 2, wq = W >> 2, tid = threadIdx.x;
  const int q = blockIdx.x * 256 + tid;
  const size_t img = (size_t)blockIdx.y * C * plane;
  const bool live = q < nq;
  const int p = q << 2, h = live ? p / W : 0, w = live ? p - h * W : 0;
  float4 d[C];
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), n2 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float4 a = *reinterpret_cast<const float4*>(gt + img + c * plane + p);
      const float4 b = *reinterpret_cast<const float4*>(gen + img + c * plane + p);
      d[c] = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
      n2.x = fmaf(d[c].x, d[c].x, n2.x); n2.y = fmaf(d[c].y, d[c].y, n2.y);
      n2.z = fmaf(d[c].z, d[c].z, n2.z); n2.w = fmaf(d[c].w, d[c].w, n2.w);
      s.x += d[c].x; s.y += d[c].y; s.z += d[c].z; s.w += d[c].w;
    }
  }
  sS[tid] = s;
  __syncthreads();
  if (!live) return;
  const float gi = g_int ? g_int[0] : 0.f, gg = g_gd ? g_gd[0] : 0.f;
  float sl0 = 0.f;
  if (w > 0) sl0 = tid > 0 ? sS[tid - 1].w : chan_sum_diff<C>(gen, gt, img + p - 1, plane);
  float4 su = make_float4(0.f, 0.f, 0.f, 0.f);
  if (h > 0) su = tid >= wq ? sS[tid - wq] : chan_sum_diff4<C>(gen, gt, img + p - W, plane);
  // forward differences of this quad's own terms (zero padding on the left / top keeps the term)
  float4 t = make_float4(sgn(s.x - sl0) + sgn(s.x - su.x), sgn(s.y - s.x) + sgn(s.y - su.y), sgn(s.z - s.y) + sgn(s.z - su.z),
                         sgn(s.w - s.z) + sgn(s.w - su.w));
  // minus the terms of the right / lower neighbours in which this pixel is the subtrahend
  t.x -= sgn(s.y - s.x); t.y -= sgn(s.z - s.y); t.z -= sgn(s.w - s.z);
  if (w + 4 < W) t.w -= sgn((tid < 255 ? sS[tid + 1].x : chan_sum_diff<C>(gen, gt, img + p + 4, plane)) - s.w);
  if (h + 1 < H) {
    const float4 sd = tid + wq < 256 ? sS[tid + wq] : chan_sum_diff4<C>(gen, gt, img + p + W, plane);
    t.x -= sgn(sd.x - s.x); t.y -= sgn(sd.y - s.y); t.z -= sgn(sd.z - s.z); t.w -= sgn(sd.w - s.w);
  }
  const float gic = gi * inv_count, ggc = gg * inv_count;
  const float4 nr = make_float4(sqrtf(n2.x), sqrtf(n2.y), sqrtf(n2.z), sqrtf(n2.w));
  const float4 ki = make_float4(nr.x > 0.f ? gic / nr.x : 0.f, nr.y > 0.f ? gic / nr.y : 0.f, nr.z > 0.f ? gic / nr.z : 0.f,
                                nr.w > 0.f ? gic / nr.w : 0.f);
  const float4 kg = make_float4(ggc * t.x, ggc * t.y, ggc * t.z, ggc * t.w);
#pragma unroll
  for (int c = 0; c < C; ++c)
    *reinterpret_cast<float4*>(grad_gen + img + c * plane + p) =
        make_float4(-(ki.x * d[c].x) - kg.x, -(ki.y * d[c].y) - kg.y, -(ki.z * d[c].z) - kg.z, -(ki.w * d[c].w) - kg.w);
}

static inline bool frame_quads_ok(const void* a, const void* b, const void* c, int W) {
  return W % 4 == 0 && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) == 0;
}

// launches the partial-sum pass of one (gen, gt) pair; *n_partials = number of (intensity, gradient) PAIRS written
static int launch_frame_partial(const float* gen, const float* gt, float* partial, int n, int C, int H, int W, cudaStream_t st,
                                int* n_partials) {
  const bool quads = frame_quads_ok(gen, gt, nullptr, W);
  const int bx = quads ? ceil_div(((int64_t)H * W) >> 2, 256) : ceil_div((int64_t)H * W, 256);
  switch (C) {
#define AMMC_FL_CASE(CC)                                                                                      \
  case CC:                                                                                                    \
    if (quads) frame_loss_partial4_kernel<CC><<<dim3(bx, n), 256, 0, st>>>(gen, gt, partial, H, W);           \
    else frame_loss_partial_kernel<CC><<<dim3(bx, n), 256, 0, st>>>(gen, gt, partial, H, W);                  \
    break;
    AMMC_FL_CASE(1) AMMC_FL_CASE(2) AMMC_FL_CASE(3) AMMC_FL_CASE(4) AMMC_FL_CASE(5) AMMC_FL_CASE(6) AMMC_FL_CASE(7) AMMC_FL_CASE(8)
#undef AMMC_FL_CASE
    default: return fail(AMMC_EINVAL, "frame losses support 1..%d channels (got %d)", LOSS_MAX_C, C);
  }
  AMMC_LAUNCH_CHECK("frame_loss_partial_kernel");
  *n_partials = bx * n;
  return 0;
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
  float* partial = (float*)workspace;
  int n_partials = 0;
  if (int rc = launch_frame_partial(gen, gt, partial, n, C, H, W, st, &n_partials)) return rc;
  frame_loss_final_kernel<<<1, 1024, 0, st>>>(partial, out2, n_partials, 1.0 / ((double)n * H * W));
  AMMC_LAUNCH_CHECK("frame_loss_final_kernel");
  return 0;
}

extern "C" int ammc_frame_losses_bwd(const float* gen, const float* gt, const float* g_int, const float* g_gd, float* grad_gen,
                                     int n, int C, int H, int W, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AMMC_REQUIRE(gen && gt && grad_gen && (g_int || g_gd) && n > 0 && H > 0 && W > 0, "bad argument");
  AMMC_REQUIRE(C >= 1 && C <= LOSS_MAX_C, "frame losses support 1..%d channels (got %d)", LOSS_MAX_C, C);
  AMMC_REQUIRE(n <= 65535, "batch %d too large for one launch", n);
  const bool quads = frame_quads_ok(gen, gt, grad_gen, W);
  const int bx = quads ? ceil_div(((int64_t)H * W) >> 2, 256) : ceil_div((int64_t)H * W, 256);
  const float inv = (float)(1.0 / ((double)n * H * W));
  switch (C) {
#define AMMC_FL_CASE(CC)                                                                                                     \
  case CC:                                                                                                                   \
    if (quads) frame_loss_bwd4_kernel<CC><<<dim3(bx, n), 256, 0, st>>>(gen, gt, g_int, g_gd, grad_gen, H, W, inv);           \
    else frame_loss_bwd_kernel<CC><<<dim3(bx, n), 256, 0, st>>>(gen, gt, g_int, g_gd, grad_gen, H, W, inv);                  \
    break;
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

// ------------------------------------------------------------------------------------------------------------------------
// Generator objective of the joint training step in one call (Twostream_vq_Loss.forward, Code/models/losses/loss_zoo.py:312-350):
//   g_loss = lam_adv * adv(d_gen) + lam_gdl * gd(rgb) + lam_flow * flow + lam_lp * int(rgb) + lam_latent * sum(latent) + lam_lp_op * int(op)
// The reference runs ~35 ATen kernels forward, as many backward, ~12 scalar kernels for the weighted sum and eight .item()
// synchronisations.  Here: four partial-sum passes (one per tensor pair) + ONE final kernel that reduces every segment in a
// fixed order and forms the weighted sum -> out8 = [g_loss, adv, flow, int, gd, int_op, latent, 0]; the backward is one scalar
// kernel (chain rule through the weighted sum) + one gradient pass per tensor that needs one.
// ------------------------------------------------------------------------------------------------------------------------
namespace ammc {

struct ObjSegments { int rgb, op, flow, adv, n_latent; };       // partial PAIRS per segment, laid out in this order

__global__ void __launch_bounds__(1024) gen_objective_final_kernel(const float* __restrict__ partial, const float* __restrict__ latent,
                                                                    ObjSegments seg, double inv_rgb, double inv_op, double inv_flow,
                                                                    double inv_adv, float lam_adv, float lam_gdl, float lam_flow,
                                                                    float lam_lp, float lam_latent, float lam_lp_op,
                                                                    float* __restrict__ out8) {
  __shared__ double red[6][32];
  double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};                 // int, gd, int_op, flow, adv, latent
  const float2* p = reinterpret_cast<const float2*>(partial);   // (a-term, b-term) pairs, 8-byte aligned segments
  for (int i = threadIdx.x; i < seg.rgb; i += 1024) { const float2 t = p[i]; v[0] += (double)t.x; v[1] += (double)t.y; }
  p += seg.rgb;
  for (int i = threadIdx.x; i < seg.op; i += 1024) v[2] += (double)p[i].x;
  p += seg.op;
  for (int i = threadIdx.x; i < seg.flow; i += 1024) v[3] += (double)p[i].x;
  p += seg.flow;
  for (int i = threadIdx.x; i < seg.adv; i += 1024) v[4] += (double)p[i].x;
  for (int i = threadIdx.x; i < seg.n_latent; i += 1024) v[5] += (double)latent[i];
#pragma unroll
  for (int q = 0; q < 6; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
    if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v[q];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s[6];
    for (int q = 0; q < 6; ++q) { s[q] = 0.0; for (int i = 0; i < 32; ++i) s[q] += red[q][i]; }
    const float g_int = (float)(s[0] * inv_rgb), g_gd = (float)(s[1] * inv_rgb), g_int_op = (float)(s[2] * inv_op);
    const float g_flow = (float)(s[3] * inv_flow), g_adv = (float)(s[4] * inv_adv), g_lat = (float)s[5];
    // same association as the reference expression (loss_zoo.py:336-339), fp32 like its tensors
    float g = lam_adv * g_adv + lam_gdl * g_gd;
    g += lam_flow * g_flow;
    g += lam_lp * g_int;
    g += lam_latent * g_lat;
    g += lam_lp_op * g_int_op;
    out8[0] = g; out8[1] = g_adv; out8[2] = g_flow; out8[3] = g_int; out8[4] = g_gd; out8[5] = g_int_op; out8[6] = g_lat;
    out8[7] = 0.f;
  }
}

// scal8 = d(sum_i g8[i] * out8[i]) / d(component): [-, adv, flow, int, gd, int_op, latent, -]
__global__ void gen_objective_scalars_kernel(const float* __restrict__ g8, float lam_adv, float lam_gdl, float lam_flow, float lam_lp,
                                             float lam_latent, float lam_lp_op, float* __restrict__ scal8) {
  if (threadIdx.x == 0) {
    const float g = g8[0];
    scal8[0] = g;
    scal8[1] = fmaf(g, lam_adv, g8[1]);
    scal8[2] = fmaf(g, lam_flow, g8[2]);
    scal8[3] = fmaf(g, lam_lp, g8[3]);
    scal8[4] = fmaf(g, lam_gdl, g8[4]);
    scal8[5] = fmaf(g, lam_lp_op, g8[5]);
    scal8[6] = fmaf(g, lam_latent, g8[6]);
    scal8[7] = 0.f;
  }
}

}  // namespace ammc

extern "C" size_t ammc_gen_objective_workspace_bytes(int n_rgb, int H_rgb, int W_rgb, int n_op, int H_op, int W_op, int64_t n_flow,
                                                     int64_t n_dgen) {
  const size_t pairs = (size_t)n_rgb * ceil_div((int64_t)H_rgb * W_rgb, 256) + (size_t)n_op * ceil_div((int64_t)H_op * W_op, 256) +
                       (size_t)elem_blocks(n_flow > 0 ? n_flow : 1) + (size_t)elem_blocks(n_dgen > 0 ? n_dgen : 1);
  return align_up(pairs * 2 * sizeof(float), 256);
}

extern "C" int ammc_gen_objective_fwd(const float* rgb_out, const float* rgb_tgt, const float* op_out, const float* op_tgt,
                                      const float* flow_pred, const float* flow_gt, const float* d_gen, const float* latent,
                                      int n_rgb, int C_rgb, int H_rgb, int W_rgb, int n_op, int C_op, int H_op, int W_op,
                                      int64_t n_flow, int64_t n_dgen, int n_latent, float lam_adv, float lam_gdl, float lam_flow,
                                      float lam_lp, float lam_latent, float lam_lp_op, float* out8, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AMMC_REQUIRE(rgb_out && rgb_tgt && op_out && op_tgt && flow_pred && flow_gt && d_gen && latent && out8, "bad argument");
  AMMC_REQUIRE(n_rgb > 0 && H_rgb > 0 && W_rgb > 0 && n_op > 0 && H_op > 0 && W_op > 0 && n_flow > 0 && n_dgen > 0 && n_latent > 0,
               "empty tensor");
  AMMC_REQUIRE(n_rgb <= 65535 && n_op <= 65535, "batch too large for one launch");
  if (!workspace || workspace_bytes < ammc_gen_objective_workspace_bytes(n_rgb, H_rgb, W_rgb, n_op, H_op, W_op, n_flow, n_dgen))
    return fail(AMMC_EWORKSPACE, "workspace too small");
  // segments follow each other in the workspace; their sizes are the partial PAIRS each pass writes
  ObjSegments seg;
  float* p_rgb = (float*)workspace;
  int rc = launch_frame_partial(rgb_out, rgb_tgt, p_rgb, n_rgb, C_rgb, H_rgb, W_rgb, st, &seg.rgb);
  if (rc) return rc;
  float* p_op = p_rgb + 2 * (size_t)seg.rgb;
  rc = launch_frame_partial(op_out, op_tgt, p_op, n_op, C_op, H_op, W_op, st, &seg.op);
  if (rc) return rc;
  seg.flow = elem_blocks(n_flow);
  seg.adv = elem_blocks(n_dgen);
  seg.n_latent = n_latent;
  float* p_flow = p_op + 2 * (size_t)seg.op;
  float* p_adv = p_flow + 2 * (size_t)seg.flow;
  if (((((uintptr_t)flow_pred) | ((uintptr_t)flow_gt)) & 15) == 0)
    elem_loss_partial_kernel<ELEM_L1, true><<<seg.flow, 256, 0, st>>>(flow_pred, flow_gt, p_flow, n_flow);
  else
    elem_loss_partial_kernel<ELEM_L1, false><<<seg.flow, 256, 0, st>>>(flow_pred, flow_gt, p_flow, n_flow);
  AMMC_LAUNCH_CHECK("elem_loss_partial_kernel");
  if ((((uintptr_t)d_gen) & 15) == 0)
    elem_loss_partial_kernel<ELEM_LSGAN_G, true><<<seg.adv, 256, 0, st>>>(d_gen, nullptr, p_adv, n_dgen);
  else
    elem_loss_partial_kernel<ELEM_LSGAN_G, false><<<seg.adv, 256, 0, st>>>(d_gen, nullptr, p_adv, n_dgen);
  AMMC_LAUNCH_CHECK("elem_loss_partial_kernel");
  gen_objective_final_kernel<<<1, 1024, 0, st>>>(p_rgb, latent, seg, 1.0 / ((double)n_rgb * H_rgb * W_rgb),
                                                1.0 / ((double)n_op * H_op * W_op), 1.0 / (double)n_flow, 1.0 / (double)n_dgen,
                                                lam_adv, lam_gdl, lam_flow, lam_lp, lam_latent, lam_lp_op, out8);
  AMMC_LAUNCH_CHECK("gen_objective_final_kernel");
  return 0;
}

extern "C" int ammc_gen_objective_bwd(const float* rgb_out, const float* rgb_tgt, const float* op_out, const float* op_tgt,
                                      const float* flow_pred, const float* flow_gt, const float* d_gen, const float* g8,
                                      int n_rgb, int C_rgb, int H_rgb, int W_rgb, int n_op, int C_op, int H_op, int W_op,
                                      int64_t n_flow, int64_t n_dgen, float lam_adv, float lam_gdl, float lam_flow, float lam_lp,
                                      float lam_latent, float lam_lp_op, float* scal8, float* grad_rgb, float* grad_op,
                                      float* grad_flow_pred, float* grad_d_gen, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AMMC_REQUIRE(g8 && scal8, "bad argument");
  gen_objective_scalars_kernel<<<1, 32, 0, st>>>(g8, lam_adv, lam_gdl, lam_flow, lam_lp, lam_latent, lam_lp_op, scal8);
  AMMC_LAUNCH_CHECK("gen_objective_scalars_kernel");
  int rc = 0;
  if (grad_rgb) {
    AMMC_REQUIRE(rgb_out && rgb_tgt, "bad argument");
    rc = ammc_frame_losses_bwd(rgb_out, rgb_tgt, scal8 + 3, scal8 + 4, grad_rgb, n_rgb, C_rgb, H_rgb, W_rgb, stream);
    if (rc) return rc;
  }
  if (grad_op) {
    AMMC_REQUIRE(op_out && op_tgt, "bad argument");
    rc = ammc_frame_losses_bwd(op_out, op_tgt, scal8 + 5, nullptr, grad_op, n_op, C_op, H_op, W_op, stream);
    if (rc) return rc;
  }
  if (grad_flow_pred) {
    rc = ammc_elem_loss_bwd(flow_pred, flow_gt, scal8 + 2, grad_flow_pred, nullptr, ELEM_L1, n_flow, stream);
    if (rc) return rc;
  }
  if (grad_d_gen) {
    rc = ammc_elem_loss_bwd(d_gen, nullptr, scal8 + 1, grad_d_gen, nullptr, ELEM_LSGAN_G, n_dgen, stream);
    if (rc) return rc;
  }
  return 0;
}
