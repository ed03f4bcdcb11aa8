// AMFT training path: batch-statistic BatchNorm (forward + backward through ReLU) and the convolution weight gradient
// on tcgen05.  Together with conv_igemm_kernel (forward and, with flipped/transposed weights, the data gradient) this is
// the autograd of reference Code/models/unet.py:8-20,956-965.
//
//   bn_stats_kernel        per-channel sum / sum of squares of a NCHW fp32 tensor (fp64 accumulation)
//   bn_finalize_kernel     mean, biased var -> scale = gamma*invstd, shift = beta - mean*scale; running stats (momentum,
//                          unbiased var) exactly as torch.nn.BatchNorm2d in training mode
//   bn_apply_pack_kernel   v = relu(y*scale + shift)  ->  NHWC bf16 hi/lo planes (next conv's A operand),
//                          NCHW bf16 hi/lo planes (wgrad operand) and/or fp32 NCHW (+ residual)
//   bn_bwd_reduce_kernel   sum g', sum g'*yhat  with g' = g * 1[relu active], yhat = (y - mean) * invstd
//   bn_bwd_apply_kernel    g_y = scale * (g' - mean(g') - yhat * mean(g' yhat))   (training)   or   scale * g' (eval BN)
//                          -> NHWC + NCHW bf16 planes; g_gamma, g_beta
//   conv_wgrad_kernel      gW[co][ci][tap] = sum_pixels gy[p][co] * x[p + tap][ci]:  GEMM with M = Cout (128/tile),
//                          N = Cin (256/tile), K = pixels (64 per block).  Both operands are the NHWC bf16 planes the
//                          forward/backward pipeline already holds; a [pixels][64 channels] TMA box is an MN-major UMMA
//                          operand (channels contiguous), so nothing is transposed.  The tap shift is a TMA box offset on
//                          the (w,h) axes with zero fill, like the forward conv.  (Pixel-contiguous NCHW operands are not
//                          usable: TMA needs a 128-byte inner box and 16-byte aligned inner coordinates -- measured.)
//                          Split-K over images, fp32 atomics into gW.
#include "common.cuh"
#include "ptx.cuh"
#include <cuda_bf16.h>

namespace ammc {

__device__ __forceinline__ void split_bf16_t(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// ------------------------------------------------------------------------------------------------
// BatchNorm forward statistics
// ------------------------------------------------------------------------------------------------
// grid (C, splits); block 256.  sums[c] += sum y, sums[C + c] += sum y^2 over the images of this split
// amax (optional, zeroed by the host): amax[c] = max |y| as the bits of a non-negative float -- behind the power-of-two scale
// of the q-format planes bn_apply_pack_kernel writes for the training path at precision 2
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ y, double* __restrict__ sums, int b, int C,
                                                        int hw, int imgs_per_split, unsigned* __restrict__ amax = nullptr) {
  __shared__ double red[2][8];
  const int c = blockIdx.x;
  const int i0 = blockIdx.y * imgs_per_split, i1 = min(b, i0 + imgs_per_split);
  double s = 0.0, s2 = 0.0;
  float am = 0.f;
  const bool vec = (hw & 3) == 0 && ((uintptr_t)y & 15) == 0;      // 16-byte loads: one 4 KB channel row per block pass
  for (int img = i0; img < i1; ++img) {
    const float* p = y + ((size_t)img * C + c) * hw;
    float fs = 0.f, fs2 = 0.f;
    if (vec) {
      const float4* p4 = reinterpret_cast<const float4*>(p);
      for (int i = threadIdx.x; i < (hw >> 2); i += 256) {
        const float4 v = __ldg(p4 + i);
        fs += (v.x + v.y) + (v.z + v.w);
        fs2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, fs2))));
        am = fmaxf(am, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
      }
    } else {
      for (int i = threadIdx.x; i < hw; i += 256) { const float v = p[i]; fs += v; fs2 = fmaf(v, v, fs2); am = fmaxf(am, fabsf(v)); }
    }
    s += (double)fs; s2 += (double)fs2;
  }
  if (amax) {
    am = warp_max(am);
    if ((threadIdx.x & 31) == 0 && am > 0.f) atomicMax(&amax[c], __float_as_uint(am));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { red[0][wid] = s; red[1][wid] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, a2 = 0.0;
    for (int i = 0; i < 8; ++i) { a += red[0][i]; a2 += red[1][i]; }
    atomicAdd(&sums[c], a);
    atomicAdd(&sums[C + c], a2);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_out, float* __restrict__ invstd_out, int C, double count,
                                   float momentum, float eps, int update_running, const unsigned* __restrict__ amax = nullptr,
                                   unsigned* __restrict__ bound_bits = nullptr) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = sums[c] / count;
  double var = sums[C + c] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - (float)mean * sc;
  mean_out[c] = (float)mean;
  invstd_out[c] = invstd;
  if (bound_bits)      // |act(y * scale + shift)| <= |scale| max|y| + |shift|: bound behind the activation's q scale
    atomicMax(bound_bits, __float_as_uint(fabsf(sc) * __uint_as_float(amax[c]) * 1.0001f + fabsf(beta[c] - (float)mean * sc)));
  if (update_running) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// eval-mode BN expressed with the same (scale, shift, mean, invstd) quadruple
__global__ void bn_eval_params_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                      float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                      float* __restrict__ invstd_out, int C, float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = 1.f / sqrtf(running_var[c] + eps);
  const float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - running_mean[c] * sc;
  mean_out[c] = running_mean[c];
  invstd_out[c] = invstd;
}

// ------------------------------------------------------------------------------------------------
// elementwise transforms with layout change: 32 (channels) x 32 (pixels) tiles through shared memory
// ------------------------------------------------------------------------------------------------
struct ApplyArgs {
  const float* y;          // [b][C][hw] fp32
  const float* g;          // gradient wrt the ReLU output (backward only) or null
  const float* scale; const float* shift; const float* mean; const float* invstd;
  const double* bsums;     // backward: [2][C] sum g', sum g' yhat
  double inv_count;        // 1 / (b*hw)
  int mode;                // 0: forward v = act(y*scale+shift); 1: backward (training BN); 2: backward (eval BN)
  int relu;
  __nv_bfloat16* nhwc;     // [2][b][hw][C] or null
  __nv_bfloat16* nchw;     // [2][b][C][hw] or null
  float* out_f32;          // [b][C][hw] or null
  const float* res;        // added to out_f32, or null
  long long plane_stride;
  // q-format copy of the NHWC output (precision 2 operand of the next forward / data-gradient conv), or null:
  __half* q16;             // [b][hw][C] fp16 plane; the e4m3 planes follow at q8 / q8 + plane_stride
  uint8_t* q8;
  const unsigned* q_bound; // bits of a bound on max |v| (bn_finalize_kernel / bn_param_grads_kernel)
  float* q_scale_out;      // scale slot of the q buffer (written by one thread)
};

// Tile = 64 channels x 32 pixels.  Phase 1 walks the NCHW side (lane = pixel: 128-byte coalesced reads of y / g / res and
// writes of out_f32 / the NCHW planes); phase 2 walks the NHWC side with one (pixel, 8-channel group) per thread: 16-byte
// stores, 128 contiguous bytes per 8 lanes (round 1 stored 2 bytes per lane = 64 bytes per warp instruction and ran at
// 40 % of the HBM peak).  C % 8 == 0 on the NHWC path (the conv engine needs C % 64 == 0 anyway).
__global__ void __launch_bounds__(256) bn_apply_pack_kernel(const ApplyArgs a, int C, int HW) {
  __shared__ float tile[64][33];
  const int img = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int pp1 = p0 + tx;
  // all loads of the thread's eight (channel, pixel) elements first -- 8 to 24 independent requests in flight -- then the
  // arithmetic and the stores
  float yv[8], gv[8], rv[8];
  bool ok[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + ty + 8 * i;
    ok[i] = c < C && pp1 < HW;
    const size_t o = ((size_t)img * C + c) * HW + pp1;
    yv[i] = ok[i] ? __ldg(a.y + o) : 0.f;
    gv[i] = (ok[i] && a.mode != 0) ? __ldg(a.g + o) : 0.f;
    rv[i] = (ok[i] && a.out_f32 && a.res) ? __ldg(a.res + o) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + ty + 8 * i;
    float v = 0.f;
    if (ok[i]) {
      const size_t o = ((size_t)img * C + c) * HW + pp1;
      const float sc = __ldg(a.scale + c);
      const float act = fmaf(yv[i], sc, __ldg(a.shift + c));
      if (a.mode == 0) {
        v = a.relu ? fmaxf(act, 0.f) : act;
      } else {
        float gp = gv[i];
        if (a.relu && !(act > 0.f)) gp = 0.f;
        if (a.mode == 1) {
          const float yhat = (yv[i] - __ldg(a.mean + c)) * __ldg(a.invstd + c);
          const float mg = (float)(a.bsums[c] * a.inv_count), mgy = (float)(a.bsums[C + c] * a.inv_count);
          v = sc * (gp - mg - yhat * mgy);
        } else {
          v = sc * gp;
        }
      }
      if (a.out_f32) a.out_f32[o] = v + rv[i];
      if (a.nchw) {
        __nv_bfloat16 hi, lo;
        split_bf16_t(v, hi, lo);
        a.nchw[o] = hi;
        a.nchw[o + a.plane_stride] = lo;
      }
    }
    tile[ty + 8 * i][tx] = v;
  }
  if (!a.nhwc && !a.q16) return;
  __syncthreads();
  const int px = threadIdx.x >> 3, cg = threadIdx.x & 7;         // pixel within the tile, 8-channel group
  const int pp = p0 + px, c = c0 + cg * 8;
  float qs = 1.f;
  if (a.q16) {
    qs = q_scale_for_bound(__uint_as_float(__ldg(a.q_bound)));
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) *a.q_scale_out = qs;
  }
  if (pp < HW && c < C) {
    const size_t o = ((size_t)img * HW + pp) * C + c;
    if (a.nhwc) {
      uint32_t hp[4], lp[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) ptx::split_pack_bf16x2(tile[cg * 8 + 2 * j][px], tile[cg * 8 + 2 * j + 1][px], hp[j], lp[j]);
      __nv_bfloat16* dst = a.nhwc + o;
      *reinterpret_cast<uint4*>(dst) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
      *reinterpret_cast<uint4*>(dst + a.plane_stride) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
    }
    if (a.q16) {           // 8 channels = 16 B of fp16 + 8 B of h8 + 8 B of l8
      uint32_t h16[4], h8[4], l8[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        ptx::split_pack_q(tile[cg * 8 + 2 * j][px] * qs, tile[cg * 8 + 2 * j + 1][px] * qs, h16[j], h8[j], l8[j]);
      *reinterpret_cast<uint4*>(a.q16 + o) = make_uint4(h16[0], h16[1], h16[2], h16[3]);
      *reinterpret_cast<uint2*>(a.q8 + o) = make_uint2(h8[0] | (h8[1] << 16), h8[2] | (h8[3] << 16));
      *reinterpret_cast<uint2*>(a.q8 + a.plane_stride + o) = make_uint2(l8[0] | (l8[1] << 16), l8[2] | (l8[3] << 16));
    }
  }
}

// Same transform for HW % 64 == 0, C % 64 == 0, no NCHW planes (the shipped shapes): tile = 64 channels x 64 pixels, the
// NCHW side is walked with 16-byte accesses (256 contiguous bytes per channel row and pass instead of 128).
__device__ __forceinline__ float bn_apply_one(const ApplyArgs& a, float yv, float gv, float sc, float sh, float mu, float is,
                                              float mg, float mgy) {
  const float act = fmaf(yv, sc, sh);
  if (a.mode == 0) return a.relu ? fmaxf(act, 0.f) : act;
  float gp = gv;
  if (a.relu && !(act > 0.f)) gp = 0.f;
  if (a.mode == 1) return sc * (gp - mg - (yv - mu) * is * mgy);
  return sc * gp;
}

__global__ void __launch_bounds__(256, 4) bn_apply_pack64_kernel(const ApplyArgs a, int C, int HW) {
  __shared__ float tile[64][65];
  const int img = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;       // 16 float4 lanes per channel row, 16 rows per pass
  float4 yv[4], gv[4], rv[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const size_t o = ((size_t)img * C + c0 + ty + 16 * i) * HW + p0 + 4 * tx;
    yv[i] = __ldg(reinterpret_cast<const float4*>(a.y + o));
    gv[i] = a.mode != 0 ? __ldg(reinterpret_cast<const float4*>(a.g + o)) : make_float4(0.f, 0.f, 0.f, 0.f);
    rv[i] = (a.out_f32 && a.res) ? __ldg(reinterpret_cast<const float4*>(a.res + o)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty + 16 * i, c = c0 + r;
    const float sc = __ldg(a.scale + c), sh = __ldg(a.shift + c);
    float mu = 0.f, is = 0.f, mg = 0.f, mgy = 0.f;
    if (a.mode == 1) {
      mu = __ldg(a.mean + c); is = __ldg(a.invstd + c);
      mg = (float)(a.bsums[c] * a.inv_count); mgy = (float)(a.bsums[C + c] * a.inv_count);
    }
    float4 v;
    v.x = bn_apply_one(a, yv[i].x, gv[i].x, sc, sh, mu, is, mg, mgy);
    v.y = bn_apply_one(a, yv[i].y, gv[i].y, sc, sh, mu, is, mg, mgy);
    v.z = bn_apply_one(a, yv[i].z, gv[i].z, sc, sh, mu, is, mg, mgy);
    v.w = bn_apply_one(a, yv[i].w, gv[i].w, sc, sh, mu, is, mg, mgy);
    if (a.out_f32) {
      const size_t o = ((size_t)img * C + c) * HW + p0 + 4 * tx;
      *reinterpret_cast<float4*>(a.out_f32 + o) = make_float4(v.x + rv[i].x, v.y + rv[i].y, v.z + rv[i].z, v.w + rv[i].w);
    }
    tile[r][4 * tx + 0] = v.x; tile[r][4 * tx + 1] = v.y; tile[r][4 * tx + 2] = v.z; tile[r][4 * tx + 3] = v.w;
  }
  if (!a.nhwc && !a.q16) return;
  __syncthreads();
  float qs = 1.f;
  if (a.q16) {
    qs = q_scale_for_bound(__uint_as_float(__ldg(a.q_bound)));
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) *a.q_scale_out = qs;
  }
  const int cg = threadIdx.x & 7, c = c0 + cg * 8;
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int px = (threadIdx.x >> 3) + 32 * hh;               // pixel within the tile
    const size_t o = ((size_t)img * HW + p0 + px) * C + c;
    if (a.nhwc) {
      uint32_t hp[4], lp[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) ptx::split_pack_bf16x2(tile[cg * 8 + 2 * j][px], tile[cg * 8 + 2 * j + 1][px], hp[j], lp[j]);
      __nv_bfloat16* dst = a.nhwc + o;
      *reinterpret_cast<uint4*>(dst) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
      *reinterpret_cast<uint4*>(dst + a.plane_stride) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
    }
    if (a.q16) {
      uint32_t h16[4], h8[4], l8[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        ptx::split_pack_q(tile[cg * 8 + 2 * j][px] * qs, tile[cg * 8 + 2 * j + 1][px] * qs, h16[j], h8[j], l8[j]);
      *reinterpret_cast<uint4*>(a.q16 + o) = make_uint4(h16[0], h16[1], h16[2], h16[3]);
      *reinterpret_cast<uint2*>(a.q8 + o) = make_uint2(h8[0] | (h8[1] << 16), h8[2] | (h8[3] << 16));
      *reinterpret_cast<uint2*>(a.q8 + a.plane_stride + o) = make_uint2(l8[0] | (l8[1] << 16), l8[2] | (l8[3] << 16));
    }
  }
}

// sums[c] += sum g', sums[C+c] += sum g' * yhat
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                                             const float* __restrict__ scale, const float* __restrict__ shift,
                                                             const float* __restrict__ mean, const float* __restrict__ invstd,
                                                             int relu, double* __restrict__ sums, int b, int C, int hw,
                                                             int imgs_per_split, unsigned* __restrict__ amax = nullptr) {
  __shared__ double red[2][8];
  const int c = blockIdx.x;
  const int i0 = blockIdx.y * imgs_per_split, i1 = min(b, i0 + imgs_per_split);
  const float sc = scale[c], sh = shift[c], mu = mean[c], is = invstd[c];
  double s = 0.0, s2 = 0.0;
  float mg = 0.f, myh = 0.f;       // max |g'|, max |yhat|: behind the q scale of the gradient planes (amax[c], amax[C + c])
  for (int img = i0; img < i1; ++img) {
    const size_t base = ((size_t)img * C + c) * hw;
    float fs = 0.f, fs2 = 0.f;
    for (int i = threadIdx.x; i < hw; i += 256) {
      const float yv = y[base + i];
      float gp = g[base + i];
      if (relu && !(fmaf(yv, sc, sh) > 0.f)) gp = 0.f;
      const float yh = (yv - mu) * is;
      fs += gp;
      fs2 = fmaf(gp, yh, fs2);
      mg = fmaxf(mg, fabsf(gp));
      myh = fmaxf(myh, fabsf(yh));
    }
    s += (double)fs; s2 += (double)fs2;
  }
  if (amax) {
    mg = warp_max(mg);
    myh = warp_max(myh);
    if ((threadIdx.x & 31) == 0) {
      if (mg > 0.f) atomicMax(&amax[c], __float_as_uint(mg));
      if (myh > 0.f) atomicMax(&amax[C + c], __float_as_uint(myh));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { red[0][wid] = s; red[1][wid] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, a2 = 0.0;
    for (int i = 0; i < 8; ++i) { a += red[0][i]; a2 += red[1][i]; }
    atomicAdd(&sums[c], a);
    atomicAdd(&sums[C + c], a2);
  }
}

// bound_bits (optional): max_c |scale_c| (max|g'| + |mean g'| + max|yhat| |mean g' yhat|) >= max |g_y| (training BN; eval BN:
// |scale_c| max|g'|) -- behind the q scale of the gradient planes
__global__ void bn_param_grads_kernel(const double* __restrict__ sums, float* __restrict__ g_gamma,
                                      float* __restrict__ g_beta, int C, const unsigned* __restrict__ amax = nullptr,
                                      const float* __restrict__ scale = nullptr, double inv_count = 0.0, int training = 1,
                                      unsigned* __restrict__ bound_bits = nullptr) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  g_beta[c] = (float)sums[c];
  g_gamma[c] = (float)sums[C + c];
  if (bound_bits) {
    const float mg = __uint_as_float(amax[c]), myh = __uint_as_float(amax[C + c]);
    float bnd = mg;
    if (training) bnd += fabsf((float)(sums[c] * inv_count)) + myh * fabsf((float)(sums[C + c] * inv_count));
    atomicMax(bound_bits, __float_as_uint(fabsf(scale[c]) * bnd * 1.0001f));
  }
}

// x fp32 (any layout) -> same layout bf16 hi/lo planes
__global__ void pack_planes_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xp, long long total) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  __nv_bfloat16 hi, lo;
  split_bf16_t(x[e], hi, lo);
  xp[e] = hi;
  xp[e + total] = lo;
}

// w [Cout][Cin][3][3] -> wp [2][Cin][9*Cout] with k = tap'*Cout + co holding w[co][ci][8 - tap']  (data-gradient conv)
__global__ void pack_weights_dgrad_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int Cout, int Cin) {
  const long long total = (long long)Cout * Cin * 9;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int co = (int)(e % Cout);
  const int tp = (int)((e / Cout) % 9);
  const int ci = (int)(e / ((long long)Cout * 9));
  __nv_bfloat16 hi, lo;
  split_bf16_t(w[((size_t)co * Cin + ci) * 9 + (8 - tp)], hi, lo);
  wp[e] = hi;
  wp[e + total] = lo;
}

// ------------------------------------------------------------------------------------------------
// weight gradient on tcgen05
// ------------------------------------------------------------------------------------------------
constexpr int WG_THREADS = 192;
constexpr int WG_BLOCK_M = 128;   // output channels  (2 boxes of 64)
constexpr int WG_BLOCK_N = 256;   // input channels   (4 boxes of 64)
constexpr int WG_BLOCK_K = 64;    // pixels per pipeline stage
constexpr int WG_BOX_BYTES = WG_BLOCK_K * 64 * 2;          // [64 px][64 ch] bf16 = 8 KB
constexpr int WG_A_BYTES = (WG_BLOCK_M / 64) * WG_BOX_BYTES;
constexpr int WG_B_BYTES = (WG_BLOCK_N / 64) * WG_BOX_BYTES;
// A stage holds the hi AND lo planes of both operands: [A hi][A lo][B hi][B lo] = 96 KB, loaded once per K block and used
// by the three MMA passes (hi,hi) (hi,lo) (lo,hi) -- 8 KB of L2 -> shared-memory traffic per MMA instead of 12 KB when
// every pass streamed K on its own.  That traffic, not the tensor pipe, bounded the kernel: 10.6 GB per launch at the
// shipped shape = 12.8 TB/s out of L2 over 0.83 ms.
constexpr int WG_STAGE_BYTES = 2 * (WG_A_BYTES + WG_B_BYTES);    // 96 KB
constexpr int WG_STAGES = 2;
constexpr int WG_BAR_OFFSET = WG_STAGES * WG_STAGE_BYTES;
constexpr int WG_SMEM = WG_BAR_OFFSET + 256 + 1024;

struct WgradParams {
  int b, H, W, Cin, Cout;
  int w_box;                // pixels of one row covered by a K block: min(W, 64)
  int rows_per_kb;          // rows per K block: 64 / W for W <= 64, else 1
  int chunks_per_row;       // ceil(W / 64): K blocks needed to cover one row (1 for W <= 64)
  int kb_per_img;           // ceil(H / rows_per_kb) * chunks_per_row
  int k_bytes;              // bytes one [pixels][64 ch] box really transfers: w_box * rows_per_kb * 128
  int tiles_m, tiles_n, splits, imgs_per_split;
  int n_pass;               // 1: hi*hi only; 3: (hi,hi) + (hi,lo) + (lo,hi)
  int block_n;              // input channels actually present in an N tile (Cin may be < 256)
  int ntaps;                // 9: 3x3 conv; 1: 1x1 conv (a plain [Cout x pixels] . [pixels x Cin] GEMM)
  float* gw;                // [Cout][Cin][ntaps]
};

__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WG_BAR_OFFSET);
  uint64_t* empty_bar = full_bar + WG_STAGES;
  uint64_t* tmem_full = empty_bar + WG_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int work_items = p.tiles_m * p.tiles_n * p.ntaps * p.splits;
  const int ntaps = p.ntaps;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    for (int s = 0; s < WG_STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tmem_full[a], 1); ptx::mbar_init(&tmem_empty[a], 128); }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 2 * WG_BLOCK_N);
    ptx::tmem_relinquish();
  }
  // K = pixels: when a box covers fewer than 64 pixels the remaining K rows of every stage are never written by TMA and
  // must contribute zero -> clear the ring once (generic-proxy writes, made visible to the async proxy by the fence)
  if (p.k_bytes < WG_BOX_BYTES) {
    for (int i = threadIdx.x * 16; i < WG_STAGES * WG_STAGE_BYTES; i += WG_THREADS * 16)
      *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    ptx::fence_proxy_async();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work item -> (split, tap, n tile, m tile).  NOTE: no by-reference lambdas in this kernel -- capturing the kernel
  // parameters moves the __grid_constant__ tensor maps to the local stack, and TMA cannot read a descriptor from there.
  const int tiles_m = p.tiles_m, tiles_n = p.tiles_n;
#define WG_DECODE(wi_, mt_, nt_, tap_, sp_)            \
  int mt_, nt_, tap_, sp_;                             \
  {                                                    \
    int w_ = (wi_);                                    \
    mt_ = w_ % tiles_m; w_ /= tiles_m;                 \
    nt_ = w_ % tiles_n; w_ /= tiles_n;                 \
    tap_ = w_ % ntaps;  sp_ = w_ / ntaps;              \
  }

  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int wi = blockIdx.x; wi < work_items; wi += gridDim.x) {
        WG_DECODE(wi, mt, nt, tap, sp)
        const int dy = ntaps == 9 ? tap / 3 - 1 : 0, dx = ntaps == 9 ? tap % 3 - 1 : 0;
        const int i0 = sp * p.imgs_per_split, i1 = min(p.b, i0 + p.imgs_per_split);
        const int planes = p.n_pass == 3 ? 2 : 1;
        for (int img = i0; img < i1; ++img) {
          for (int kb = 0; kb < p.kb_per_img; ++kb) {
            const int h0 = (kb / p.chunks_per_row) * p.rows_per_kb, w0 = (kb % p.chunks_per_row) * 64;
            ptx::mbar_wait(&empty_bar[s], ph ^ 1, 21);
            ptx::mbar_expect_tx(&full_bar[s], planes * (WG_BLOCK_M / 64 + WG_BLOCK_N / 64) * p.k_bytes);
            uint8_t* dst = smem + s * WG_STAGE_BYTES;
            for (int pl = 0; pl < planes; ++pl) {
#pragma unroll
              for (int i = 0; i < WG_BLOCK_M / 64; ++i)
                ptx::tma_load_5d(dst + pl * WG_A_BYTES + i * WG_BOX_BYTES, &tmA, &full_bar[s], mt * WG_BLOCK_M + i * 64, w0, h0,
                                 img, pl);
#pragma unroll
              for (int i = 0; i < WG_BLOCK_N / 64; ++i)
                ptx::tma_load_5d(dst + 2 * WG_A_BYTES + pl * WG_B_BYTES + i * WG_BOX_BYTES, &tmB, &full_bar[s],
                                 nt * WG_BLOCK_N + i * 64, w0 + dx, h0 + dy, img, pl);
            }
            if (++s == WG_STAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_mn(1, WG_BLOCK_M, WG_BLOCK_N);
      int s = 0; uint32_t ph = 0;
      int it = 0;
      for (int wi = blockIdx.x; wi < work_items; wi += gridDim.x, ++it) {
        WG_DECODE(wi, mt, nt, tap, sp)
        (void)mt; (void)nt; (void)tap;
        const int i0 = sp * p.imgs_per_split, i1 = min(p.b, i0 + p.imgs_per_split);
        const int kb_total = (i1 - i0) * p.kb_per_img;
        const bool three = p.n_pass == 3;
        const int acc = it & 1;
        const uint32_t acc_ph = (it >> 1) & 1;
        ptx::mbar_wait(&tmem_empty[acc], acc_ph ^ 1, 22);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * WG_BLOCK_N;
        for (int kb = 0; kb < kb_total; ++kb) {
          ptx::mbar_wait(&full_bar[s], ph, 23);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem + s * WG_STAGE_BYTES);
          const uint64_t a_hi = ptx::umma_desc_mn_sw128(a_addr, WG_BOX_BYTES);
          const uint64_t a_lo = ptx::umma_desc_mn_sw128(a_addr + WG_A_BYTES, WG_BOX_BYTES);
          const uint64_t b_hi = ptx::umma_desc_mn_sw128(a_addr + 2 * WG_A_BYTES, WG_BOX_BYTES);
          const uint64_t b_lo = ptx::umma_desc_mn_sw128(a_addr + 2 * WG_A_BYTES + WG_B_BYTES, WG_BOX_BYTES);
#pragma unroll
          for (int k4 = 0; k4 < WG_BLOCK_K / 16; ++k4) {
            // 16 pixels (K) = 16 rows of 128 B = 2048 B further down each box: +128 in the (>>4) address field
            ptx::mma_f16_ss(d_tmem, a_hi + 128 * k4, b_hi + 128 * k4, idesc, (kb | k4) != 0 ? 1u : 0u);
            if (three) {
              ptx::mma_f16_ss(d_tmem, a_hi + 128 * k4, b_lo + 128 * k4, idesc, 1u);
              ptx::mma_f16_ss(d_tmem, a_lo + 128 * k4, b_hi + 128 * k4, idesc, 1u);
            }
          }
          ptx::mma_commit(&empty_bar[s]);
          if (kb == kb_total - 1) ptx::mma_commit(&tmem_full[acc]);
          if (++s == WG_STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;                 // accumulator row = output channel within the tile
    int it = 0;
    for (int wi = blockIdx.x; wi < work_items; wi += gridDim.x, ++it) {
      WG_DECODE(wi, mt, nt, tap, sp)
      (void)sp;
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      const int co = mt * WG_BLOCK_M + r;
      ptx::mbar_wait(&tmem_full[acc], acc_ph, 24);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * WG_BLOCK_N;
#pragma unroll 1
      for (int c32 = 0; c32 < WG_BLOCK_N / 32; ++c32) {
        if (c32 * 32 >= p.block_n) break;        // uniform: Cin tile narrower than 256
        uint32_t v[32];
        ptx::tmem_ld_32x32(taddr + c32 * 32, v);
        ptx::tmem_ld_wait();
        if (co < p.Cout) {
          const int ci0 = nt * WG_BLOCK_N + c32 * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (ci0 + j < p.Cin) atomicAdd(&p.gw[((size_t)co * p.Cin + ci0 + j) * ntaps + tap], __uint_as_float(v[j]));
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tmem_empty[acc]);
    }
  }
#undef WG_DECODE
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 2 * WG_BLOCK_N);
  }
}

int make_map_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                  const uint32_t* box);   // amft_conv.cu

#ifdef AMMC_DEBUG_PROBES
// Debug: load ONE 5-D box into shared memory with TMA and dump the raw bytes (layout / fault investigations).
__global__ void tma_probe_kernel(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, int c3, int c4,
                                 uint32_t box_bytes, uint8_t* out, uint32_t out_bytes) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + out_bytes);
  for (uint32_t i = threadIdx.x; i < out_bytes; i += blockDim.x) smem[i] = 0xEE;
  __syncthreads();
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_mbar_init();
    ptx::fence_proxy_async();
    ptx::mbar_expect_tx(bar, box_bytes);
    ptx::tma_load_5d(smem, &tm, bar, c0, c1, c2, c3, c4);
    ptx::mbar_wait(bar, 0, 31);
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < out_bytes; i += blockDim.x) out[i] = smem[i];
}
#endif  // AMMC_DEBUG_PROBES

}  // namespace ammc

namespace ammc { AMMC_DEFINE_TIMEOUT_READER(timeout_reader_train) }

using namespace ammc;

#ifdef AMMC_DEBUG_PROBES
extern "C" int ammc_debug_tma_probe(const void* base, const int64_t* dims5, const int64_t* strides4_bytes,
                                    const int* box5, const int* coords5, void* out, int out_bytes, void* stream) {
  uint64_t d[5], s[4];
  uint32_t bx[5];
  uint64_t box_elems = 1;
  for (int i = 0; i < 5; ++i) { d[i] = (uint64_t)dims5[i]; bx[i] = (uint32_t)box5[i]; box_elems *= bx[i]; }
  for (int i = 0; i < 4; ++i) s[i] = (uint64_t)strides4_bytes[i];
  CUtensorMap tm;
  if (int rc = make_map_bf16(&tm, base, 5, d, s, bx)) return rc;
  AMMC_REQUIRE(out_bytes >= (int)(box_elems * 2) && out_bytes <= 200 * 1024, "bad out_bytes");
  AMMC_CUDA_CHECK(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, out_bytes + 2048));
  tma_probe_kernel<<<1, 128, out_bytes + 2048, (cudaStream_t)stream>>>(tm, coords5[0], coords5[1], coords5[2], coords5[3],
                                                                        coords5[4], (uint32_t)(box_elems * 2),
                                                                        (uint8_t*)out, (uint32_t)out_bytes);
  AMMC_LAUNCH_CHECK("tma_probe_kernel");
  return 0;
}
#endif  // AMMC_DEBUG_PROBES

extern "C" int ammc_bn_batch_stats_staged(const float* y, const float* gamma, const float* beta, float* running_mean,
                                          float* running_var, float* scale, float* shift, float* mean, float* invstd,
                                          void* workspace, size_t workspace_bytes, int b, int C, int h, int w,
                                          float momentum, float eps, int training, int stage, double count_total,
                                          void* stream);

extern "C" int ammc_bn_batch_stats(const float* y, const float* gamma, const float* beta, float* running_mean,
                                   float* running_var, float* scale, float* shift, float* mean, float* invstd,
                                   void* workspace, size_t workspace_bytes, int b, int C, int h, int w, float momentum,
                                   float eps, int training, void* stream) {
  return ammc_bn_batch_stats_staged(y, gamma, beta, running_mean, running_var, scale, shift, mean, invstd, workspace,
                                    workspace_bytes, b, C, h, w, momentum, eps, training, 0, 0.0, stream);
}

// stage 0: everything in one call (per-rank statistics).  Data-parallel training with GLOBAL-batch statistics (what a
// single-GPU run of the reference computes, unet.py:11-16): stage 1 leaves the per-channel (sum, sum of squares) as 2*C
// doubles in the workspace, the caller all-reduces them, stage 2 finishes with `count_total` = elements per channel over
// all ranks.
extern "C" int ammc_bn_batch_stats_staged(const float* y, const float* gamma, const float* beta, float* running_mean,
                                          float* running_var, float* scale, float* shift, float* mean, float* invstd,
                                          void* workspace, size_t workspace_bytes, int b, int C, int h, int w,
                                          float momentum, float eps, int training, int stage, double count_total,
                                          void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AMMC_REQUIRE(gamma && beta && running_mean && running_var && scale && shift && mean && invstd && C > 0, "bad argument");
  AMMC_REQUIRE(stage >= 0 && stage <= 2, "stage must be 0 (all), 1 (local sums) or 2 (finalize)");
  if (!training) {
    bn_eval_params_kernel<<<ceil_div(C, 128), 128, 0, st>>>(gamma, beta, running_mean, running_var, scale, shift, mean,
                                                            invstd, C, eps);
    AMMC_LAUNCH_CHECK("bn_eval_params_kernel");
    return 0;
  }
  AMMC_REQUIRE(y && b > 0 && h > 0 && w > 0, "bad argument");
  if (!workspace || workspace_bytes < (size_t)2 * C * sizeof(double)) return fail(AMMC_EWORKSPACE, "workspace too small");
  double* sums = (double*)workspace;
  if (stage != 2) {
    AMMC_CUDA_CHECK(cudaMemsetAsync(sums, 0, (size_t)2 * C * sizeof(double), st));
    const int splits = max(1, min(b, ceil_div(4 * num_sms(), C)));
    const int per = ceil_div(b, splits);
    bn_stats_kernel<<<dim3(C, ceil_div(b, per)), 256, 0, st>>>(y, sums, b, C, h * w, per);
    AMMC_LAUNCH_CHECK("bn_stats_kernel");
    if (stage == 1) return 0;
  }
  const double count = stage == 2 ? count_total : (double)b * h * w;
  AMMC_REQUIRE(count >= 1.0, "count_total must be the number of elements per channel over all ranks");
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>(sums, gamma, beta, running_mean, running_var, scale, shift, mean,
                                                       invstd, C, count, momentum, eps, 1);
  AMMC_LAUNCH_CHECK("bn_finalize_kernel");
  return 0;
}

static int launch_apply(const ApplyArgs& a, int b, int C, int h, int w, cudaStream_t st) {
  AMMC_REQUIRE(b <= 65535, "batch %d too large for one launch", b);
  const int HW = h * w;
  AMMC_REQUIRE(!a.nhwc || C % 8 == 0, "NHWC plane output needs C %% 8 == 0 (got %d)", C);
  const bool aligned16 = (((uintptr_t)a.y | (uintptr_t)a.g | (uintptr_t)a.res | (uintptr_t)a.out_f32) & 15) == 0;
  if (HW % 64 == 0 && C % 64 == 0 && !a.nchw && aligned16) {
    bn_apply_pack64_kernel<<<dim3(HW / 64, C / 64, b), 256, 0, st>>>(a, C, HW);
    AMMC_LAUNCH_CHECK("bn_apply_pack64_kernel");
    return 0;
  }
  bn_apply_pack_kernel<<<dim3(ceil_div(HW, 32), ceil_div(C, 64), b), 256, 0, st>>>(a, C, HW);
  AMMC_LAUNCH_CHECK("bn_apply_pack_kernel");
  return 0;
}

extern "C" int ammc_bn_apply(const float* y, const float* scale, const float* shift, int relu, void* out_nhwc_planes,
                             void* out_nchw_planes, float* out_f32, const float* res, int b, int C, int h, int w,
                             void* stream) {
  AMMC_REQUIRE(y && scale && shift && (out_nhwc_planes || out_nchw_planes || out_f32), "bad argument");
  ApplyArgs a{};
  a.y = y; a.scale = scale; a.shift = shift; a.mode = 0; a.relu = relu;
  a.nhwc = (__nv_bfloat16*)out_nhwc_planes; a.nchw = (__nv_bfloat16*)out_nchw_planes; a.out_f32 = out_f32; a.res = res;
  a.plane_stride = (long long)b * C * h * w;
  return launch_apply(a, b, C, h, w, (cudaStream_t)stream);
}

extern "C" int ammc_bn_backward_staged(const float* g, const float* y, const float* scale, const float* shift,
                                       const float* mean, const float* invstd, int relu, int training, void* gy_nhwc_planes,
                                       void* gy_nchw_planes, float* g_gamma, float* g_beta, void* workspace,
                                       size_t workspace_bytes, int b, int C, int h, int w, int stage, double count_total,
                                       void* stream);

extern "C" int ammc_bn_backward(const float* g, const float* y, const float* scale, const float* shift, const float* mean,
                                const float* invstd, int relu, int training, void* gy_nhwc_planes, void* gy_nchw_planes,
                                float* g_gamma, float* g_beta, void* workspace, size_t workspace_bytes, int b, int C, int h,
                                int w, void* stream) {
  return ammc_bn_backward_staged(g, y, scale, shift, mean, invstd, relu, training, gy_nhwc_planes, gy_nchw_planes, g_gamma,
                                 g_beta, workspace, workspace_bytes, b, C, h, w, 0, 0.0, stream);
}

// stage 0: one call.  Global-batch statistics: stage 1 reduces this rank's (sum g', sum g' * yhat) into the workspace and
// writes the LOCAL parameter gradients (they are all-reduced with the other gradients); the caller all-reduces the 2*C
// doubles; stage 2 applies the BatchNorm backward with the global sums and `count_total`.
extern "C" int ammc_bn_backward_staged(const float* g, const float* y, const float* scale, const float* shift,
                                       const float* mean, const float* invstd, int relu, int training, void* gy_nhwc_planes,
                                       void* gy_nchw_planes, float* g_gamma, float* g_beta, void* workspace,
                                       size_t workspace_bytes, int b, int C, int h, int w, int stage, double count_total,
                                       void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AMMC_REQUIRE(g && y && scale && shift && mean && invstd && g_gamma && g_beta && (gy_nhwc_planes || gy_nchw_planes),
               "bad argument");
  AMMC_REQUIRE(stage >= 0 && stage <= 2, "stage must be 0 (all), 1 (local sums) or 2 (apply)");
  if (!workspace || workspace_bytes < (size_t)2 * C * sizeof(double)) return fail(AMMC_EWORKSPACE, "workspace too small");
  double* sums = (double*)workspace;
  if (stage != 2) {
    AMMC_CUDA_CHECK(cudaMemsetAsync(sums, 0, (size_t)2 * C * sizeof(double), st));
    const int splits = max(1, min(b, ceil_div(4 * num_sms(), C)));
    const int per = ceil_div(b, splits);
    bn_bwd_reduce_kernel<<<dim3(C, ceil_div(b, per)), 256, 0, st>>>(g, y, scale, shift, mean, invstd, relu, sums, b, C,
                                                                    h * w, per);
    AMMC_LAUNCH_CHECK("bn_bwd_reduce_kernel");
    bn_param_grads_kernel<<<ceil_div(C, 128), 128, 0, st>>>(sums, g_gamma, g_beta, C);
    AMMC_LAUNCH_CHECK("bn_param_grads_kernel");
    if (stage == 1) return 0;
  }
  ApplyArgs a{};
  a.y = y; a.g = g; a.scale = scale; a.shift = shift; a.mean = mean; a.invstd = invstd; a.bsums = sums;
  a.inv_count = 1.0 / (stage == 2 ? count_total : (double)b * h * w);
  a.mode = training ? 1 : 2; a.relu = relu;
  a.nhwc = (__nv_bfloat16*)gy_nhwc_planes; a.nchw = (__nv_bfloat16*)gy_nchw_planes;
  a.plane_stride = (long long)b * C * h * w;
  return launch_apply(a, b, C, h, w, st);
}

// ------------------------------------------------------------------------------------------------
// q-format variants (training at precision 2: the forward and data-gradient convs take fp16 + e4m3 operands).  The BN
// kernels already see every value that ends up in an operand plane, so the power-of-two scale of the q planes comes out
// of their reductions as a BOUND (loose bounds cost nothing): per-channel max|y| in the statistics pass,
// max|g'| and max|yhat| in the backward reduction.  Per-rank statistics only (no staged form).
//   workspace: [2C doubles: sums][2C uint32: maxima][uint32: bound bits]  ->  ammc_bn_q_workspace_bytes(C)
// ------------------------------------------------------------------------------------------------
extern "C" size_t ammc_bn_q_workspace_bytes(int C) { return (size_t)2 * C * 8 + (size_t)2 * C * 4 + 16; }

extern "C" int ammc_bn_batch_stats_q(const float* y, const float* gamma, const float* beta, float* running_mean,
                                     float* running_var, float* scale, float* shift, float* mean, float* invstd,
                                     void* workspace, size_t workspace_bytes, int b, int C, int h, int w, float momentum,
                                     float eps, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AMMC_REQUIRE(y && gamma && beta && running_mean && running_var && scale && shift && mean && invstd && C > 0 && b > 0,
               "bad argument");
  if (!workspace || workspace_bytes < ammc_bn_q_workspace_bytes(C)) return fail(AMMC_EWORKSPACE, "workspace too small");
  double* sums = (double*)workspace;
  unsigned* amax = (unsigned*)(sums + 2 * C);
  unsigned* bound = amax + 2 * C;
  AMMC_CUDA_CHECK(cudaMemsetAsync(workspace, 0, ammc_bn_q_workspace_bytes(C), st));
  const int splits = max(1, min(b, ceil_div(4 * num_sms(), C)));
  const int per = ceil_div(b, splits);
  bn_stats_kernel<<<dim3(C, ceil_div(b, per)), 256, 0, st>>>(y, sums, b, C, h * w, per, amax);
  AMMC_LAUNCH_CHECK("bn_stats_kernel");
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>(sums, gamma, beta, running_mean, running_var, scale, shift, mean,
                                                       invstd, C, (double)b * h * w, momentum, eps, 1, amax, bound);
  AMMC_LAUNCH_CHECK("bn_finalize_kernel");
  return 0;
}

// relu(y * scale + shift) -> bf16 hi/lo NHWC planes (operand of the weight gradient; may be NULL) and the q buffer
// `out_q` (ammc_q_act_bytes); `workspace` is the one ammc_bn_batch_stats_q filled
extern "C" int ammc_bn_apply_q(const float* y, const float* scale, const float* shift, int relu, void* out_nhwc_planes,
                               void* out_q, const void* workspace, int b, int C, int h, int w, void* stream) {
  AMMC_REQUIRE(y && scale && shift && out_q && workspace && C % 8 == 0, "bad argument");
  const long long n = (long long)b * C * h * w;
  ApplyArgs a{};
  a.y = y; a.scale = scale; a.shift = shift; a.mode = 0; a.relu = relu;
  a.nhwc = (__nv_bfloat16*)out_nhwc_planes;
  a.plane_stride = n;
  a.q16 = (__half*)out_q; a.q8 = (uint8_t*)out_q + 2 * n;
  a.q_bound = (const unsigned*)((const double*)workspace + 2 * C) + 2 * C;
  a.q_scale_out = (float*)((uint8_t*)out_q + 4 * n);
  return launch_apply(a, b, C, h, w, (cudaStream_t)stream);
}

// gradient through ReLU + BN -> g_y as bf16 hi/lo NHWC planes (weight gradient; may be NULL) and as q planes (data gradient)
extern "C" int ammc_bn_backward_q(const float* g, const float* y, const float* scale, const float* shift, const float* mean,
                                  const float* invstd, int relu, int training, void* gy_nhwc_planes, void* gy_q,
                                  float* g_gamma, float* g_beta, void* workspace, size_t workspace_bytes, int b, int C, int h,
                                  int w, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  AMMC_REQUIRE(g && y && scale && shift && mean && invstd && g_gamma && g_beta && gy_q && C % 8 == 0 && b > 0, "bad argument");
  if (!workspace || workspace_bytes < ammc_bn_q_workspace_bytes(C)) return fail(AMMC_EWORKSPACE, "workspace too small");
  double* sums = (double*)workspace;
  unsigned* amax = (unsigned*)(sums + 2 * C);
  unsigned* bound = amax + 2 * C;
  AMMC_CUDA_CHECK(cudaMemsetAsync(workspace, 0, ammc_bn_q_workspace_bytes(C), st));
  const int splits = max(1, min(b, ceil_div(4 * num_sms(), C)));
  const int per = ceil_div(b, splits);
  const double inv_count = 1.0 / ((double)b * h * w);
  bn_bwd_reduce_kernel<<<dim3(C, ceil_div(b, per)), 256, 0, st>>>(g, y, scale, shift, mean, invstd, relu, sums, b, C, h * w,
                                                                  per, amax);
  AMMC_LAUNCH_CHECK("bn_bwd_reduce_kernel");
  bn_param_grads_kernel<<<ceil_div(C, 128), 128, 0, st>>>(sums, g_gamma, g_beta, C, amax, scale, inv_count, training, bound);
  AMMC_LAUNCH_CHECK("bn_param_grads_kernel");
  const long long n = (long long)b * C * h * w;
  ApplyArgs a{};
  a.y = y; a.g = g; a.scale = scale; a.shift = shift; a.mean = mean; a.invstd = invstd; a.bsums = sums;
  a.inv_count = inv_count;
  a.mode = training ? 1 : 2; a.relu = relu;
  a.nhwc = (__nv_bfloat16*)gy_nhwc_planes;
  a.plane_stride = n;
  a.q16 = (__half*)gy_q; a.q8 = (uint8_t*)gy_q + 2 * n;
  a.q_bound = bound;
  a.q_scale_out = (float*)((uint8_t*)gy_q + 4 * n);
  return launch_apply(a, b, C, h, w, st);
}

namespace ammc {
int pack_planes_f32(const float* x, void* xp, long long n, cudaStream_t st) {
  pack_planes_kernel<<<ceil_div(n, 256), 256, 0, st>>>(x, (__nv_bfloat16*)xp, n);
  AMMC_LAUNCH_CHECK("pack_planes_kernel");
  return 0;
}
}  // namespace ammc

extern "C" int ammc_pack_planes(const float* x, void* xp, int64_t n, void* stream) {
  AMMC_REQUIRE(x && xp && n > 0, "bad argument");
  pack_planes_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)xp, n);
  AMMC_LAUNCH_CHECK("pack_planes_kernel");
  return 0;
}

extern "C" int ammc_pack_conv_weights_dgrad(const float* w, void* wp, int Cout, int Cin, void* stream) {
  AMMC_REQUIRE(w && wp && Cout > 0 && Cin > 0, "bad argument");
  const long long total = (long long)Cout * Cin * 9;
  pack_weights_dgrad_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)wp, Cout, Cin);
  AMMC_LAUNCH_CHECK("pack_weights_dgrad_kernel");
  return 0;
}

namespace ammc {
// gw [Cout][Cin][ntaps] = sum over pixels of gy[px][Cout] (x) x[px + tap][Cin]; operands are NHWC bf16 hi/lo planes.
// Channel counts need not fill the 128 x 256 tile: TMA zero-fills the channel boxes past Cout / Cin.
int conv_wgrad_run(const void* gy_nhwc_planes, const void* x_nhwc_planes, float* gw, int b, int Cin, int Cout, int h, int w,
                   int ntaps, int precision, cudaStream_t st) {
  AMMC_REQUIRE(gy_nhwc_planes && x_nhwc_planes && gw && b > 0, "bad argument");
  AMMC_REQUIRE(ntaps == 9 || ntaps == 1, "ntaps must be 9 or 1");
  if (precision != 1 && precision != 3) return fail(AMMC_EINVAL, "precision must be 1 or 3 (got %d)", precision);
  if (Cin % 64 != 0 || Cout % 64 != 0)
    return fail(AMMC_EUNSUPPORTED, "tcgen05 wgrad needs Cin and Cout multiples of 64 (got %d, %d)", Cin, Cout);
  if (w > 128)
    return fail(AMMC_EUNSUPPORTED, "tcgen05 wgrad supports feature maps up to 128 pixels wide (got %d)", w);
  WgradParams p;
  p.b = b; p.H = h; p.W = w; p.Cin = Cin; p.Cout = Cout;
  p.w_box = min(w, 64);
  p.rows_per_kb = w <= 64 ? min(h, 64 / w) : 1;
  p.chunks_per_row = ceil_div(w, 64);
  p.kb_per_img = ceil_div(h, p.rows_per_kb) * p.chunks_per_row;
  p.k_bytes = p.w_box * p.rows_per_kb * 128;
  p.tiles_m = ceil_div(Cout, WG_BLOCK_M);
  p.tiles_n = ceil_div(Cin, WG_BLOCK_N);
  p.block_n = min(Cin, WG_BLOCK_N);
  p.ntaps = ntaps;
  const int base_items = p.tiles_m * p.tiles_n * ntaps;
  // split-K over images: the kernel's time is (rounds of work items over the SMs) x (images per item), so take the split
  // count that minimises that product -- a tail round with a few items costs as much as a full one (b = 64, 72 base
  // items: 2 splits = 144 items in ONE round of 32 images; the old "2 items per SM" rule gave 5 splits = 360 items in
  // three rounds of 13) -- and among equals the smallest one (fewer atomics)
  {
    long long best_cost = -1;
    int best = 1;
    for (int sp = 1; sp <= min(b, 256); ++sp) {
      const int ips = ceil_div(b, sp), eff = ceil_div(b, ips);
      const long long cost = (long long)ceil_div((long long)base_items * eff, num_sms()) * ips;
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = eff; }
    }
    p.splits = best;
  }
  p.imgs_per_split = ceil_div(b, p.splits);
  p.splits = ceil_div(b, p.imgs_per_split);
  p.n_pass = precision;
  p.gw = gw;
  AMMC_CUDA_CHECK(cudaMemsetAsync(gw, 0, (size_t)Cout * Cin * ntaps * sizeof(float), st));
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[5] = {(uint64_t)Cout, (uint64_t)w, (uint64_t)h, (uint64_t)b, 2};
    uint64_t strides[4] = {(uint64_t)Cout * 2, (uint64_t)w * Cout * 2, (uint64_t)h * w * Cout * 2,
                           (uint64_t)b * h * w * Cout * 2};
    uint32_t box[5] = {64, (uint32_t)p.w_box, (uint32_t)p.rows_per_kb, 1, 1};
    if (int rc = make_map_bf16(&tmA, gy_nhwc_planes, 5, dims, strides, box)) return rc;
  }
  {
    uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)w, (uint64_t)h, (uint64_t)b, 2};
    uint64_t strides[4] = {(uint64_t)Cin * 2, (uint64_t)w * Cin * 2, (uint64_t)h * w * Cin * 2,
                           (uint64_t)b * h * w * Cin * 2};
    uint32_t box[5] = {64, (uint32_t)p.w_box, (uint32_t)p.rows_per_kb, 1, 1};
    if (int rc = make_map_bf16(&tmB, x_nhwc_planes, 5, dims, strides, box)) return rc;
  }
  static bool configured[64] = {false};
  int dev = 0;
  AMMC_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    AMMC_CUDA_CHECK(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
    configured[dev] = true;
  }
  const int items = base_items * p.splits;
  conv_wgrad_kernel<<<min(num_sms(), items), WG_THREADS, WG_SMEM, st>>>(tmA, tmB, p);
  AMMC_LAUNCH_CHECK("conv_wgrad_kernel");
  return 0;
}
}  // namespace ammc

extern "C" int ammc_conv3x3_wgrad(const void* gy_nhwc_planes, const void* x_nhwc_planes, float* gw, int b, int Cin,
                                  int Cout, int h, int w, int precision, void* stream) {
  return conv_wgrad_run(gy_nhwc_planes, x_nhwc_planes, gw, b, Cin, Cout, h, w, 9, precision, (cudaStream_t)stream);
}

extern "C" int ammc_conv1x1_wgrad(const void* gy_nhwc_planes, const void* x_nhwc_planes, float* gw, int b, int Cin,
                                  int Cout, int h, int w, int precision, void* stream) {
  return conv_wgrad_run(gy_nhwc_planes, x_nhwc_planes, gw, b, Cin, Cout, h, w, 1, precision, (cudaStream_t)stream);
}
