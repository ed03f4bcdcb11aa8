// Data-movement kernels of the U-Net encoder/decoder around the memory path (reference
// Code/models/unet.py:23-59 inconv/down/up).  All activations between layers are NHWC bf16 hi/lo planes -- the operand
// format of the tcgen05 conv engine (amft_conv.cu) -- so these kernels are pure HBM streams: 16-byte vector accesses,
// one thread per 8 channels of one output pixel.
#include "common.cuh"
#include <cuda_bf16.h>

namespace ammc {

__device__ __forceinline__ void split2(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

__device__ __forceinline__ void load8(const __nv_bfloat16* hi, const __nv_bfloat16* lo, float (&v)[8]) {
  const uint4 a = *reinterpret_cast<const uint4*>(hi), c = *reinterpret_cast<const uint4*>(lo);
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, cw[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    // bf16 -> fp32 is a 16-bit shift; hi + lo is exact in fp32 (lo is the rounding residual of hi)
    v[2 * j] = __uint_as_float(aw[j] << 16) + __uint_as_float(cw[j] << 16);
    v[2 * j + 1] = __uint_as_float(aw[j] & 0xffff0000u) + __uint_as_float(cw[j] & 0xffff0000u);
  }
}

// MaxPool2d(2), floor mode (unet.py:33).  in [2][b,h,w,in_cs] window [in_c_off, +C) -> out [2][b,h/2,w/2,C]
__global__ void __launch_bounds__(256) maxpool2_planes_kernel(const __nv_bfloat16* __restrict__ in, int in_cs, int in_c_off,
                                                               long long in_plane, __nv_bfloat16* __restrict__ out,
                                                               long long out_plane, int b, int h, int w, int C) {
  const int ho = h >> 1, wo = w >> 1, c8n = C >> 3;
  const long long total = (long long)b * ho * wo * c8n;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(e % c8n);
    long long pix = e / c8n;
    const int x = (int)(pix % wo); pix /= wo;
    const int y = (int)(pix % ho);
    const int img = (int)(pix / ho);
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const size_t o = (((size_t)img * h + 2 * y + dy) * w + 2 * x + dx) * in_cs + in_c_off + c8 * 8;
        float v[8];
        load8(in + o, in + o + in_plane, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
      }
    uint32_t hp[4], lp[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat16 h0, l0, h1, l1;
      split2(m[2 * j], h0, l0);
      split2(m[2 * j + 1], h1, l1);
      hp[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      lp[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    const size_t oo = (((size_t)img * ho + y) * wo + x) * C + c8 * 8;
    *reinterpret_cast<uint4*>(out + oo) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
    *reinterpret_cast<uint4*>(out + oo + out_plane) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
  }
}

// x [b][C][HW] fp32 -> xp [2][b][HW][C_pad] bf16, zero in the pad channels.  A block transposes a [64 ch][32 px] tile
// through shared memory (coalesced row loads of the real channels only); every thread then emits 8 channels of one pixel as
// one 16-byte store per plane.  C_pad must be a multiple of 8.
__global__ void __launch_bounds__(256) pack_nhwc_padded_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xp,
                                                                int C, int C_pad, int HW, long long plane_stride) {
  __shared__ float tile[64][33];
  const int img = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = ty; r < 64; r += 8) {
    const int c = c0 + r, pp = p0 + tx;
    tile[r][tx] = (c < C && pp < HW) ? x[((size_t)img * C + c) * HW + pp] : 0.f;
  }
  __syncthreads();
  const int px = threadIdx.x >> 3, cg = threadIdx.x & 7;
  const int pp = p0 + px, c = c0 + cg * 8;
  if (pp < HW && c < C_pad) {
    uint32_t hp[4], lp[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat16 h0, l0, h1, l1;
      split2(tile[cg * 8 + 2 * j][px], h0, l0);
      split2(tile[cg * 8 + 2 * j + 1][px], h1, l1);
      hp[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      lp[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    __nv_bfloat16* dst = xp + ((size_t)img * HW + pp) * C_pad + c;
    *reinterpret_cast<uint4*>(dst) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
    *reinterpret_cast<uint4*>(dst + plane_stride) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
  }
}

// planes window -> fp32 NCHW (hi + lo), the inverse of the pack; 32x32 tiles through shared memory
__global__ void __launch_bounds__(256) unpack_nhwc_kernel(const __nv_bfloat16* __restrict__ xp, int cs, int c_off,
                                                           long long plane_stride, float* __restrict__ x, int C, int HW) {
  __shared__ float tile[32][33];
  const int img = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int pp = p0 + r, c = c0 + tx;
    float v = 0.f;
    if (pp < HW && c < C) {
      const size_t o = ((size_t)img * HW + pp) * cs + c_off + c;
      v = __bfloat162float(xp[o]) + __bfloat162float(xp[o + plane_stride]);
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, pp = p0 + tx;
    if (c < C && pp < HW) x[((size_t)img * C + c) * HW + pp] = tile[tx][r];
  }
}

// w [Cout][Cin][taps] fp32 -> wp [2][Cout_pad][taps*Cin_pad] bf16, k = tap*Cin_pad + cin, zeros in the padding
__global__ void pack_weights_padded_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int Cout, int Cin,
                                           int Cout_pad, int Cin_pad, int taps) {
  const long long total = (long long)Cout_pad * Cin_pad * taps;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int cin = (int)(e % Cin_pad);
  const int tap = (int)((e / Cin_pad) % taps);
  const int co = (int)(e / ((long long)Cin_pad * taps));
  const float v = (co < Cout && cin < Cin) ? w[((size_t)co * Cin + cin) * taps + tap] : 0.f;
  __nv_bfloat16 hi, lo;
  split2(v, hi, lo);
  wp[e] = hi;
  wp[e + total] = lo;
}

// ConvTranspose2d weight [Cin][Cout][2][2] -> wp [2][4*Cout][Cin], row n = (dy*2+dx)*Cout + co
__global__ void pack_convT_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int Cin, int Cout) {
  const long long total = 4LL * Cout * Cin;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int ci = (int)(e % Cin);
  const int n = (int)(e / Cin);
  const int co = n % Cout, q4 = n / Cout;
  __nv_bfloat16 hi, lo;
  split2(w[((size_t)ci * Cout + co) * 4 + q4], hi, lo);
  wp[e] = hi;
  wp[e + total] = lo;
}

}  // namespace ammc

using namespace ammc;

extern "C" int ammc_maxpool2_planes(const void* in_planes, int in_cs, int in_c_off, void* out_planes, int b, int h, int w,
                                    int C, void* stream) {
  AMMC_REQUIRE(in_planes && out_planes && b > 0 && h >= 2 && w >= 2 && C > 0, "bad argument");
  if (in_cs <= 0) in_cs = C;
  AMMC_REQUIRE(C % 8 == 0 && in_cs % 8 == 0 && in_c_off % 8 == 0 && in_c_off >= 0 && in_c_off + C <= in_cs,
               "channel window [%d, %d) of a %d-channel buffer must be 8-channel aligned", in_c_off, in_c_off + C, in_cs);
  const long long total = (long long)b * (h / 2) * (w / 2) * (C / 8);
  const long long want = (total + 255) / 256, cap = (long long)num_sms() * 16;
  const int grid = (int)(want < cap ? want : cap);
  maxpool2_planes_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)in_planes, in_cs, in_c_off, (long long)b * h * w * in_cs, (__nv_bfloat16*)out_planes,
      (long long)b * (h / 2) * (w / 2) * C, b, h, w, C);
  AMMC_LAUNCH_CHECK("maxpool2_planes_kernel");
  return 0;
}

extern "C" int ammc_pack_nhwc_padded(const float* x, void* xp, int b, int C, int C_pad, int h, int w, void* stream) {
  AMMC_REQUIRE(x && xp && b > 0 && C > 0 && C_pad >= C && h > 0 && w > 0, "bad argument");
  AMMC_REQUIRE(C_pad % 8 == 0, "C_pad must be a multiple of 8 (got %d)", C_pad);
  AMMC_REQUIRE(b <= 65535, "batch %d too large for one pack launch", b);
  const int HW = h * w;
  pack_nhwc_padded_kernel<<<dim3(ceil_div(HW, 32), ceil_div(C_pad, 64), b), 256, 0, (cudaStream_t)stream>>>(
      x, (__nv_bfloat16*)xp, C, C_pad, HW, (long long)b * HW * C_pad);
  AMMC_LAUNCH_CHECK("pack_nhwc_padded_kernel");
  return 0;
}

extern "C" int ammc_unpack_nhwc(const void* xp, int cs, int c_off, float* x, int b, int C, int h, int w, void* stream) {
  AMMC_REQUIRE(x && xp && b > 0 && C > 0 && h > 0 && w > 0, "bad argument");
  if (cs <= 0) cs = C;
  AMMC_REQUIRE(c_off >= 0 && c_off + C <= cs && b <= 65535, "bad channel window / batch");
  const int HW = h * w;
  unpack_nhwc_kernel<<<dim3(ceil_div(HW, 32), ceil_div(C, 32), b), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)xp, cs, c_off, (long long)b * HW * cs, x, C, HW);
  AMMC_LAUNCH_CHECK("unpack_nhwc_kernel");
  return 0;
}

extern "C" int ammc_pack_conv_weights_padded(const float* w, void* wp, int Cout, int Cin, int Cout_pad, int Cin_pad,
                                             int taps, void* stream) {
  AMMC_REQUIRE(w && wp && Cout > 0 && Cin > 0 && Cout_pad >= Cout && Cin_pad >= Cin && (taps == 9 || taps == 1),
               "bad argument");
  const long long total = (long long)Cout_pad * Cin_pad * taps;
  pack_weights_padded_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)wp, Cout, Cin,
                                                                                      Cout_pad, Cin_pad, taps);
  AMMC_LAUNCH_CHECK("pack_weights_padded_kernel");
  return 0;
}

extern "C" int ammc_pack_convt_weights(const float* w, void* wp, int Cin, int Cout, void* stream) {
  AMMC_REQUIRE(w && wp && Cin > 0 && Cout > 0, "bad argument");
  const long long total = 4LL * Cout * Cin;
  pack_convT_weights_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)wp, Cin, Cout);
  AMMC_LAUNCH_CHECK("pack_convT_weights_kernel");
  return 0;
}
