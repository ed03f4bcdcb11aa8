// Frame / flow preprocessing in front of the generator (reference
// Code/dataset/two_stream_dataset.py:72-99 _load_frame / _load_op, transforms at 501-509).
//   frames: decoded BGR uint8 [n,h0,w0,3] -> BGR2RGB -> cv2.resize (INTER_LINEAR, 8-bit fixed-point path) -> ToTensor (/255)
//           -> Normalize(0.5, 0.5) -> fp32 [n,3,H,W]
//   flow:   .flo payload fp32 [n,h0,w0,2] -> cv2.resize (float path) -> ch0 = u*1.0/H, ch1 = ch0/W (the loader's quirk: the
//           resized v component is never used) -> fp32 [n,2,H,W]
// Bit-exact with the reference's loaders (tests/golden/preprocess.npz): every float operation is an explicit _rn intrinsic so
// nothing is contracted into an FMA, the tap positions use the same double -> float sequence as OpenCV's resize.cpp.
// Pure HBM/L2 streams: one thread per output pixel, coalesced fp32 plane writes; the uint8 sources (4x smaller than the fp32
// tensors the host used to upload) are read through L1/L2.
#include "common.cuh"
#include "ptx.cuh"
#include <cuda_bf16.h>

namespace ammc {

struct Tap { int i0, i1; float w0, w1; };

// cv2 resize.cpp: f = float((d + 0.5) * scale - 0.5); s = floor(f); f -= s.  Along x the weight is clamped at the borders,
// along y only the indices are.
__device__ __forceinline__ Tap linear_tap(int d, double scale, int src, bool clamp_weights) {
  float f = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);
  int s = __float2int_rd(f);
  f = __fsub_rn(f, (float)s);
  if (clamp_weights) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= src - 1) { f = 0.f; s = src - 1; }
  }
  Tap t;
  t.w0 = __fsub_rn(1.f, f);
  t.w1 = f;
  t.i0 = min(max(s, 0), src - 1);
  t.i1 = min(max(s + 1, 0), src - 1);
  return t;
}

__global__ void __launch_bounds__(256) preprocess_frames_u8_kernel(const uint8_t* __restrict__ in, float* __restrict__ out,
                                                                    int n, int h0, int w0, int H, int W, double sx,
                                                                    double sy) {
  // ToTensor + Normalize of every possible byte value, computed once per block with the reference's operation order:
  // the per-pixel work is then integer arithmetic and one shared-memory lookup
  __shared__ float lut[256];
  lut[threadIdx.x] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)threadIdx.x, 255.f), 0.5f), 0.5f);
  __syncthreads();
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const Tap tx = linear_tap(x, sx, w0, true), ty = linear_tap(y, sy, h0, false);
  const int a0 = __float2int_rn(__fmul_rn(tx.w0, 2048.f)), a1 = __float2int_rn(__fmul_rn(tx.w1, 2048.f));
  const int b0 = __float2int_rn(__fmul_rn(ty.w0, 2048.f)), b1 = __float2int_rn(__fmul_rn(ty.w1, 2048.f));
  const size_t plane = (size_t)H * W;
  for (int img = blockIdx.z; img < n; img += gridDim.z) {     // the taps are reused for every image of the z stride
    const uint8_t* p = in + (size_t)img * h0 * w0 * 3;
    const uint8_t* r0 = p + (size_t)ty.i0 * w0 * 3;
    const uint8_t* r1 = p + (size_t)ty.i1 * w0 * 3;
    float* o = out + (size_t)img * 3 * plane + (size_t)y * W + x;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int cs = 2 - c;                                   // BGR -> RGB
      const int s0 = (int)r0[tx.i0 * 3 + cs] * a0 + (int)r0[tx.i1 * 3 + cs] * a1;
      const int s1 = (int)r1[tx.i0 * 3 + cs] * a0 + (int)r1[tx.i1 * 3 + cs] * a1;
      int v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
      v = min(max(v, 0), 255);
      o[c * plane] = lut[v];
    }
  }
}

__global__ void __launch_bounds__(256) preprocess_flow_kernel(const float* __restrict__ in, float* __restrict__ out, int n,
                                                               int h0, int w0, int H, int W, double sx, double sy) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const Tap tx = linear_tap(x, sx, w0, true), ty = linear_tap(y, sy, h0, false);
  for (int img = blockIdx.z; img < n; img += gridDim.z) {
    const float* p = in + (size_t)img * h0 * w0 * 2;
    const float* r0 = p + (size_t)ty.i0 * w0 * 2;
    const float* r1 = p + (size_t)ty.i1 * w0 * 2;
    // channel 0 only: two_stream_dataset.py:94-95 overwrites channel 1 with (scaled channel 0) / W
    const float s0 = __fadd_rn(__fmul_rn(r0[tx.i0 * 2], tx.w0), __fmul_rn(r0[tx.i1 * 2], tx.w1));
    const float s1 = __fadd_rn(__fmul_rn(r1[tx.i0 * 2], tx.w0), __fmul_rn(r1[tx.i1 * 2], tx.w1));
    const float u = __fadd_rn(__fmul_rn(s0, ty.w0), __fmul_rn(s1, ty.w1));
    const float c0 = __fdiv_rn(__fmul_rn(u, 1.0f), (float)H);
    const float c1 = __fdiv_rn(c0, (float)W);
    const size_t o = ((size_t)img * 2 * H + y) * W + x;
    out[o] = c0;
    out[o + (size_t)H * W] = c1;
  }
}

// bf16 -> fp32, 16 bytes in / 32 bytes out per thread and step (the bf16 feature-I/O boundary of BASELINE configs[2])
__global__ void __launch_bounds__(256) cast_bf16_f32_kernel(const uint16_t* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long n8 = n >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + i);
    float4 a, b;
    a.x = __uint_as_float(v.x << 16); a.y = __uint_as_float(v.x & 0xffff0000u);
    a.z = __uint_as_float(v.y << 16); a.w = __uint_as_float(v.y & 0xffff0000u);
    b.x = __uint_as_float(v.z << 16); b.y = __uint_as_float(v.z & 0xffff0000u);
    b.z = __uint_as_float(v.w << 16); b.w = __uint_as_float(v.w & 0xffff0000u);
    reinterpret_cast<float4*>(dst)[2 * i] = a;
    reinterpret_cast<float4*>(dst)[2 * i + 1] = b;
  }
  if (blockIdx.x == 0)
    for (long long i = (n8 << 3) + threadIdx.x; i < n; i += blockDim.x) dst[i] = __uint_as_float((uint32_t)src[i] << 16);
}

// fp32 -> bf16 (round to nearest even), 32 bytes in / 16 bytes out per thread and step
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, long long n) {
  const long long n8 = n >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    uint4 o;
    o.x = ptx::pack_bf16x2(a.x, a.y); o.y = ptx::pack_bf16x2(a.z, a.w);
    o.z = ptx::pack_bf16x2(b.x, b.y); o.w = ptx::pack_bf16x2(b.z, b.w);
    reinterpret_cast<uint4*>(dst)[i] = o;
  }
  if (blockIdx.x == 0)
    for (long long i = (n8 << 3) + threadIdx.x; i < n; i += blockDim.x) dst[i] = __bfloat16_as_ushort(__float2bfloat16_rn(src[i]));
}

}  // namespace ammc

using namespace ammc;

extern "C" int ammc_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream) {
  AMMC_REQUIRE(src && dst && n > 0, "bad argument");
  AMMC_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0, "16-byte aligned buffers required");
  cast_f32_bf16_kernel<<<min(ceil_div(n, 2048), 148 * 16), 256, 0, (cudaStream_t)stream>>>(src, (uint16_t*)dst, n);
  AMMC_LAUNCH_CHECK("cast_f32_bf16_kernel");
  return 0;
}

extern "C" int ammc_cast_bf16_f32(const void* src, float* dst, int64_t n, void* stream) {
  AMMC_REQUIRE(src && dst && n > 0, "bad argument");
  AMMC_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0, "16-byte aligned buffers required");
  cast_bf16_f32_kernel<<<min(ceil_div(n, 2048), 148 * 16), 256, 0, (cudaStream_t)stream>>>((const uint16_t*)src, dst, n);
  AMMC_LAUNCH_CHECK("cast_bf16_f32_kernel");
  return 0;
}

static int check_dims(int n, int h0, int w0, int H, int W) {
  AMMC_REQUIRE(n > 0 && h0 > 0 && w0 > 0 && H > 0 && W > 0, "bad shape n=%d src=%dx%d dst=%dx%d", n, h0, w0, H, W);
  AMMC_REQUIRE((long long)h0 * w0 <= (1LL << 28) && (long long)H * W <= (1LL << 28), "frame too large");
  return 0;
}

extern "C" int ammc_preprocess_frames_u8(const uint8_t* frames_bgr, float* out, int n, int h0, int w0, int H, int W,
                                         void* stream) {
  AMMC_REQUIRE(frames_bgr && out, "null pointer argument");
  if (int rc = check_dims(n, h0, w0, H, W)) return rc;
  const double sx = 1.0 / ((double)W / (double)w0), sy = 1.0 / ((double)H / (double)h0);
  dim3 grid(ceil_div(W, 32), ceil_div(H, 8), n < 16 ? n : 16);
  preprocess_frames_u8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(frames_bgr, out, n, h0, w0, H, W, sx, sy);
  AMMC_LAUNCH_CHECK("preprocess_frames_u8_kernel");
  return 0;
}

extern "C" int ammc_preprocess_flow(const float* flow, float* out, int n, int h0, int w0, int H, int W, void* stream) {
  AMMC_REQUIRE(flow && out, "null pointer argument");
  if (int rc = check_dims(n, h0, w0, H, W)) return rc;
  const double sx = 1.0 / ((double)W / (double)w0), sy = 1.0 / ((double)H / (double)h0);
  dim3 grid(ceil_div(W, 32), ceil_div(H, 8), n < 16 ? n : 16);
  preprocess_flow_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(flow, out, n, h0, w0, H, W, sx, sy);
  AMMC_LAUNCH_CHECK("preprocess_flow_kernel");
  return 0;
}
