// 4-lane "team" helpers shared by the generic fp32 addressing kernel (mem_simt.cu) and the refine / re-scan kernels
// of the tensor-core path (addr_tc.cu).  Both paths call exactly these functions, so norms, exact distances, gathered
// rows and commit partials are bit-identical by construction.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace ammc {

__device__ __forceinline__ float team_sum4(float s) {
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  return s;
}

__device__ __forceinline__ float f4_comp(const float4& v, int c) { return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w)); }

// ||z||^2: lane `part` accumulates the elements d = part (mod 4) in ascending order, then the team adds the partials.
__device__ __forceinline__ float team_zn2(const float* __restrict__ zr, int D, int part, bool valid) {
  float s = 0.f;
  if (valid) {
    if ((D & 3) == 0) {
      const float4* z4 = reinterpret_cast<const float4*>(zr);
      for (int i = 0; i < (D >> 2); ++i) { const float c = f4_comp(__ldg(z4 + i), part); s = fmaf(c, c, s); }
    } else {
      for (int d = part; d < D; d += 4) { const float v = zr[d]; s = fmaf(v, v, s); }
    }
  }
  return team_sum4(s);
}

// z . e accumulated with fmaf in ascending d (the order the generic kernel's smem-tiled GEMM uses)
__device__ __forceinline__ float exact_dot(const float* __restrict__ zr, const float* __restrict__ er, int D) {
  float acc = 0.f;
  if ((D & 3) == 0) {
    const float4* z4 = reinterpret_cast<const float4*>(zr);
    const float4* e4 = reinterpret_cast<const float4*>(er);
    const int n4 = D >> 2;
    int i = 0;
    for (; i + 8 <= n4; i += 8) {       // 16 loads in flight per lane, the fmaf chain itself is unchanged (ascending d)
      float4 a[8], b[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { a[j] = __ldg(z4 + i + j); b[j] = __ldg(e4 + i + j); }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc = fmaf(a[j].x, b[j].x, acc); acc = fmaf(a[j].y, b[j].y, acc);
        acc = fmaf(a[j].z, b[j].z, acc); acc = fmaf(a[j].w, b[j].w, acc);
      }
    }
    for (; i < n4; ++i) {
      const float4 a = __ldg(z4 + i), b = __ldg(e4 + i);
      acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
    }
  } else {
    for (int d = 0; d < D; ++d) acc = fmaf(zr[d], er[d], acc);
  }
  return acc;
}

// dist = (||z||^2 - 2 z.e) + ||e||^2, association of reference Code/models/unet.py:283-288
__device__ __forceinline__ float exact_dist(float zn2, float dot, float e2) {
  return __fadd_rn(__fsub_rn(zn2, 2.f * dot), e2);
}

// Outputs of one query row, produced cooperatively by its 4-lane team (lanes hold identical `ids`).
//   q1 = z + (e_top1 - z) (unet.py:311), read = concat of the K items (unet.py:295-297), per-pixel SSE (unet.py:310),
//   EMA statistics (unet.py:298-302).  Lane `part` owns the 16-byte chunks i = part (mod 4).
template <int K>
__device__ __forceinline__ void team_emit_row(const float* __restrict__ zr, const float* __restrict__ bank_t,
                                              const int (&ids)[K], int64_t n, int D, int M, int part, bool valid,
                                              float* __restrict__ read, float* __restrict__ q1,
                                              int64_t* __restrict__ idx, float* __restrict__ sse_px,
                                              float* __restrict__ counts, float* __restrict__ embed_sum,
                                              __nv_bfloat16* __restrict__ read_planes = nullptr,
                                              long long read_plane_stride = 0) {
  float sse = 0.f;
  if (valid) {
    if (part == 0) {
#pragma unroll
      for (int i = 0; i < K; ++i) idx[n * K + i] = (int64_t)ids[i];
    }
    const float* e1 = bank_t + (size_t)ids[0] * D;
    if ((D & 3) == 0) {
      const float4* z4 = reinterpret_cast<const float4*>(zr);
      const float4* e4 = reinterpret_cast<const float4*>(e1);
      float4* q4 = reinterpret_cast<float4*>(q1 + n * D);
      for (int i = part; i < (D >> 2); i += 4) {
        const float4 zv = __ldg(z4 + i), ev = __ldg(e4 + i);
        float4 df = make_float4(ev.x - zv.x, ev.y - zv.y, ev.z - zv.z, ev.w - zv.w);
        q4[i] = make_float4(zv.x + df.x, zv.y + df.y, zv.z + df.z, zv.w + df.w);
        sse = fmaf(df.x, df.x, sse); sse = fmaf(df.y, df.y, sse); sse = fmaf(df.z, df.z, sse); sse = fmaf(df.w, df.w, sse);
        if (embed_sum) {
          atomicAdd(&embed_sum[(size_t)(4 * i + 0) * M + ids[0]], zv.x);
          atomicAdd(&embed_sum[(size_t)(4 * i + 1) * M + ids[0]], zv.y);
          atomicAdd(&embed_sum[(size_t)(4 * i + 2) * M + ids[0]], zv.z);
          atomicAdd(&embed_sum[(size_t)(4 * i + 3) * M + ids[0]], zv.w);
        }
      }
      if (read) {
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const float4* er = reinterpret_cast<const float4*>(bank_t + (size_t)ids[j] * D);
          float4* rr = reinterpret_cast<float4*>(read + (n * K + j) * D);
          for (int i = part; i < (D >> 2); i += 4) rr[i] = __ldg(er + i);
        }
      }
      if (read_planes) {   // bf16 hi/lo split of the read: the A operand of the tensor-core `dec` GEMM
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const float4* er = reinterpret_cast<const float4*>(bank_t + (size_t)ids[j] * D);
          __nv_bfloat16* hp = read_planes + (n * K + j) * D;
          for (int i = part; i < (D >> 2); i += 4) {
            const float4 v = __ldg(er + i);
            const __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y);
            const __nv_bfloat16 h2 = __float2bfloat16_rn(v.z), h3 = __float2bfloat16_rn(v.w);
            const __nv_bfloat16 l0 = __float2bfloat16_rn(v.x - __bfloat162float(h0));
            const __nv_bfloat16 l1 = __float2bfloat16_rn(v.y - __bfloat162float(h1));
            const __nv_bfloat16 l2 = __float2bfloat16_rn(v.z - __bfloat162float(h2));
            const __nv_bfloat16 l3 = __float2bfloat16_rn(v.w - __bfloat162float(h3));
            uint2 ph, pl;
            ph.x = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            ph.y = (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16);
            pl.x = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
            pl.y = (uint32_t)__bfloat16_as_ushort(l2) | ((uint32_t)__bfloat16_as_ushort(l3) << 16);
            *reinterpret_cast<uint2*>(hp + 4 * i) = ph;
            *reinterpret_cast<uint2*>(hp + read_plane_stride + 4 * i) = pl;
          }
        }
      }
    } else {
      for (int d = part; d < D; d += 4) {
        const float zv = zr[d], ev = e1[d];
        const float df = ev - zv;
        q1[n * D + d] = zv + df;
        sse = fmaf(df, df, sse);
        if (embed_sum) atomicAdd(&embed_sum[(size_t)d * M + ids[0]], zv);
      }
      if (read) {
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const float* er = bank_t + (size_t)ids[j] * D;
          float* rr = read + (n * K + j) * D;
          for (int d = part; d < D; d += 4) rr[d] = er[d];
        }
      }
    }
  }
  sse = team_sum4(sse);
  if (part == 0 && valid) {
    sse_px[n] = sse;
    if (counts) atomicAdd(&counts[ids[0]], 1.f);
  }
}

}  // namespace ammc
