/*
 * ammc_b200_debug.h -- hardware-behaviour probes behind the kernel designs.  NOT part of the product ABI: these entry points
 * exist only in libammc_b200_debug.so (`python -m ammcnet_aaai2021_b200.build --debug`, sources compiled with
 * -DAMMC_DEBUG_PROBES); libammc_b200.so does not contain them.  tools/fp8_probe.py, tools/desc_probe.py and tools/tma_probe.py
 * bind them through ammcnet_aaai2021_b200/_capi_debug.py.
 */
#ifndef AMMC_B200_DEBUG_H_
#define AMMC_B200_DEBUG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* clock64 ticks for `iters` x 4 back-to-back MMAs (M = 128, N = n) of kind::f16 / bf16 (fp8 = 0, K = 16) or
 * kind::f8f6f4 / e4m3 (fp8 = 1, K = 32) on one SM.  cycles: device long long[1]. */
int ammc_debug_mma_rate(long long* cycles, int fp8, int n, int iters, void* stream);
/* kind::f8f6f4 (e4m3) MMAs chained into kind::f16 MMAs through scale-input-d on one 128x64 tile (csrc/probes.cu):
 * mode 0: out = (a8.b8^T) * 2^-12 + a16.b16^T; mode 1: out = a8.b8^T; mode 2: out = a16.b16^T. */
int ammc_debug_fp8_probe(const void* a8, const void* b8, const void* a16, const void* b16, float* out, int mode, void* stream);
/* UMMA K-major SWIZZLE_128B descriptor starting at an arbitrary 128-byte row (see csrc/halo_conv.cu) */
int ammc_debug_desc_probe(const void* a, const void* b, float* out, int rows, int row_off, int base_off, void* stream);
/* TMA-load one 5-D bf16 box (128B swizzle, zero OOB fill) and dump the raw shared-memory bytes to `out`. */
int ammc_debug_tma_probe(const void* base, const int64_t* dims5, const int64_t* strides4_bytes, const int* box5,
                         const int* coords5, void* out, int out_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AMMC_B200_DEBUG_H_ */
