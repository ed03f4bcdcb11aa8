// Library plumbing: error reporting, device checks.
#include "common.cuh"
#include <string.h>

namespace ammc {

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!cached[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace ammc

namespace ammc {
int timeout_reader_conv(int*);
int timeout_reader_addr(int*);
int timeout_reader_train(int*);
int timeout_reader_enc(int*);
int timeout_reader_halo(int*);
int timeout_reader_front(int*);
}

// Synchronises the device and reports whether a tcgen05 pipeline wait hit its bound.  A wait that does executes `trap`
// (ptx.cuh): the synchronise below then fails and this returns AMMC_ECUDA -- the normal way a broken pipeline surfaces.
// Return 1 with out4 = {kernel family (1 conv, 2 addressing, 3 training, 4 enc, 5 halo conv, 6 fused memory front), tag,
// block, thread} is kept for the case that the record could still be read; 0 = all pipelines healthy.
extern "C" int ammc_pipeline_check(int* out4) {
  if (cudaDeviceSynchronize() != cudaSuccess) return ammc::fail(AMMC_ECUDA, "device synchronize failed: %s",
                                                               cudaGetErrorString(cudaGetLastError()));
  int rec[4];
  int (*readers[6])(int*) = {ammc::timeout_reader_conv, ammc::timeout_reader_addr, ammc::timeout_reader_train,
                             ammc::timeout_reader_enc, ammc::timeout_reader_halo, ammc::timeout_reader_front};
  for (int i = 0; i < 6; ++i) {
    if (readers[i](rec) != 0) return ammc::fail(AMMC_ECUDA, "cannot read the watchdog record");
    if (rec[0]) {
      if (out4) { out4[0] = i + 1; out4[1] = rec[1]; out4[2] = rec[2]; out4[3] = rec[3]; }
      return 1;
    }
  }
  return 0;
}

extern "C" int ammc_version(void) { return 100; }

extern "C" const char* ammc_last_error(void) { return ammc::err_buf(); }

extern "C" int ammc_device_supported(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}
