// Library-level entry points: error reporting, version, device check, launch counter.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

namespace dcv {
static thread_local char g_error[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
}  // namespace dcv

extern "C" {

int dcv_abi_version(void) { return DCV_ABI_VERSION; }

const char* dcv_last_error(void) { return dcv::g_error; }

uint64_t dcv_launch_count(void) { return dcv::g_launches.load(std::memory_order_relaxed); }

int dcv_device_check(void) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    dcv::set_error("no CUDA device: %s (this library has no CPU path)", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return 1;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    dcv::set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev, major, minor);
    return 1;
  }
  return 0;
}

}  // extern "C"
