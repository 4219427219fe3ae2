// Error reporting + device check for the C ABI (include/hosnerf_b200.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace hos {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_arch() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_res = HOS_ERR_ARCH;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDevice failed: %s (no CUDA device? this library has no CPU fallback)",
              cudaGetErrorString(e));
    return HOS_ERR_CUDA;
  }
  if (dev == cached_dev) {
    if (cached_res != HOS_OK) set_error("device %d is not sm_100 (B200)", dev);
    return cached_res;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  cached_dev = dev;
  cached_res = (major == 10 && minor == 0) ? HOS_OK : HOS_ERR_ARCH;
  if (cached_res != HOS_OK)
    set_error("device %d is sm_%d%d; libhosnerf_b200 is built for sm_100a only", dev, major, minor);
  return cached_res;
}

}  // namespace hos

extern "C" {

const char* hos_last_error(void) { return hos::g_err; }

int hos_version(void) { return 100; }

int hos_device_check(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || device < 0 || device >= count) {
    hos::set_error("device %d not available (%s)", device,
                   e == cudaSuccess ? "out of range" : cudaGetErrorString(e));
    return HOS_ERR_CUDA;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
  if (major != 10 || minor != 0) {
    hos::set_error("device %d is sm_%d%d; need sm_100 (B200)", device, major, minor);
    return HOS_ERR_ARCH;
  }
  return HOS_OK;
}

}  // extern "C"
