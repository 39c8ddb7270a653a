// Error plumbing, device check and the launch counter behind the C ABI.
#include <stdarg.h>

#include "common.cuh"

namespace btsb {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// There is no CPU or other-arch path: everything but compute capability 10.x is refused.
int check_device() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_rc = 0;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no CUDA device: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return BTSB_EARCH;
  }
  if (dev == cached_dev) {
    if (cached_rc) set_error("device %d is not compute capability 10.x (sm_100a required)", dev);
    return cached_rc;
  }
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) {
    set_error("cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
    return BTSB_ECUDA;
  }
  cached_dev = dev;
  cached_rc = (major == 10) ? BTSB_OK : BTSB_EARCH;
  if (cached_rc) set_error("device %d is compute capability %d.x; btsbot_b200 is built for sm_100a only", dev, major);
  return cached_rc;
}

}  // namespace btsb

extern "C" int btsb_version(void) { return 100; }
extern "C" const char* btsb_last_error_string(void) { return btsb::g_err; }
extern "C" int btsb_device_ok(void) { return btsb::check_device(); }
extern "C" uint64_t btsb_launch_count(void) { return btsb::g_launches.load(std::memory_order_relaxed); }
