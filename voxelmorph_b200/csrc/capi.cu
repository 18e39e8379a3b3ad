// Error text, version string, launch counter and small device queries of the C ABI.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace vxm {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: kernel launch failed: %s", what, cudaGetErrorString(e));
    return VXM_ERR_CUDA;
  }
  count_launch(1);
  return VXM_OK;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace vxm

extern "C" const char* vxm_last_error(void) { return vxm::g_err; }
extern "C" const char* vxm_version(void) { return "vxm_b200 0.1 sm_100a"; }
extern "C" uint64_t vxm_launch_count(void) { return vxm::g_launches.load(std::memory_order_relaxed); }
