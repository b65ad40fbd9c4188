// Error channel, launch accounting and misc. entry points of libmodest_b200.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace modest {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return MODEST_ERR_CUDA;
}

void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

}  // namespace modest

extern "C" int modest_abi_version(void) { return MODEST_ABI_VERSION; }
extern "C" const char* modest_last_error(void) { return modest::g_err; }
extern "C" int64_t modest_launch_count(void) { return (int64_t)modest::g_launches.load(); }
