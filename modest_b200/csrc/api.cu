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

// Stage A plumbing for the streaming engine: many small host -> device copies in one call (a batch of
// a drive needs up to ~800 frames of ~1 MB that live in separate pinned host buffers; one Python-level
// copy per frame costs 30-80 us of host time each, this loop 2-3 us).
extern "C" int modest_upload_frames(const void* const* h_src, void* const* d_dst, const int64_t* n_bytes, int n, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MODEST_REQUIRE(n >= 0, "upload_frames: negative count");
  if (n == 0) return MODEST_OK;
  MODEST_REQUIRE(h_src && d_dst && n_bytes, "upload_frames: null pointer argument");
  for (int i = 0; i < n; ++i) {
    MODEST_REQUIRE(n_bytes[i] >= 0 && (n_bytes[i] == 0 || (h_src[i] && d_dst[i])), "upload_frames: bad entry %d", i);
    if (n_bytes[i]) MODEST_CUDA(cudaMemcpyAsync(d_dst[i], h_src[i], (size_t)n_bytes[i], cudaMemcpyHostToDevice, stream));
  }
  return MODEST_OK;
}
