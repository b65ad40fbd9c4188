// Shared helpers for libmodest_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <nvtx3/nvToolsExt.h>

#include "../../include/modest_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libmodest_b200 is written for sm_100a (B200) only"
#endif

namespace modest {

// ---- error channel (thread-local text, integer codes across the ABI) -----------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define MODEST_CUDA(call)                                             \
  do {                                                                \
    cudaError_t e__ = (call);                                         \
    if (e__ != cudaSuccess) return ::modest::cuda_fail(e__, #call);   \
  } while (0)

#define MODEST_LAUNCH_CHECK(name)                                     \
  do {                                                                \
    cudaError_t e__ = cudaGetLastError();                             \
    if (e__ != cudaSuccess) return ::modest::cuda_fail(e__, name);    \
  } while (0)

#define MODEST_REQUIRE(cond, ...)                                     \
  do {                                                                \
    if (!(cond)) {                                                    \
      ::modest::set_error(__VA_ARGS__);                               \
      return MODEST_ERR_ARG;                                          \
    }                                                                 \
  } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Bump allocator over the caller-provided workspace; the library never cudaMallocs.
struct Arena {
  char* base;
  size_t cap;
  size_t used;
  Arena(void* p, size_t bytes) : base(static_cast<char*>(p)), cap(bytes), used(0) {}
  template <typename T>
  T* take(size_t n) {
    size_t off = align_up(used, 256);
    used = off + n * sizeof(T);
    return reinterpret_cast<T*>(base + off);
  }
  bool ok() const { return base != nullptr && used <= cap; }
};

int sm_count();

// NVTX range over one stage's launch sequence (host side; free when no profiler listens): ncu /
// nsys timelines show which reference stage (SURVEY.md 8(a) letter) a group of kernels belongs to.
struct StageRange {
  explicit StageRange(const char* name) { nvtxRangePushA(name); }
  ~StageRange() { nvtxRangePop(); }
};

// ---- device helpers --------------------------------------------------------------------------
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// order-preserving map float -> uint (for radix selection / atomicMin on floats)
__device__ __forceinline__ unsigned f32_ordered(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_from_ordered(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ unsigned long long f64_ordered(double d) {
  unsigned long long u = (unsigned long long)__double_as_longlong(d);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double f64_from_ordered(unsigned long long u) {
  return __longlong_as_double((long long)((u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u));
}

// squared distance exactly as scipy's cKDTree / sklearn's KDTree evaluate it on f32-originated
// coordinates: widen to f64, ((dx*dx) + dy*dy) + dz*dz with every product and sum rounded
// separately (no FMA contraction).
__device__ __forceinline__ double sqdist_f64_seq(float ax, float ay, float az, float bx, float by, float bz) {
  double dx = __dsub_rn((double)ax, (double)bx);
  double dy = __dsub_rn((double)ay, (double)by);
  double dz = __dsub_rn((double)az, (double)bz);
  double s = __dmul_rn(dx, dx);
  s = __dadd_rn(s, __dmul_rn(dy, dy));
  s = __dadd_rn(s, __dmul_rn(dz, dz));
  return s;
}

}  // namespace modest
