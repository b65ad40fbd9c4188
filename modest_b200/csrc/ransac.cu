// Stage E: RANSAC ground plane (sklearn RANSACRegressor().fit(xy, z) as called by
// estimate_plane(), utils/pointcloud_utils.py:44-65; the trial loop is
// sklearn/linear_model/_ransac.py:447-560, un-vendored third-party code restated here).
//
// One scan = one column of work in every kernel (blockIdx.y or blockIdx.x = scan):
//   1. plane_candidates: order-preserving compaction of the points with z < max_hs strictly
//      inside the x/y range (the triples index into that compacted list, so order matters).
//   2. mad_threshold:    residual_threshold = median(|z - median(z)|), float32, by two block
//      radix selections (numpy's even-length rule: mean of the two middle values in f32).
//   3. hypotheses:       plane through each 3-point minimal set (exact in f64, rounded to the
//      float32 model sklearn's LinearRegression holds).  Triples come from the host (parity
//      mode: drawn with numpy's global RandomState exactly like sklearn) or from a
//      counter-based device generator (throughput mode).
//   4. score:            every hypothesis against every candidate: inlier count and the sums
//      needed for R^2, points held in registers, warp-shuffle reductions, one atomic per
//      (warp, hypothesis).
//   5. select:           sequential replay of sklearn's accept / dynamic-max-trials rules.
//   6. refit:            least squares on the consensus set (centred normal equations, f64),
//      rounded to f32 like sklearn's coef_/intercept_, normalised into [a,b,c,d] with c > 0.
#include "common.cuh"

namespace modest {
extern void note_launch(int n);

// ---- scalar-type traits: float for the seed-mask planes, double for the road planes ------------
template <typename T> struct Ord;
template <> struct Ord<float> {
  using U = unsigned;
  static __device__ __forceinline__ U enc(float v) { return f32_ordered(v); }
  static __device__ __forceinline__ float dec(U u) { return f32_from_ordered(u); }
};
template <> struct Ord<double> {
  using U = unsigned long long;
  static __device__ __forceinline__ U enc(double v) { return f64_ordered(v); }
  static __device__ __forceinline__ double dec(U u) { return f64_from_ordered(u); }
};
__device__ __forceinline__ float t_abs(float v) { return fabsf(v); }
__device__ __forceinline__ double t_abs(double v) { return fabs(v); }
__device__ __forceinline__ float t_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double t_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float t_mean2(float a, float b) { return __fmul_rn(__fadd_rn(a, b), 0.5f); }
__device__ __forceinline__ double t_mean2(double a, double b) { return __dmul_rn(__dadd_rn(a, b), 0.5); }

// ---- block-wide k-th smallest of T values (k 0-based), 1024 threads ----------------------------
template <typename T, typename F>
__device__ T block_kth_smallest(int n, int k, F value, unsigned* hist /*[256] smem*/, unsigned* sel /*[2] smem*/) {
  using U = typename Ord<T>::U;
  U prefix = 0, mask = 0;
  int kk = k;
  for (int shift = 8 * (int)sizeof(T) - 8; shift >= 0; shift -= 8) {
    for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const U u = Ord<T>::enc(value(i));
      if ((u & mask) == prefix) atomicAdd(&hist[(unsigned)(u >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned acc = 0;
      int b = 0;
      for (; b < 256; ++b) {
        if (acc + hist[b] > (unsigned)kk) break;
        acc += hist[b];
      }
      sel[0] = (unsigned)b;
      sel[1] = acc;
    }
    __syncthreads();
    prefix |= (U)sel[0] << shift;
    mask |= (U)255u << shift;
    kk -= (int)sel[1];
    __syncthreads();
  }
  return Ord<T>::dec(prefix);
}

// numpy.median: middle element, or the mean (in T) of the two middle ones.
template <typename T, typename F>
__device__ T block_median(int n, F value, unsigned* hist, unsigned* sel) {
  if (n <= 0) return (T)0;
  const T hi = block_kth_smallest<T>(n, n / 2, value, hist, sel);
  if (n & 1) return hi;
  const T lo = block_kth_smallest<T>(n, n / 2 - 1, value, hist, sel);
  return t_mean2(lo, hi);
}

// ---- 1. ordered compaction of plane candidates -------------------------------------------------
__global__ void __launch_bounds__(1024) plane_candidates_kernel(
    const float* __restrict__ ptc, int stride, const int64_t* __restrict__ off, float max_hs,
    float x_lo, float x_hi, float y_lo, float y_hi, float* __restrict__ cand, int32_t* __restrict__ n_cand) {
  const int s = blockIdx.x;
  const int64_t beg = off[s];
  const int n = (int)(off[s + 1] - beg);
  float* out = cand + 3 * beg;
  __shared__ int warp_cnt[32];
  __shared__ int tile_total;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int base = 0;
  for (int t0 = 0; t0 < n; t0 += 1024) {
    const int i = t0 + threadIdx.x;
    float x = 0, y = 0, z = 0;
    bool keep = false;
    if (i < n) {
      const float* p = ptc + (size_t)stride * (beg + i);
      x = p[0]; y = p[1]; z = p[2];
      keep = (z < max_hs) && (x > x_lo) && (x < x_hi) && (y > y_lo) && (y < y_hi);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[w] = __popc(bal);
    __syncthreads();
    if (w == 0) {
      int c = warp_cnt[lane], inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
      }
      warp_cnt[lane] = inc - c;
      if (lane == 31) tile_total = inc;
    }
    __syncthreads();
    if (keep) {
      const int pos = base + warp_cnt[w] + __popc(bal & ((1u << lane) - 1u));
      out[3 * pos] = x; out[3 * pos + 1] = y; out[3 * pos + 2] = z;
    }
    base += tile_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) n_cand[s] = base;
}

// ---- 2. MAD threshold ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024) mad_threshold_kernel(
    const T* __restrict__ cand, const int64_t* __restrict__ off, const int32_t* __restrict__ n_cand,
    T* __restrict__ thr) {
  const int s = blockIdx.x;
  const T* z = cand + 3 * off[s] + 2;
  const int n = n_cand[s];
  __shared__ unsigned hist[256];
  __shared__ unsigned sel[2];
  const T med = block_median<T>(n, [&](int i) { return z[3 * i]; }, hist, sel);
  const T mad = block_median<T>(n, [&](int i) { return t_abs(t_sub(z[3 * i], med)); }, hist, sel);
  if (threadIdx.x == 0) thr[s] = mad;
}

// ---- 3. hypotheses ------------------------------------------------------------------------------
template <typename T> struct HypT { T a, b, c; int valid; int pad; };

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

template <typename T>
__global__ void ransac_hypotheses_kernel(
    const T* __restrict__ cand, const int64_t* __restrict__ off, const int32_t* __restrict__ n_cand,
    const int32_t* __restrict__ triples, uint64_t seed, const int64_t* __restrict__ scan_keys, int H,
    HypT<T>* __restrict__ hyps, int32_t* __restrict__ triples_out) {
  const int s = blockIdx.x;
  const int n = n_cand[s];
  const T* p = cand + 3 * off[s];
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    HypT<T> out = {(T)0, (T)0, (T)0, 0, 0};
    int id[3] = {0, 0, 0};
    bool ok = n >= 3;
    if (ok) {
      if (triples) {
        for (int k = 0; k < 3; ++k) id[k] = triples[((size_t)s * H + h) * 3 + k];
        ok = id[0] >= 0 && id[1] >= 0 && id[2] >= 0 && id[0] < n && id[1] < n && id[2] < n;
      } else {   // device generator: three distinct indices from a counter-based hash of
                 // (seed, scan key, trial) -- a scan's draws do not depend on its batch slot
        const uint64_t key = scan_keys ? (uint64_t)scan_keys[s] : (uint64_t)s;
        uint64_t ctr = splitmix64(seed ^ (0xD1B54A32D192ED03ull * (key + 1))) + (uint64_t)h * 64u;
        int got = 0;
        for (int it = 0; it < 64 && got < 3; ++it) {
          const uint64_t r = splitmix64(ctr + it);
          const int j = (int)(((r >> 32) * (uint64_t)n) >> 32);
          bool dup = false;
          for (int k = 0; k < got; ++k) dup |= (id[k] == j);
          if (!dup) id[got++] = j;
        }
        ok = got == 3;
      }
    }
    if (ok) {
      // z = a x + b y + c through three points, centred like LinearRegression.fit, f64
      double X[3], Y[3], Z[3];
      for (int k = 0; k < 3; ++k) { X[k] = p[3 * id[k]]; Y[k] = p[3 * id[k] + 1]; Z[k] = p[3 * id[k] + 2]; }
      const double mx = (X[0] + X[1] + X[2]) / 3.0, my = (Y[0] + Y[1] + Y[2]) / 3.0, mz = (Z[0] + Z[1] + Z[2]) / 3.0;
      double sxx = 0, sxy = 0, syy = 0, sxz = 0, syz = 0;
      for (int k = 0; k < 3; ++k) {
        const double dx = X[k] - mx, dy = Y[k] - my, dz = Z[k] - mz;
        sxx += dx * dx; sxy += dx * dy; syy += dy * dy; sxz += dx * dz; syz += dy * dz;
      }
      const double det = sxx * syy - sxy * sxy;
      if (fabs(det) > 1e-30 * (sxx * syy + 1e-300)) {
        const double a = (sxz * syy - syz * sxy) / det, b = (syz * sxx - sxz * sxy) / det;
        out.a = (T)a; out.b = (T)b;
        out.c = (T)(mz - a * mx - b * my);
        out.valid = 1;
      }
    }
    hyps[(size_t)s * H + h] = out;
    if (triples_out) for (int k = 0; k < 3; ++k) triples_out[((size_t)s * H + h) * 3 + k] = id[k];
  }
}

// ---- 4. score all hypotheses --------------------------------------------------------------------
struct HypStat { int count; int pad; double ss_res, sum_z, sum_zz; };

constexpr int kPtsPerThread = 8;

// model evaluation X @ coef + intercept exactly as numpy/OpenBLAS round it (measured on this
// image, tests/test_host_logic.py): sgemv gives fma(y,b, x*a), dgemv gives fma(x,a, y*b)
__device__ __forceinline__ float predict(const HypT<float>& h, float x, float y) {
  return __fadd_rn(fmaf(y, h.b, __fmul_rn(x, h.a)), h.c);
}
__device__ __forceinline__ double predict(const HypT<double>& h, double x, double y) {
  return __dadd_rn(__fma_rn(x, h.a, __dmul_rn(y, h.b)), h.c);
}

template <typename T>
__global__ void __launch_bounds__(256) ransac_score_kernel(
    const T* __restrict__ cand, const int64_t* __restrict__ off, const int32_t* __restrict__ n_cand,
    const T* __restrict__ thr, const HypT<T>* __restrict__ hyps, int H, HypStat* __restrict__ stats) {
  const int s = blockIdx.y;
  const int n = n_cand[s];
  const int chunk = 256 * kPtsPerThread;
  const int start = blockIdx.x * chunk;
  if (start >= n) return;
  extern __shared__ __align__(16) unsigned char sh_raw[];
  HypT<T>* sh_h = reinterpret_cast<HypT<T>*>(sh_raw);
  for (int h = threadIdx.x; h < H; h += blockDim.x) sh_h[h] = hyps[(size_t)s * H + h];
  __syncthreads();
  const T* p = cand + 3 * off[s];
  const T t = thr[s];
  T x[kPtsPerThread], y[kPtsPerThread], z[kPtsPerThread];
  bool live[kPtsPerThread];
#pragma unroll
  for (int k = 0; k < kPtsPerThread; ++k) {
    const int i = start + k * 256 + threadIdx.x;
    live[k] = i < n;
    x[k] = live[k] ? p[3 * i] : (T)0;
    y[k] = live[k] ? p[3 * i + 1] : (T)0;
    z[k] = live[k] ? p[3 * i + 2] : (T)0;
  }
  const int lane = threadIdx.x & 31;
  for (int h = 0; h < H; ++h) {
    const HypT<T> hy = sh_h[h];
    if (!hy.valid) continue;
    int cnt = 0;
    double ssr = 0.0, sz = 0.0, szz = 0.0;
#pragma unroll
    for (int k = 0; k < kPtsPerThread; ++k) {
      const T r = t_abs(t_sub(z[k], predict(hy, x[k], y[k])));
      if (live[k] && r <= t) {
        ++cnt;
        const double rd = (double)r, zd = (double)z[k];
        ssr += rd * rd; sz += zd; szz += zd * zd;
      }
    }
    cnt = warp_sum(cnt);
    if (cnt) {     // warp-uniform after the reduction
      ssr = warp_sum(ssr); sz = warp_sum(sz); szz = warp_sum(szz);
      if (lane == 0) {
        HypStat* st = stats + (size_t)s * H + h;
        atomicAdd(&st->count, cnt);
        atomicAdd(&st->ss_res, ssr);
        atomicAdd(&st->sum_z, sz);
        atomicAdd(&st->sum_zz, szz);
      }
    }
  }
}

// ---- 5. sequential accept / early-stop replay (sklearn/_ransac.py:459-556) ----------------------
__device__ double dynamic_max_trials(int n_inliers, int n_samples) {
  const double eps = 2.220446049250313e-16;
  const double ratio = (double)n_inliers / (double)n_samples;
  const double nom = fmax(eps, 1.0 - 0.99);
  const double denom = fmax(eps, 1.0 - pow(ratio, 3.0));
  if (denom == 1.0) return 1e300;
  return fabs(ceil(log(nom) / log(denom)));
}

template <typename T>
__global__ void ransac_select_kernel(const HypStat* __restrict__ stats, const HypT<T>* __restrict__ hyps,
                                     const int32_t* __restrict__ n_cand, int H, int n_scans,
                                     int32_t* __restrict__ info /* (S,4): n_cand, n_trials, best, best_inliers */) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_scans) return;
  const int n = n_cand[s];
  int best_n = 1, best = -1, trials = 0;
  double best_score = -1e300, max_trials = (double)H;
  while ((double)trials < max_trials && trials < H) {
    const int i = trials++;
    if (!hyps[(size_t)s * H + i].valid) continue;
    const HypStat st = stats[(size_t)s * H + i];
    if (st.count < best_n) continue;
    const double ss_tot = st.sum_zz - st.sum_z * st.sum_z / (double)st.count;
    const double score = ss_tot > 0.0 ? 1.0 - st.ss_res / ss_tot : (st.ss_res == 0.0 ? 1.0 : 0.0);
    if (st.count == best_n && score < best_score) continue;
    best_n = st.count; best_score = score; best = i;
    max_trials = fmin(max_trials, dynamic_max_trials(best_n, n));
  }
  info[4 * s + 0] = n;
  info[4 * s + 1] = trials;
  info[4 * s + 2] = best;
  info[4 * s + 3] = best >= 0 ? best_n : 0;
}

// ---- 6. least-squares refit on the consensus set -> plane ---------------------------------------
__device__ double block_sum(double v, double* sh /*[32]*/) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = lane < (int)(blockDim.x >> 5) ? sh[lane] : 0.0;
    r = warp_sum(r);
    if (lane == 0) sh[0] = r;
  }
  __syncthreads();
  r = sh[0];
  return r;
}

// mode 0: seed-mask plane of utils/pointcloud_utils.py:53-62  -> -[a, b, -1, c]/|(a,b,-1)|  (c > 0)
// mode 1: road plane of data_preprocessing/RANSAC.py:44-52    ->  [a, -1, b, c]/|(a,-1,b)|   (features are
//         rect x and z, target rect y); fewer than 5 candidates -> the script's default [0,-1,0,1.65]
template <typename T>
__global__ void __launch_bounds__(1024) ransac_refit_kernel(
    const T* __restrict__ cand, const int64_t* __restrict__ off, const int32_t* __restrict__ n_cand,
    const T* __restrict__ thr, const HypT<T>* __restrict__ hyps, int H, const int32_t* __restrict__ info,
    double* __restrict__ plane /* (S,4) */, double* __restrict__ model /* (S,3) a,b,c or NULL */,
    uint8_t* __restrict__ inlier_mask /* per candidate, at off[s], or NULL */, int mode) {
  const int s = blockIdx.x;
  const int n = n_cand[s];
  const int best = info[4 * s + 2];
  const T* p = cand + 3 * off[s];
  __shared__ double sh[32];
  if (mode == 1 && n < 5) {
    if (threadIdx.x == 0) { plane[4 * s] = 0.0; plane[4 * s + 1] = -1.0; plane[4 * s + 2] = 0.0; plane[4 * s + 3] = 1.65; }
    return;
  }
  if (best < 0) {
    if (threadIdx.x < 4) plane[4 * s + threadIdx.x] = __longlong_as_double(0x7ff8000000000000ll);
    return;
  }
  const HypT<T> hy = hyps[(size_t)s * H + best];
  const T t = thr[s];
  double cnt = 0, sx = 0, sy = 0, sz = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const T x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
    const bool in = t_abs(t_sub(z, predict(hy, x, y))) <= t;
    if (inlier_mask) inlier_mask[off[s] + i] = in ? 1 : 0;
    if (in) { cnt += 1.0; sx += x; sy += y; sz += z; }
  }
  cnt = block_sum(cnt, sh); sx = block_sum(sx, sh); sy = block_sum(sy, sh); sz = block_sum(sz, sh);
  const double mx = sx / cnt, my = sy / cnt, mz = sz / cnt;
  double sxx = 0, sxy = 0, syy = 0, sxz = 0, syz = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const T x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
    if (t_abs(t_sub(z, predict(hy, x, y))) <= t) {
      const double dx = x - mx, dy = y - my, dz = z - mz;
      sxx += dx * dx; sxy += dx * dy; syy += dy * dy; sxz += dx * dz; syz += dy * dz;
    }
  }
  sxx = block_sum(sxx, sh); sxy = block_sum(sxy, sh); syy = block_sum(syy, sh);
  sxz = block_sum(sxz, sh); syz = block_sum(syz, sh);
  if (threadIdx.x == 0) {
    const double det = sxx * syy - sxy * sxy;
    const double a = (sxz * syy - syz * sxy) / det, b = (syz * sxx - sxz * sxy) / det;
    // sklearn keeps coef_/intercept_ in the input dtype
    const T af = (T)a, bf = (T)b;
    const T cf = (T)(mz - a * mx - b * my);
    if (model) { model[3 * s] = af; model[3 * s + 1] = bf; model[3 * s + 2] = cf; }
    const double wa = af, wb = bf;
    const double nrm = sqrt(wa * wa + wb * wb + 1.0);
    if (mode == 0) {
      plane[4 * s + 0] = -(wa / nrm);
      plane[4 * s + 1] = -(wb / nrm);
      plane[4 * s + 2] = -(-1.0 / nrm);
      plane[4 * s + 3] = -((double)cf / nrm);
    } else {
      plane[4 * s + 0] = wa / nrm;
      plane[4 * s + 1] = -1.0 / nrm;
      plane[4 * s + 2] = wb / nrm;
      plane[4 * s + 3] = (double)cf / nrm;
    }
  }
}

// road-plane candidates (data_preprocessing/RANSAC.py:31-37): rect coordinates in f64, keep
// min_h < y < max_h, -10 < z < 70, -20 < x < 20 in order; candidate rows are (x, z, y)
struct RoadCalib { double v2c[12]; double r0[9]; };

__global__ void __launch_bounds__(1024) road_candidates_kernel(
    const float* __restrict__ ptc, int stride, const int64_t* __restrict__ off, const RoadCalib* __restrict__ calibs,
    double min_h, double max_h, double* __restrict__ cand, int32_t* __restrict__ n_cand) {
  const int s = blockIdx.x;
  const int64_t beg = off[s];
  const int n = (int)(off[s + 1] - beg);
  const RoadCalib cb = calibs[s];
  double* out = cand + 3 * beg;
  __shared__ int warp_cnt[32];
  __shared__ int tile_total;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int base = 0;
  for (int t0 = 0; t0 < n; t0 += 1024) {
    const int i = t0 + threadIdx.x;
    double r[3] = {0, 0, 0};
    bool keep = false;
    if (i < n) {
      const float* p = ptc + (size_t)stride * (beg + i);
      double ref[3];
#pragma unroll
      for (int k = 0; k < 3; ++k)      // [p,1] @ V2C^T then R0 @ ref, dgemm's k-ordered fused multiply-adds
        ref[k] = __fma_rn(1.0, cb.v2c[4 * k + 3], __fma_rn((double)p[2], cb.v2c[4 * k + 2],
                          __fma_rn((double)p[1], cb.v2c[4 * k + 1], __dmul_rn((double)p[0], cb.v2c[4 * k]))));
#pragma unroll
      for (int k = 0; k < 3; ++k)
        r[k] = __fma_rn(cb.r0[3 * k + 2], ref[2], __fma_rn(cb.r0[3 * k + 1], ref[1], __dmul_rn(cb.r0[3 * k], ref[0])));
      keep = (r[1] > min_h) && (r[1] < max_h) && (r[2] > -10.0) && (r[2] < 70.0) && (r[0] > -20.0) && (r[0] < 20.0);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[w] = __popc(bal);
    __syncthreads();
    if (w == 0) {
      int c = warp_cnt[lane], inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
      }
      warp_cnt[lane] = inc - c;
      if (lane == 31) tile_total = inc;
    }
    __syncthreads();
    if (keep) {
      const int pos = base + warp_cnt[w] + __popc(bal & ((1u << lane) - 1u));
      out[3 * pos] = r[0]; out[3 * pos + 1] = r[2]; out[3 * pos + 2] = r[1];
    }
    base += tile_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) n_cand[s] = base;
}

template <typename T>
static int ransac_fit_impl(const T* d_cand, const int64_t* d_off, const int32_t* d_n_cand, const T* d_thr, int n_scans,
                           int64_t max_points, const int32_t* d_triples, uint64_t seed, const int64_t* d_scan_keys,
                           int max_trials, double* d_plane, double* d_model, int32_t* d_info, int32_t* d_triples_out, uint8_t* d_inlier_mask, void* d_ws,
                           size_t ws_bytes, cudaStream_t stream, int mode) {
  Arena ar(d_ws, ws_bytes);
  HypT<T>* hyps = ar.take<HypT<T>>((size_t)n_scans * max_trials);
  HypStat* stats = ar.take<HypStat>((size_t)n_scans * max_trials);
  MODEST_REQUIRE(ar.ok(), "workspace too small for the requested sizes");
  MODEST_CUDA(cudaMemsetAsync(stats, 0, sizeof(HypStat) * (size_t)n_scans * max_trials, stream));
  ransac_hypotheses_kernel<T><<<n_scans, 128, 0, stream>>>(d_cand, d_off, d_n_cand, d_triples, seed, d_scan_keys, max_trials,
                                                          hyps, d_triples_out);
  MODEST_LAUNCH_CHECK("ransac_hypotheses_kernel");
  const int chunk = 256 * kPtsPerThread;
  dim3 grid((unsigned)((max_points + chunk - 1) / chunk), n_scans);
  if (grid.x == 0) grid.x = 1;
  ransac_score_kernel<T><<<grid, 256, sizeof(HypT<T>) * max_trials, stream>>>(d_cand, d_off, d_n_cand, d_thr, hyps,
                                                                             max_trials, stats);
  MODEST_LAUNCH_CHECK("ransac_score_kernel");
  ransac_select_kernel<T><<<(n_scans + 63) / 64, 64, 0, stream>>>(stats, hyps, d_n_cand, max_trials, n_scans, d_info);
  MODEST_LAUNCH_CHECK("ransac_select_kernel");
  ransac_refit_kernel<T><<<n_scans, 1024, 0, stream>>>(d_cand, d_off, d_n_cand, d_thr, hyps, max_trials, d_info, d_plane,
                                                     d_model, d_inlier_mask, mode);
  MODEST_LAUNCH_CHECK("ransac_refit_kernel");
  note_launch(4);
  return MODEST_OK;
}

}  // namespace modest

using namespace modest;

extern "C" int modest_plane_candidates_batch(const float* d_ptc, int point_stride, const int64_t* d_off,
                                             int n_scans, float max_hs, float x_lo, float x_hi, float y_lo,
                                             float y_hi, float* d_cand, int32_t* d_n_cand, float* d_thr,
                                             void* stream_) {
  modest::StageRange nvtx_("modest:E plane candidates");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_scans <= 0) return MODEST_OK;
  MODEST_REQUIRE(d_ptc && d_off && d_cand && d_n_cand && d_thr, "plane_candidates: null pointer argument");
  MODEST_REQUIRE(point_stride >= 3, "plane_candidates: point_stride %d < 3", point_stride);
  plane_candidates_kernel<<<n_scans, 1024, 0, stream>>>(d_ptc, point_stride, d_off, max_hs, x_lo, x_hi, y_lo,
                                                        y_hi, d_cand, d_n_cand);
  MODEST_LAUNCH_CHECK("plane_candidates_kernel");
  mad_threshold_kernel<float><<<n_scans, 1024, 0, stream>>>(d_cand, d_off, d_n_cand, d_thr);
  MODEST_LAUNCH_CHECK("mad_threshold_kernel");
  note_launch(2);
  return MODEST_OK;
}

extern "C" size_t modest_ransac_workspace_bytes(int n_scans, int max_trials) {
  return align_up(sizeof(HypT<double>) * (size_t)n_scans * max_trials, 256) +
         align_up(sizeof(HypStat) * (size_t)n_scans * max_trials, 256) + 512;
}

static int fit_args_ok(const void* d_cand, const void* d_off, const void* d_n_cand, const void* d_thr, const void* d_plane,
                       const void* d_info, const void* d_ws, int n_scans, int max_trials, size_t ws_bytes) {
  MODEST_REQUIRE(d_cand && d_off && d_n_cand && d_thr && d_plane && d_info && d_ws, "ransac_fit: null pointer argument");
  MODEST_REQUIRE(max_trials >= 1 && max_trials <= 1024, "ransac_fit: max_trials %d out of range", max_trials);
  MODEST_REQUIRE(ws_bytes >= modest_ransac_workspace_bytes(n_scans, max_trials), "ransac_fit: workspace too small");
  MODEST_REQUIRE(n_scans <= 65535, "ransac_fit: more than 65535 scans in one launch");
  return MODEST_OK;
}

extern "C" int modest_ransac_fit_batch(const float* d_cand, const int64_t* d_off, const int32_t* d_n_cand,
                                       const float* d_thr, int n_scans, int64_t max_points,
                                       const int32_t* d_triples, uint64_t seed, const int64_t* d_scan_keys,
                                       int max_trials, double* d_plane, double* d_model, int32_t* d_info,
                                       int32_t* d_triples_out, uint8_t* d_inlier_mask, void* d_ws,
                                       size_t ws_bytes, void* stream_) {
  modest::StageRange nvtx_("modest:E RANSAC fit");
  if (n_scans <= 0) return MODEST_OK;
  const int rc = fit_args_ok(d_cand, d_off, d_n_cand, d_thr, d_plane, d_info, d_ws, n_scans, max_trials, ws_bytes);
  if (rc != MODEST_OK) return rc;
  return ransac_fit_impl<float>(d_cand, d_off, d_n_cand, d_thr, n_scans, max_points, d_triples, seed, d_scan_keys, max_trials,
                                d_plane, d_model, d_info, d_triples_out, d_inlier_mask, d_ws, ws_bytes,
                                static_cast<cudaStream_t>(stream_), 0);
}

extern "C" int modest_road_candidates_batch(const float* d_ptc, int point_stride, const int64_t* d_off,
                                            const double* d_calib, int n_scans, double min_h, double max_h,
                                            double* d_cand, int32_t* d_n_cand, double* d_thr, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_scans <= 0) return MODEST_OK;
  MODEST_REQUIRE(d_ptc && d_off && d_calib && d_cand && d_n_cand && d_thr, "road_candidates: null pointer argument");
  MODEST_REQUIRE(point_stride >= 3, "road_candidates: point_stride %d < 3", point_stride);
  road_candidates_kernel<<<n_scans, 1024, 0, stream>>>(d_ptc, point_stride, d_off,
                                                       reinterpret_cast<const RoadCalib*>(d_calib), min_h, max_h, d_cand,
                                                       d_n_cand);
  MODEST_LAUNCH_CHECK("road_candidates_kernel");
  mad_threshold_kernel<double><<<n_scans, 1024, 0, stream>>>(d_cand, d_off, d_n_cand, d_thr);
  MODEST_LAUNCH_CHECK("mad_threshold_kernel");
  note_launch(2);
  return MODEST_OK;
}

extern "C" int modest_road_plane_fit_batch(const double* d_cand, const int64_t* d_off, const int32_t* d_n_cand,
                                           const double* d_thr, int n_scans, int64_t max_points,
                                           const int32_t* d_triples, uint64_t seed, int max_trials,
                                           double* d_plane, int32_t* d_info, void* d_ws, size_t ws_bytes,
                                           void* stream_) {
  modest::StageRange nvtx_("modest:f-3 road plane fit");
  if (n_scans <= 0) return MODEST_OK;
  const int rc = fit_args_ok(d_cand, d_off, d_n_cand, d_thr, d_plane, d_info, d_ws, n_scans, max_trials, ws_bytes);
  if (rc != MODEST_OK) return rc;
  return ransac_fit_impl<double>(d_cand, d_off, d_n_cand, d_thr, n_scans, max_points, d_triples, seed, nullptr, max_trials,
                                 d_plane, nullptr, d_info, nullptr, nullptr, d_ws, ws_bytes,
                                 static_cast<cudaStream_t>(stream_), 1);
}
