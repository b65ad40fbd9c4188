// Stages J-M: cluster filtering, closeness-to-edge box fitting, volume gate, relabelling.
//
// Reference behaviour:
//   J  filter_labels / is_valid_cluster        utils/clustering_utils.py:94-135
//   K  Calibration.project_velo_to_rect        utils/kitti_util.py:293-329
//   L  get_obj(..., 'closeness_to_edge')       utils/pointcloud_utils.py:167-216,278-317
//   M  volume gate + label compaction          generate_mask.py:92-103
//
// Layout: cluster c of scan s is slot s*max_clusters + c in every per-cluster array.
#include "common.cuh"

namespace modest {
extern void note_launch(int n);

struct ClusterStat {
  int count;
  int valid;
  unsigned long long dmin, dmax;   // ordered-u64 images of the f64 signed ground distances
};

__device__ __forceinline__ double plane_distance2(float x, float y, float z, const double* pl, double nrm) {
  // numpy hands (N,3) @ (3,) to OpenBLAS dgemv, whose kernel rounds as fma(z,c, fma(x,a, y*b))
  double d = __fma_rn((double)z, pl[2], __fma_rn((double)x, pl[0], __dmul_rn((double)y, pl[1])));
  d = __dadd_rn(d, pl[3]);
  return __ddiv_rn(d, nrm);
}

// ---- J.1 per-cluster count / min / max ground distance ---------------------------------------
__global__ void __launch_bounds__(256) cluster_stats_kernel(
    const float* __restrict__ ptc, int stride, const int64_t* __restrict__ off, const int32_t* __restrict__ labels,
    const double* __restrict__ planes, int max_clusters, ClusterStat* __restrict__ stats, int32_t* __restrict__ flags) {
  const int s = blockIdx.y;
  const int64_t beg = off[s];
  const int n = (int)(off[s + 1] - beg);
  double pl[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) pl[k] = planes[4 * s + k];
  const double nrm = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(pl[0], pl[0]), __dmul_rn(pl[1], pl[1])), __dmul_rn(pl[2], pl[2])));
  ClusterStat* st = stats + (size_t)s * max_clusters;
  const int lane = threadIdx.x & 31;
  // consecutive returns of a beam mostly belong to one cluster: when every labelled lane of the warp
  // holds the same cluster the warp reduces first and issues three atomics instead of 96 (the big
  // clusters serialised thousands of atomics on one record)
  for (int base = blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < n; base += gridDim.x * blockDim.x) {   // warp-uniform
    const int i = base + lane;
    int c = i < n ? labels[beg + i] : -1;
    if (c >= max_clusters) { atomicOr(flags, 4); c = -1; }
    unsigned long long d = 0ull;
    if (c >= 0) {
      const float* p = ptc + (size_t)stride * (beg + i);
      d = f64_ordered(plane_distance2(p[0], p[1], p[2], pl, nrm));
    }
    const unsigned vm = __ballot_sync(0xffffffffu, c >= 0);
    if (vm == 0u) continue;
    const int c0 = __shfl_sync(0xffffffffu, c, __ffs(vm) - 1);
    if (__all_sync(0xffffffffu, c < 0 || c == c0)) {
      unsigned long long lo = c >= 0 ? d : 0xffffffffffffffffull, hi = c >= 0 ? d : 0ull;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, lo, o), b = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = a < lo ? a : lo;
        hi = b > hi ? b : hi;
      }
      if (lane == 0) {
        atomicAdd(&st[c0].count, __popc(vm));
        atomicMin(&st[c0].dmin, lo);
        atomicMax(&st[c0].dmax, hi);
      }
    } else if (c >= 0) {
      atomicAdd(&st[c].count, 1);
      atomicMin(&st[c].dmin, d);
      atomicMax(&st[c].dmax, d);
    }
  }
}

__global__ void cluster_stats_init_kernel(ClusterStat* __restrict__ stats, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    stats[i].count = 0; stats[i].valid = 0;
    stats[i].dmin = 0xffffffffffffffffull; stats[i].dmax = 0ull;
  }
}

// ---- J.2 exclusive scan of cluster sizes per scan, then scatter member data -------------------
__global__ void __launch_bounds__(1024) cluster_offsets_kernel(const ClusterStat* __restrict__ stats, int max_clusters,
                                                              const int32_t* __restrict__ n_clusters,
                                                              int32_t* __restrict__ cl_off /* (S, max_clusters+1) */,
                                                              int32_t* __restrict__ cl_fill) {
  const int s = blockIdx.x;
  const int C = min(n_clusters[s], max_clusters);
  const ClusterStat* st = stats + (size_t)s * max_clusters;
  int32_t* o = cl_off + (size_t)s * (max_clusters + 1);
  __shared__ int warp_tot[32];
  __shared__ int tile_total;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int carry = 0;
  for (int t0 = 0; t0 < C; t0 += 1024) {
    const int c = t0 + threadIdx.x;
    const int v = c < C ? st[c].count : 0;
    int inc = v;
#pragma unroll
    for (int of = 1; of < 32; of <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, inc, of);
      if (lane >= of) inc += u;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    if (w == 0) {
      int t = warp_tot[lane], ti = t;
#pragma unroll
      for (int of = 1; of < 32; of <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, ti, of);
        if (lane >= of) ti += u;
      }
      warp_tot[lane] = ti - t;
      if (lane == 31) tile_total = ti;
    }
    __syncthreads();
    if (c < C) { o[c] = carry + warp_tot[w] + inc - v; cl_fill[(size_t)s * max_clusters + c] = 0; }
    carry += tile_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) o[C] = carry;
}

// member arrays are packed per scan at off[s]; cluster c's members at off[s] + cl_off[c]
__global__ void __launch_bounds__(256) cluster_scatter_kernel(
    const int64_t* __restrict__ off, const int32_t* __restrict__ labels, int max_clusters,
    const int32_t* __restrict__ cl_off, int32_t* __restrict__ cl_fill, int32_t* __restrict__ members) {
  const int s = blockIdx.y;
  const int64_t beg = off[s];
  const int n = (int)(off[s + 1] - beg);
  const int lane = threadIdx.x & 31;
  for (int base = blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < n; base += gridDim.x * blockDim.x) {   // warp-uniform
    const int i = base + lane;
    int c = i < n ? labels[beg + i] : -1;
    if (c >= max_clusters) c = -1;
    const unsigned vm = __ballot_sync(0xffffffffu, c >= 0);
    if (vm == 0u) continue;
    const int c0 = __shfl_sync(0xffffffffu, c, __ffs(vm) - 1);
    if (__all_sync(0xffffffffu, c < 0 || c == c0)) {                 // one cluster in the warp: one atomic
      int first = 0;
      if (lane == 0) first = atomicAdd(&cl_fill[(size_t)s * max_clusters + c0], __popc(vm));
      first = __shfl_sync(0xffffffffu, first, 0);
      if (c >= 0) members[beg + cl_off[(size_t)s * (max_clusters + 1) + c0] + first + __popc(vm & ((1u << lane) - 1u))] = i;
    } else if (c >= 0) {
      const int pos = cl_off[(size_t)s * (max_clusters + 1) + c] + atomicAdd(&cl_fill[(size_t)s * max_clusters + c], 1);
      members[beg + pos] = i;
    }
  }
}

// ---- J.3 percentile gate + the other gates, one CTA per cluster --------------------------------
struct FilterCfg {
  int min_points;
  double max_min_height, min_max_height;
  float q;                      // percentile/100 in float32 (numpy: q / f32(100))
  float min_percentile_pp;      // compared as float32 > python float -> widened to f64
  double min_percentile_pp_d;
};

template <typename F>
__device__ float block256_kth_smallest(int n, int k, F value, unsigned* hist, unsigned* sel) {
  unsigned prefix = 0, mask = 0;
  int kk = k;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      unsigned u = f32_ordered(value(i));
      if ((u & mask) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned acc = 0;
      int b = 0;
      for (; b < 256; ++b) {
        if (acc + hist[b] > (unsigned)kk) break;
        acc += hist[b];
      }
      sel[0] = (unsigned)b; sel[1] = acc;
    }
    __syncthreads();
    prefix |= sel[0] << shift;
    mask |= 255u << shift;
    kk -= (int)sel[1];
    __syncthreads();
  }
  return f32_from_ordered(prefix);
}

__global__ void __launch_bounds__(256) cluster_validate_kernel(
    const int64_t* __restrict__ off, const float* __restrict__ pp, int max_clusters, const int32_t* __restrict__ n_clusters,
    const int32_t* __restrict__ cl_off, const int32_t* __restrict__ members, FilterCfg cfg, ClusterStat* __restrict__ stats) {
  const int s = blockIdx.y;
  const int C = min(n_clusters[s], max_clusters);
  __shared__ unsigned hist[256];
  __shared__ unsigned sel[2];
  for (int c = blockIdx.x; c < C; c += gridDim.x) {
    ClusterStat* st = stats + (size_t)s * max_clusters + c;
    const int n = st->count;
    bool ok = n >= cfg.min_points;
    if (ok) {
      const double dmin = f64_from_ordered(st->dmin), dmax = f64_from_ordered(st->dmax);
      ok = !(dmin > cfg.max_min_height) && !(dmax < cfg.min_max_height);
    }
    if (ok) {   // block-uniform
      // numpy 2.x percentile on a float32 vector, method 'linear': everything in float32
      const int32_t* mem = members + off[s] + cl_off[(size_t)s * (max_clusters + 1) + c];
      const float* v = pp + off[s];
      const float vi = __fmul_rn((float)(n - 1), cfg.q);
      const int lo = (int)floorf(vi);
      const int hi = min(lo + 1, n - 1);
      const float g = __fsub_rn(vi, (float)lo);
      const float a = block256_kth_smallest(n, lo, [&](int i) { return v[mem[i]]; }, hist, sel);
      const float b = hi == lo ? a : block256_kth_smallest(n, hi, [&](int i) { return v[mem[i]]; }, hist, sel);
      const float d = __fsub_rn(b, a);
      float r;
      if (g >= 0.5f) r = __fsub_rn(b, __fmul_rn(d, __fsub_rn(1.0f, g)));
      else r = __fadd_rn(a, __fmul_rn(d, g));
      ok = !((double)r > cfg.min_percentile_pp_d);
    }
    if (threadIdx.x == 0) st->valid = ok ? 1 : 0;
    __syncthreads();
  }
}

// ---- J.4 relabel: valid clusters -> 1..K in id order (0 = everything else) --------------------
__global__ void cluster_relabel_kernel(const int64_t* __restrict__ off, int max_clusters, const int32_t* __restrict__ n_clusters,
                                       const ClusterStat* __restrict__ stats, int32_t* __restrict__ new_id /* (S,max_clusters) */,
                                       int32_t* __restrict__ n_valid, int32_t* __restrict__ has_noise, int n_scans,
                                       int32_t* __restrict__ valid_list /* (S,max_valid) cluster id per rank */, int max_valid) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_scans) return;
  const int C = min(n_clusters[s], max_clusters);
  const ClusterStat* st = stats + (size_t)s * max_clusters;
  long long covered = 0;
  for (int c = 0; c < C; ++c) if (st[c].valid) covered += st[c].count;
  const int n = (int)(off[s + 1] - off[s]);
  // sorted(set(labels)): -1 exists iff some point is not in a valid cluster
  const int noise = covered < n ? 1 : 0;
  int k = 0;
  for (int c = 0; c < C; ++c) {
    if (st[c].valid) {
      new_id[(size_t)s * max_clusters + c] = k + noise;
      if (k < max_valid) valid_list[(size_t)s * max_valid + k] = c;
      ++k;
    }
    else new_id[(size_t)s * max_clusters + c] = noise ? 0 : -1;
  }
  n_valid[s] = k;
  has_noise[s] = noise;
}

__global__ void __launch_bounds__(256) apply_relabel_kernel(const int64_t* __restrict__ off, const int32_t* __restrict__ labels_in,
                                                            int max_clusters, const int32_t* __restrict__ new_id,
                                                            int32_t* __restrict__ labels_out) {
  const int s = blockIdx.y;
  const int64_t beg = off[s];
  const int n = (int)(off[s + 1] - beg);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int c = labels_in[beg + i];
    labels_out[beg + i] = (c < 0 || c >= max_clusters) ? 0 : new_id[(size_t)s * max_clusters + c];
  }
}

// ---- K+L box fit -------------------------------------------------------------------------------
struct CalibDev { double v2c[12]; double r0[9]; };

__device__ __forceinline__ void velo_to_rect(const CalibDev& cb, float x, float y, float z, double* r) {
  // ref = [p,1] @ V2C^T ; rect = R0 @ ref       (kitti_util.py:293-329, f64 throughout)
  // both products go through dgemm, which accumulates the k terms with fused multiply-adds
  // in k order (checked against numpy + OpenBLAS in tests/test_host_numerics.py)
  double ref[3];
#pragma unroll
  for (int k = 0; k < 3; ++k)
    ref[k] = __fma_rn(1.0, cb.v2c[4 * k + 3], __fma_rn((double)z, cb.v2c[4 * k + 2],
                      __fma_rn((double)y, cb.v2c[4 * k + 1], __dmul_rn((double)x, cb.v2c[4 * k]))));
#pragma unroll
  for (int k = 0; k < 3; ++k)
    r[k] = __fma_rn(cb.r0[3 * k + 2], ref[2], __fma_rn(cb.r0[3 * k + 1], ref[1], __dmul_rn(cb.r0[3 * k], ref[0])));
}

// rect coordinates of every point (f64 x,y,z), packed per scan
__global__ void __launch_bounds__(256) rect_coords_kernel(const float* __restrict__ ptc, int stride, const int64_t* __restrict__ off,
                                                          const CalibDev* __restrict__ calibs, double* __restrict__ rect) {
  const int s = blockIdx.y;
  const int64_t beg = off[s];
  const int n = (int)(off[s + 1] - beg);
  const CalibDev cb = calibs[s];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float* p = ptc + (size_t)stride * (beg + i);
    double r[3];
    velo_to_rect(cb, p[0], p[1], p[2], r);
    rect[3 * (beg + i)] = r[0]; rect[3 * (beg + i) + 1] = r[1]; rect[3 * (beg + i) + 2] = r[2];
  }
}

struct BoxRec {            // per cluster slot
  double t[3], l, w, h, ry, volume;   // the SimpleNamespace fields of pointcloud_utils.py:310-316
  double area, ymin;                  // footprint area, min rect-y of the cluster
  double cosr, sinr;                  // cos(ry), sin(ry) used by the footprint test
  int keep, pad;
};

__device__ __forceinline__ void project(double x, double z, double c, double s, double* px, double* py) {
  // cluster_ptc @ [[c, s], [-s, c]].T  -- dgemm rounding: fma(second term, first product)
  *px = __fma_rn(z, s, __dmul_rn(x, c));
  *py = __fma_rn(z, c, __dmul_rn(x, -s));
}

constexpr int kFitThreads = 256;
constexpr int kCandCap = 64;            // f64 re-scoring handles at most this many near-maximal angles
constexpr float kBetaMargin = 5e-3f;    // >> 2x the relative error of the float32 pre-pass
constexpr int kAngleChunks = 15;        // the pre-pass splits the 901 angles over this many CTAs per cluster

// ---- L.0 float32 pre-pass of the closeness score over all search angles -------------------------
// box_center32: member coordinates of every valid cluster, centred on the cluster's first point
// in f64 and only then rounded, so float32 keeps ~1e-6 m accuracy whatever the range; stored
// cluster-contiguously (same layout as `members`).
__global__ void __launch_bounds__(256) box_center32_kernel(
    const int64_t* __restrict__ off, const double* __restrict__ rect, int max_clusters, const ClusterStat* __restrict__ stats,
    const int32_t* __restrict__ cl_off, const int32_t* __restrict__ members, const int32_t* __restrict__ n_valid,
    const int32_t* __restrict__ valid_list, int max_valid, float2* __restrict__ xz32) {
  const int s = blockIdx.z, rank = blockIdx.y;
  if (rank >= min(n_valid[s], max_valid)) return;
  const int c = valid_list[(size_t)s * max_valid + rank];
  if (stats[(size_t)s * max_clusters + c].valid != 1) return;        // certified to fail the volume gate: not fitted
  const int n = stats[(size_t)s * max_clusters + c].count;
  const int64_t mbase = off[s] + cl_off[(size_t)s * (max_clusters + 1) + c];
  const int32_t* mem = members + mbase;
  const double* R = rect + 3 * off[s];
  const double x0 = R[3 * mem[0]], z0 = R[3 * mem[0] + 2];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int m = mem[i];
    xz32[mbase + i] = make_float2((float)(R[3 * m] - x0), (float)(R[3 * m + 2] - z0));
  }
}

// ---- L.-1 clusters that cannot pass the volume gate are not fitted at all -------------------------
// A few clusters per scan are huge (strips of ground beyond plane_estimate.range, long walls:
// thousands of points) and are thrown away by the volume gate of generate_mask.py:95 after the
// 901-heading search, which they dominate.  This kernel certifies "volume > max_volume" from
// three facts that hold for EVERY heading the search could choose, so the search is skipped
// without changing any result:
//   * a rectangle that contains a triangle has at least twice its area, so the footprint area is
//     >= |AB| * dist(C, AB) for any three cluster points A, B, C (here: the farthest pair of the
//     four axis-extreme points and the point farthest from their line);
//   * a point with cluster points in all four open quadrants around it (1 mm margin) lies inside
//     their convex hull, hence strictly inside every rectangle that contains the cluster, hence
//     inside the footprint test of get_lowest_point_rect: bottom >= its y, h >= y - ymin;
//   * volume = area * h.
// ClusterStat::valid becomes 2 for a certified cluster: still valid for filter_labels' output,
// skipped by the fit, dropped by the finalize step exactly as the gate would drop it.
struct ArgD { double v; int i; };
__device__ __forceinline__ ArgD block_argmax(double v, int i, ArgD* sh /*[32]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = ArgD{v, i};
  __syncthreads();
  ArgD r = sh[0];
  for (int k = 1; k < (int)(blockDim.x >> 5); ++k)
    if (sh[k].v > r.v || (sh[k].v == r.v && sh[k].i < r.i)) r = sh[k];
  return r;
}

__global__ void __launch_bounds__(256) box_prereject_kernel(
    const int64_t* __restrict__ off, const double* __restrict__ rect, int max_clusters, ClusterStat* __restrict__ stats,
    const int32_t* __restrict__ cl_off, const int32_t* __restrict__ members, const int32_t* __restrict__ n_valid,
    const int32_t* __restrict__ valid_list, int max_valid, double max_volume) {
  const int s = blockIdx.y, rank = blockIdx.x;
  if (rank >= min(n_valid[s], max_valid)) return;
  const int c = valid_list[(size_t)s * max_valid + rank];
  ClusterStat* st = stats + (size_t)s * max_clusters + c;
  const int n = st->count;
  if (n < 64 || !(max_volume > 0.0) || max_volume > 1e200) return;     // small clusters are cheap to fit
  const int32_t* mem = members + off[s] + cl_off[(size_t)s * (max_clusters + 1) + c];
  const double* R = rect + 3 * off[s];
  __shared__ ArgD sh[32];
  const double kNegInf = -1e300;
  // axis-extreme points and the y range
  double kx0 = kNegInf, kx1 = kNegInf, kz0 = kNegInf, kz1 = kNegInf, ky0 = kNegInf, ky1 = kNegInf;
  int ix0 = 0, ix1 = 0, iz0 = 0, iz1 = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int m = mem[i];
    const double x = R[3 * m], y = R[3 * m + 1], z = R[3 * m + 2];
    if (-x > kx0) { kx0 = -x; ix0 = m; }
    if (x > kx1) { kx1 = x; ix1 = m; }
    if (-z > kz0) { kz0 = -z; iz0 = m; }
    if (z > kz1) { kz1 = z; iz1 = m; }
    ky0 = fmax(ky0, -y); ky1 = fmax(ky1, y);
  }
  const ArgD ax0 = block_argmax(kx0, ix0, sh), ax1 = block_argmax(kx1, ix1, sh);
  const ArgD az0 = block_argmax(kz0, iz0, sh), az1 = block_argmax(kz1, iz1, sh);
  const double ymin = -block_argmax(ky0, 0, sh).v, ymax = block_argmax(ky1, 0, sh).v;
  const int ext[4] = {ax0.i, ax1.i, az0.i, az1.i};
  double best = -1.0, Ax = 0, Az = 0, Bx = 0, Bz = 0;
  for (int a = 0; a < 4; ++a)
    for (int b = a + 1; b < 4; ++b) {
      const double dx = R[3 * ext[a]] - R[3 * ext[b]], dz = R[3 * ext[a] + 2] - R[3 * ext[b] + 2];
      const double d2 = dx * dx + dz * dz;
      if (d2 > best) { best = d2; Ax = R[3 * ext[a]]; Az = R[3 * ext[a] + 2]; Bx = R[3 * ext[b]]; Bz = R[3 * ext[b] + 2]; }
    }
  if (!(best > 0.0)) return;
  // |AB| * dist(C, AB) = max |cross(B - A, C - A)|
  double kc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int m = mem[i];
    kc = fmax(kc, fabs((Bx - Ax) * (R[3 * m + 2] - Az) - (Bz - Az) * (R[3 * m] - Ax)));
  }
  const double area_lb = block_argmax(kc, 0, sh).v * (1.0 - 1e-9);
  if (!(area_lb > 0.0)) return;
  const double h_need = max_volume / area_lb * (1.0 + 1e-6) + 1e-9;
  if (!(ymax - ymin > h_need)) return;                                // block-uniform
  // among the points low enough, the one nearest the middle of the cluster is the likeliest interior point
  const double cx = 0.5 * (ax1.v - ax0.v), cz = 0.5 * (az1.v - az0.v);
  const double ex = fmax(ax1.v + ax0.v, 1e-9), ez = fmax(az1.v + az0.v, 1e-9);
  double kp = kNegInf;
  int ip = -1;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int m = mem[i];
    if (R[3 * m + 1] - ymin > h_need) {
      const double d = -(fabs(R[3 * m] - cx) / ex + fabs(R[3 * m + 2] - cz) / ez);
      if (d > kp) { kp = d; ip = m; }
    }
  }
  const ArgD cand = block_argmax(kp, ip < 0 ? 0x7fffffff : ip, sh);
  if (!(cand.v > kNegInf)) return;
  const double px = R[3 * cand.i], pz = R[3 * cand.i + 2], delta = 1e-3;
  int q = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int m = mem[i];
    const double x = R[3 * m], z = R[3 * m + 2];
    if (x > px + delta && z > pz + delta) q |= 1;
    if (x < px - delta && z > pz + delta) q |= 2;
    if (x < px - delta && z < pz - delta) q |= 4;
    if (x > px + delta && z < pz - delta) q |= 8;
  }
  // (__syncthreads_or tells whether ANY thread's predicate is non-zero: one call per quadrant)
  const int q1 = __syncthreads_or(q & 1), q2 = __syncthreads_or(q & 2), q3 = __syncthreads_or(q & 4), q4 = __syncthreads_or(q & 8);
  if (q1 && q2 && q3 && q4 && threadIdx.x == 0) st->valid = 2;
}

// grid (kAngleChunks, valid-cluster rank, scan).  One THREAD per search angle, the cluster's points
// broadcast from shared memory: every lane is busy whatever the cluster size and no heading needs a
// warp reduction (one warp per angle with lanes over the points spent 28 % of its instructions on
// the shuffles of the four extrema and the sum, and ran 30-60-point clusters with half the lanes).
// The extrema are order-independent; the float32 sum is only used to shortlist angles (0.5 %).
constexpr int kBetaThreads = 64;
constexpr int kBetaStage = 512;
__global__ void __launch_bounds__(kBetaThreads) box_beta32_kernel(
    const int64_t* __restrict__ off, int max_clusters, const ClusterStat* __restrict__ stats, const int32_t* __restrict__ cl_off,
    const int32_t* __restrict__ n_valid, const int32_t* __restrict__ valid_list, int max_valid, const float2* __restrict__ xz32,
    const double* __restrict__ trig, int n_angles, float d0, float* __restrict__ beta32) {
  const int s = blockIdx.z, rank = blockIdx.y;
  if (rank >= min(n_valid[s], max_valid)) return;
  const int c = valid_list[(size_t)s * max_valid + rank];
  if (stats[(size_t)s * max_clusters + c].valid != 1) return;        // certified to fail the volume gate: not fitted
  const int n = stats[(size_t)s * max_clusters + c].count;
  const float2* __restrict__ pts = xz32 + off[s] + cl_off[(size_t)s * (max_clusters + 1) + c];
  const int per = (n_angles + gridDim.x - 1) / gridDim.x;
  const int a0 = blockIdx.x * per, a1 = min(n_angles, a0 + per);
  float* out = beta32 + ((size_t)s * max_valid + rank) * n_angles;
  __shared__ float2 sp[kBetaStage];
  auto stage = [&](int c0, int m) {
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) sp[i] = __ldg(pts + c0 + i);
    __syncthreads();
  };
  const bool resident = n <= kBetaStage;                              // staged once, read by both passes
  if (resident) stage(0, n);
  for (int ab = a0; ab < a1; ab += blockDim.x) {                      // block-uniform
    const int a = min(ab + (int)threadIdx.x, a1 - 1);
    const float cs = (float)trig[a], sn = (float)trig[n_angles + a];
    float lox = 3e38f, hix = -3e38f, loy = 3e38f, hiy = -3e38f;
    for (int c0 = 0; c0 < n; c0 += kBetaStage) {
      const int m = min(kBetaStage, n - c0);
      if (!resident) stage(c0, m);
      for (int i = 0; i < m; ++i) {
        const float2 p = sp[i];
        const float px = fmaf(p.y, sn, p.x * cs), py = fmaf(p.y, cs, -p.x * sn);
        lox = fminf(lox, px); hix = fmaxf(hix, px); loy = fminf(loy, py); hiy = fmaxf(hiy, py);
      }
    }
    float beta = 0.f;
    for (int c0 = 0; c0 < n; c0 += kBetaStage) {
      const int m = min(kBetaStage, n - c0);
      if (!resident) stage(c0, m);
      for (int i = 0; i < m; ++i) {
        const float2 p = sp[i];
        const float px = fmaf(p.y, sn, p.x * cs), py = fmaf(p.y, cs, -p.x * sn);
        const float dx = fminf(px - lox, hix - px), dy = fminf(py - loy, hiy - py);
        beta += __fdividef(1.0f, fmaxf(fminf(dx, dy), d0));
      }
    }
    if (ab + (int)threadIdx.x < a1) out[ab + threadIdx.x] = beta;
  }
}

__global__ void __launch_bounds__(kFitThreads) box_fit_kernel(
    const int64_t* __restrict__ off, const double* __restrict__ rect, int max_clusters, const int32_t* __restrict__ n_clusters,
    const ClusterStat* __restrict__ stats, const int32_t* __restrict__ cl_off, const int32_t* __restrict__ members,
    const double* __restrict__ trig /* (4, n_angles): cos, sin, cos(+pi/2), sin(+pi/2) */, int n_angles,
    const double* __restrict__ angles /* (2, n_angles): angle, angle+pi/2 */, double d0, BoxRec* __restrict__ boxes,
    const int32_t* __restrict__ new_id, const int32_t* __restrict__ has_noise, const float* __restrict__ beta32,
    int max_valid) {
  const int s = blockIdx.y;
  const int C = min(n_clusters[s], max_clusters);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = kFitThreads / 32;
  __shared__ double sh_beta[kFitThreads / 32];
  __shared__ int sh_idx[kFitThreads / 32];
  __shared__ double sh_red[4][kFitThreads / 32];
  __shared__ int sh_ncand;
  __shared__ int sh_cand[kCandCap];
  for (int c = blockIdx.x; c < C; c += gridDim.x) {
    const ClusterStat* st = stats + (size_t)s * max_clusters + c;
    BoxRec* bx = boxes + (size_t)s * max_clusters + c;
    if (st->valid != 1) { if (threadIdx.x == 0) bx->keep = 0; continue; }   // invalid, or certified to fail the volume gate
    const int n = st->count;
    const int32_t* mem = members + off[s] + cl_off[(size_t)s * (max_clusters + 1) + c];
    const double* R = rect + 3 * off[s];
    // ---- heading search -----------------------------------------------------------------
    // beta64(a): the reference's score at search angle a, f64, one warp
    auto beta64 = [&](int a) {
      const double cs = trig[a], sn = trig[n_angles + a];
      double lox = 1e300, hix = -1e300, loy = 1e300, hiy = -1e300;
      for (int i = lane; i < n; i += 32) {
        const int m = mem[i];
        double px, py;
        project(R[3 * m], R[3 * m + 2], cs, sn, &px, &py);
        lox = fmin(lox, px); hix = fmax(hix, px); loy = fmin(loy, py); hiy = fmax(hiy, py);
      }
      lox = warp_min(lox); loy = warp_min(loy); hix = warp_max(hix); hiy = warp_max(hiy);
      double beta = 0.0;
      for (int i = lane; i < n; i += 32) {
        const int m = mem[i];
        double px, py;
        project(R[3 * m], R[3 * m + 2], cs, sn, &px, &py);
        const double dx = fmin(__dsub_rn(px, lox), __dsub_rn(hix, px));
        const double dy = fmin(__dsub_rn(py, loy), __dsub_rn(hiy, py));
        beta += __ddiv_rn(1.0, fmax(fmin(dx, dy), d0));
      }
      return warp_sum(beta);
    };
    double best_beta = -1e300;
    int best_idx = 0x7fffffff;
    // The float32 pre-pass (box_beta32_kernel) scored every angle; only angles whose f32 score is
    // within kBetaMargin of the f32 maximum can hold the f64 maximum (the f32 evaluation error is
    // far below the margin), so only those are re-scored in f64.  Falls back to the full f64
    // scan when the cluster has no pre-pass slot or too many angles are that close.
    const int rank = new_id[(size_t)s * max_clusters + c] - has_noise[s];
    bool full_scan = !(beta32 && rank >= 0 && rank < max_valid);
    if (!full_scan) {
      const float* b32 = beta32 + ((size_t)s * max_valid + rank) * n_angles;
      float m32 = -3.0e38f;
      for (int a = threadIdx.x; a < n_angles; a += kFitThreads) m32 = fmaxf(m32, b32[a]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m32 = fmaxf(m32, __shfl_xor_sync(0xffffffffu, m32, o));
      __syncthreads();
      if (threadIdx.x == 0) sh_ncand = 0;
      if (lane == 0) sh_red[0][w] = (double)m32;
      __syncthreads();
      float mx = (float)sh_red[0][0];
      for (int k = 1; k < nw; ++k) mx = fmaxf(mx, (float)sh_red[0][k]);
      const float thr = mx * (1.0f - kBetaMargin);
      for (int a = threadIdx.x; a < n_angles; a += kFitThreads)
        if (b32[a] >= thr) {
          const int k = atomicAdd(&sh_ncand, 1);
          if (k < kCandCap) sh_cand[k] = a;
        }
      __syncthreads();
      const int nc = sh_ncand;
      if (nc > kCandCap || nc == 0) full_scan = true;      // block-uniform
      else
        for (int ci = w; ci < nc; ci += nw) {
          const int a = sh_cand[ci];
          const double beta = beta64(a);
          if (beta > best_beta || (beta == best_beta && a < best_idx)) { best_beta = beta; best_idx = a; }
        }
    }
    if (full_scan) {
      best_beta = -1e300; best_idx = 0x7fffffff;
      for (int a = w; a < n_angles; a += nw) {
        const double beta = beta64(a);
        if (beta > best_beta) { best_beta = beta; best_idx = a; }   // ascending a: first strict max wins
      }
    }
    __syncthreads();
    if (lane == 0) { sh_beta[w] = best_beta; sh_idx[w] = best_idx; }
    __syncthreads();
    int sel = 0;
    {
      double bb = -1e300;
      int bi = 0x7fffffff;
      for (int k = 0; k < nw; ++k)
        if (sh_beta[k] > bb || (sh_beta[k] == bb && sh_idx[k] < bi)) { bb = sh_beta[k]; bi = sh_idx[k]; }
      sel = bi;
    }
    // ---- extents at the chosen heading (block-wide) ----
    int variant = 0;
    double cs = trig[sel], sn = trig[n_angles + sel];
    double ext[4];
    for (int pass = 0; pass < 2; ++pass) {
      double lox = 1e300, hix = -1e300, loy = 1e300, hiy = -1e300;
      for (int i = threadIdx.x; i < n; i += kFitThreads) {
        const int m = mem[i];
        double px, py;
        project(R[3 * m], R[3 * m + 2], cs, sn, &px, &py);
        lox = fmin(lox, px); hix = fmax(hix, px); loy = fmin(loy, py); hiy = fmax(hiy, py);
      }
      lox = warp_min(lox); loy = warp_min(loy); hix = warp_max(hix); hiy = warp_max(hiy);
      __syncthreads();
      if (lane == 0) { sh_red[0][w] = lox; sh_red[1][w] = hix; sh_red[2][w] = loy; sh_red[3][w] = hiy; }
      __syncthreads();
      lox = sh_red[0][0]; hix = sh_red[1][0]; loy = sh_red[2][0]; hiy = sh_red[3][0];
      for (int k = 1; k < nw; ++k) {
        lox = fmin(lox, sh_red[0][k]); hix = fmax(hix, sh_red[1][k]);
        loy = fmin(loy, sh_red[2][k]); hiy = fmax(hiy, sh_red[3][k]);
      }
      ext[0] = lox; ext[1] = hix; ext[2] = loy; ext[3] = hiy;
      if (pass == 0 && __dsub_rn(hix, lox) < __dsub_rn(hiy, loy)) {
        variant = 1;                                   // angle += pi/2, recompute
        cs = trig[2 * n_angles + sel]; sn = trig[3 * n_angles + sel];
      } else break;
    }
    // min rect-y of the cluster
    double ymin = 1e300;
    for (int i = threadIdx.x; i < n; i += kFitThreads) ymin = fmin(ymin, R[3 * mem[i] + 1]);
    ymin = warp_min(ymin);
    __syncthreads();
    if (lane == 0) sh_red[0][w] = ymin;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 1; k < nw; ++k) ymin = fmin(ymin, sh_red[0][k]);
      const double lox = ext[0], hix = ext[1], loy = ext[2], hiy = ext[3];
      const double area = __dmul_rn(__dsub_rn(hix, lox), __dsub_rn(hiy, loy));
      // corners = [[hix,loy],[lox,loy],[lox,hiy],[hix,hiy]] @ [[c,s],[-s,c]]
      auto corner = [&](double a, double b, double* ox, double* oy) {
        *ox = __fma_rn(b, -sn, __dmul_rn(a, cs));
        *oy = __fma_rn(b, cs, __dmul_rn(a, sn));
      };
      double c0x, c0y, c1x, c1y, c2x, c2y, c3x, c3y;
      corner(hix, loy, &c0x, &c0y); corner(lox, loy, &c1x, &c1y);
      corner(lox, hiy, &c2x, &c2y); corner(hix, hiy, &c3x, &c3y);
      const double ang = angles[variant * n_angles + sel];
      auto norm2 = [](double ax, double ay) { return sqrt(__fma_rn(ay, ay, __dmul_rn(ax, ax))); };   // sqrt(ddot)
      bx->l = norm2(__dsub_rn(c0x, c1x), __dsub_rn(c0y, c1y));
      bx->w = norm2(__dsub_rn(c0x, c3x), __dsub_rn(c0y, c3y));
      bx->t[0] = __ddiv_rn(__dadd_rn(c0x, c2x), 2.0);
      bx->t[2] = __ddiv_rn(__dadd_rn(c0y, c2y), 2.0);
      bx->t[1] = 0.0;
      bx->ry = -ang;
      bx->area = area;
      bx->ymin = ymin;
      bx->cosr = cs;          // cos(-angle) = cos(angle)
      bx->sinr = -sn;         // sin(-angle) = -sin(angle)
      bx->h = 0.0; bx->volume = 0.0;
      bx->keep = 1;
    }
    __syncthreads();
  }
}

// lowest scan point strictly inside each footprint (pointcloud_utils.py:278-290): every point
// is tested against every valid box of its scan; boxes cached in shared memory.
struct Foot { double cx, cz, c, s, hl, hw; int slot; int pad; };
constexpr int kFootCap = 128;

__global__ void __launch_bounds__(256) box_bottom_kernel(
    const int64_t* __restrict__ off, const double* __restrict__ rect, int max_clusters, const int32_t* __restrict__ n_clusters,
    const ClusterStat* __restrict__ stats, const BoxRec* __restrict__ boxes, unsigned long long* __restrict__ bottom) {
  const int s = blockIdx.y;
  const int C = min(n_clusters[s], max_clusters);
  const int64_t beg = off[s];
  const int n = (int)(off[s + 1] - beg);
  __shared__ Foot feet[kFootCap];
  __shared__ int nfeet;
  for (int c0 = 0; c0 < C; c0 += kFootCap) {        // usually a single chunk
    __syncthreads();
    if (threadIdx.x == 0) nfeet = 0;
    __syncthreads();
    for (int c = c0 + threadIdx.x; c < min(C, c0 + kFootCap); c += blockDim.x) {
      if (stats[(size_t)s * max_clusters + c].valid == 1) {
        const BoxRec& b = boxes[(size_t)s * max_clusters + c];
        const int k = atomicAdd(&nfeet, 1);
        feet[k] = Foot{b.t[0], b.t[2], b.cosr, b.sinr, __ddiv_rn(b.l, 2.0), __ddiv_rn(b.w, 2.0), c, 0};
      }
    }
    __syncthreads();
    const int nf = nfeet;
    if (nf == 0) continue;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const double x = rect[3 * (beg + i)], y = rect[3 * (beg + i) + 1], z = rect[3 * (beg + i) + 2];
      for (int k = 0; k < nf; ++k) {
        const Foot f = feet[k];
        const double dx = __dsub_rn(x, f.cx), dz = __dsub_rn(z, f.cz);
        // (ptc_xz - c) @ [[cos, -sin],[sin, cos]].T
        const double lx = __fma_rn(dz, -f.s, __dmul_rn(dx, f.c));
        const double ly = __fma_rn(dz, f.c, __dmul_rn(dx, f.s));
        if (lx > -f.hl && lx < f.hl && ly > -f.hw && ly < f.hw)
          atomicMax(&bottom[(size_t)s * max_clusters + f.slot], f64_ordered(y));
      }
    }
  }
}

struct VolumeCfg { double min_volume, max_volume; };

// finish boxes (h, t.y, volume, keep), renumber surviving clusters, emit compact box rows.
// One CTA per scan: the keep decisions are independent per cluster, the box order (cluster id
// order) comes from a block-wide ordered scan of the keep flags.
__global__ void __launch_bounds__(256) box_finalize_kernel(
    const int64_t* __restrict__ off, int max_clusters, const int32_t* __restrict__ n_clusters,
    const ClusterStat* __restrict__ stats, BoxRec* __restrict__ boxes, const unsigned long long* __restrict__ bottom, VolumeCfg cfg,
    const int32_t* __restrict__ new_id, const int32_t* __restrict__ has_noise, int32_t* __restrict__ final_id /* (S,max_clusters+1) by filtered id */,
    double* __restrict__ out_boxes /* (S,max_boxes,8) */, int max_boxes, int32_t* __restrict__ n_boxes, int32_t* __restrict__ flags, int n_scans) {
  const int s = blockIdx.x;
  if (s >= n_scans) return;
  const int C = min(n_clusters[s], max_clusters);
  int32_t* fid = final_id + (size_t)s * (max_clusters + 1);
  const int n = (int)(off[s + 1] - off[s]);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __shared__ long long sh_cov[8];
  __shared__ int sh_cnt[8];
  __shared__ int sh_base;
  // pass 1: decide keep
  long long covered = 0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const ClusterStat st = stats[(size_t)s * max_clusters + c];
    if (!st.valid) continue;
    BoxRec& b = boxes[(size_t)s * max_clusters + c];
    if (st.valid == 2) { b.keep = 0; continue; }                                 // volume certified > max_volume
    if (new_id[(size_t)s * max_clusters + c] <= 0) { b.keep = 0; continue; }     // id 0 is background for generate_mask.py:93
    const unsigned long long bo = bottom[(size_t)s * max_clusters + c];
    if (bo == 0ull) { b.keep = 0; atomicOr(flags, 8); continue; }                // empty footprint (reference would raise)
    const double bot = f64_from_ordered(bo);
    b.t[1] = bot;
    b.h = __dsub_rn(bot, b.ymin);
    b.volume = __dmul_rn(b.area, b.h);
    b.keep = (b.volume > cfg.min_volume && b.volume < cfg.max_volume) ? 1 : 0;
    if (b.keep) covered += st.count;
  }
  covered = warp_sum(covered);
  if (lane == 0) sh_cov[w] = covered;
  if (threadIdx.x == 0) { fid[0] = 0; sh_base = 0; }
  __syncthreads();                                                               // also publishes b.keep to the block
  covered = 0;
  for (int k = 0; k < 8; ++k) covered += sh_cov[k];
  // sorted(set(labels_filtered)) after rejected clusters were set to 0 (generate_mask.py:100-103)
  const int zero_present = (has_noise[s] || covered < n) ? 1 : 0;
  // pass 2: kept boxes in cluster order
  for (int c0 = 0; c0 < C; c0 += blockDim.x) {
    const int c = c0 + threadIdx.x;
    bool valid = false, keep = false;
    int filt = -1;
    if (c < C && stats[(size_t)s * max_clusters + c].valid) {
      valid = true;
      filt = new_id[(size_t)s * max_clusters + c];                               // 1..K (or 0..K-1 without noise)
      keep = boxes[(size_t)s * max_clusters + c].keep != 0;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) sh_cnt[w] = __popc(bal);
    __syncthreads();
    int k = sh_base + __popc(bal & ((1u << lane) - 1u));
    for (int j = 0; j < w; ++j) k += sh_cnt[j];
    if (valid) {
      if (keep) {
        if (k < max_boxes) {
          const BoxRec& b = boxes[(size_t)s * max_clusters + c];
          double* o = out_boxes + ((size_t)s * max_boxes + k) * 8;
          o[0] = b.t[0]; o[1] = b.t[1]; o[2] = b.t[2]; o[3] = b.l; o[4] = b.w; o[5] = b.h; o[6] = b.ry; o[7] = b.volume;
        } else atomicOr(flags, 16);
        if (filt >= 0 && filt <= max_clusters) fid[filt] = k + zero_present;
      } else if (filt >= 0 && filt <= max_clusters) fid[filt] = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int j = 0; j < 8; ++j) t += sh_cnt[j]; sh_base += t; }
    __syncthreads();
  }
  if (threadIdx.x == 0) n_boxes[s] = min(sh_base, max_boxes);
}

__global__ void __launch_bounds__(256) apply_final_kernel(const int64_t* __restrict__ off, const int32_t* __restrict__ labels_filtered,
                                                          int max_clusters, const int32_t* __restrict__ final_id,
                                                          int32_t* __restrict__ labels_out) {
  const int s = blockIdx.y;
  const int64_t beg = off[s];
  const int n = (int)(off[s + 1] - beg);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int f = labels_filtered[beg + i];
    labels_out[beg + i] = (f <= 0 || f > max_clusters) ? 0 : final_id[(size_t)s * (max_clusters + 1) + f];
  }
}


// ---- f-1: PP-score percentile of the scan points inside each (detector) box ---------------------
// filter_by_ppscore() of combine_labels.py:42-60: footprint test in the box frame (same dgemm
// rounding as get_lowest_point_rect) AND t.y - h < y <= t.y, then np.percentile(pp[inside], q)
// in numpy 2.x float32 arithmetic.  One CTA per box; pass 1 collects the member pp values into
// a per-box slice of `vals` (capacity = points of the scan), pass 2 selects.
__global__ void __launch_bounds__(256) box_pp_percentile_kernel(
    const int64_t* __restrict__ off, const double* __restrict__ rect, const float* __restrict__ pp,
    const double* __restrict__ boxes /* (S,max_boxes,8) t.x t.y t.z l w h ry _ */, const int32_t* __restrict__ n_boxes,
    const double* __restrict__ box_trig /* (S,max_boxes,2) cos(ry), sin(ry) from the host libm, or NULL */,
    int max_boxes, float q, float* __restrict__ vals /* (S,max_boxes) slices of max_points floats */,
    int64_t max_points, float* __restrict__ pct_out, int32_t* __restrict__ cnt_out) {
  const int s = blockIdx.y, k = blockIdx.x;
  if (k >= n_boxes[s]) return;
  const int64_t beg = off[s];
  const int n = (int)(off[s + 1] - beg);
  const double* b = boxes + ((size_t)s * max_boxes + k) * 8;
  const double tx = b[0], ty = b[1], tz = b[2], l = b[3], w = b[4], h = b[5], ry = b[6];
  // numpy's cos/sin of ry when the caller supplies them (bit-faithful footprints), else libdevice's
  const double c = box_trig ? box_trig[((size_t)s * max_boxes + k) * 2] : cos(ry);
  const double sn = box_trig ? box_trig[((size_t)s * max_boxes + k) * 2 + 1] : sin(ry);
  const double hl = __ddiv_rn(l, 2.0), hw = __ddiv_rn(w, 2.0), ylo = __dsub_rn(ty, h);
  float* mine = vals + ((size_t)s * max_boxes + k) * (size_t)max_points;
  __shared__ int s_cnt;
  __shared__ unsigned hist[256];
  __shared__ unsigned sel[2];
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double x = rect[3 * (beg + i)], y = rect[3 * (beg + i) + 1], z = rect[3 * (beg + i) + 2];
    const double dx = __dsub_rn(x, tx), dz = __dsub_rn(z, tz);
    const double lx = __fma_rn(dz, -sn, __dmul_rn(dx, c));
    const double ly = __fma_rn(dz, c, __dmul_rn(dx, sn));
    if (lx > -hl && lx < hl && ly > -hw && ly < hw && y > ylo && y <= ty) mine[atomicAdd(&s_cnt, 1)] = pp[beg + i];
  }
  __syncthreads();
  const int m = s_cnt;
  float r = 0.f;
  if (m > 0) {
    const float vi = __fmul_rn((float)(m - 1), q);
    const int lo = (int)floorf(vi);
    const int hi = min(lo + 1, m - 1);
    const float g = __fsub_rn(vi, (float)lo);
    const float a = block256_kth_smallest(m, lo, [&](int i) { return mine[i]; }, hist, sel);
    const float bb = hi == lo ? a : block256_kth_smallest(m, hi, [&](int i) { return mine[i]; }, hist, sel);
    const float d = __fsub_rn(bb, a);
    r = g >= 0.5f ? __fsub_rn(bb, __fmul_rn(d, __fsub_rn(1.0f, g))) : __fadd_rn(a, __fmul_rn(d, g));
  }
  if (threadIdx.x == 0) { pct_out[(size_t)s * max_boxes + k] = r; cnt_out[(size_t)s * max_boxes + k] = m; }
}

}  // namespace modest

using namespace modest;

static const int kMaxAngles = 1024;
static int prepass_slots(int max_clusters) { return max_clusters < 384 ? max_clusters : 384; }

extern "C" size_t modest_filter_workspace_bytes(int n_scans, int64_t n_points_total, int max_clusters) {
  size_t b = 0;
  auto add = [&](size_t bytes) { b = align_up(b, 256) + bytes; };
  add(sizeof(ClusterStat) * (size_t)n_scans * max_clusters);
  add(sizeof(int32_t) * (size_t)n_scans * (max_clusters + 1));   // cl_off
  add(sizeof(int32_t) * (size_t)n_scans * max_clusters);         // cl_fill
  add(sizeof(int32_t) * (size_t)n_points_total);                 // members
  add(sizeof(int32_t) * (size_t)n_scans * max_clusters);         // new_id
  add(sizeof(int32_t) * (size_t)n_scans * 2);                    // n_valid, has_noise
  add(sizeof(double) * 3 * (size_t)n_points_total);              // rect
  add(sizeof(BoxRec) * (size_t)n_scans * max_clusters);
  add(sizeof(unsigned long long) * (size_t)n_scans * max_clusters);   // bottom
  add(sizeof(int32_t) * (size_t)n_scans * (max_clusters + 1));   // final_id
  add(sizeof(CalibDev) * (size_t)n_scans);
  add(sizeof(float) * (size_t)n_scans * prepass_slots(max_clusters) * kMaxAngles);   // beta32
  add(sizeof(int32_t) * (size_t)n_scans * prepass_slots(max_clusters));              // valid_list
  add(sizeof(float2) * (size_t)n_points_total);                                      // centred f32 member coords
  return b + 256;
}

// h_gates: {min_points, max_min_height, min_max_height, q(=percentile/100 in f32), min_percentile_pp_score,
//           min_volume, max_volume, d0}
extern "C" int modest_filter_and_fit_batch(
    const float* d_ptc, int point_stride, const int64_t* d_off, const float* d_pp, const int32_t* d_labels,
    const int32_t* d_n_clusters, const double* d_planes, const double* d_calib /* (S,21): V2C(12) R0(9) */,
    const double* d_rect_in /* (NP,3) rect coords or NULL */,
    int n_scans, int64_t n_points_total, int64_t max_points, int max_clusters, int max_boxes, const double* h_gates,
    const double* d_trig, const double* d_angles, int n_angles, int32_t* d_labels_filtered, int32_t* d_labels_final,
    double* d_boxes, int32_t* d_n_boxes, int32_t* d_n_valid, int32_t* d_flags, void* d_ws, size_t ws_bytes, void* stream_) {
  modest::StageRange nvtx_("modest:J-M cluster gates + box fit");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_scans <= 0) return MODEST_OK;
  MODEST_REQUIRE(d_ptc && d_off && d_pp && d_labels && d_n_clusters && d_planes && (d_calib || d_rect_in) && h_gates && d_trig &&
                     d_angles && d_labels_filtered && d_labels_final && d_boxes && d_n_boxes && d_n_valid && d_flags && d_ws,
                 "filter_and_fit: null pointer argument");
  MODEST_REQUIRE(max_clusters >= 1 && max_boxes >= 1 && n_angles >= 1, "filter_and_fit: bad capacity arguments");
  MODEST_REQUIRE(ws_bytes >= modest_filter_workspace_bytes(n_scans, n_points_total, max_clusters),
                 "filter_and_fit: workspace too small");
  MODEST_REQUIRE(n_scans <= 65535, "filter_and_fit: more than 65535 scans in one launch");
  Arena ar(d_ws, ws_bytes);
  ClusterStat* stats = ar.take<ClusterStat>((size_t)n_scans * max_clusters);
  int32_t* cl_off = ar.take<int32_t>((size_t)n_scans * (max_clusters + 1));
  int32_t* cl_fill = ar.take<int32_t>((size_t)n_scans * max_clusters);
  int32_t* members = ar.take<int32_t>(n_points_total);
  int32_t* new_id = ar.take<int32_t>((size_t)n_scans * max_clusters);
  int32_t* nv_hn = ar.take<int32_t>((size_t)n_scans * 2);
  double* rect_ws = ar.take<double>(3 * (size_t)n_points_total);
  const double* rect = d_rect_in ? d_rect_in : rect_ws;
  BoxRec* boxes = ar.take<BoxRec>((size_t)n_scans * max_clusters);
  unsigned long long* bottom = ar.take<unsigned long long>((size_t)n_scans * max_clusters);
  int32_t* final_id = ar.take<int32_t>((size_t)n_scans * (max_clusters + 1));
  CalibDev* calibs = ar.take<CalibDev>(n_scans);
  const int max_valid = prepass_slots(max_clusters);
  float* beta32 = ar.take<float>((size_t)n_scans * max_valid * kMaxAngles);
  int32_t* valid_list = ar.take<int32_t>((size_t)n_scans * max_valid);
  float2* xz32 = ar.take<float2>(n_points_total);
  MODEST_REQUIRE(ar.ok(), "workspace too small for the requested sizes");
  int32_t* has_noise = nv_hn + n_scans;

  FilterCfg fc;
  fc.min_points = (int)h_gates[0];
  fc.max_min_height = h_gates[1];
  fc.min_max_height = h_gates[2];
  fc.q = (float)h_gates[3];
  fc.min_percentile_pp = (float)h_gates[4];
  fc.min_percentile_pp_d = h_gates[4];
  VolumeCfg vc{h_gates[5], h_gates[6]};
  const double d0 = h_gates[7];

  if (!d_rect_in)
    MODEST_CUDA(cudaMemcpyAsync(calibs, d_calib, sizeof(CalibDev) * (size_t)n_scans, cudaMemcpyDeviceToDevice, stream));
  MODEST_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int32_t), stream));
  MODEST_CUDA(cudaMemsetAsync(bottom, 0, sizeof(unsigned long long) * (size_t)n_scans * max_clusters, stream));
  int pblocks = (int)((max_points + 255) / 256);
  if (pblocks < 1) pblocks = 1;
  if (pblocks > 1024) pblocks = 1024;
  const dim3 pgrid(pblocks, n_scans);
  cluster_stats_init_kernel<<<256, 256, 0, stream>>>(stats, (size_t)n_scans * max_clusters);
  MODEST_LAUNCH_CHECK("cluster_stats_init_kernel");
  cluster_stats_kernel<<<pgrid, 256, 0, stream>>>(d_ptc, point_stride, d_off, d_labels, d_planes, max_clusters, stats, d_flags);
  MODEST_LAUNCH_CHECK("cluster_stats_kernel");
  cluster_offsets_kernel<<<n_scans, 1024, 0, stream>>>(stats, max_clusters, d_n_clusters, cl_off, cl_fill);
  MODEST_LAUNCH_CHECK("cluster_offsets_kernel");
  cluster_scatter_kernel<<<pgrid, 256, 0, stream>>>(d_off, d_labels, max_clusters, cl_off, cl_fill, members);
  MODEST_LAUNCH_CHECK("cluster_scatter_kernel");
  const int cblocks = max_clusters < 512 ? max_clusters : 512;
  cluster_validate_kernel<<<dim3(cblocks, n_scans), 256, 0, stream>>>(d_off, d_pp, max_clusters, d_n_clusters, cl_off, members, fc, stats);
  MODEST_LAUNCH_CHECK("cluster_validate_kernel");
  cluster_relabel_kernel<<<(n_scans + 63) / 64, 64, 0, stream>>>(d_off, max_clusters, d_n_clusters, stats, new_id, d_n_valid, has_noise, n_scans,
                                                                valid_list, max_valid);
  MODEST_LAUNCH_CHECK("cluster_relabel_kernel");
  apply_relabel_kernel<<<pgrid, 256, 0, stream>>>(d_off, d_labels, max_clusters, new_id, d_labels_filtered);
  MODEST_LAUNCH_CHECK("apply_relabel_kernel");
  if (!d_rect_in) {
    rect_coords_kernel<<<pgrid, 256, 0, stream>>>(d_ptc, point_stride, d_off, calibs, rect_ws);
    MODEST_LAUNCH_CHECK("rect_coords_kernel");
  }
  MODEST_REQUIRE(n_angles <= kMaxAngles, "filter_and_fit: more than %d search angles", kMaxAngles);
  // beta32 rows are n_angles wide (n_angles <= kMaxAngles, the stride the workspace was sized for)
  box_prereject_kernel<<<dim3(max_valid, n_scans), 256, 0, stream>>>(d_off, rect, max_clusters, stats, cl_off, members, d_n_valid,
                                                                    valid_list, max_valid, vc.max_volume);
  MODEST_LAUNCH_CHECK("box_prereject_kernel");
  box_center32_kernel<<<dim3(4, max_valid, n_scans), 256, 0, stream>>>(d_off, rect, max_clusters, stats, cl_off, members, d_n_valid,
                                                                      valid_list, max_valid, xz32);
  MODEST_LAUNCH_CHECK("box_center32_kernel");
  box_beta32_kernel<<<dim3(kAngleChunks, max_valid, n_scans), kBetaThreads, 0, stream>>>(d_off, max_clusters, stats, cl_off, d_n_valid, valid_list,
                                                                             max_valid, xz32, d_trig, n_angles, (float)d0, beta32);
  MODEST_LAUNCH_CHECK("box_beta32_kernel");
  box_fit_kernel<<<dim3(cblocks, n_scans), kFitThreads, 0, stream>>>(d_off, rect, max_clusters, d_n_clusters, stats, cl_off, members,
                                                                   d_trig, n_angles, d_angles, d0, boxes, new_id, has_noise, beta32,
                                                                   max_valid);
  MODEST_LAUNCH_CHECK("box_fit_kernel");
  box_bottom_kernel<<<pgrid, 256, 0, stream>>>(d_off, rect, max_clusters, d_n_clusters, stats, boxes, bottom);
  MODEST_LAUNCH_CHECK("box_bottom_kernel");
  box_finalize_kernel<<<n_scans, 256, 0, stream>>>(d_off, max_clusters, d_n_clusters, stats, boxes, bottom, vc, new_id, has_noise,
                                                             final_id, d_boxes, max_boxes, d_n_boxes, d_flags, n_scans);
  MODEST_LAUNCH_CHECK("box_finalize_kernel");
  apply_final_kernel<<<pgrid, 256, 0, stream>>>(d_off, d_labels_filtered, max_clusters, final_id, d_labels_final);
  MODEST_LAUNCH_CHECK("apply_final_kernel");
  note_launch(15);
  return MODEST_OK;
}

extern "C" size_t modest_box_pp_workspace_bytes(int n_scans, int64_t n_points_total, int64_t max_points, int max_boxes) {
  return align_up(sizeof(double) * 3 * (size_t)n_points_total, 256) + align_up(sizeof(CalibDev) * (size_t)n_scans, 256) +
         align_up(sizeof(float) * (size_t)n_scans * max_boxes * (size_t)max_points, 256) + 1024;
}

extern "C" int modest_box_pp_percentile_batch(const float* d_ptc, int point_stride, const int64_t* d_off, const float* d_pp,
                                              const double* d_calib, const double* d_rect_in, const double* d_boxes,
                                              const double* d_box_trig, const int32_t* d_n_boxes, int n_scans, int64_t n_points_total, int64_t max_points,
                                              int max_boxes, double q_f32, float* d_percentile, int32_t* d_count, void* d_ws,
                                              size_t ws_bytes, void* stream_) {
  modest::StageRange nvtx_("modest:f-1 in-box PP percentile");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_scans <= 0) return MODEST_OK;
  MODEST_REQUIRE(d_ptc && d_off && d_pp && (d_calib || d_rect_in) && d_boxes && d_n_boxes && d_percentile && d_count && d_ws,
                 "box_pp_percentile: null pointer argument");
  MODEST_REQUIRE(max_boxes >= 1 && n_scans <= 65535, "box_pp_percentile: bad sizes");
  MODEST_REQUIRE(ws_bytes >= modest_box_pp_workspace_bytes(n_scans, n_points_total, max_points, max_boxes),
                 "box_pp_percentile: workspace too small");
  Arena ar(d_ws, ws_bytes);
  double* rect_ws = ar.take<double>(3 * (size_t)n_points_total);
  CalibDev* calibs = ar.take<CalibDev>(n_scans);
  float* vals = ar.take<float>((size_t)n_scans * max_boxes * (size_t)max_points);
  MODEST_REQUIRE(ar.ok(), "workspace too small for the requested sizes");
  const double* rect = d_rect_in ? d_rect_in : rect_ws;
  int pblocks = (int)((max_points + 255) / 256);
  if (pblocks < 1) pblocks = 1;
  if (!d_rect_in) {
    MODEST_CUDA(cudaMemcpyAsync(calibs, d_calib, sizeof(CalibDev) * (size_t)n_scans, cudaMemcpyDeviceToDevice, stream));
    rect_coords_kernel<<<dim3(pblocks, n_scans), 256, 0, stream>>>(d_ptc, point_stride, d_off, calibs, rect_ws);
    MODEST_LAUNCH_CHECK("rect_coords_kernel");
  }
  box_pp_percentile_kernel<<<dim3(max_boxes, n_scans), 256, 0, stream>>>(d_off, rect, d_pp, d_boxes, d_n_boxes, d_box_trig, max_boxes,
                                                                        (float)q_f32, vals, max_points, d_percentile, d_count);
  MODEST_LAUNCH_CHECK("box_pp_percentile_kernel");
  note_launch(d_rect_in ? 1 : 2);
  return MODEST_OK;
}
