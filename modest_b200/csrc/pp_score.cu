// Stage C+D: persistence-point score (neighbour counts over T traversals + entropy).
//
// Reference behaviour: pre_compute_pp_score.py:54-75 (count_neighbors, compute_ephe_score) on
// trees built at :188-190.  Design (DESIGN.md "PP score"): the 60k-point query scan is binned
// into a per-scan 2-D grid over (x,y) with cell edge just above the search radius (counting
// sort: histogram -> scan -> scatter); the history -- 16x larger -- is then streamed through
// exactly once, each history point probing the 3x3 neighbouring columns (three contiguous row
// segments of the sorted query) and adding 1 to count[q][t] for every query point within the
// radius.  Distances are decided in f32 when they are clear of the sphere surface and
// re-evaluated in sequential f64 (the arithmetic of cKDTree) inside a thin shell around it, so
// counts are bit-exact.  A last kernel turns the (N,T) counts into the normalised entropy.
#include "common.cuh"

namespace modest {
extern void note_launch(int n);

struct PPScanMeta {      // per scan, device resident
  float x0, y0;          // grid origin
  float inv_cell;        // 1 / cell edge
  int   pad;
};

__device__ __forceinline__ int cell_coord(float v, float origin, float inv_cell) {
  // monotone in v (one rounded subtract, one rounded multiply, floor) -- that is all the
  // neighbour search needs; see DESIGN.md for the |cell(q) - cell(h)| <= 1 argument.
  return __float2int_rd(__fmul_rn(__fsub_rn(v, origin), inv_cell));
}
// ints per scan in the cell table: G*G cells + sentinel, padded so every scan stays 16-B aligned
__host__ __device__ __forceinline__ size_t cell_stride(int G) { return (size_t)G * G + 4; }
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ---- 1. per-scan query bounding box -> grid origin -------------------------------------------
__global__ void __launch_bounds__(1024) pp_query_origin_kernel(
    const float* __restrict__ q_xyz, const int64_t* __restrict__ q_off, PPScanMeta* __restrict__ meta,
    int G, float cell) {
  const int s = blockIdx.x;
  const int64_t beg = q_off[s], end = q_off[s + 1];
  float lox = 3.0e38f, loy = 3.0e38f, hix = -3.0e38f, hiy = -3.0e38f;
  for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
    float x = q_xyz[3 * i], y = q_xyz[3 * i + 1];
    lox = fminf(lox, x); hix = fmaxf(hix, x);
    loy = fminf(loy, y); hiy = fmaxf(hiy, y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lox = fminf(lox, __shfl_xor_sync(0xffffffffu, lox, o));
    loy = fminf(loy, __shfl_xor_sync(0xffffffffu, loy, o));
    hix = fmaxf(hix, __shfl_xor_sync(0xffffffffu, hix, o));
    hiy = fmaxf(hiy, __shfl_xor_sync(0xffffffffu, hiy, o));
  }
  __shared__ float sh[4][32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sh[0][w] = lox; sh[1][w] = loy; sh[2][w] = hix; sh[3][w] = hiy; }
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    lox = l < nw ? sh[0][l] : 3.0e38f;  loy = l < nw ? sh[1][l] : 3.0e38f;
    hix = l < nw ? sh[2][l] : -3.0e38f; hiy = l < nw ? sh[3][l] : -3.0e38f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lox = fminf(lox, __shfl_xor_sync(0xffffffffu, lox, o));
      loy = fminf(loy, __shfl_xor_sync(0xffffffffu, loy, o));
      hix = fmaxf(hix, __shfl_xor_sync(0xffffffffu, hix, o));
      hiy = fmaxf(hiy, __shfl_xor_sync(0xffffffffu, hiy, o));
    }
    if (l == 0) {
      PPScanMeta m;
      if (end <= beg) { lox = loy = hix = hiy = 0.f; }
      const float half = 0.5f * cell * (float)G;
      m.x0 = 0.5f * (lox + hix) - half;
      m.y0 = 0.5f * (loy + hiy) - half;
      m.inv_cell = 1.0f / cell;
      m.pad = 0;
      meta[s] = m;
    }
  }
}

// ---- 2. histogram of query points per cell ---------------------------------------------------
__global__ void __launch_bounds__(256) pp_query_hist_kernel(
    const float* __restrict__ q_xyz, const int64_t* __restrict__ q_off,
    const PPScanMeta* __restrict__ meta, int* __restrict__ cells, int G) {
  const int s = blockIdx.y;
  const int64_t beg = q_off[s], n = q_off[s + 1] - beg;
  const PPScanMeta m = meta[s];
  int* c = cells + (size_t)s * cell_stride(G);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float* p = q_xyz + 3 * (beg + i);
    int cx = clampi(cell_coord(p[0], m.x0, m.inv_cell), 0, G - 1);
    int cy = clampi(cell_coord(p[1], m.y0, m.inv_cell), 0, G - 1);
    atomicAdd(&c[cy * G + cx], 1);
  }
}

// ---- 3. in-place inclusive scan of the G*G cell counts, one CTA per scan ---------------------
__global__ void __launch_bounds__(1024) pp_cell_scan_kernel(int* __restrict__ cells, int G) {
  const size_t ncell = (size_t)G * G;
  int* c = cells + (size_t)blockIdx.x * (ncell + 1);
  __shared__ int warp_excl[32];
  __shared__ int tile_total;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int carry = 0;                       // running prefix, identical in every thread
  for (size_t base = 0; base < ncell; base += 4096) {
    const size_t i = base + 4 * (size_t)threadIdx.x;
    int4 v = make_int4(0, 0, 0, 0);
    if (i + 3 < ncell) v = *reinterpret_cast<const int4*>(c + i);
    else {
      if (i < ncell) v.x = c[i];
      if (i + 1 < ncell) v.y = c[i + 1];
      if (i + 2 < ncell) v.z = c[i + 2];
    }
    v.y += v.x; v.z += v.y; v.w += v.z;
    int incl = v.w;                    // inclusive scan of per-thread totals within the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_excl[w] = incl;
    __syncthreads();
    if (w == 0) {
      const int t = warp_excl[lane];
      int ti = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, ti, o);
        if (lane >= o) ti += u;
      }
      warp_excl[lane] = ti - t;
      if (lane == 31) tile_total = ti;
    }
    __syncthreads();
    const int off = carry + warp_excl[w] + (incl - v.w);
    v.x += off; v.y += off; v.z += off; v.w += off;
    if (i + 3 < ncell) *reinterpret_cast<int4*>(c + i) = v;
    else {
      if (i < ncell) c[i] = v.x;
      if (i + 1 < ncell) c[i + 1] = v.y;
      if (i + 2 < ncell) c[i + 2] = v.z;
    }
    carry += tile_total;
    __syncthreads();                   // warp_excl / tile_total are rewritten next tile
  }
  if (threadIdx.x == 0) c[ncell] = carry;   // sentinel: total number of query points
}

// ---- 4. scatter query points into cell order (x,y,z,original index) --------------------------
__global__ void __launch_bounds__(256) pp_query_scatter_kernel(
    const float* __restrict__ q_xyz, const int64_t* __restrict__ q_off,
    const PPScanMeta* __restrict__ meta, int* __restrict__ cells, float4* __restrict__ sorted, int G) {
  const int s = blockIdx.y;
  const int64_t beg = q_off[s], n = q_off[s + 1] - beg;
  const PPScanMeta m = meta[s];
  int* c = cells + (size_t)s * cell_stride(G);
  float4* out = sorted + beg;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float* p = q_xyz + 3 * (beg + i);
    const float x = p[0], y = p[1], z = p[2];
    int cx = clampi(cell_coord(x, m.x0, m.inv_cell), 0, G - 1);
    int cy = clampi(cell_coord(y, m.y0, m.inv_cell), 0, G - 1);
    // the scanned array holds the END of each cell; counting down leaves the START behind
    int pos = atomicSub(&c[cy * G + cx], 1) - 1;
    out[pos] = make_float4(x, y, z, __int_as_float((int)i));
  }
}

// ---- 5. stream the history once: probe 3 row segments, count hits ----------------------------
// blockIdx.y = global traversal index g; the scan it belongs to comes from trav_scan[g].
__global__ void __launch_bounds__(256) pp_count_kernel(
    const float* __restrict__ h_xyz, const int64_t* __restrict__ h_off,
    const int32_t* __restrict__ trav_scan, const int32_t* __restrict__ trav_off,
    const int64_t* __restrict__ q_off, const int64_t* __restrict__ count_off,
    const PPScanMeta* __restrict__ meta, const int* __restrict__ cells,
    const float4* __restrict__ sorted, int* __restrict__ counts, int G, float r2f, float band,
    double r2) {
  const int g = blockIdx.y;
  const int s = trav_scan[g];
  const int t = g - trav_off[s];
  const int T = trav_off[s + 1] - trav_off[s];
  const int64_t hbeg = h_off[g], hn = h_off[g + 1] - hbeg;
  const PPScanMeta m = meta[s];
  const int* __restrict__ c = cells + (size_t)s * cell_stride(G);
  const float4* __restrict__ qs = sorted + q_off[s];
  int* __restrict__ cnt = counts + count_off[s] + t;
  const float lo = r2f - band, hi = r2f + band;

  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hn; i += (int64_t)gridDim.x * blockDim.x) {
    const float* p = h_xyz + 3 * (hbeg + i);
    const float hx = __ldg(p), hy = __ldg(p + 1), hz = __ldg(p + 2);
    const int cx = cell_coord(hx, m.x0, m.inv_cell);
    const int cy = cell_coord(hy, m.y0, m.inv_cell);
    const int xa = clampi(cx - 1, 0, G - 1), xb = clampi(cx + 1, 0, G - 1);
    const int ya = clampi(cy - 1, 0, G - 1), yb = clampi(cy + 1, 0, G - 1);
    for (int y = ya; y <= yb; ++y) {
      const int kb = __ldg(c + y * G + xa);
      const int ke = __ldg(c + y * G + xb + 1);
      for (int k = kb; k < ke; ++k) {
        const float4 q = __ldg(qs + k);
        const float dx = q.x - hx, dy = q.y - hy, dz = q.z - hz;
        const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        bool hit = d2 < lo;
        if (!hit && d2 <= hi) hit = sqdist_f64_seq(q.x, q.y, q.z, hx, hy, hz) <= r2;
        if (hit) atomicAdd(cnt + (size_t)__float_as_int(q.w) * T, 1);
      }
    }
  }
}

// ---- 6. entropy over traversals ---------------------------------------------------------------
// numpy reduces the contiguous T axis with its pairwise-sum kernel: for T >= 8 eight running
// partial sums combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail sequentially
// (blocks above 128 elements are split recursively; T is small here so that never happens).
template <typename F>
__device__ __forceinline__ double numpy_pairwise_sum(int n, F term) {
  if (n < 8) {
    double r = 0.0;   // numpy starts from -0.0 ... identical for our non-negative-or-mixed terms
    for (int i = 0; i < n; ++i) r = __dadd_rn(r, term(i));
    return r;
  }
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = term(j);
  int i = 8;
  for (; i + 8 <= n; i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], term(i + j));
  }
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __dadd_rn(res, term(i));
  return res;
}

__global__ void __launch_bounds__(256) pp_entropy_kernel(
    const int* __restrict__ counts, const int64_t* __restrict__ q_off,
    const int64_t* __restrict__ count_off, const int32_t* __restrict__ trav_off,
    float* __restrict__ pp) {
  const int s = blockIdx.y;
  const int T = trav_off[s + 1] - trav_off[s];
  const int64_t qbeg = q_off[s], n = q_off[s + 1] - qbeg;
  const int* __restrict__ c = counts + count_off[s];
  const double logT = log((double)T);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int* row = c + i * T;
    long long tot = 0;
    for (int t = 0; t < T; ++t) tot += row[t];
    const double denom = __dadd_rn((double)tot, 1e-8);
    const double acc = numpy_pairwise_sum(T, [&](int t) {
      const double P = __ddiv_rn((double)row[t], denom);
      return __dmul_rn(-P, log(__dadd_rn(P, 1e-8)));
    });
    pp[qbeg + i] = (float)__ddiv_rn(acc, logT);
  }
}

__global__ void pp_trav_scan_kernel(const int32_t* __restrict__ trav_off, int n_scans, int32_t* __restrict__ trav_scan) {
  const int s = blockIdx.x;
  if (s >= n_scans) return;
  for (int g = trav_off[s] + threadIdx.x; g < trav_off[s + 1]; g += blockDim.x) trav_scan[g] = s;
}

}  // namespace modest

using namespace modest;

static const float kCellSlack = 1.001f;   // cell edge = radius * slack, see cell_coord()

extern "C" size_t modest_pp_workspace_bytes(int n_scans, int64_t n_query_total, int64_t n_count_total,
                                            int grid_dim) {
  if (grid_dim <= 0) grid_dim = 512;
  size_t b = 0;
  auto add = [&](size_t bytes) { b = align_up(b, 256) + bytes; };
  add(sizeof(PPScanMeta) * (size_t)n_scans);
  add(sizeof(int) * (size_t)n_scans * cell_stride(grid_dim));
  add(sizeof(float4) * (size_t)n_query_total);
  add(sizeof(int) * (size_t)n_count_total);
  add(sizeof(int32_t) * (size_t)(n_count_total > 0 ? 1 << 20 : 1 << 20));  // trav_scan (<= 1M traversals)
  return b + 256;
}

extern "C" int modest_pp_score_batch(const float* d_query_xyz, const int64_t* d_q_off,
                                     const float* d_hist_xyz, const int64_t* d_h_off,
                                     const int32_t* d_trav_off, int n_scans, int n_trav_total,
                                     int64_t n_query_total, int64_t n_count_total,
                                     int64_t max_query_points, int64_t max_trav_points, double radius,
                                     int grid_dim, int32_t* d_counts, const int64_t* d_count_off,
                                     float* d_pp, void* d_ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (grid_dim <= 0) grid_dim = 512;
  MODEST_REQUIRE(n_scans >= 0 && n_trav_total >= 0, "pp_score: negative sizes");
  if (n_scans == 0 || n_query_total == 0) return MODEST_OK;
  MODEST_REQUIRE(d_query_xyz && d_q_off && d_h_off && d_trav_off && d_count_off && d_pp && d_ws,
                 "pp_score: null pointer argument");
  MODEST_REQUIRE(n_trav_total <= (1 << 20), "pp_score: more than 2^20 traversals in one batch");
  MODEST_REQUIRE(radius > 0.0 && radius < 1e3, "pp_score: radius %g out of range", radius);
  MODEST_REQUIRE(grid_dim >= 8 && grid_dim <= 4096 && grid_dim % 4 == 0,
                 "pp_score: grid_dim %d must be a multiple of 4 in [8,4096]", grid_dim);
  MODEST_REQUIRE(ws_bytes >= modest_pp_workspace_bytes(n_scans, n_query_total, n_count_total, grid_dim),
                 "pp_score: workspace too small (%zu bytes given)", ws_bytes);
  MODEST_REQUIRE(max_query_points < (1ll << 31), "pp_score: a scan has >= 2^31 points");

  const int G = grid_dim;
  const size_t ncell1 = cell_stride(G);
  Arena ar(d_ws, ws_bytes);
  PPScanMeta* meta = ar.take<PPScanMeta>(n_scans);
  int* cells = ar.take<int>((size_t)n_scans * ncell1);
  float4* sorted = ar.take<float4>(n_query_total);
  int* counts_ws = ar.take<int>(n_count_total);
  int32_t* trav_scan = ar.take<int32_t>(1 << 20);
  int* counts = d_counts ? d_counts : counts_ws;

  const float cell = (float)radius * kCellSlack;
  const double r2 = radius * radius;
  const float r2f = (float)r2;
  const float band = 1e-5f * r2f;   // ~100x the f32 evaluation error of d2 for d2 ~ r2

  MODEST_CUDA(cudaMemsetAsync(cells, 0, sizeof(int) * (size_t)n_scans * ncell1, stream));
  MODEST_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)n_count_total, stream));

  pp_trav_scan_kernel<<<n_scans, 32, 0, stream>>>(d_trav_off, n_scans, trav_scan);
  MODEST_LAUNCH_CHECK("pp_trav_scan_kernel");
  pp_query_origin_kernel<<<n_scans, 1024, 0, stream>>>(d_query_xyz, d_q_off, meta, G, cell);
  MODEST_LAUNCH_CHECK("pp_query_origin_kernel");
  const int qblocks = (int)((max_query_points + 255) / 256);
  dim3 qgrid(qblocks > 0 ? qblocks : 1, n_scans);
  pp_query_hist_kernel<<<qgrid, 256, 0, stream>>>(d_query_xyz, d_q_off, meta, cells, G);
  MODEST_LAUNCH_CHECK("pp_query_hist_kernel");
  pp_cell_scan_kernel<<<n_scans, 1024, 0, stream>>>(cells, G);
  MODEST_LAUNCH_CHECK("pp_cell_scan_kernel");
  pp_query_scatter_kernel<<<qgrid, 256, 0, stream>>>(d_query_xyz, d_q_off, meta, cells, sorted, G);
  MODEST_LAUNCH_CHECK("pp_query_scatter_kernel");
  if (n_trav_total > 0 && max_trav_points > 0) {
    int64_t hb = (max_trav_points + 255) / 256;
    if (hb > 65535) hb = 65535;
    dim3 hgrid((unsigned)hb, n_trav_total);
    MODEST_REQUIRE(n_trav_total <= 65535, "pp_score: more than 65535 traversals in one launch");
    pp_count_kernel<<<hgrid, 256, 0, stream>>>(d_hist_xyz, d_h_off, trav_scan, d_trav_off, d_q_off,
                                               d_count_off, meta, cells, sorted, counts, G, r2f, band, r2);
    MODEST_LAUNCH_CHECK("pp_count_kernel");
  }
  MODEST_REQUIRE(n_scans <= 65535, "pp_score: more than 65535 scans in one launch");
  pp_entropy_kernel<<<qgrid, 256, 0, stream>>>(counts, d_q_off, d_count_off, d_trav_off, d_pp);
  MODEST_LAUNCH_CHECK("pp_entropy_kernel");
  note_launch(7);
  return MODEST_OK;
}
