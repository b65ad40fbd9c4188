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
#include "grid2d.cuh"

namespace modest {
extern void note_launch(int n);

// ---- 5. stream the history once: probe 3 row segments, count hits ----------------------------
// blockIdx.y = global traversal index g; the scan it belongs to comes from trav_scan[g].
__global__ void __launch_bounds__(256) pp_count_kernel(
    const float* __restrict__ h_xyz, const int64_t* __restrict__ h_off,
    const int32_t* __restrict__ trav_scan, const int32_t* __restrict__ trav_off,
    const int64_t* __restrict__ q_off, const int64_t* __restrict__ count_off,
    const GridMeta* __restrict__ meta, const int* __restrict__ cells,
    const float4* __restrict__ sorted, int* __restrict__ counts, int G, float r2f, float band,
    double r2) {
  const int g = blockIdx.y;
  const int s = trav_scan[g];
  const int t = g - trav_off[s];
  const int T = trav_off[s + 1] - trav_off[s];
  const int64_t hbeg = h_off[g], hn = h_off[g + 1] - hbeg;
  const GridMeta m = meta[s];
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


// ---- stage B: rigid transform of scan frames into the fixed frame ------------------------------
// transform_points() (utils/pointcloud_utils.py:11-19) is [p,1] @ Tr^T in float32 through
// BLAS sgemm; its kernels accumulate the 4-term dot product with fused multiply-adds in k
// order, which is what this kernel does.  remove_center() (pre_compute_pp_score.py:48-52) is
// folded in: a removed point becomes NaN, which no distance test can ever accept.
__global__ void __launch_bounds__(256) transform_frames_kernel(
    const float* __restrict__ in, int stride, const int64_t* __restrict__ frame_off, const float* __restrict__ T,
    int remove_center, float cx0, float cx1, float cy0, float cy1, float* __restrict__ out) {
  const int f = blockIdx.y;
  const int64_t beg = frame_off[f], n = frame_off[f + 1] - beg;
  float t[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) t[k] = T[16 * f + k];
  const float nanv = __int_as_float(0x7fc00000);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float* p = in + (size_t)stride * (beg + i);
    const float x = p[0], y = p[1], z = p[2];
    float* o = out + 3 * (beg + i);
    if (remove_center && x < cx1 && x >= cx0 && y < cy1 && y >= cy0) {
      o[0] = nanv; o[1] = nanv; o[2] = nanv;
      continue;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j)
      o[j] = fmaf(1.0f, t[4 * j + 3], fmaf(z, t[4 * j + 2], fmaf(y, t[4 * j + 1], __fmul_rn(x, t[4 * j]))));
  }
}

}  // namespace modest

using namespace modest;

// ---- optional timing of the dominant kernel (bench.py's roofline line) ---------------------------
// A ring of CUDA event pairs recorded on the launching stream around pp_count_kernel; read back
// after the caller has synchronised.  Off by default.
static cudaEvent_t g_prof_ev[2 * 256];
static int g_prof_slots = 0;
static long long g_prof_calls = 0;

extern "C" int modest_pp_profile_enable(int n_slots) {
  if (n_slots < 0 || n_slots > 256) { set_error("pp_profile_enable: n_slots %d out of range [0,256]", n_slots); return MODEST_ERR_ARG; }
  for (int i = 0; i < 2 * g_prof_slots; ++i) cudaEventDestroy(g_prof_ev[i]);
  g_prof_slots = 0;
  g_prof_calls = 0;
  for (int i = 0; i < 2 * n_slots; ++i) MODEST_CUDA(cudaEventCreate(&g_prof_ev[i]));
  g_prof_slots = n_slots;
  return MODEST_OK;
}

// ms per recorded pp_count launch (most recent min(calls, slots)); returns the number written
extern "C" int modest_pp_profile_read(float* h_ms, int max_out) {
  int n = (int)(g_prof_calls < g_prof_slots ? g_prof_calls : g_prof_slots);
  if (n > max_out) n = max_out;
  for (int i = 0; i < n; ++i) {
    const int slot = (int)((g_prof_calls - 1 - i) % g_prof_slots);
    if (cudaEventElapsedTime(&h_ms[i], g_prof_ev[2 * slot], g_prof_ev[2 * slot + 1]) != cudaSuccess) { h_ms[i] = -1.f; cudaGetLastError(); }
  }
  return n;
}

static const float kCellSlack = 1.001f;   // cell edge = radius * slack, see cell_coord()

extern "C" size_t modest_pp_workspace_bytes(int n_scans, int64_t n_query_total, int64_t n_count_total,
                                            int grid_dim) {
  if (grid_dim <= 0) grid_dim = 512;
  size_t b = 0;
  auto add = [&](size_t bytes) { b = align_up(b, 256) + bytes; };
  add(sizeof(GridMeta) * (size_t)n_scans);
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
  GridMeta* meta = ar.take<GridMeta>(n_scans);
  int* cells = ar.take<int>((size_t)n_scans * ncell1);
  float4* sorted = ar.take<float4>(n_query_total);
  int* counts_ws = ar.take<int>(n_count_total);
  int32_t* trav_scan = ar.take<int32_t>(1 << 20);
  int* counts = d_counts ? d_counts : counts_ws;

  const float cell = (float)radius * kCellSlack;
  const double r2 = radius * radius;
  const float r2f = (float)r2;
  const float band = 1e-5f * r2f;   // ~100x the f32 evaluation error of d2 for d2 ~ r2

  MODEST_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)n_count_total, stream));

  pp_trav_scan_kernel<<<n_scans, 32, 0, stream>>>(d_trav_off, n_scans, trav_scan);
  MODEST_LAUNCH_CHECK("pp_trav_scan_kernel");
  {
    int rc = grid2d_build(d_query_xyz, 3, d_q_off, nullptr, n_scans, max_query_points, cell, G, meta, cells,
                          sorted, stream);
    if (rc != MODEST_OK) return rc;
  }
  const int qblocks = (int)((max_query_points + 255) / 256);
  dim3 qgrid(qblocks > 0 ? qblocks : 1, n_scans);
  if (n_trav_total > 0 && max_trav_points > 0) {
    int64_t hb = (max_trav_points + 255) / 256;
    if (hb > 65535) hb = 65535;
    dim3 hgrid((unsigned)hb, n_trav_total);
    MODEST_REQUIRE(n_trav_total <= 65535, "pp_score: more than 65535 traversals in one launch");
    const int slot = g_prof_slots ? (int)(g_prof_calls % g_prof_slots) : -1;
    if (slot >= 0) cudaEventRecord(g_prof_ev[2 * slot], stream);
    pp_count_kernel<<<hgrid, 256, 0, stream>>>(d_hist_xyz, d_h_off, trav_scan, d_trav_off, d_q_off,
                                               d_count_off, meta, cells, sorted, counts, G, r2f, band, r2);
    MODEST_LAUNCH_CHECK("pp_count_kernel");
    if (slot >= 0) { cudaEventRecord(g_prof_ev[2 * slot + 1], stream); ++g_prof_calls; }
  }
  MODEST_REQUIRE(n_scans <= 65535, "pp_score: more than 65535 scans in one launch");
  pp_entropy_kernel<<<qgrid, 256, 0, stream>>>(counts, d_q_off, d_count_off, d_trav_off, d_pp);
  MODEST_LAUNCH_CHECK("pp_entropy_kernel");
  note_launch(3);
  return MODEST_OK;
}

extern "C" int modest_transform_frames_batch(const float* d_in, int point_stride, const int64_t* d_frame_off,
                                             const float* d_T, int n_frames, int64_t max_frame_points,
                                             int remove_center, const float* h_center_box, float* d_out,
                                             void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_frames <= 0 || max_frame_points <= 0) return MODEST_OK;
  MODEST_REQUIRE(d_in && d_frame_off && d_T && d_out, "transform_frames: null pointer argument");
  MODEST_REQUIRE(point_stride >= 3, "transform_frames: point_stride %d < 3", point_stride);
  MODEST_REQUIRE(n_frames <= 65535, "transform_frames: more than 65535 frames in one launch");
  MODEST_REQUIRE(!remove_center || h_center_box, "transform_frames: remove_center without a box");
  int64_t blocks = (max_frame_points + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  const float* c = h_center_box;
  transform_frames_kernel<<<dim3((unsigned)blocks, n_frames), 256, 0, stream>>>(
      d_in, point_stride, d_frame_off, d_T, remove_center, c ? c[0] : 0.f, c ? c[1] : 0.f, c ? c[2] : 0.f,
      c ? c[3] : 0.f, d_out);
  MODEST_LAUNCH_CHECK("transform_frames_kernel");
  note_launch(1);
  return MODEST_OK;
}
