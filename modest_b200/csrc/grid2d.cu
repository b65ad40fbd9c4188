// Counting sort of each scan's points into a per-scan 2-D cell grid (see grid2d.cuh).
#include "grid2d.cuh"

namespace modest {
extern void note_launch(int n);

__device__ __forceinline__ int scan_count(const int64_t* off, const int32_t* cnt, int s) {
  return cnt ? cnt[s] : (int)(off[s + 1] - off[s]);
}

// ---- 1. per-scan query bounding box -> grid origin -------------------------------------------
__global__ void __launch_bounds__(1024) grid_origin_kernel(
    const float* __restrict__ q_xyz, int stride, const int64_t* __restrict__ q_off,
    const int32_t* __restrict__ cnt, GridMeta* __restrict__ meta, int G, float cell) {
  const int s = blockIdx.x;
  const int64_t beg = q_off[s], end = beg + scan_count(q_off, cnt, s);
  float lox = 3.0e38f, loy = 3.0e38f, hix = -3.0e38f, hiy = -3.0e38f;
  for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
    float x = q_xyz[(size_t)stride * i], y = q_xyz[(size_t)stride * i + 1];
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
      GridMeta m;
      if (end <= beg) { lox = loy = hix = hiy = 0.f; }
      const float half = 0.5f * cell * (float)G;
      m.x0 = 0.5f * (lox + hix) - half;
      m.y0 = 0.5f * (loy + hiy) - half;
      m.inv_cell = 1.0f / cell;
      m.n = (int)(end - beg);
      meta[s] = m;
    }
  }
}

// ---- 2. histogram of query points per cell ---------------------------------------------------
__global__ void __launch_bounds__(256) grid_hist_kernel(
    const float* __restrict__ q_xyz, int stride, const int64_t* __restrict__ q_off,
    const GridMeta* __restrict__ meta, int* __restrict__ cells, int G) {
  const int s = blockIdx.y;
  const GridMeta m = meta[s];
  const int64_t beg = q_off[s], n = m.n;
  int* c = cells + (size_t)s * cell_stride(G);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float* p = q_xyz + (size_t)stride * (beg + i);
    int cx = clampi(cell_coord(p[0], m.x0, m.inv_cell), 0, G - 1);
    int cy = clampi(cell_coord(p[1], m.y0, m.inv_cell), 0, G - 1);
    atomicAdd(&c[cy * G + cx], 1);
  }
}

// ---- 3. in-place inclusive scan of the G*G cell counts, one CTA per scan ---------------------
__global__ void __launch_bounds__(1024) grid_cell_scan_kernel(int* __restrict__ cells, int G) {
  const size_t ncell = (size_t)G * G;
  int* c = cells + (size_t)blockIdx.x * cell_stride(G);
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
__global__ void __launch_bounds__(256) grid_scatter_kernel(
    const float* __restrict__ q_xyz, int stride, const int64_t* __restrict__ q_off,
    const GridMeta* __restrict__ meta, int* __restrict__ cells, float4* __restrict__ sorted, int G) {
  const int s = blockIdx.y;
  const GridMeta m = meta[s];
  const int64_t beg = q_off[s], n = m.n;
  int* c = cells + (size_t)s * cell_stride(G);
  float4* out = sorted + beg;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float* p = q_xyz + (size_t)stride * (beg + i);
    const float x = p[0], y = p[1], z = p[2];
    int cx = clampi(cell_coord(x, m.x0, m.inv_cell), 0, G - 1);
    int cy = clampi(cell_coord(y, m.y0, m.inv_cell), 0, G - 1);
    // the scanned array holds the END of each cell; counting down leaves the START behind
    int pos = atomicSub(&c[cy * G + cx], 1) - 1;
    out[pos] = make_float4(x, y, z, __int_as_float((int)i));
  }
}

int grid2d_build(const float* pts, int stride, const int64_t* off, const int32_t* cnt, int n_scans,
                 int64_t max_points, float cell, int G, GridMeta* meta, int* cells, float4* sorted,
                 cudaStream_t stream) {
  if (n_scans <= 0) return MODEST_OK;
  MODEST_REQUIRE(G >= 8 && G <= 4096 && G % 4 == 0, "grid2d: G=%d must be a multiple of 4 in [8,4096]", G);
  MODEST_REQUIRE(n_scans <= 65535, "grid2d: more than 65535 scans in one launch");
  MODEST_REQUIRE(max_points < (1ll << 31), "grid2d: a scan has >= 2^31 points");
  MODEST_CUDA(cudaMemsetAsync(cells, 0, sizeof(int) * (size_t)n_scans * cell_stride(G), stream));
  grid_origin_kernel<<<n_scans, 1024, 0, stream>>>(pts, stride, off, cnt, meta, G, cell);
  MODEST_LAUNCH_CHECK("grid_origin_kernel");
  int qblocks = (int)((max_points + 255) / 256);
  if (qblocks < 1) qblocks = 1;
  if (qblocks > 4096) qblocks = 4096;
  dim3 qgrid(qblocks, n_scans);
  grid_hist_kernel<<<qgrid, 256, 0, stream>>>(pts, stride, off, meta, cells, G);
  MODEST_LAUNCH_CHECK("grid_hist_kernel");
  grid_cell_scan_kernel<<<n_scans, 1024, 0, stream>>>(cells, G);
  MODEST_LAUNCH_CHECK("grid_cell_scan_kernel");
  grid_scatter_kernel<<<qgrid, 256, 0, stream>>>(pts, stride, off, meta, cells, sorted, G);
  MODEST_LAUNCH_CHECK("grid_scatter_kernel");
  note_launch(4);
  return MODEST_OK;
}

}  // namespace modest
