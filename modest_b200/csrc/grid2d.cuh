// Per-scan 2-D cell grid over (x,y): counting sort of a scan's points into row-major cell
// order.  Shared by the PP-score stage (cell edge ~ 0.3 m) and the kNN-graph stage (~0.5 m).
#pragma once
#include "common.cuh"

namespace modest {

struct GridMeta {        // per scan, device resident
  float x0, y0;          // grid origin
  float inv_cell;        // 1 / cell edge
  int   n;               // number of points binned
};

// ints per scan in the cell table: G*G cells + sentinel, padded so every scan stays 16-B aligned
__host__ __device__ __forceinline__ size_t cell_stride(int G) { return (size_t)G * G + 4; }

__device__ __forceinline__ int cell_coord(float v, float origin, float inv_cell) {
  // monotone in v (one rounded subtract, one rounded multiply, floor) -- that is all the
  // neighbour search needs; see DESIGN.md for the |cell(q) - cell(h)| <= 1 argument.
  return __float2int_rd(__fmul_rn(__fsub_rn(v, origin), inv_cell));
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Bins the points of every scan.  Scan s owns rows [off[s], off[s]+n_s) of `pts` (row =
// `stride` floats, xyz first) with n_s = cnt ? cnt[s] : off[s+1]-off[s].
//   meta   (n_scans)                 out
//   cells  (n_scans * cell_stride)   out: cells[c] = first sorted position of cell c, cells[G*G] = n_s
//   sorted (rows of pts, at off[s])  out: float4(x, y, z, original index bits)
// Enqueues 1 memset + 4 kernels on `stream`.
int grid2d_build(const float* pts, int stride, const int64_t* off, const int32_t* cnt, int n_scans,
                 int64_t max_points, float cell, int G, GridMeta* meta, int* cells, float4* sorted,
                 cudaStream_t stream);

}  // namespace modest
