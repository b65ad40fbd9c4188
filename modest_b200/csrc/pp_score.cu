// Stage C+D: persistence-point score (neighbour counts over T traversals + entropy), and stage B
// (frame transform).
//
// Reference behaviour: pre_compute_pp_score.py:54-75 (count_neighbors, compute_ephe_score) on
// the cKDTrees built at :188-190.
//
// Design (DESIGN.md "PP score").  The query scan (60k points) is the small side, the history
// (T traversals, 16x larger and up) the big one, so the query is indexed and the history is
// streamed through exactly once:
//   * query index, per scan: a G x G table of (x,y) columns with cell edge just above the
//     search radius; each column record holds a 32-bit occupancy mask over z-cells of the same
//     edge and the index of the column's first occupied (column,z) cell in a compact array
//     `zc` of start positions into the query sorted by (y, x, z) cell.  A 3x3x3 neighbourhood
//     probe is then 9 record loads + popcounts, and touches only occupied cells.
//   * history pass: one thread per history point; candidates from the 27 neighbouring cells
//     are tested in f32 and re-tested in sequential f64 (the arithmetic of cKDTree) only
//     inside a thin shell around the sphere surface, so the counts are bit-exact.  Hits are
//     accumulated with one red.global.add per (query, traversal) hit into counts laid out
//     [traversal][sorted query position] (spatial neighbours share sectors).
//   * entropy pass: per sorted query position, T counts -> H in f64 -> pp[original index].
#include <algorithm>

#include "grid2d.cuh"

namespace modest {
extern void note_launch(int n);

constexpr int kZCells = 32;

struct PPMeta {            // per scan, device resident
  float x0, y0, z0, inv_cell;
  int n;                   // query points
  int m;                   // occupied (column, z) cells
  int pad0, pad1;
};

__device__ __forceinline__ size_t pp_col_stride(int G) { return (size_t)G * G; }
// scan s's slice of the compact-cell array: 16-B aligned, room for N_s + 1 entries
__device__ __forceinline__ size_t zc_offset(const int64_t* q_off, int s) { return (size_t)((q_off[s] + 3) & ~3ll) + 8 * (size_t)s; }
__device__ __forceinline__ unsigned below_mask(int bit) { return bit >= 32 ? 0xffffffffu : ((1u << bit) - 1u); }

struct PPCell { int col; int cz; };
__device__ __forceinline__ PPCell pp_cell_of(float x, float y, float z, const PPMeta& m, int G) {
  const int cx = clampi(cell_coord(x, m.x0, m.inv_cell), 0, G - 1);
  const int cy = clampi(cell_coord(y, m.y0, m.inv_cell), 0, G - 1);
  const int cz = clampi(cell_coord(z, m.z0, m.inv_cell), 0, kZCells - 1);
  return PPCell{cy * G + cx, cz};
}

// ---- 1. bounding box of the query -> grid origin ------------------------------------------------
__global__ void __launch_bounds__(1024) pp_origin_kernel(const float* __restrict__ q_xyz, const int64_t* __restrict__ q_off,
                                                        PPMeta* __restrict__ meta, int G, float cell) {
  const int s = blockIdx.x;
  const int64_t beg = q_off[s], end = q_off[s + 1];
  float lox = 3.0e38f, loy = 3.0e38f, loz = 3.0e38f, hix = -3.0e38f, hiy = -3.0e38f;
  for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
    const float x = q_xyz[3 * i], y = q_xyz[3 * i + 1], z = q_xyz[3 * i + 2];
    lox = fminf(lox, x); hix = fmaxf(hix, x); loy = fminf(loy, y); hiy = fmaxf(hiy, y); loz = fminf(loz, z);
  }
  __shared__ float sh[5][32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lox = fminf(lox, __shfl_xor_sync(0xffffffffu, lox, o)); loy = fminf(loy, __shfl_xor_sync(0xffffffffu, loy, o));
    loz = fminf(loz, __shfl_xor_sync(0xffffffffu, loz, o));
    hix = fmaxf(hix, __shfl_xor_sync(0xffffffffu, hix, o)); hiy = fmaxf(hiy, __shfl_xor_sync(0xffffffffu, hiy, o));
  }
  if (l == 0) { sh[0][w] = lox; sh[1][w] = loy; sh[2][w] = hix; sh[3][w] = hiy; sh[4][w] = loz; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
      lox = fminf(lox, sh[0][k]); loy = fminf(loy, sh[1][k]); hix = fmaxf(hix, sh[2][k]); hiy = fmaxf(hiy, sh[3][k]);
      loz = fminf(loz, sh[4][k]);
    }
    if (end <= beg) { lox = loy = hix = hiy = loz = 0.f; }
    const float half = 0.5f * cell * (float)G;
    PPMeta m;
    m.x0 = 0.5f * (lox + hix) - half;
    m.y0 = 0.5f * (loy + hiy) - half;
    m.z0 = loz;
    m.inv_cell = 1.0f / cell;
    m.n = (int)(end - beg);
    m.m = 0; m.pad0 = m.pad1 = 0;
    meta[s] = m;
  }
}

// ---- 2. z-occupancy mask per column -------------------------------------------------------------
__global__ void __launch_bounds__(256) pp_mask_kernel(const float* __restrict__ q_xyz, const int64_t* __restrict__ q_off,
                                                      const PPMeta* __restrict__ meta, int2* __restrict__ cols, int G) {
  const int s = blockIdx.y;
  const PPMeta m = meta[s];
  const int64_t beg = q_off[s];
  int2* c = cols + (size_t)s * pp_col_stride(G);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m.n; i += gridDim.x * blockDim.x) {
    const float* p = q_xyz + 3 * (beg + i);
    const PPCell cc = pp_cell_of(p[0], p[1], p[2], m, G);
    atomicOr(reinterpret_cast<unsigned*>(&c[cc.col].y), 1u << cc.cz);
  }
}

// ---- 3. exclusive scan of popc(mask) over the columns: tile sums, their scan, apply --------------
constexpr int kColTile = 4096;     // columns per CTA (1024 threads x 4)

__global__ void __launch_bounds__(1024) pp_col_tilesum_kernel(const int2* __restrict__ cols, int G, int* __restrict__ tile_sums,
                                                             int tiles_per_scan) {
  const int s = blockIdx.y, tile = blockIdx.x;
  const size_t ncol = pp_col_stride(G);
  const int2* c = cols + (size_t)s * ncol;
  int v = 0;
  for (int k = 0; k < 4; ++k) {
    const size_t i = (size_t)tile * kColTile + k * 1024 + threadIdx.x;
    if (i < ncol) v += __popc((unsigned)c[i].y);
  }
  v = warp_sum(v);
  __shared__ int sh[32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    int t = sh[threadIdx.x];
    t = warp_sum(t);
    if (threadIdx.x == 0) tile_sums[s * tiles_per_scan + tile] = t;
  }
}

__global__ void pp_col_tilescan_kernel(int* __restrict__ tile_sums, int tiles_per_scan, PPMeta* __restrict__ meta) {
  const int s = blockIdx.x;
  int* t = tile_sums + s * tiles_per_scan;
  const int lane = threadIdx.x;
  int carry = 0;
  for (int b = 0; b < tiles_per_scan; b += 32) {
    const int i = b + lane;
    const int v = i < tiles_per_scan ? t[i] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (i < tiles_per_scan) t[i] = carry + inc - v;
    carry += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (lane == 0) meta[s].m = carry;
}

__global__ void __launch_bounds__(1024) pp_col_apply_kernel(int2* __restrict__ cols, int G, const int* __restrict__ tile_sums,
                                                           int tiles_per_scan) {
  const int s = blockIdx.y, tile = blockIdx.x;
  const size_t ncol = pp_col_stride(G);
  int2* c = cols + (size_t)s * ncol;
  const size_t i0 = (size_t)tile * kColTile + 4 * (size_t)threadIdx.x;      // 4 consecutive columns per thread
  int pc[4], tot = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) { pc[k] = (i0 + k < ncol) ? __popc((unsigned)c[i0 + k].y) : 0; tot += pc[k]; }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  __shared__ int wex[32];
  if (lane == 31) wex[w] = inc;
  __syncthreads();
  if (w == 0) {
    int t = wex[lane], ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, ti, o);
      if (lane >= o) ti += u;
    }
    wex[lane] = ti - t;
  }
  __syncthreads();
  int base = tile_sums[s * tiles_per_scan + tile] + wex[w] + inc - tot;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (i0 + k < ncol) c[i0 + k].x = base;
    base += pc[k];
  }
}

// ---- 4. points per occupied (column, z) cell ------------------------------------------------------
__global__ void __launch_bounds__(256) pp_cellcount_kernel(const float* __restrict__ q_xyz, const int64_t* __restrict__ q_off,
                                                           const PPMeta* __restrict__ meta, const int2* __restrict__ cols,
                                                           int* __restrict__ zc, int G) {
  const int s = blockIdx.y;
  const PPMeta m = meta[s];
  const int64_t beg = q_off[s];
  const int2* c = cols + (size_t)s * pp_col_stride(G);
  int* z = zc + zc_offset(q_off, s);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m.n; i += gridDim.x * blockDim.x) {
    const float* p = q_xyz + 3 * (beg + i);
    const PPCell cc = pp_cell_of(p[0], p[1], p[2], m, G);
    const int2 rec = c[cc.col];
    atomicAdd(&z[rec.x + __popc((unsigned)rec.y & below_mask(cc.cz))], 1);
  }
}

// ---- 5. inclusive scan of the compact cell counts (one CTA per scan; <= N entries) ---------------
__global__ void __launch_bounds__(1024) pp_cellscan_kernel(int* __restrict__ zc, const int64_t* __restrict__ q_off,
                                                           const PPMeta* __restrict__ meta) {
  const int s = blockIdx.x;
  int* c = zc + zc_offset(q_off, s);
  const int ncell = meta[s].m;
  __shared__ int warp_excl[32];
  __shared__ int tile_total;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int carry = 0;
  for (int base = 0; base < ncell; base += 4096) {
    const int i = base + 4 * threadIdx.x;
    int4 v = make_int4(0, 0, 0, 0);
    if (i + 3 < ncell) v = *reinterpret_cast<const int4*>(c + i);
    else {
      if (i < ncell) v.x = c[i];
      if (i + 1 < ncell) v.y = c[i + 1];
      if (i + 2 < ncell) v.z = c[i + 2];
    }
    v.y += v.x; v.z += v.y; v.w += v.z;
    int incl = v.w;
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
    __syncthreads();
  }
  if (threadIdx.x == 0) c[ncell] = carry;     // sentinel = number of query points
}

// ---- 6. scatter the query into (y, x, z)-cell order ------------------------------------------------
__global__ void __launch_bounds__(256) pp_scatter_kernel(const float* __restrict__ q_xyz, const int64_t* __restrict__ q_off,
                                                         const PPMeta* __restrict__ meta, const int2* __restrict__ cols,
                                                         int* __restrict__ zc, float4* __restrict__ sorted, int G) {
  const int s = blockIdx.y;
  const PPMeta m = meta[s];
  const int64_t beg = q_off[s];
  const int2* c = cols + (size_t)s * pp_col_stride(G);
  int* z = zc + zc_offset(q_off, s);
  float4* out = sorted + beg;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m.n; i += gridDim.x * blockDim.x) {
    const float* p = q_xyz + 3 * (beg + i);
    const float x = p[0], y = p[1], zz = p[2];
    const PPCell cc = pp_cell_of(x, y, zz, m, G);
    const int2 rec = c[cc.col];
    // zc holds the END of every cell after the scan; counting down leaves the START behind
    const int pos = atomicSub(&z[rec.x + __popc((unsigned)rec.y & below_mask(cc.cz))], 1) - 1;
    out[pos] = make_float4(x, y, zz, __int_as_float(i));
  }
}

// ---- 7. stream the history once -------------------------------------------------------------------
// blockIdx.y = global traversal index g; the scan it belongs to comes from trav_scan[g].
// A warp takes 32 consecutive history points.  Phase 1: every lane looks up the (<= 9) non-empty
// neighbouring columns of its point and leaves their candidate ranges in shared memory.
// Phase 2: the warp's candidates (about 5 per point on average, but heavy-tailed: a point next
// to a wall or pole sees 10x more) are dealt out evenly -- lane l tests the l-th slice of the
// concatenated candidate sequence, whichever point it belongs to -- so the SIMT lanes stay busy.
constexpr int kCountWarps = 8;

__global__ void __launch_bounds__(kCountWarps * 32, 6) pp_count_kernel(
    const float* __restrict__ h_xyz, const int64_t* __restrict__ h_off, const int32_t* __restrict__ trav_scan,
    const int32_t* __restrict__ trav_off, const int64_t* __restrict__ q_off, const int64_t* __restrict__ count_off,
    const PPMeta* __restrict__ meta, const int2* __restrict__ cols, const int* __restrict__ zc,
    const float4* __restrict__ sorted, int* __restrict__ counts, int G, float r2f, float band, double r2) {
  const int g = blockIdx.y;
  const int s = trav_scan[g];
  const int t = g - trav_off[s];
  const int64_t hbeg = h_off[g], hn = h_off[g + 1] - hbeg;
  const PPMeta m = meta[s];
  const int2* __restrict__ c = cols + (size_t)s * pp_col_stride(G);
  const int* __restrict__ z = zc + zc_offset(q_off, s);
  const float4* __restrict__ qs = sorted + q_off[s];
  int* __restrict__ cnt = counts + count_off[s] + (size_t)t * m.n;       // [traversal][sorted position]
  const float lo = r2f - band, hi = r2f + band;

  __shared__ int s_rb[kCountWarps][32 * 9];
  __shared__ int s_re[kCountWarps][32 * 9];
  __shared__ float s_pt[kCountWarps][3][32];
  __shared__ int s_nr[kCountWarps][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int* rb = s_rb[w];
  int* re = s_re[w];

  const int64_t warp0 = ((int64_t)blockIdx.x * kCountWarps + w) * 32;
  const int64_t stride = (int64_t)gridDim.x * kCountWarps * 32;
  for (int64_t base = warp0; base < hn; base += stride) {          // warp-uniform trip count
    const int64_t i = base + lane;
    // ---- phase 1 ----
    int nr = 0, total = 0;
    float hx = 0.f, hy = 0.f, hz = 0.f;
    if (i < hn) {
      const float* p = h_xyz + 3 * (hbeg + i);
      hx = __ldg(p); hy = __ldg(p + 1); hz = __ldg(p + 2);
      const int cx = cell_coord(hx, m.x0, m.inv_cell);
      const int cy = cell_coord(hy, m.y0, m.inv_cell);
      const int cz = cell_coord(hz, m.z0, m.inv_cell);
      const int xa = clampi(cx - 1, 0, G - 1), xb = clampi(cx + 1, 0, G - 1);
      const int ya = clampi(cy - 1, 0, G - 1), yb = clampi(cy + 1, 0, G - 1);
      const int za = clampi(cz - 1, 0, kZCells - 1), zb = clampi(cz + 1, 0, kZCells - 1);
      const unsigned below_a = below_mask(za), upto_b = below_mask(zb + 1);
      const unsigned want = upto_b & ~below_a;
      // staged so that the 9 record loads, then the (up to 18) cell-start loads, are each in
      // flight together instead of one dependent round trip per column
      int2 rec[9];
      bool ok[9];
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const int y = ya + j / 3, x = xa + j % 3;
        ok[j] = (y <= yb) && (x <= xb);
        rec[j] = ok[j] ? __ldg(c + y * G + x) : make_int2(0, 0);
      }
      int kb[9], ke[9];
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const unsigned msk = (unsigned)rec[j].y;
        ok[j] = ok[j] && (msk & want);
        kb[j] = rec[j].x + __popc(msk & below_a);      // compact-cell indices for now
        ke[j] = rec[j].x + __popc(msk & upto_b);
      }
#pragma unroll
      for (int j = 0; j < 9; ++j)
        if (ok[j]) { kb[j] = __ldg(z + kb[j]); ke[j] = __ldg(z + ke[j]); }
#pragma unroll
      for (int j = 0; j < 9; ++j)
        if (ok[j]) {
          rb[lane * 9 + j] = kb[j];
          re[lane * 9 + j] = ke[j];
          total += ke[j] - kb[j];
          nr |= 1 << j;                                  // nr is the 9-bit mask of live ranges
        }
    }
    s_nr[w][lane] = nr;
    s_pt[w][0][lane] = hx; s_pt[w][1][lane] = hy; s_pt[w][2][lane] = hz;
    int incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    const int excl = incl - total;
    const int T = __shfl_sync(0xffffffffu, incl, 31);
    __syncwarp();
    if (T == 0) continue;                                        // warp-uniform
    // ---- phase 2 ----
    const int S = (T + 31) >> 5;
    const int my_begin = min(lane * S, T), my_end = min(T, my_begin + S);
    int lo_l = 0, hi_l = 31;                                      // largest lane L with excl[L] <= my_begin
#pragma unroll
    for (int it = 0; it < 5; ++it) {
      const int mid = (lo_l + hi_l + 1) >> 1;
      const int v = __shfl_sync(0xffffffffu, excl, mid);
      if (v <= my_begin) lo_l = mid; else hi_l = mid - 1;
    }
    int src = lo_l;
    int off = my_begin - __shfl_sync(0xffffffffu, excl, src);
    int left = my_end - my_begin;
    if (left > 0) {
      unsigned live = (unsigned)s_nr[w][src];                     // ranges of `src` not yet consumed
      int j = __ffs(live) - 1;
      live &= live - 1;
      int k = rb[src * 9 + j], ke = re[src * 9 + j];
      while (off >= ke - k) {                                     // skip whole ranges of this point
        off -= ke - k;
        j = __ffs(live) - 1;
        live &= live - 1;
        k = rb[src * 9 + j]; ke = re[src * 9 + j];
      }
      k += off;
      float px = s_pt[w][0][src], py = s_pt[w][1][src], pz = s_pt[w][2][src];
      while (true) {
        const float4 q = __ldg(qs + k);
        const float dx = q.x - px, dy = q.y - py, dz = q.z - pz;
        const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        bool hit = d2 < lo;
        if (!hit && d2 <= hi) hit = sqdist_f64_seq(q.x, q.y, q.z, px, py, pz) <= r2;
        if (hit) atomicAdd(cnt + k, 1);
        if (--left == 0) break;
        if (++k >= ke) {
          if (live == 0) {                                        // next point that has candidates
            do { ++src; live = (unsigned)s_nr[w][src]; } while (live == 0);
            px = s_pt[w][0][src]; py = s_pt[w][1][src]; pz = s_pt[w][2][src];
          }
          j = __ffs(live) - 1;
          live &= live - 1;
          k = rb[src * 9 + j]; ke = re[src * 9 + j];
        }
      }
    }
    __syncwarp();                                                 // ranges are rewritten next chunk
  }
}

// ---- 8. entropy over traversals -------------------------------------------------------------------
// The reference's count array is np.stack(cols).T (pre_compute_pp_score.py:59-60), i.e. F-ordered,
// and P and -P*log(P) inherit that layout, so numpy's .sum(axis=1) walks a strided axis and adds
// the T terms one after the other in t order (measured on numpy 2.3.5: the sequential sum equals
// numpy's on 100 % of rows at T = 16, the pairwise scheme of a contiguous axis on 48 %).
template <typename F>
__device__ __forceinline__ double numpy_strided_sum(int n, F term) {
  double r = 0.0;
  for (int i = 0; i < n; ++i) r = __dadd_rn(r, term(i));
  return r;
}

// One term of the entropy, -P ln(P + 1e-8) with P = count / (total + 1e-8): a function of the two
// integers only.  The terms of all (total <= kEntS, count <= total) pairs are tabulated once per
// call by the arithmetic below (4 656 logarithms instead of ~40 per query point; a scan's points
// see ~28 neighbours in total on average, so nearly every row takes the table): same bits.
constexpr int kEntS = 96;
constexpr int kEntTabLen = (kEntS + 1) * (kEntS + 2) / 2;
__device__ __forceinline__ double pp_entropy_term(int ct, double denom) {
  if (ct == 0) return 0.0;                           // -0.0 * ln(1e-8) is exactly +0.0: skip the log
  const double P = __ddiv_rn((double)ct, denom);
  return __dmul_rn(-P, log(__dadd_rn(P, 1e-8)));
}
__global__ void pp_entropy_table_kernel(double* __restrict__ tab) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kEntTabLen) return;
  int tot = 0;
  while ((tot + 1) * (tot + 2) / 2 <= i) ++tot;      // i = tot (tot + 1) / 2 + ct, ct <= tot
  const int ct = i - tot * (tot + 1) / 2;
  tab[i] = pp_entropy_term(ct, __dadd_rn((double)tot, 1e-8));
}

// H of one query point from its T neighbour counts (compute_ephe_score, :68-75)
template <typename C>
__device__ __forceinline__ float pp_entropy_of(int T, double logT, C count_of, const double* __restrict__ tab = nullptr) {
  long long tot = 0;
  for (int t = 0; t < T; ++t) tot += count_of(t);
  if (tab != nullptr && tot <= kEntS) {
    const double* row = tab + (int)tot * ((int)tot + 1) / 2;
    const double acc = numpy_strided_sum(T, [&](int t) { return __ldg(row + count_of(t)); });
    return (float)__ddiv_rn(acc, logT);
  }
  const double denom = __dadd_rn((double)tot, 1e-8);
  const double acc = numpy_strided_sum(T, [&](int t) { return pp_entropy_term(count_of(t), denom); });
  return (float)__ddiv_rn(acc, logT);
}

__global__ void __launch_bounds__(256) pp_entropy_kernel(
    const int* __restrict__ counts, const float4* __restrict__ sorted, const int64_t* __restrict__ q_off,
    const int64_t* __restrict__ count_off, const int32_t* __restrict__ trav_off, float* __restrict__ pp,
    int32_t* __restrict__ counts_out /* (N,T) row-major per scan at count_off, or NULL */, const double* __restrict__ tab) {
  const int s = blockIdx.y;
  const int T = trav_off[s + 1] - trav_off[s];
  const int64_t qbeg = q_off[s];
  const int n = (int)(q_off[s + 1] - qbeg);
  const int* __restrict__ c = counts + count_off[s];
  const double logT = log((double)T);
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int orig = __float_as_int(sorted[qbeg + k].w);
    pp[qbeg + orig] = pp_entropy_of(T, logT, [&](int t) { return c[(size_t)t * n + k]; }, tab);
    if (counts_out)
      for (int t = 0; t < T; ++t) counts_out[count_off[s] + (size_t)orig * T + t] = c[(size_t)t * n + k];
  }
}

// ==== tiled path (round 2): coarse spatial partition + per-tile shared-memory join ================
// The history pass above probes a global hash grid from every history point; what bounds it is
// the number of scattered global accesses per point.  The tiled path makes every
// scattered access a shared-memory access:
//   * the (x,y) cell grid is cut into tiles of TW x TW columns.  K1 derives, from the query
//     index, how many query points every tile holds (and where: TW contiguous segments of the
//     (y,x,z)-sorted query);
//   * A0/A1 (pp_hist_tile_kernel) stream the history twice, coalesced: count, then scatter
//     float4{x,y,z,traversal} records into one contiguous bin per tile that holds query points.
//     A history point goes to its own tile and to every neighbouring tile whose one-cell halo it
//     touches (1.56 copies on average for TW = 8), so that a tile's bin holds EVERY history point
//     that can be within the radius of one of the tile's query points.  The atomics of the own
//     tile are aggregated per warp (lidar order is azimuth-coherent: a warp's 32 points fall
//     into a few tiles);
//   * the join (pp_join_kernel, persistent CTAs pulling work items heaviest first) keeps a tile's
//     query points and their per-traversal counters in shared memory, streams the tile's bin in
//     chunks of kChunk records, counting-sorts every chunk by (column, z-cell) inside shared
//     memory (u16 table over the (TW+2)^2 columns x the z-cells the tile's queries can reach,
//     rank from the histogram atomic), then every query lane walks the cells around it.  Tiles
//     with few query points spread each query's candidates over several lanes.  Counts are
//     complete when the bin is consumed, so the entropy is written from shared memory: no global
//     count array, no global atomics per hit.
// Work is done in groups of scans small enough for the bins to stay L2-resident between the
// scatter and the join.
constexpr int kJoinThreads = 256;                       // CTA items: one CTA per item, kChunk records per pass
constexpr int kChunk = 2048;
constexpr int kWarpThreads = 32;                        // warp items (small tiles): one warp per item
constexpr int kWarpChunk = 256;
constexpr int kPtsPerThreadJ = 8;                       // records a thread holds while a chunk is sorted
constexpr int kTagShift = 8;                            // a record's w = traversal << kTagShift
constexpr int kJoinMaxT = 32;                           // traversals per scan the tiled path accepts
constexpr int kHeavyBin = 2 * kChunk;                   // bins at least this long: 64 queries per CTA item
constexpr int kQCapCta = kJoinThreads, kQCapHeavy = 64, kQCapWarp = kWarpThreads;
constexpr int kWarpMaxRecords = 2048, kWarpMaxQueries = 64;   // tiles up to this size are warp items
constexpr int kFullColumnZ = 4;                         // z-extent up to which a column is taken whole

template <int TW>
struct PPItemT {               // one unit of join work: <= 256 query points of one tile
  int scan;
  short tx, ty;
  int k0, nq;                  // k0 = rank of the item's first query inside the tile's query list
  long long boff;              // first record of the tile's bin
  double logT;                 // ln(traversals of the scan)
  int bcnt, pad;
  int seg_start[TW];           // the tile's query list = TW segments of the sorted query (per row)
  int seg_len[TW];
};

struct PPGroupCtr {            // per group of scans, zeroed before use
  unsigned long long total;    // bin records reserved so far
  int n_cta, n_warp;           // CTA items are appended at the front, warp items at the back of the item array
  int next_cta, next_warp;     // work counters of the two join kernels
};

template <int TW, int NTHR> struct TileGeo {
  static constexpr int WW = TW + 2;                                       // tile + halo, in columns
  static constexpr int kCellsMax = WW * WW * kZCells;
  static constexpr int kWptMax = (((kCellsMax + 2) / 2 + NTHR - 1) / NTHR) | 1;
  static constexpr int kTblWords = kWptMax * NTHR;
};

// K1: query points per tile and their segments in the sorted query
template <int TW>
__global__ void __launch_bounds__(256) pp_tile_query_kernel(
    const int2* __restrict__ cols, const int* __restrict__ zc, const int64_t* __restrict__ q_off,
    const PPMeta* __restrict__ meta, int G, int NT1, int* __restrict__ tile_nq, int2* __restrict__ tile_seg) {
  const int s = blockIdx.y;
  const int NT = NT1 * NT1;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= NT) return;
  const int n = meta[s].n;
  const int2* __restrict__ c = cols + (size_t)s * pp_col_stride(G);
  const int* __restrict__ z = zc + zc_offset(q_off, s);
  const int ty = t / NT1, tx = t - ty * NT1;
  const int ncol = G * G;
  auto qstart = [&](int lin) { return lin < ncol ? __ldg(z + __ldg(&c[lin].x)) : n; };
  int total = 0;
  int2 seg[TW];
#pragma unroll
  for (int r = 0; r < TW; ++r) {
    const int lin0 = (ty * TW + r) * G + tx * TW;
    const int a = qstart(lin0), b = qstart(lin0 + TW);
    seg[r] = make_int2(a, b - a);
    total += b - a;
  }
  tile_nq[(size_t)s * NT + t] = total;
  if (total > 0) {
    int2* o = tile_seg + ((size_t)s * NT + t) * TW;
#pragma unroll
    for (int r = 0; r < TW; ++r) o[r] = seg[r];
  }
}

// A0 / A1: stream the history of the group's traversals; count (SCATTER = false) or write
// (SCATTER = true) one record per (point, tile that needs it).  A record's w is the traversal
// index << kTagShift.
template <int TW, bool SCATTER>
__global__ void __launch_bounds__(256) pp_hist_tile_kernel(
    const float* __restrict__ h_xyz, const int64_t* __restrict__ h_off, const int32_t* __restrict__ trav_scan,
    const int32_t* __restrict__ trav_off, int g0, const PPMeta* __restrict__ meta, const int* __restrict__ tile_nq,
    int* __restrict__ tile_cnt, const long long* __restrict__ tile_boff, float4* __restrict__ bins, int G, int NT1) {
  const int g = g0 + blockIdx.y;
  const int s = trav_scan[g];
  const float tagw = __int_as_float((g - trav_off[s]) << kTagShift);
  const int64_t hbeg = h_off[g], hn = h_off[g + 1] - hbeg;
  const PPMeta m = meta[s];
  const int NT = NT1 * NT1;
  const int* __restrict__ nq = tile_nq + (size_t)s * NT;
  int* __restrict__ cnt = tile_cnt + (size_t)s * NT;
  const long long* __restrict__ boff = tile_boff + (size_t)s * NT;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int64_t warp0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t base = warp0; base < hn; base += stride) {            // warp-uniform trip count
    const int64_t i = base + lane;
    float x = 0.f, y = 0.f, z = 0.f;
    int key0 = -1, key1 = -1, key2 = -1, key3 = -1;
    if (i < hn) {
      const float* p = h_xyz + 3 * (hbeg + i);
      x = __ldg(p); y = __ldg(p + 1); z = __ldg(p + 2);
      if (x == x && y == y && z == z) {                              // removed (NaN) rows never count
        const int cx = clampi(cell_coord(x, m.x0, m.inv_cell), 0, G - 1);
        const int cy = clampi(cell_coord(y, m.y0, m.inv_cell), 0, G - 1);
        const int tx = cx / TW, ty = cy / TW, lx = cx - tx * TW, ly = cy - ty * TW;
        const int nx = (lx == 0 && tx > 0) ? tx - 1 : ((lx == TW - 1 && tx < NT1 - 1) ? tx + 1 : -1);
        const int ny = (ly == 0 && ty > 0) ? ty - 1 : ((ly == TW - 1 && ty < NT1 - 1) ? ty + 1 : -1);
        key0 = ty * NT1 + tx;
        if (__ldg(nq + key0) == 0) key0 = -1;                          // nobody to count for over there
        if (nx >= 0 && __ldg(nq + ty * NT1 + nx) != 0) key1 = ty * NT1 + nx;
        if (ny >= 0 && __ldg(nq + ny * NT1 + tx) != 0) key2 = ny * NT1 + tx;
        if (nx >= 0 && ny >= 0 && __ldg(nq + ny * NT1 + nx) != 0) key3 = ny * NT1 + nx;
      }
    }
    // own tile: one atomic per (warp, tile); halo copies (a quarter of the lanes): one each.
    // All four are issued before any result is used.
    const unsigned grp = __match_any_sync(0xffffffffu, key0);
    const int leader = __ffs(grp) - 1;
    if (SCATTER) {
      int first = 0, p1 = 0, p2 = 0, p3 = 0;
      if (key0 >= 0 && lane == leader) first = atomicAdd(cnt + key0, __popc(grp));
      if (key1 >= 0) p1 = atomicAdd(cnt + key1, 1);
      if (key2 >= 0) p2 = atomicAdd(cnt + key2, 1);
      if (key3 >= 0) p3 = atomicAdd(cnt + key3, 1);
      first = __shfl_sync(0xffffffffu, first, leader);
      const float4 rec = make_float4(x, y, z, tagw);
      if (key0 >= 0) bins[boff[key0] + first + __popc(grp & lt)] = rec;
      if (key1 >= 0) bins[boff[key1] + p1] = rec;
      if (key2 >= 0) bins[boff[key2] + p2] = rec;
      if (key3 >= 0) bins[boff[key3] + p3] = rec;
    } else {
      if (key0 >= 0 && lane == leader) atomicAdd(cnt + key0, __popc(grp));
      if (key1 >= 0) atomicAdd(cnt + key1, 1);
      if (key2 >= 0) atomicAdd(cnt + key2, 1);
      if (key3 >= 0) atomicAdd(cnt + key3, 1);
    }
  }
}

// K3: bin offsets and work items.  One thread per tile; bins and items may come in any order, so
// a warp reserves the space of its 32 tiles with one atomic per quantity.  Small tiles become
// warp items (one warp sorts and joins the whole tile), the others CTA items.
template <int TW>
__global__ void __launch_bounds__(256) pp_tile_plan_kernel(
    int s0, int NT1, const int* __restrict__ tile_nq, const int* __restrict__ tile_hcnt, const int2* __restrict__ tile_seg,
    const int32_t* __restrict__ trav_off, long long* __restrict__ tile_boff, PPItemT<TW>* __restrict__ items, int item_cap,
    PPGroupCtr* __restrict__ ctr) {
  const int s = s0 + blockIdx.y;
  const int NT = NT1 * NT1;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int q = 0, h = 0;
  if (t < NT) {
    q = tile_nq[(size_t)s * NT + t];
    if (q > 0) h = tile_hcnt[(size_t)s * NT + t];
  }
  const bool warp_item = h <= kWarpMaxRecords && q <= kWarpMaxQueries;
  const int cap = warp_item ? kQCapWarp : (h >= kHeavyBin ? kQCapHeavy : kQCapCta);
  const int n_it = q > 0 ? (q + cap - 1) / cap : 0;
  int ih = h, iC = warp_item ? 0 : n_it, iW = warp_item ? n_it : 0;
  const int sh = ih, sC = iC, sW = iW;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int uh = __shfl_up_sync(0xffffffffu, ih, o), uC = __shfl_up_sync(0xffffffffu, iC, o),
              uW = __shfl_up_sync(0xffffffffu, iW, o);
    if (lane >= o) { ih += uh; iC += uC; iW += uW; }
  }
  long long bh = 0;
  int bC = 0, bW = 0;
  if (lane == 31) {
    if (ih) bh = (long long)atomicAdd(&ctr->total, (unsigned long long)ih);
    if (iC) bC = atomicAdd(&ctr->n_cta, iC);
    if (iW) bW = atomicAdd(&ctr->n_warp, iW);
  }
  bh = __shfl_sync(0xffffffffu, bh, 31) + ih - sh;
  bC = __shfl_sync(0xffffffffu, bC, 31) + iC - sC;
  bW = __shfl_sync(0xffffffffu, bW, 31) + iW - sW;
  if (q <= 0) return;
  tile_boff[(size_t)s * NT + t] = bh;
  const int2* seg = tile_seg + ((size_t)s * NT + t) * TW;
  PPItemT<TW> it;
  it.scan = s; it.ty = (short)(t / NT1); it.tx = (short)(t - (t / NT1) * NT1);
  it.boff = bh; it.bcnt = h; it.pad = 0;
  it.logT = log((double)(trav_off[s + 1] - trav_off[s]));
#pragma unroll
  for (int r = 0; r < TW; ++r) { const int2 sg = seg[r]; it.seg_start[r] = sg.x; it.seg_len[r] = sg.y; }
  for (int k0 = 0; k0 < q; k0 += cap) {
    const int idx = warp_item ? item_cap - 1 - (bW++) : bC++;
    if (idx < 0 || idx >= item_cap) continue;                        // cannot happen: capacity is an upper bound
    it.k0 = k0; it.nq = min(cap, q - k0);
    items[idx] = it;
  }
}

// K5: the join.  Persistent blocks of NTHR threads (256: CTA items from the front of the item
// array, 32: warp items from its back); dynamic shared memory = chunk | cell table | counters.
template <int TW, int NTHR, int CHUNK>
__global__ void __launch_bounds__(NTHR, NTHR == 32 ? 16 : 3) pp_join_kernel(
    const PPItemT<TW>* __restrict__ items, int item_cap, PPGroupCtr* __restrict__ ctr, const float4* __restrict__ sorted,
    const int64_t* __restrict__ q_off, const PPMeta* __restrict__ meta, const int32_t* __restrict__ trav_off,
    const int64_t* __restrict__ count_off, const float4* __restrict__ bins, int G, float r2f, float band, double r2,
    float* __restrict__ pp, int32_t* __restrict__ counts_out) {
  using Geo = TileGeo<TW, NTHR>;
  constexpr bool kWarp = NTHR == 32;
  constexpr int kTermSlots = CHUNK * 2;                // doubles that fit the chunk buffer
  constexpr int WW = Geo::WW;
  extern __shared__ __align__(16) unsigned char s_dyn[];             // chunk | cell table | counters
  float4* s_h = reinterpret_cast<float4*>(s_dyn);
  unsigned* s_tbl = reinterpret_cast<unsigned*>(s_dyn + sizeof(float4) * CHUNK);
  int* s_cnt = reinterpret_cast<int*>(s_dyn + sizeof(float4) * CHUNK + sizeof(unsigned) * Geo::kTblWords);   // [traversal][query]
  __shared__ int s_warp[8];
  __shared__ int s_work, s_zlo, s_zhi;
  __shared__ int s_qorig[NTHR];
  __shared__ PPItemT<TW> s_item;
  const unsigned short* s_start = reinterpret_cast<const unsigned short*>(s_tbl);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float lo = r2f - band, hi = r2f + band;
  const int n_items = kWarp ? ctr->n_warp : ctr->n_cta;
  while (true) {
    if (tid == 0) { s_work = atomicAdd(kWarp ? &ctr->next_warp : &ctr->next_cta, 1); s_zlo = kZCells; s_zhi = -1; }
    __syncthreads();
    const int wk = s_work;
    if (wk >= n_items) break;
    const PPItemT<TW>* gi = items + (kWarp ? item_cap - 1 - wk : wk);
    if (tid < (int)(sizeof(PPItemT<TW>) / 4)) reinterpret_cast<int*>(&s_item)[tid] = reinterpret_cast<const int*>(gi)[tid];
    __syncthreads();
    const int s = s_item.scan, nq = s_item.nq;
    const PPMeta m = meta[s];
    const int T = trav_off[s + 1] - trav_off[s];
    const int wx0 = s_item.tx * TW - 1, wy0 = s_item.ty * TW - 1;     // window origin in cells
    const long long boff = s_item.boff;
    const int bcnt = s_item.bcnt;
    // lanes per query: tiles with few query points spread a query's candidates over 2..32 lanes
    int lsh = 0;
    while (lsh < 5 && (2 << lsh) * nq <= NTHR) ++lsh;
    const int lpq = 1 << lsh;
    const int qi = tid >> lsh, sub = tid & (lpq - 1);
    const bool has_q = qi < nq;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    int qcol = 0, qz = 0;
    if (has_q) {
      int k = s_item.k0 + qi, r = 0;
      while (r < TW - 1 && k >= s_item.seg_len[r]) { k -= s_item.seg_len[r]; ++r; }
      q = __ldg(sorted + q_off[s] + s_item.seg_start[r] + k);
      const int cx = clampi(cell_coord(q.x, m.x0, m.inv_cell), 0, G - 1);
      const int cy = clampi(cell_coord(q.y, m.y0, m.inv_cell), 0, G - 1);
      qz = clampi(cell_coord(q.z, m.z0, m.inv_cell), 0, kZCells - 1);
      qcol = (cy - wy0) * WW + (cx - wx0);                          // an interior column of the window
      if (sub == 0) {
        s_qorig[qi] = __float_as_int(q.w);
        atomicMin(&s_zlo, qz);
        atomicMax(&s_zhi, qz);
      }
    }
    for (int t = 0; t < T; ++t) s_cnt[t * NTHR + tid] = 0;
    __syncthreads();
    // z-cells a record must lie in to matter to one of these queries
    const int zlo = max(s_zlo - 1, 0), zhi = min(s_zhi + 1, kZCells - 1);
    const int nz = zhi - zlo + 1;
    const bool full_col = nz <= kFullColumnZ;                        // short columns are taken whole: 3 ranges per query
    const int ncell = WW * WW * nz;
    const int wpt = (((ncell + 2) / 2 + NTHR - 1) / NTHR) | 1;   // table words per thread (odd: no bank conflicts)
    const int za = max(qz - 1, zlo) - zlo, zb1 = min(qz + 1, zhi) + 1 - zlo;
    int* cntq = s_cnt + qi;

    for (int c0 = 0; c0 < bcnt; c0 += CHUNK) {
      const int n = min(CHUNK, bcnt - c0);
      // ---- counting sort of the chunk by fine cell, in shared memory ----
      for (int i = tid; i < wpt * NTHR; i += NTHR) s_tbl[i] = 0u;
      __syncthreads();
      float4 h[kPtsPerThreadJ];
      int cr[kPtsPerThreadJ];                                        // cell | rank << 14
#pragma unroll
      for (int j = 0; j < kPtsPerThreadJ; ++j) {
        const int i = tid + j * NTHR;
        if (i < n) h[j] = __ldg(bins + boff + c0 + i);
      }
#pragma unroll
      for (int j = 0; j < kPtsPerThreadJ; ++j) {
        const int i = tid + j * NTHR;
        cr[j] = -1;
        if (i < n) {
          const int cz = clampi(cell_coord(h[j].z, m.z0, m.inv_cell), 0, kZCells - 1);
          if (cz >= zlo && cz <= zhi) {
            const int cx = clampi(cell_coord(h[j].x, m.x0, m.inv_cell), 0, G - 1);
            const int cy = clampi(cell_coord(h[j].y, m.y0, m.inv_cell), 0, G - 1);
            const int wx = clampi(cx - wx0, 0, WW - 1), wy = clampi(cy - wy0, 0, WW - 1);
            const int cell = (wy * WW + wx) * nz + (cz - zlo);
            const unsigned old = atomicAdd(&s_tbl[cell >> 1], (cell & 1) ? 0x10000u : 1u);
            cr[j] = cell | (int)(((cell & 1) ? (old >> 16) : (old & 0xffffu)) << 14);
          }
        }
      }
      __syncthreads();
      {   // exclusive scan of the u16 cell counts (two per word), wpt consecutive words per thread
        int sum = 0;
        for (int i = 0; i < wpt; ++i) {
          const unsigned v = s_tbl[tid * wpt + i];
          sum += (int)(v & 0xffffu) + (int)(v >> 16);
        }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int u = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += u;
        }
        int run = incl - sum;
        if (!kWarp) {
          if (lane == 31) s_warp[wid] = incl;
          __syncthreads();
          for (int k = 0; k < wid; ++k) run += s_warp[k];
        }
        for (int i = 0; i < wpt; ++i) {
          const unsigned v = s_tbl[tid * wpt + i];
          const unsigned c_lo = v & 0xffffu, c_hi = v >> 16;
          s_tbl[tid * wpt + i] = (unsigned)run | ((unsigned)(run + c_lo) << 16);
          run += (int)(c_lo + c_hi);
        }
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < kPtsPerThreadJ; ++j)
        if (cr[j] >= 0) s_h[s_start[cr[j] & 0x3fff] + (cr[j] >> 14)] = h[j];
      __syncthreads();
      // ---- every query walks the cells around it; its lanes interleave over the concatenation ----
      if (has_q) {
        int skip = sub;
        auto walk = [&](int kb, int ke) {
          int c = kb + skip;
          while (c < ke) {
            const float4 p = s_h[c];
            const float ddx = q.x - p.x, ddy = q.y - p.y, ddz = q.z - p.z;
            const float d2 = fmaf(ddz, ddz, fmaf(ddy, ddy, ddx * ddx));
            if (d2 <= hi && (d2 < lo || sqdist_f64_seq(q.x, q.y, q.z, p.x, p.y, p.z) <= r2))
              atomicAdd(cntq + (__float_as_int(p.w) >> kTagShift) * NTHR, 1);
            c += lpq;
          }
          skip = c - ke;
        };
        if (full_col) {
#pragma unroll
          for (int dy = -1; dy <= 1; ++dy) {
            const int b0 = (qcol + dy * WW - 1) * nz;
            walk(s_start[b0], s_start[b0 + 3 * nz]);
          }
        } else {
#pragma unroll
          for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
              const int b0 = (qcol + dy * WW + dx) * nz;
              walk(s_start[b0 + za], s_start[b0 + zb1]);
            }
        }
      }
      __syncthreads();                                               // table and chunk are rewritten next pass
    }
    // ---- counts complete: entropy straight from shared memory ----
    const double logT = s_item.logT;
    if (nq * T <= kTermSlots) {
      // one (query, traversal) term per thread and trip: all lanes busy with the f64 log
      double* s_denom = reinterpret_cast<double*>(s_tbl);
      double* s_term = reinterpret_cast<double*>(s_h);
      if (tid < nq) {
        long long tot = 0;
        for (int t = 0; t < T; ++t) tot += s_cnt[t * NTHR + tid];
        s_denom[tid] = __dadd_rn((double)tot, 1e-8);
      }
      __syncthreads();
      for (int p = tid; p < nq * T; p += NTHR) {
        const int t = p / nq, k = p - t * nq;
        const int ct = s_cnt[t * NTHR + k];
        double term = 0.0;                                           // -0.0 * ln(1e-8) is exactly +0.0
        if (ct != 0) {
          const double P = __ddiv_rn((double)ct, s_denom[k]);
          term = __dmul_rn(-P, log(__dadd_rn(P, 1e-8)));
        }
        s_term[p] = term;
      }
      __syncthreads();
      if (tid < nq) {
        double acc = 0.0;
        for (int t = 0; t < T; ++t) acc = __dadd_rn(acc, s_term[t * nq + tid]);
        pp[q_off[s] + s_qorig[tid]] = (float)__ddiv_rn(acc, logT);
      }
    } else if (tid < nq) {
      pp[q_off[s] + s_qorig[tid]] = pp_entropy_of(T, logT, [&](int t) { return s_cnt[t * NTHR + tid]; });
    }
    if (counts_out && tid < nq)
      for (int t = 0; t < T; ++t)
        counts_out[count_off[s] + (size_t)s_qorig[tid] * T + t] = s_cnt[t * NTHR + tid];
    __syncthreads();                                                 // s_item / s_cnt / s_tbl are rewritten by the next item
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


// Stage B for the streaming engine: the raw frames stay where they are (a device
// cache of velodyne/*.bin contents, one allocation per frame) and are addressed by pointer; every
// job (source frame, 4x4, destination row) writes xyz rows of `out_stride` floats -- 3 for the PP
// inputs (query / history in the fixed frame), 4 to assemble a batch's raw scans (identity, the
// intensity column copied).  Frame f of the launch is blockIdx.y.
struct FrameJob {            // 96 bytes, built on the host
  const float* src;          // (n, src_stride) f32 rows [x,y,z,...]
  long long dst_row;         // first output row
  int n, flags;              // flags: 1 = remove_center, 2 = copy rows unchanged (no transform)
  float T[16];               // row-major 4x4 (float32, as get_relative_pose returns it)
  long long pad;
};

__global__ void __launch_bounds__(256) transform_gather_kernel(
    const FrameJob* __restrict__ jobs, int src_stride, int out_stride, float cx0, float cx1, float cy0, float cy1,
    float* __restrict__ out) {
  const FrameJob jb = jobs[blockIdx.y];
  const float nanv = __int_as_float(0x7fc00000);
  const bool copy = jb.flags & 2, rc = jb.flags & 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < jb.n; i += gridDim.x * blockDim.x) {
    const float* p = jb.src + (size_t)src_stride * i;
    float* o = out + (size_t)out_stride * (jb.dst_row + i);
    const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
    if (copy) {
      o[0] = x; o[1] = y; o[2] = z;
      for (int k = 3; k < out_stride; ++k) o[k] = k < src_stride ? __ldg(p + k) : 0.f;
      continue;
    }
    if (rc && x < cx1 && x >= cx0 && y < cy1 && y >= cy0) {
      o[0] = nanv; o[1] = nanv; o[2] = nanv;
      continue;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j)     // same rounding as transform_frames_kernel (sgemm's k-ordered FMAs)
      o[j] = fmaf(1.0f, jb.T[4 * j + 3], fmaf(z, jb.T[4 * j + 2], fmaf(y, jb.T[4 * j + 1], __fmul_rn(x, jb.T[4 * j]))));
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
static const int kMaxTraversals = 1 << 16;

static int col_tiles(int G) { return (int)(((size_t)G * G + kColTile - 1) / kColTile); }

// ---- grouping of the tiled path (host) -------------------------------------------------------------
// Consecutive scans are processed together while their history stays below `group_points`
// points, so that a group's bins (written by the scatter, read by the join) stay L2-resident.
// A history point is copied into at most 4 tiles, which bounds a group's bin records.
static const int64_t kDefaultGroupPoints = 2 * 1000 * 1000;
static const int kMaxBinCopies = 4;

static int64_t scan_hist_points(const int64_t* h_h_off, const int32_t* h_trav_off, int s) {
  return h_h_off[h_trav_off[s + 1]] - h_h_off[h_trav_off[s]];
}
// end (exclusive) of the group that starts at scan s0
static int group_end(const int64_t* h_h_off, const int32_t* h_trav_off, int n_scans, int s0, int64_t group_points) {
  int64_t pts = 0;
  int s = s0;
  while (s < n_scans) {
    const int64_t p = scan_hist_points(h_h_off, h_trav_off, s);
    if (s > s0 && pts + p > group_points) break;
    pts += p;
    ++s;
  }
  return s;
}

static size_t pp_item_capacity(int n_scans, int64_t n_query_total, int NT) {
  return (size_t)n_scans * NT + (size_t)(n_query_total / kQCapWarp) + (size_t)n_scans + 1;
}

constexpr int kTWMin = 8, kTWMax = 8;             // tile edge in cells (the kernels are templates on it)

struct PPTiledWs {
  int* tile_nq;                 // [3][n_scans][NT]: query points, history records, scatter cursor
  long long* tile_boff;
  int2* tile_seg;
  void* items;
  PPGroupCtr* ctrs;
  float4* bins;
};
struct PPTiledArgs {
  const float* d_hist_xyz; const int64_t* d_h_off; const int32_t* d_trav_off; const int64_t* d_q_off;
  const int64_t* d_count_off; float* d_pp; int32_t* d_counts;
  const int64_t* h_q_off; const int64_t* h_h_off; const int32_t* h_trav_off;
  int n_scans; int64_t n_query_total; int64_t group_points; int64_t bin_records; int t_max; int G;
  float r2f, band; double r2;
  const PPMeta* meta; const int2* cols; const int* zc; const float4* sorted; const int32_t* trav_scan;
};

template <int TW>
static int pp_tiled_run(const PPTiledArgs& a, const PPTiledWs& w, cudaStream_t stream, int* launched) {
  const int G = a.G, NT1 = G / TW, NT = NT1 * NT1, n_scans = a.n_scans;
  int* tile_nq = w.tile_nq;
  int* tile_hcnt = tile_nq + (size_t)n_scans * NT;
  int* tile_fill = tile_hcnt + (size_t)n_scans * NT;
  PPItemT<TW>* items = static_cast<PPItemT<TW>*>(w.items);
  const size_t cta_fixed = sizeof(float4) * kChunk + sizeof(unsigned) * TileGeo<TW, kJoinThreads>::kTblWords;
  const size_t warp_fixed = sizeof(float4) * kWarpChunk + sizeof(unsigned) * TileGeo<TW, kWarpThreads>::kTblWords;
  static bool attr_set = false;
  if (!attr_set) {
    MODEST_CUDA(cudaFuncSetAttribute(pp_join_kernel<TW, kJoinThreads, kChunk>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(cta_fixed + kJoinMaxT * kJoinThreads * sizeof(int))));
    MODEST_CUDA(cudaFuncSetAttribute(pp_join_kernel<TW, kWarpThreads, kWarpChunk>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(warp_fixed + kJoinMaxT * kWarpThreads * sizeof(int))));
    attr_set = true;
  }
  MODEST_CUDA(cudaMemsetAsync(tile_hcnt, 0, sizeof(int) * (size_t)n_scans * NT * 2, stream));   // + tile_fill
  MODEST_CUDA(cudaMemsetAsync(w.ctrs, 0, sizeof(PPGroupCtr) * (size_t)n_scans, stream));
  pp_tile_query_kernel<TW><<<dim3((NT + 255) / 256, n_scans), 256, 0, stream>>>(a.cols, a.zc, a.d_q_off, a.meta, G, NT1, tile_nq,
                                                                               w.tile_seg);
  MODEST_LAUNCH_CHECK("pp_tile_query_kernel");
  ++*launched;
  const int join_ctas = sm_count() * 3, join_warps = sm_count() * 16;
  const size_t cta_smem = cta_fixed + (size_t)std::max(a.t_max, 1) * kJoinThreads * sizeof(int);
  const size_t warp_smem = warp_fixed + (size_t)std::max(a.t_max, 1) * kWarpThreads * sizeof(int);
  int grp = 0;
  for (int s0 = 0; s0 < n_scans; ++grp) {
    const int s1 = group_end(a.h_h_off, a.h_trav_off, n_scans, s0, a.group_points);
    const int g0 = a.h_trav_off[s0], g1 = a.h_trav_off[s1];
    const int64_t pts = a.h_h_off[g1] - a.h_h_off[g0];
    MODEST_REQUIRE(kMaxBinCopies * pts + 16 <= a.bin_records,
                   "pp_score: bin_records %lld too small for a group of %lld history points (use modest_pp_bin_records)",
                   (long long)a.bin_records, (long long)pts);
    int64_t gmax = 0, gq = 0;
    for (int g = g0; g < g1; ++g) gmax = std::max(gmax, (int64_t)(a.h_h_off[g + 1] - a.h_h_off[g]));
    for (int s = s0; s < s1; ++s) gq += a.h_q_off[s + 1] - a.h_q_off[s];
    const int item_cap = (int)pp_item_capacity(s1 - s0, gq, NT);
    int64_t hb = std::min<int64_t>(std::max<int64_t>((gmax + 255) / 256, 1), 4096);
    const dim3 hgrid((unsigned)hb, std::max(g1 - g0, 1));
    const bool any_hist = g1 > g0 && pts > 0;
    if (any_hist) {
      pp_hist_tile_kernel<TW, false><<<hgrid, 256, 0, stream>>>(a.d_hist_xyz, a.d_h_off, a.trav_scan, a.d_trav_off, g0, a.meta,
                                                               tile_nq, tile_hcnt, nullptr, nullptr, G, NT1);
      MODEST_LAUNCH_CHECK("pp_hist_tile_kernel<count>");
      ++*launched;
    }
    pp_tile_plan_kernel<TW><<<dim3((NT + 255) / 256, s1 - s0), 256, 0, stream>>>(s0, NT1, tile_nq, tile_hcnt, w.tile_seg, a.d_trav_off,
                                                                                w.tile_boff, items, item_cap, w.ctrs + grp);
    MODEST_LAUNCH_CHECK("pp_tile_plan_kernel");
    if (any_hist) {
      pp_hist_tile_kernel<TW, true><<<hgrid, 256, 0, stream>>>(a.d_hist_xyz, a.d_h_off, a.trav_scan, a.d_trav_off, g0, a.meta,
                                                              tile_nq, tile_fill, w.tile_boff, w.bins, G, NT1);
      MODEST_LAUNCH_CHECK("pp_hist_tile_kernel<scatter>");
      ++*launched;
    }
    // scans without any history still get their (zero) scores from the join; the big tiles
    // (CTA items) start first, the many small ones (warp items) fill in around them
    pp_join_kernel<TW, kJoinThreads, kChunk><<<join_ctas, kJoinThreads, cta_smem, stream>>>(
        items, item_cap, w.ctrs + grp, a.sorted, a.d_q_off, a.meta, a.d_trav_off, a.d_count_off, w.bins, G, a.r2f, a.band, a.r2,
        a.d_pp, a.d_counts);
    MODEST_LAUNCH_CHECK("pp_join_kernel<cta>");
    pp_join_kernel<TW, kWarpThreads, kWarpChunk><<<join_warps, kWarpThreads, warp_smem, stream>>>(
        items, item_cap, w.ctrs + grp, a.sorted, a.d_q_off, a.meta, a.d_trav_off, a.d_count_off, w.bins, G, a.r2f, a.band, a.r2,
        a.d_pp, a.d_counts);
    MODEST_LAUNCH_CHECK("pp_join_kernel<warp>");
    *launched += 3;
    s0 = s1;
  }
  return MODEST_OK;
}

extern "C" int64_t modest_pp_bin_records(const int64_t* h_h_off, const int32_t* h_trav_off, int n_scans,
                                         int64_t group_points) {
  if (!h_h_off || !h_trav_off || n_scans <= 0) return 0;
  if (group_points <= 0) group_points = kDefaultGroupPoints;
  int64_t worst = 0;
  for (int s0 = 0; s0 < n_scans;) {
    const int s1 = group_end(h_h_off, h_trav_off, n_scans, s0, group_points);
    const int64_t pts = h_h_off[h_trav_off[s1]] - h_h_off[h_trav_off[s0]];
    if (pts > worst) worst = pts;
    s0 = s1;
  }
  return kMaxBinCopies * worst + 16;
}



extern "C" size_t modest_pp_workspace_bytes(int n_scans, int64_t n_query_total, int64_t n_count_total, int grid_dim,
                                            int64_t bin_records) {
  if (grid_dim <= 0) grid_dim = 512;
  size_t b = 0;
  auto add = [&](size_t bytes) { b = align_up(b, 256) + bytes; };
  add(sizeof(PPMeta) * (size_t)n_scans);
  add(sizeof(int2) * (size_t)n_scans * grid_dim * grid_dim);                 // column records
  add(sizeof(int) * ((size_t)n_query_total + 8 * (size_t)n_scans + 8));      // compact cell starts
  add(sizeof(int) * (size_t)n_scans * col_tiles(grid_dim));                  // tile sums
  add(sizeof(float4) * (size_t)n_query_total);                               // sorted query
  add(sizeof(int32_t) * (size_t)kMaxTraversals);                             // traversal -> scan
  add(sizeof(double) * (size_t)kEntTabLen);                                  // entropy terms of small (total, count) pairs
  // legacy path: counts [t][pos]; tiled path: the bins -- never both in one call
  const size_t legacy = sizeof(int) * (size_t)n_count_total, tiled = sizeof(float4) * (size_t)(bin_records > 0 ? bin_records : 0);
  add(legacy > tiled ? legacy : tiled);
  if (bin_records > 0 && grid_dim % kTWMax == 0) {
    const int NT = (grid_dim / kTWMin) * (grid_dim / kTWMin);
    add(sizeof(int) * (size_t)n_scans * NT * 3);                             // tile_nq, tile_hcnt, tile_fill
    add(sizeof(long long) * (size_t)n_scans * NT);                           // tile_boff
    add(sizeof(int2) * (size_t)n_scans * NT * kTWMin);                       // tile_seg
    add(sizeof(PPItemT<kTWMax>) * pp_item_capacity(n_scans, n_query_total, NT));
    add(sizeof(PPGroupCtr) * (size_t)n_scans);
  }
  return b + 256;
}

extern "C" int modest_pp_score_batch(const float* d_query_xyz, const int64_t* d_q_off,
                                     const float* d_hist_xyz, const int64_t* d_h_off,
                                     const int32_t* d_trav_off, int n_scans, int n_trav_total,
                                     int64_t n_query_total, int64_t n_count_total,
                                     int64_t max_query_points, int64_t max_trav_points, double radius,
                                     int grid_dim, int32_t* d_counts, const int64_t* d_count_off,
                                     float* d_pp, const int64_t* h_q_off, const int64_t* h_h_off,
                                     const int32_t* h_trav_off, int64_t group_points, int64_t bin_records,
                                     void* d_ws, size_t ws_bytes, void* stream_) {
  modest::StageRange nvtx_("modest:C,D PP score");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (grid_dim <= 0) grid_dim = 512;
  MODEST_REQUIRE(n_scans >= 0 && n_trav_total >= 0, "pp_score: negative sizes");
  if (n_scans == 0 || n_query_total == 0) return MODEST_OK;
  MODEST_REQUIRE(d_query_xyz && d_q_off && d_h_off && d_trav_off && d_count_off && d_pp && d_ws,
                 "pp_score: null pointer argument");
  MODEST_REQUIRE(n_trav_total < kMaxTraversals, "pp_score: more than 65535 traversals in one launch");
  MODEST_REQUIRE(n_scans <= 65535, "pp_score: more than 65535 scans in one launch");
  MODEST_REQUIRE(radius > 0.0 && radius < 1e3, "pp_score: radius %g out of range", radius);
  MODEST_REQUIRE(grid_dim >= 8 && grid_dim <= 4096 && grid_dim % 4 == 0,
                 "pp_score: grid_dim %d must be a multiple of 4 in [8,4096]", grid_dim);
  MODEST_REQUIRE(ws_bytes >= modest_pp_workspace_bytes(n_scans, n_query_total, n_count_total, grid_dim, bin_records),
                 "pp_score: workspace too small (%zu bytes given)", ws_bytes);
  MODEST_REQUIRE(max_query_points < (1ll << 30), "pp_score: a scan has >= 2^30 points");

  // which history pass: the tiled one needs the host copies of the offset tables (to cut the
  // batch into groups) and at most kJoinMaxT traversals per scan; everything else takes the
  // global-hash pass (group_points < 0 forces it)
  bool tiled = h_q_off && h_h_off && h_trav_off && group_points >= 0 && bin_records > 0 && grid_dim % kTWMax == 0 &&
               n_trav_total > 0 && max_trav_points > 0;
  int t_max = 0;
  if (tiled) {
    for (int s = 0; s < n_scans; ++s) t_max = std::max(t_max, (int)(h_trav_off[s + 1] - h_trav_off[s]));
    tiled = t_max <= kJoinMaxT;
  }
  if (tiled && group_points == 0) group_points = kDefaultGroupPoints;

  const int G = grid_dim;
  const size_t ncol = (size_t)G * G;
  const int tiles = col_tiles(G);
  Arena ar(d_ws, ws_bytes);
  PPMeta* meta = ar.take<PPMeta>(n_scans);
  int2* cols = ar.take<int2>((size_t)n_scans * ncol);
  const size_t zc_len = (size_t)n_query_total + 8 * (size_t)n_scans + 8;
  int* zc = ar.take<int>(zc_len);
  int* tile_sums = ar.take<int>((size_t)n_scans * tiles);
  float4* sorted = ar.take<float4>(n_query_total);
  int32_t* trav_scan = ar.take<int32_t>(kMaxTraversals);
  double* ent_tab = ar.take<double>(kEntTabLen);
  const size_t legacy_b = sizeof(int) * (size_t)n_count_total, tiled_b = sizeof(float4) * (size_t)(bin_records > 0 ? bin_records : 0);
  char* big = ar.take<char>(legacy_b > tiled_b ? legacy_b : tiled_b);
  int* counts = reinterpret_cast<int*>(big);
  float4* bins = reinterpret_cast<float4*>(big);
  PPTiledWs tw = {};
  if (tiled) {
    const int NTs = (G / kTWMin) * (G / kTWMin);                              // sized for the finer tiling
    tw.tile_nq = ar.take<int>((size_t)n_scans * NTs * 3);
    tw.tile_boff = ar.take<long long>((size_t)n_scans * NTs);
    tw.tile_seg = ar.take<int2>((size_t)n_scans * NTs * kTWMin);
    tw.items = ar.take<PPItemT<kTWMax>>(pp_item_capacity(n_scans, n_query_total, NTs));
    tw.ctrs = ar.take<PPGroupCtr>(n_scans);
    tw.bins = bins;
  }
  MODEST_REQUIRE(ar.ok(), "workspace too small for the requested sizes");

  const float cell = (float)radius * kCellSlack;
  const double r2 = radius * radius;
  const float r2f = (float)r2;
  const float band = 1e-5f * r2f;   // ~100x the f32 evaluation error of d2 for d2 ~ r2

  MODEST_CUDA(cudaMemsetAsync(cols, 0, sizeof(int2) * (size_t)n_scans * ncol, stream));
  MODEST_CUDA(cudaMemsetAsync(zc, 0, sizeof(int) * zc_len, stream));
  if (!tiled) MODEST_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)n_count_total, stream));

  int qblocks = (int)((max_query_points + 255) / 256);
  if (qblocks < 1) qblocks = 1;
  if (qblocks > 2048) qblocks = 2048;
  const dim3 qgrid(qblocks, n_scans);
  pp_trav_scan_kernel<<<n_scans, 32, 0, stream>>>(d_trav_off, n_scans, trav_scan);
  MODEST_LAUNCH_CHECK("pp_trav_scan_kernel");
  pp_origin_kernel<<<n_scans, 1024, 0, stream>>>(d_query_xyz, d_q_off, meta, G, cell);
  MODEST_LAUNCH_CHECK("pp_origin_kernel");
  pp_mask_kernel<<<qgrid, 256, 0, stream>>>(d_query_xyz, d_q_off, meta, cols, G);
  MODEST_LAUNCH_CHECK("pp_mask_kernel");
  pp_col_tilesum_kernel<<<dim3(tiles, n_scans), 1024, 0, stream>>>(cols, G, tile_sums, tiles);
  MODEST_LAUNCH_CHECK("pp_col_tilesum_kernel");
  pp_col_tilescan_kernel<<<n_scans, 32, 0, stream>>>(tile_sums, tiles, meta);
  MODEST_LAUNCH_CHECK("pp_col_tilescan_kernel");
  pp_col_apply_kernel<<<dim3(tiles, n_scans), 1024, 0, stream>>>(cols, G, tile_sums, tiles);
  MODEST_LAUNCH_CHECK("pp_col_apply_kernel");
  pp_cellcount_kernel<<<qgrid, 256, 0, stream>>>(d_query_xyz, d_q_off, meta, cols, zc, G);
  MODEST_LAUNCH_CHECK("pp_cellcount_kernel");
  pp_cellscan_kernel<<<n_scans, 1024, 0, stream>>>(zc, d_q_off, meta);
  MODEST_LAUNCH_CHECK("pp_cellscan_kernel");
  pp_scatter_kernel<<<qgrid, 256, 0, stream>>>(d_query_xyz, d_q_off, meta, cols, zc, sorted, G);
  MODEST_LAUNCH_CHECK("pp_scatter_kernel");
  int n_launched = 9;
  const int slot = g_prof_slots ? (int)(g_prof_calls % g_prof_slots) : -1;

  if (tiled) {
    if (slot >= 0) cudaEventRecord(g_prof_ev[2 * slot], stream);
    PPTiledArgs ta = {d_hist_xyz, d_h_off, d_trav_off, d_q_off, d_count_off, d_pp, d_counts, h_q_off, h_h_off, h_trav_off,
                      n_scans, n_query_total, group_points, bin_records, t_max, G, r2f, band, r2, meta, cols, zc, sorted, trav_scan};
    int launched = 0;
    const int rc = pp_tiled_run<kTWMin>(ta, tw, stream, &launched);
    if (rc != MODEST_OK) return rc;
    if (slot >= 0) { cudaEventRecord(g_prof_ev[2 * slot + 1], stream); ++g_prof_calls; }
    note_launch(n_launched + launched);
    return MODEST_OK;
  }

  if (n_trav_total > 0 && max_trav_points > 0) {
    int64_t hb = (max_trav_points + 255) / 256;
    if (hb > 65535) hb = 65535;
    const dim3 hgrid((unsigned)hb, n_trav_total);   // 8 warps x 32 points per CTA per trip
    if (slot >= 0) cudaEventRecord(g_prof_ev[2 * slot], stream);
    pp_count_kernel<<<hgrid, 256, 0, stream>>>(d_hist_xyz, d_h_off, trav_scan, d_trav_off, d_q_off, d_count_off, meta,
                                               cols, zc, sorted, counts, G, r2f, band, r2);
    MODEST_LAUNCH_CHECK("pp_count_kernel");
    if (slot >= 0) { cudaEventRecord(g_prof_ev[2 * slot + 1], stream); ++g_prof_calls; }
    ++n_launched;
  }
  pp_entropy_table_kernel<<<(kEntTabLen + 255) / 256, 256, 0, stream>>>(ent_tab);
  MODEST_LAUNCH_CHECK("pp_entropy_table_kernel");
  pp_entropy_kernel<<<qgrid, 256, 0, stream>>>(counts, sorted, d_q_off, d_count_off, d_trav_off, d_pp, d_counts, ent_tab);
  MODEST_LAUNCH_CHECK("pp_entropy_kernel");
  note_launch(n_launched + 2);
  return MODEST_OK;
}

extern "C" int modest_transform_frames_batch(const float* d_in, int point_stride, const int64_t* d_frame_off,
                                             const float* d_T, int n_frames, int64_t max_frame_points,
                                             int remove_center, const float* h_center_box, float* d_out,
                                             void* stream_) {
  modest::StageRange nvtx_("modest:B frame transform");
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

extern "C" int modest_transform_gather_batch(const void* d_jobs, int n_jobs, int src_stride, int out_stride,
                                             int64_t max_frame_points, const float* h_center_box, float* d_out,
                                             void* stream_) {
  modest::StageRange nvtx_("modest:B frame transform (gather)");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_jobs <= 0 || max_frame_points <= 0) return MODEST_OK;
  static_assert(sizeof(FrameJob) == 96, "FrameJob is part of the ABI (96 bytes)");
  MODEST_REQUIRE(d_jobs && d_out, "transform_gather: null pointer argument");
  MODEST_REQUIRE(src_stride >= 3 && out_stride >= 3 && out_stride <= 8, "transform_gather: strides %d / %d", src_stride, out_stride);
  int done = 0;
  const float* c = h_center_box;
  int64_t blocks = (max_frame_points + 255) / 256;
  if (blocks > 64) blocks = 64;                      // a frame is ~60k points: 64 CTAs x 4 trips; the grid is wide in y
  while (done < n_jobs) {                            // gridDim.y <= 65535
    const int n = std::min(n_jobs - done, 65535);
    transform_gather_kernel<<<dim3((unsigned)blocks, n), 256, 0, stream>>>(
        static_cast<const FrameJob*>(d_jobs) + done, src_stride, out_stride, c ? c[0] : 0.f, c ? c[1] : 0.f, c ? c[2] : 0.f,
        c ? c[3] : 0.f, d_out);
    MODEST_LAUNCH_CHECK("transform_gather_kernel");
    note_launch(1);
    done += n;
  }
  return MODEST_OK;
}
