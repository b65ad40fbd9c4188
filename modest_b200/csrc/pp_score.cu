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

// H of one query point from its T neighbour counts (compute_ephe_score, :68-75)
template <typename C>
__device__ __forceinline__ float pp_entropy_of(int T, double logT, C count_of) {
  long long tot = 0;
  for (int t = 0; t < T; ++t) tot += count_of(t);
  const double denom = __dadd_rn((double)tot, 1e-8);
  const double acc = numpy_strided_sum(T, [&](int t) {
    const int ct = count_of(t);
    if (ct == 0) return 0.0;                         // -0.0 * ln(1e-8) is exactly +0.0: skip the log
    const double P = __ddiv_rn((double)ct, denom);
    return __dmul_rn(-P, log(__dadd_rn(P, 1e-8)));
  });
  return (float)__ddiv_rn(acc, logT);
}

__global__ void __launch_bounds__(256) pp_entropy_kernel(
    const int* __restrict__ counts, const float4* __restrict__ sorted, const int64_t* __restrict__ q_off,
    const int64_t* __restrict__ count_off, const int32_t* __restrict__ trav_off, float* __restrict__ pp,
    int32_t* __restrict__ counts_out /* (N,T) row-major per scan at count_off, or NULL */) {
  const int s = blockIdx.y;
  const int T = trav_off[s + 1] - trav_off[s];
  const int64_t qbeg = q_off[s];
  const int n = (int)(q_off[s + 1] - qbeg);
  const int* __restrict__ c = counts + count_off[s];
  const double logT = log((double)T);
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int orig = __float_as_int(sorted[qbeg + k].w);
    pp[qbeg + orig] = pp_entropy_of(T, logT, [&](int t) { return c[(size_t)t * n + k]; });
    if (counts_out)
      for (int t = 0; t < T; ++t) counts_out[count_off[s] + (size_t)orig * T + t] = c[(size_t)t * n + k];
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
static const int kMaxTraversals = 1 << 16;

static int col_tiles(int G) { return (int)(((size_t)G * G + kColTile - 1) / kColTile); }

extern "C" size_t modest_pp_workspace_bytes(int n_scans, int64_t n_query_total, int64_t n_count_total, int grid_dim) {
  if (grid_dim <= 0) grid_dim = 512;
  size_t b = 0;
  auto add = [&](size_t bytes) { b = align_up(b, 256) + bytes; };
  add(sizeof(PPMeta) * (size_t)n_scans);
  add(sizeof(int2) * (size_t)n_scans * grid_dim * grid_dim);                 // column records
  add(sizeof(int) * ((size_t)n_query_total + 8 * (size_t)n_scans + 8));      // compact cell starts
  add(sizeof(int) * (size_t)n_scans * col_tiles(grid_dim));                  // tile sums
  add(sizeof(float4) * (size_t)n_query_total);                               // sorted query
  add(sizeof(int) * (size_t)n_count_total);                                  // counts [t][pos]
  add(sizeof(int32_t) * (size_t)kMaxTraversals);                             // traversal -> scan
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
  MODEST_REQUIRE(n_trav_total < kMaxTraversals, "pp_score: more than 65535 traversals in one launch");
  MODEST_REQUIRE(n_scans <= 65535, "pp_score: more than 65535 scans in one launch");
  MODEST_REQUIRE(radius > 0.0 && radius < 1e3, "pp_score: radius %g out of range", radius);
  MODEST_REQUIRE(grid_dim >= 8 && grid_dim <= 4096 && grid_dim % 4 == 0,
                 "pp_score: grid_dim %d must be a multiple of 4 in [8,4096]", grid_dim);
  MODEST_REQUIRE(ws_bytes >= modest_pp_workspace_bytes(n_scans, n_query_total, n_count_total, grid_dim),
                 "pp_score: workspace too small (%zu bytes given)", ws_bytes);
  MODEST_REQUIRE(max_query_points < (1ll << 30), "pp_score: a scan has >= 2^30 points");

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
  int* counts = ar.take<int>(n_count_total);
  int32_t* trav_scan = ar.take<int32_t>(kMaxTraversals);
  MODEST_REQUIRE(ar.ok(), "workspace too small for the requested sizes");

  const float cell = (float)radius * kCellSlack;
  const double r2 = radius * radius;
  const float r2f = (float)r2;
  const float band = 1e-5f * r2f;   // ~100x the f32 evaluation error of d2 for d2 ~ r2

  MODEST_CUDA(cudaMemsetAsync(cols, 0, sizeof(int2) * (size_t)n_scans * ncol, stream));
  MODEST_CUDA(cudaMemsetAsync(zc, 0, sizeof(int) * zc_len, stream));
  MODEST_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)n_count_total, stream));

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
  if (n_trav_total > 0 && max_trav_points > 0) {
    int64_t hb = (max_trav_points + 255) / 256;
    if (hb > 65535) hb = 65535;
    const dim3 hgrid((unsigned)hb, n_trav_total);   // 8 warps x 32 points per CTA per trip
    const int slot = g_prof_slots ? (int)(g_prof_calls % g_prof_slots) : -1;
    if (slot >= 0) cudaEventRecord(g_prof_ev[2 * slot], stream);
    pp_count_kernel<<<hgrid, 256, 0, stream>>>(d_hist_xyz, d_h_off, trav_scan, d_trav_off, d_q_off, d_count_off, meta,
                                               cols, zc, sorted, counts, G, r2f, band, r2);
    MODEST_LAUNCH_CHECK("pp_count_kernel");
    if (slot >= 0) { cudaEventRecord(g_prof_ev[2 * slot + 1], stream); ++g_prof_calls; }
    ++n_launched;
  }
  pp_entropy_kernel<<<qgrid, 256, 0, stream>>>(counts, sorted, d_q_off, d_count_off, d_trav_off, d_pp, d_counts);
  MODEST_LAUNCH_CHECK("pp_entropy_kernel");
  note_launch(n_launched + 1);
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
