// Stages F-I: ground/range masks, mutual-kNN ^ radius graph with |delta pp| weights, DBSCAN.
//
// Reference behaviour:
//   F  above_plane / distance_to_plane        utils/pointcloud_utils.py:68-81
//   G  limit_range mask                       generate_mask.py:61-65
//   H  precompute_affinity_matrix(..., 'radius_mutual_knn', 'l1', k, radius)
//                                             utils/clustering_utils.py:32-48 (sklearn
//      kneighbors_graph / radius_neighbors_graph: f64 squared distances on f32 coordinates,
//      self excluded, j in kNN_k(i) and i in kNN_k(j) and d2 <= radius^2)
//   I  sklearn DBSCAN(metric='precomputed')   generate_mask.py:77-81; semantics of
//      sklearn/cluster/_dbscan.py:427-463 + _dbscan_inner.pyx (un-vendored dependency).
//
// Data layout (per scan s, rows at off[s] in every per-point array):
//   kept[i]   float4 (x, y, z, pp) of the i-th surviving point, original order preserved
//   kept_idx  index of that point in the scan;  n_kept[s] on the device
//   rk2[i]    f64: squared distance to the k-th nearest other point (+inf when fewer than k
//             other points lie within the radius -- then every one of them is a k-neighbour)
//   knn       (row, k) i32 candidate lists: { j != i : d2(i,j) <= min(rk2[i], radius^2) }
//   nbr / nbr_w / nbr_cnt: the mutual edges of row i and their f32 weights |pp_i - pp_j|
#include "grid2d.cuh"

namespace modest {
extern void note_launch(int n);

// ---- F+G: masks and order-preserving compaction ----------------------------------------------
struct MaskCfg {
  double offset;
  float only_x_lo, only_x_hi, only_y_lo, only_y_hi;   // plane_estimate.range (strict both sides)
  int use_only_range;
  float lim_x_lo, lim_x_hi, lim_y_lo, lim_y_hi;       // limit_range ( lo < v <= hi )
};

__device__ __forceinline__ double plane_distance(float x, float y, float z, const double* pl, double nrm) {
  // ptc(f32) @ plane[:3](f64) + plane[3], then / ||n||   (pointcloud_utils.py:76-81)
  // numpy hands the (N,3) @ (3,) product to OpenBLAS dgemv, whose kernel rounds as
  // fma(z,c, fma(x,a, y*b)) (checked against numpy in tests/test_host_numerics.py)
  double d = __fma_rn((double)z, pl[2], __fma_rn((double)x, pl[0], __dmul_rn((double)y, pl[1])));
  d = __dadd_rn(d, pl[3]);
  return __ddiv_rn(d, nrm);
}
__device__ __forceinline__ double plane_norm(const double* pl) {
  return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(pl[0], pl[0]), __dmul_rn(pl[1], pl[1])), __dmul_rn(pl[2], pl[2])));
}

__global__ void __launch_bounds__(1024) ground_mask_compact_kernel(
    const float* __restrict__ ptc, int stride, const int64_t* __restrict__ off, const float* __restrict__ pp,
    const double* __restrict__ planes, MaskCfg cfg, float4* __restrict__ kept, int32_t* __restrict__ kept_idx,
    int32_t* __restrict__ n_kept, uint8_t* __restrict__ mask_out) {
  const int s = blockIdx.x;
  const int64_t beg = off[s];
  const int n = (int)(off[s + 1] - beg);
  double pl[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) pl[k] = planes[4 * s + k];
  const double nrm = plane_norm(pl);
  __shared__ int warp_cnt[32];
  __shared__ int tile_total;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int base = 0;
  for (int t0 = 0; t0 < n; t0 += 1024) {
    const int i = t0 + threadIdx.x;
    float x = 0, y = 0, z = 0, v = 0;
    bool keep = false;
    if (i < n) {
      const float* p = ptc + (size_t)stride * (beg + i);
      x = p[0]; y = p[1]; z = p[2];
      v = pp[beg + i];
      bool drop = plane_distance(x, y, z, pl, nrm) < cfg.offset;
      if (cfg.use_only_range)
        drop = drop && (x < cfg.only_x_hi) && (x > cfg.only_x_lo) && (y < cfg.only_y_hi) && (y > cfg.only_y_lo);
      const bool in_lim = (x <= cfg.lim_x_hi) && (x > cfg.lim_x_lo) && (y <= cfg.lim_y_hi) && (y > cfg.lim_y_lo);
      keep = !drop && in_lim;
      if (mask_out) mask_out[beg + i] = keep ? 1 : 0;
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
      kept[beg + pos] = make_float4(x, y, z, v);
      kept_idx[beg + pos] = i;
    }
    base += tile_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) n_kept[s] = base;
}

// ---- H.1: k-th neighbour distance and candidate lists, one warp per point ---------------------
constexpr int kBins = 256;
constexpr int kMaxDepth = 6;
// More than k points at or inside the k-th neighbour distance (exact ties, e.g. duplicate or
// zero-filled returns): the row keeps the first k in window order, like a k-nearest query that
// breaks ties by traversal order, and is marked so that the mutual test looks the partner up in
// the row instead of trusting the distance alone (the mutual graph must stay symmetric).
constexpr int kTruncatedRow = 1 << 30;

template <typename T>
struct BinChain {            // nested linear binning of d2: level l keeps bin sel[l] of [lo[l], lo[l]+256/scale[l])
  T lo[kMaxDepth], scale[kMaxDepth];
  int sel[kMaxDepth];
  int depth;
};

template <typename T>
__device__ __forceinline__ int bin_of(T d2, T lo, T scale) {
  const T t = (d2 - lo) * scale;
  int b = (int)t;
  return b < 0 ? 0 : (b > kBins - 1 ? kBins - 1 : b);
}

// true iff d2 passes every closed level of the chain; *last = bin at the open level `depth`
template <typename T>
__device__ __forceinline__ bool chain_bin(const BinChain<T>& c, T d2, int* last) {
  for (int l = 0; l < c.depth; ++l)
    if (bin_of(d2, c.lo[l], c.scale[l]) != c.sel[l]) return false;
  *last = bin_of(d2, c.lo[c.depth], c.scale[c.depth]);
  return true;
}

template <typename F>
__device__ __forceinline__ void for_each_candidate(const float4* __restrict__ sorted, const int* __restrict__ cells,
                                                   int G, int cx, int cy, int L, int self_pos, F f) {
  const int lane = threadIdx.x & 31;
  const int xa = clampi(cx - L, 0, G - 1), xb = clampi(cx + L, 0, G - 1);
  const int ya = clampi(cy - L, 0, G - 1), yb = clampi(cy + L, 0, G - 1);
  // lane r fetches the position range of window row r (windows are at most 32 rows tall), so
  // the row loop below never waits on a dependent load of its own
  int rkb = 0, rke = 0;
  if (ya + lane <= yb) {
    rkb = __ldg(cells + (ya + lane) * G + xa);
    rke = __ldg(cells + (ya + lane) * G + xb + 1);
  }
  for (int y = ya; y <= yb; ++y) {
    const int kb = __shfl_sync(0xffffffffu, rkb, y - ya), ke = __shfl_sync(0xffffffffu, rke, y - ya);
    for (int k0 = kb; k0 < ke; k0 += 32) {
      const int k = k0 + lane;
      const bool live = k < ke && k != self_pos;
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) q = __ldg(sorted + k);
      f(live, q, k);         // called convergently by the whole warp
    }
  }
}

// Exact k-th smallest of the values a `scan` enumerates (those <= R2), by nested 256-bin
// histograms over d2 followed by an exact ranking of the <= 32 members of the final bin.
// `scan(f)` must call f(live, d2, j) convergently for every candidate slot; it may be replayed.
// Returns false when fewer than k_nn values are <= R2 (then *rk2 is untouched).
template <typename T, typename Scan>
__device__ __forceinline__ bool kth_by_histogram(Scan scan, int k_nn, T R2, int* hist, int lane, T* rk2,
                                                 int32_t* flags) {
  BinChain<T> ch;
  ch.depth = 0;
  ch.lo[0] = (T)0;
  ch.scale[0] = (T)kBins / (R2 * (T)(1.0 + 1e-6) + (T)1e-30);
  int need = k_nn;
  for (int round = 0; round < kMaxDepth; ++round) {
    __syncwarp();                                   // the previous round's reads of hist[] are done
    for (int b = lane; b < kBins; b += 32) hist[b] = 0;
    __syncwarp();
    scan([&](bool live, T d2, int) {
      int b;
      if (live && d2 <= R2 && chain_bin(ch, d2, &b)) atomicAdd(&hist[b], 1);
    });
    __syncwarp();
    int loc[8], tot = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { loc[j] = hist[lane * 8 + j]; tot += loc[j]; }
    int inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    const int total = __shfl_sync(0xffffffffu, inc, 31);
    if (round == 0 && total < k_nn) return false;
    int exc = inc - tot, selbin = -1, below = 0, inbin = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (selbin < 0 && exc + loc[j] >= need && need > exc) { selbin = lane * 8 + j; below = exc; inbin = loc[j]; }
      exc += loc[j];
    }
    const unsigned who = __ballot_sync(0xffffffffu, selbin >= 0);
    const int src = __ffs(who) - 1;
    selbin = __shfl_sync(0xffffffffu, selbin, src);
    below = __shfl_sync(0xffffffffu, below, src);
    inbin = __shfl_sync(0xffffffffu, inbin, src);
    need -= below;
    ch.sel[ch.depth] = selbin;
    if (inbin <= 32 || round == kMaxDepth - 1) {
      const int closed = ch.depth + 1;
      T mine = (T)__longlong_as_double(0x7ff0000000000000ll);
      int have = 0;
      scan([&](bool live, T d2, int) {
        bool in = live && d2 <= R2;
        if (in)
          for (int l = 0; l < closed; ++l) in = in && (bin_of(d2, ch.lo[l], ch.scale[l]) == ch.sel[l]);
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        const int slot = have + __popc(bal & ((1u << lane) - 1u));
        for (unsigned rem = bal; rem; rem &= rem - 1) {
          const int srcl = __ffs(rem) - 1;
          const int dst = __shfl_sync(0xffffffffu, slot, srcl);
          const T v = __shfl_sync(0xffffffffu, d2, srcl);
          if (lane == dst) mine = v;
        }
        have += __popc(bal);
      });
      if (inbin > 32 && lane == 0) atomicOr(flags, 1);          // unresolved tie block
      if (have == 1) { *rk2 = __shfl_sync(0xffffffffu, mine, 0); return true; }
      int rank = 0;
      for (int l2 = 0; l2 < 32; ++l2) {
        const T v = __shfl_sync(0xffffffffu, mine, l2);
        rank += (v < mine) || (v == mine && l2 < lane);
      }
      const unsigned hitl = __ballot_sync(0xffffffffu, rank == need - 1 && lane < have);
      const int srcl = hitl ? __ffs(hitl) - 1 : 0;
      *rk2 = __shfl_sync(0xffffffffu, mine, srcl);
      return true;
    }
    const T wbin = (T)1 / ch.scale[ch.depth];
    ch.lo[ch.depth + 1] = ch.lo[ch.depth] + (T)selbin * wbin;
    ch.scale[ch.depth + 1] = ch.scale[ch.depth] * (T)kBins;
    ch.depth += 1;
  }
  return true;
}

constexpr int kKnnWarps = 4;
constexpr int kMaxLevels = 4;
struct KnnLevels { int n; int min_pop; int L[kMaxLevels]; double r2[kMaxLevels]; };   // windows of +-L cells, radius^2 they cover
constexpr int kListCap = 1024;      // (f32 d2, sorted position) pairs cached per warp between the passes

// One warp per point.  Pass A sweeps the cell window once, evaluating d2 in float32 and caching
// (d2, position) of everything inside the window radius in shared memory.  The k-th smallest
// f32 value v is then found on the cached list; because the f32 evaluation error eps is known,
// the exact (f64) k-th value lies among the few entries within a band of v, every entry
// clearly below the band is a neighbour, every entry clearly above is not, and only the band
// entries are re-evaluated in sequential f64 (sklearn's arithmetic).  rk2 and the neighbour
// sets are therefore exactly those of the all-f64 algorithm (kept as the overflow fallback).
//
// This general kernel handles any neighbourhood; in the batched path it only sees the points the
// fast kernel below deferred (`queue` = their sorted positions, `queue_cnt` per scan).
__global__ void __launch_bounds__(kKnnWarps * 32) knn_select_kernel(
    const float4* __restrict__ sorted_all, const int* __restrict__ cells_all, const GridMeta* __restrict__ meta,
    const int64_t* __restrict__ off, int G, int k_nn, double r2_max, KnnLevels lv,
    double* __restrict__ rk2_all, int32_t* __restrict__ knn_all, int32_t* __restrict__ knn_cnt_all,
    int32_t* __restrict__ flags, const int32_t* __restrict__ queue_all, const int32_t* __restrict__ queue_cnt) {
  const int s = blockIdx.y;
  const GridMeta m = meta[s];
  const int n = queue_all ? queue_cnt[s] : m.n;
  const int32_t* __restrict__ queue = queue_all ? queue_all + off[s] : nullptr;
  const float4* __restrict__ sorted = sorted_all + off[s];
  const int* __restrict__ cells = cells_all + (size_t)s * cell_stride(G);
  __shared__ int hist_sh[kKnnWarps][kBins];
  __shared__ float list_d2[kKnnWarps][kListCap];
  __shared__ int list_pos[kKnnWarps][kListCap];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int* hist = hist_sh[wib];
  float* ld2 = list_d2[wib];
  int* lpos = list_pos[wib];
  const int warps_per_grid = gridDim.x * kKnnWarps;
  const double kInf = __longlong_as_double(0x7ff0000000000000ll);

  for (int item = blockIdx.x * kKnnWarps + wib; item < n; item += warps_per_grid) {
    const int pos = queue ? queue[item] : item;
    const float4 p = sorted[pos];
    const int i = __float_as_int(p.w);
    const int cx = cell_coord(p.x, m.x0, m.inv_cell), cy = cell_coord(p.y, m.y0, m.inv_cell);
    double rk2 = kInf;
    int32_t* out = knn_all + ((size_t)off[s] + i) * k_nn;
    int emitted = 0;
    auto exact_d2 = [&](int kpos) {
      const float4 q = __ldg(sorted + kpos);
      return sqdist_f64_seq(p.x, p.y, p.z, q.x, q.y, q.z);
    };
    for (int level = 0; level < lv.n; ++level) {
      const int L = lv.L[level];
      const double R2 = lv.r2[level];
      const bool last = level == lv.n - 1;
      if (!last) {             // cheap reject: fewer than k+1 points in the whole window
        const int xa = clampi(cx - L, 0, G - 1), xb = clampi(cx + L, 0, G - 1);
        const int ya = clampi(cy - L, 0, G - 1), yb = clampi(cy + L, 0, G - 1);
        int tot = 0;
        if (ya + lane <= yb) tot = __ldg(cells + (ya + lane) * G + xb + 1) - __ldg(cells + (ya + lane) * G + xa);
        tot = warp_sum(tot);                             // windows are at most 32 rows tall
        if (tot <= lv.min_pop) continue;
      }
      // f32 evaluation error of d2 (coordinates up to ~1e2 m, d2 <= R2): a few ulps of d2 plus
      // the rounding of the coordinate differences; eps is a safe absolute bound
      const float eps = 4e-6f * (float)R2 + 1e-9f;
      const float band = 8.0f * eps;
      const float R2f_list = (float)R2 + band;          // cache everything that could be within R2
      // ---- pass A ----
      int cnt = 0;
      for_each_candidate(sorted, cells, G, cx, cy, L, pos, [&](bool live, const float4& q, int kpos) {
        float d2 = 0.f;
        bool in = false;
        if (live) {
          const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
          d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
          in = d2 <= R2f_list;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        const int slot = cnt + __popc(bal & ((1u << lane) - 1u));
        if (in && slot < kListCap) { ld2[slot] = d2; lpos[slot] = kpos; }
        cnt += __popc(bal);
      });
      __syncwarp();
      if (cnt > kListCap) {
        // ---- overflow (very dense neighbourhood): all-f64 algorithm, window replayed per pass ----
        auto scan_grid = [&](auto f) {
          for_each_candidate(sorted, cells, G, cx, cy, L, pos, [&](bool live, const float4& q, int) {
            f(live, live ? sqdist_f64_seq(p.x, p.y, p.z, q.x, q.y, q.z) : 0.0, __float_as_int(q.w));
          });
        };
        const bool found = kth_by_histogram<double>(scan_grid, k_nn, R2, hist, lane, &rk2, flags);
        if (!found && !last) continue;
        const double cut = fmin(rk2, r2_max);
        scan_grid([&](bool live, double d2, int j) {
          const bool in = live && d2 <= cut;
          const unsigned bal = __ballot_sync(0xffffffffu, in);
          const int slot = emitted + __popc(bal & ((1u << lane) - 1u));
          if (in && slot < k_nn) out[slot] = j;
          emitted += __popc(bal);
        });
        break;
      }
      auto scan_list = [&](auto f) {
        for (int e0 = 0; e0 < cnt; e0 += 32) {
          const int e = e0 + lane;
          const bool live = e < cnt;
          f(live, live ? ld2[e] : 0.f, live ? lpos[e] : 0);
        }
      };
      // number of entries certainly / possibly within R2 (exactly)
      int sure = 0;
      scan_list([&](bool live, float d2, int) {
        sure += __popc(__ballot_sync(0xffffffffu, live && d2 < (float)R2 - band));
      });
      double cut = r2_max;              // exact cut-off for emission
      float v32 = 0.f;
      // fine level: only when k points are certainly inside the fine radius (the window then
      // provably holds the k nearest); radius level: whenever k candidates are cached
      if (sure >= k_nn || (last && cnt >= k_nn)) {
        // find the k-th smallest on the f32 keys, then exactly inside the band around it
        float v = 0.f;
        kth_by_histogram<float>(scan_list, k_nn, R2f_list, hist, lane, &v, flags);
        v32 = v;
        int n_low = 0, n_band = 0;
        double mine = kInf;
        double bmin = kInf, bmax = -kInf;                  // over ALL band entries (tie blocks > 32)
        scan_list([&](bool live, float d2, int kpos) {
          const bool low = live && d2 < v32 - band;
          const bool inb = live && !low && d2 <= v32 + band;
          n_low += __popc(__ballot_sync(0xffffffffu, low));
          const unsigned bal = __ballot_sync(0xffffffffu, inb);
          const int slot = n_band + __popc(bal & ((1u << lane) - 1u));
          const double ex = inb ? exact_d2(kpos) : 0.0;
          if (inb) { bmin = fmin(bmin, ex); bmax = fmax(bmax, ex); }
          for (unsigned rem = bal; rem; rem &= rem - 1) {
            const int srcl = __ffs(rem) - 1;
            const int dst = __shfl_sync(0xffffffffu, slot, srcl);
            const double vv = __shfl_sync(0xffffffffu, ex, srcl);
            if (lane == dst) mine = vv;
          }
          n_band += __popc(bal);
        });
        const int need = k_nn - n_low;                     // 1-based rank inside the band
        bool tie_block = false;                            // > 32 band entries, all exactly equal
        if (n_band > 32) {
          bmin = warp_min(bmin); bmax = warp_max(bmax);
          tie_block = bmin == bmax;
          if (!tie_block && lane == 0) atomicOr(flags, 1);
        }
        if (tie_block) {
          rk2 = bmin;                                      // every rank inside the block has this value
        } else {
          int rank = 0;
          if (n_band > 1) {
            const int nb = min(n_band, 32);                    // lanes beyond hold +inf and rank nobody down
            for (int l2 = 0; l2 < nb; ++l2) {
              const double vv = __shfl_sync(0xffffffffu, mine, l2);
              rank += (vv < mine) || (vv == mine && l2 < lane);
            }
          }
          const unsigned hitl = __ballot_sync(0xffffffffu, rank == need - 1 && lane < min(n_band, 32));
          if (hitl) rk2 = __shfl_sync(0xffffffffu, mine, __ffs(hitl) - 1);
          else if (lane == 0) atomicOr(flags, 1);
        }
        cut = fmin(rk2, r2_max);
      } else if (!last) {
        continue;      // (also the borderline case: the radius level, whose window contains this one, decides)
      }
      // ---- emission from the list: sure-below, sure-above, or exact inside the band around cut ----
      const float cutf = (float)cut;
      scan_list([&](bool live, float d2, int kpos) {
        bool in = false;
        if (live) {
          if (d2 < cutf - band) in = true;
          else if (d2 <= cutf + band) in = exact_d2(kpos) <= cut;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        const int slot = emitted + __popc(bal & ((1u << lane) - 1u));
        if (in && slot < k_nn) out[slot] = __float_as_int(__ldg(sorted + kpos).w);
        emitted += __popc(bal);
      });
      break;
    }
    if (lane == 0) {
      if (emitted > k_nn) { atomicOr(flags, 2); emitted = k_nn | kTruncatedRow; }   // ties beyond k: first k kept
      rk2_all[off[s] + i] = rk2;
      knn_cnt_all[off[s] + i] = emitted;
    }
    __syncwarp();
  }
}

// ---- H.1b: the common case, tuned for occupancy ---------------------------------------------------
// Same algorithm as knn_select_kernel for neighbourhoods that fit its compact scratch: list
// entries are (f32 d2, u16 {window row, offset in the row}) and the histogram packs two 16-bit
// counters per word, 6.5 KB per warp instead of 9, and the all-f64 fallback lives in the general
// kernel, which keeps this one at 64 registers: 8 CTAs per SM instead of 6.  Points whose
// window has a row of more than 4095 points, or more than kListCap candidates inside the radius,
// are appended to `queue` for the general kernel.
constexpr int kRowBits = 12;

__device__ __forceinline__ int bin_of_f(float d2, float lo, float scale) {
  const int b = (int)((d2 - lo) * scale);
  return b < 0 ? 0 : (b > kBins - 1 ? kBins - 1 : b);
}

// k-th smallest (1-based `k_nn`) of the cnt >= k_nn floats in ld2 (all <= R2).  Entry e belongs to
// lane e % 32; bit e / 32 of the lane's `alive` word says whether it passed every closed level of
// the nested 256-bin histograms, so a pass evaluates at most the previous and the open level.
__device__ __forceinline__ float kth_in_list(const float* ld2, int cnt, int k_nn, float R2, unsigned* hist32, int lane,
                                             int32_t* flags) {
  const int nchunks = (cnt + 31) >> 5;
  unsigned alive = 0u;
  for (int c = 0; c < nchunks; ++c) alive |= (unsigned)(c * 32 + lane < cnt) << c;
  float lo = 0.f, scale = (float)kBins / (R2 * (float)(1.0 + 1e-6) + 1e-30f);
  float lo_prev = 0.f, scale_prev = 0.f;
  int sel_prev = -1;
  int need = k_nn;
  for (int round = 0; round < kMaxDepth; ++round) {
#pragma unroll
    for (int j = 0; j < kBins / 64; ++j) hist32[lane + 32 * j] = 0u;
    __syncwarp();
    for (int c = 0; c < nchunks; ++c) {
      if (!((alive >> c) & 1u)) continue;
      const float d2 = ld2[c * 32 + lane];
      if (sel_prev >= 0 && bin_of_f(d2, lo_prev, scale_prev) != sel_prev) { alive &= ~(1u << c); continue; }
      const int b = bin_of_f(d2, lo, scale);
      atomicAdd(&hist32[b >> 1], 1u << ((b & 1) * 16));
    }
    __syncwarp();
    int loc[8], tot = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned w = hist32[lane * 4 + j];
      loc[2 * j] = (int)(w & 0xffffu); loc[2 * j + 1] = (int)(w >> 16);
      tot += loc[2 * j] + loc[2 * j + 1];
    }
    int inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    int exc = inc - tot, selbin = -1, below = 0, inbin = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (selbin < 0 && exc + loc[j] >= need && need > exc) { selbin = lane * 8 + j; below = exc; inbin = loc[j]; }
      exc += loc[j];
    }
    const unsigned who = __ballot_sync(0xffffffffu, selbin >= 0);
    const int src = __ffs(who) - 1;
    selbin = __shfl_sync(0xffffffffu, selbin, src);
    below = __shfl_sync(0xffffffffu, below, src);
    inbin = __shfl_sync(0xffffffffu, inbin, src);
    need -= below;
    if (inbin <= 32 || round == kMaxDepth - 1) {
      float mine = __int_as_float(0x7f800000);
      int have = 0;
      for (int c = 0; c < nchunks; ++c) {                      // warp-uniform trip count
        const float d2 = ((alive >> c) & 1u) ? ld2[c * 32 + lane] : 0.f;
        const bool in = ((alive >> c) & 1u) && bin_of_f(d2, lo, scale) == selbin;
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        const int slot = have + __popc(bal & ((1u << lane) - 1u));
        for (unsigned rem = bal; rem; rem &= rem - 1) {
          const int srcl = __ffs(rem) - 1;
          const int dst = __shfl_sync(0xffffffffu, slot, srcl);
          const float v = __shfl_sync(0xffffffffu, d2, srcl);
          if (lane == dst) mine = v;
        }
        have += __popc(bal);
      }
      if (inbin > 32 && lane == 0) atomicOr(flags, 1);          // unresolved tie block
      if (have == 1) return __shfl_sync(0xffffffffu, mine, 0);
      int rank = 0;
      const int nh = min(have, 32);                            // lanes beyond hold +inf and rank nobody down
      for (int l2 = 0; l2 < nh; ++l2) {
        const float v = __shfl_sync(0xffffffffu, mine, l2);
        rank += (v < mine) || (v == mine && l2 < lane);
      }
      const unsigned hitl = __ballot_sync(0xffffffffu, rank == need - 1 && lane < have);
      return __shfl_sync(0xffffffffu, mine, hitl ? __ffs(hitl) - 1 : 0);
    }
    lo_prev = lo; scale_prev = scale; sel_prev = selbin;
    lo = lo + (float)selbin * (1.f / scale);
    scale = scale * (float)kBins;
  }
  return 0.f;     // not reached: the last round always returns
}

__global__ void __launch_bounds__(kKnnWarps * 32, 8) knn_select_fast_kernel(
    const float4* __restrict__ sorted_all, const int* __restrict__ cells_all, const GridMeta* __restrict__ meta,
    const int64_t* __restrict__ off, int G, int k_nn, double r2_max, KnnLevels lv,
    double* __restrict__ rk2_all, int32_t* __restrict__ knn_all, int32_t* __restrict__ knn_cnt_all,
    int32_t* __restrict__ flags, int32_t* __restrict__ queue_all, int32_t* __restrict__ queue_cnt) {
  const int s = blockIdx.y;
  const GridMeta m = meta[s];
  const int n = m.n;
  const float4* __restrict__ sorted = sorted_all + off[s];
  const int* __restrict__ cells = cells_all + (size_t)s * cell_stride(G);
  int32_t* __restrict__ queue = queue_all + off[s];
  __shared__ unsigned hist_sh[kKnnWarps][kBins / 2];
  __shared__ float list_d2[kKnnWarps][kListCap];
  __shared__ unsigned short list_pos[kKnnWarps][kListCap];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  unsigned* hist32 = hist_sh[wib];
  float* ld2 = list_d2[wib];
  unsigned short* lpos = list_pos[wib];
  const int warps_per_grid = gridDim.x * kKnnWarps;
  const double kInf = __longlong_as_double(0x7ff0000000000000ll);

  for (int pos = blockIdx.x * kKnnWarps + wib; pos < n; pos += warps_per_grid) {
    const float4 p = sorted[pos];
    const int i = __float_as_int(p.w);
    const int cx = cell_coord(p.x, m.x0, m.inv_cell), cy = cell_coord(p.y, m.y0, m.inv_cell);
    double rk2 = kInf;
    int32_t* out = knn_all + ((size_t)off[s] + i) * k_nn;
    int emitted = 0;
    bool defer = false;
    for (int level = 0; level < lv.n; ++level) {
      const int L = lv.L[level];
      const double R2 = lv.r2[level];
      const bool last = level == lv.n - 1;
      const int xa = clampi(cx - L, 0, G - 1), xb = clampi(cx + L, 0, G - 1);
      const int ya = clampi(cy - L, 0, G - 1), yb = clampi(cy + L, 0, G - 1);
      const int nrows = yb - ya + 1;                           // <= 31 (checked by the launcher)
      int rkb = 0, rlen = 0;                                   // lane r: position range of window row r
      if (lane < nrows) {
        rkb = __ldg(cells + (ya + lane) * G + xa);
        rlen = __ldg(cells + (ya + lane) * G + xb + 1) - rkb;
      }
      if (!last && warp_sum(rlen) <= lv.min_pop) continue;     // too few points for a fine level
      if (__any_sync(0xffffffffu, rlen >= (1 << kRowBits))) { defer = true; break; }
      const float eps = 4e-6f * (float)R2 + 1e-9f;             // f32 evaluation error bound, as in the general kernel
      const float band = 8.0f * eps;
      const float R2f_list = (float)R2 + band;
      // the rows' position ranges concatenated (rows hold ~19 points here: row-by-row trips would run
      // half empty).  The non-empty rows are compacted into the low lanes -- lane k holds the first
      // flat index and the first position of the k-th non-empty row -- so that a trip finds every
      // lane's row with one OR-reduction of the rows' start lanes and two popcounts
      int incl = rlen;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      const unsigned ne_rows = __ballot_sync(0xffffffffu, rlen > 0);
      const int n_ne = __popc(ne_rows);
      const unsigned lane_le = 0xffffffffu >> (31 - lane);
      // compaction through the (idle) histogram words: row r -> lane popc(non-empty rows below r)
      int* scr = reinterpret_cast<int*>(hist32);
      if (rlen > 0) {
        const int k = __popc(ne_rows & (lane_le >> 1));
        scr[k] = incl - rlen;
        scr[32 + k] = rkb;
      }
      __syncwarp();
      const int cexcl = lane < n_ne ? scr[lane] : 0x7fffffff;
      const int crkb = lane < n_ne ? scr[32 + lane] : 0;
      __syncwarp();
      auto position = [&](unsigned short e) { return __shfl_sync(0xffffffffu, crkb, e >> kRowBits) + (int)(e & ((1u << kRowBits) - 1u)); };
      // ---- pass A: cache (d2, row|offset) of everything that could be within R2; count the sure ones ----
      int cnt = 0, sure = 0;
      {
        for (int f0 = 0; f0 < total; f0 += 32) {
          const int idx = f0 + lane;
          const int sl = cexcl - f0;                                // lane of this trip at which my row starts
          const unsigned starts = __reduce_or_sync(0xffffffffu, (sl > 0 && sl < 32) ? 1u << sl : 0u);
          const int row = __popc(__ballot_sync(0xffffffffu, cexcl <= f0)) - 1 + __popc(starts & lane_le);
          const int o = idx - __shfl_sync(0xffffffffu, cexcl, row);
          const int kq = __shfl_sync(0xffffffffu, crkb, row) + o;
          float d2 = 0.f;
          bool in = false;
          if (idx < total && kq != pos) {
            const float4 q = __ldg(sorted + kq);
            const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
            d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            in = d2 <= R2f_list;
          }
          const unsigned bal = __ballot_sync(0xffffffffu, in);
          const int slot = cnt + __popc(bal & lt);
          if (in && slot < kListCap) { ld2[slot] = d2; lpos[slot] = (unsigned short)((row << kRowBits) | o); }
          cnt += __popc(bal);
          sure += (in && d2 < (float)R2 - band) ? 1 : 0;         // per lane; summed once below
        }
        sure = __reduce_add_sync(0xffffffffu, sure);
      }
      __syncwarp();
      if (cnt > kListCap) { defer = true; break; }
      auto exact_d2 = [&](int kpos) {
        const float4 q = __ldg(sorted + kpos);
        return sqdist_f64_seq(p.x, p.y, p.z, q.x, q.y, q.z);
      };
      double cut = r2_max;              // exact cut-off for emission
      // fine level: only when k points are certainly inside the fine radius (the window then
      // provably holds the k nearest); radius level: whenever k candidates are cached
      if (sure >= k_nn || (last && cnt >= k_nn)) {
        const float v32 = kth_in_list(ld2, cnt, k_nn, R2f_list, hist32, lane, flags);
        // When the k-th distance lies clearly inside the graph radius, everything clearly below
        // it is a neighbour and is emitted in this same pass; the few entries inside the band
        // around it are kept in registers, ranked exactly, and emitted right after.
        const bool fused = v32 + band < (float)r2_max - band;           // warp-uniform
        int n_low = 0, n_band = 0;
        double mine = kInf;
        double bmin = kInf, bmax = -kInf;                  // over ALL band entries (tie blocks > 32)
        int mine_pos = 0;
        for (int e0 = 0; e0 < cnt; e0 += 32) {
          const int e = e0 + lane;
          const bool live = e < cnt;
          const float d2 = live ? ld2[e] : 0.f;
          const unsigned short code = live ? lpos[e] : (unsigned short)0;
          const int kpos = position(code);
          const bool low = live && d2 < v32 - band;
          const bool inb = live && !low && d2 <= v32 + band;
          const unsigned bal_low = __ballot_sync(0xffffffffu, low);
          n_low += __popc(bal_low);
          if (fused) {
            const int slot = emitted + __popc(bal_low & lt);
            if (low && slot < k_nn) out[slot] = __float_as_int(__ldg(sorted + kpos).w);
            emitted += __popc(bal_low);
          }
          const unsigned bal = __ballot_sync(0xffffffffu, inb);
          const int slot = n_band + __popc(bal & lt);
          const double ex = inb ? exact_d2(kpos) : 0.0;
          if (inb) { bmin = fmin(bmin, ex); bmax = fmax(bmax, ex); }
          for (unsigned rem = bal; rem; rem &= rem - 1) {
            const int srcl = __ffs(rem) - 1;
            const int dst = __shfl_sync(0xffffffffu, slot, srcl);
            const double vv = __shfl_sync(0xffffffffu, ex, srcl);
            const int pp_ = __shfl_sync(0xffffffffu, kpos, srcl);
            if (lane == dst) { mine = vv; mine_pos = pp_; }
          }
          n_band += __popc(bal);
        }
        const int need = k_nn - n_low;                     // 1-based rank inside the band
        bool tie_block = false;                            // > 32 band entries, all exactly equal
        if (n_band > 32) {
          bmin = warp_min(bmin); bmax = warp_max(bmax);
          tie_block = bmin == bmax;
          if (!tie_block && lane == 0) atomicOr(flags, 1);
        }
        if (tie_block) {
          rk2 = bmin;                                      // every rank inside the block has this value
        } else {
          int rank = 0;
          if (n_band > 1) {
            const int nb = min(n_band, 32);                    // lanes beyond hold +inf and rank nobody down
            for (int l2 = 0; l2 < nb; ++l2) {
              const double vv = __shfl_sync(0xffffffffu, mine, l2);
              rank += (vv < mine) || (vv == mine && l2 < lane);
            }
          }
          const unsigned hitl = __ballot_sync(0xffffffffu, rank == need - 1 && lane < min(n_band, 32));
          if (hitl) rk2 = __shfl_sync(0xffffffffu, mine, __ffs(hitl) - 1);
          else if (lane == 0) atomicOr(flags, 1);
        }
        cut = fmin(rk2, r2_max);
        if (fused && n_band > 32) emitted = 0;             // the registers hold 32 of the band: emit from the list
        if (fused && n_band <= 32) {
          const bool in = lane < min(n_band, 32) && mine <= cut;
          const unsigned bal = __ballot_sync(0xffffffffu, in);
          const int slot = emitted + __popc(bal & lt);
          if (in && slot < k_nn) out[slot] = __float_as_int(__ldg(sorted + mine_pos).w);
          emitted += __popc(bal);
          break;
        }
      } else if (!last) {
        continue;      // (also the borderline case: the radius level, whose window contains this one, decides)
      }
      // ---- emission from the list: sure-below, sure-above, or exact inside the band around cut ----
      const float cutf = (float)cut;
      for (int e0 = 0; e0 < cnt; e0 += 32) {
        const int e = e0 + lane;
        const bool live = e < cnt;
        const float d2 = live ? ld2[e] : 0.f;
        const unsigned short code = live ? lpos[e] : (unsigned short)0;
        const int kpos = position(code);
        bool in = false;
        if (live) {
          if (d2 < cutf - band) in = true;
          else if (d2 <= cutf + band) in = exact_d2(kpos) <= cut;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        const int slot = emitted + __popc(bal & lt);
        if (in && slot < k_nn) out[slot] = __float_as_int(__ldg(sorted + kpos).w);
        emitted += __popc(bal);
      }
      break;
    }
    if (defer) {                                               // warp-uniform
      if (lane == 0) queue[atomicAdd(&queue_cnt[s], 1)] = pos;
    } else if (lane == 0) {
      if (emitted > k_nn) { atomicOr(flags, 2); emitted = k_nn | kTruncatedRow; }   // ties beyond k: first k kept
      rk2_all[off[s] + i] = rk2;
      knn_cnt_all[off[s] + i] = emitted;
    }
    __syncwarp();
  }
}

// ---- H.2: mutual test + L1 pp weights -------------------------------------------------------------
// Row i keeps j iff i is also within j's k-th neighbour distance.  When the caller announces the
// DBSCAN radius (eps_first >= 0), the edges with (double)w <= eps_first are written first and
// their number goes to nbr_eps_cnt: the DBSCAN kernels then touch only that prefix of every row
// and never read the weights (order within a row carries no meaning; rows of more than 96 slots
// are left unpartitioned and get nbr_eps_cnt = -1).
constexpr int kMutualChunks = 3;        // 32-lane chunks held in registers for the partition
constexpr int kMutualLanes = 4;         // lanes per row on the fused (eps-edges only) path: 385 (one warp) / 327 (8) / 312 (4) / 349 (2) us per 24 scans

__global__ void __launch_bounds__(256) mutual_edges_kernel(
    const float4* __restrict__ kept, const int64_t* __restrict__ off, const int32_t* __restrict__ n_kept,
    int k_nn, const double* __restrict__ rk2, const int32_t* __restrict__ knn, const int32_t* __restrict__ knn_cnt,
    int32_t* __restrict__ nbr, float* __restrict__ nbr_w, int32_t* __restrict__ nbr_cnt, double eps_first,
    int32_t* __restrict__ nbr_eps_cnt) {
  const int s = blockIdx.y;
  const int n = n_kept[s];
  const int64_t base = off[s];
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const bool partition = nbr_eps_cnt != nullptr && eps_first >= 0.0 && k_nn <= 32 * kMutualChunks;
  if (partition && nbr_w == nullptr) {
    // the fused path (eps-edges only): kMutualLanes lanes per row, several rows per warp -- the pass is
    // a chain of dependent gathers (row -> neighbour's point, k-th distance, flags), so rows in flight
    // count for more than full lanes (as in the DBSCAN forest passes)
    const int gl = lane & (kMutualLanes - 1), sub = lane / kMutualLanes;
    const unsigned gmask = ((kMutualLanes == 32 ? 0u : (1u << kMutualLanes)) - 1u) << (sub * kMutualLanes);
    for (int iw = warp * (32 / kMutualLanes); iw < n; iw += nwarps * (32 / kMutualLanes)) {    // warp-uniform
      const int i = iw + sub;
      const bool act = i < n;
      float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
      int m = 0;
      if (act) { p = kept[base + i]; m = knn_cnt[base + i] & ~kTruncatedRow; }
      const int32_t* row = knn + (size_t)(base + i) * k_nn;
      int32_t* orow = nbr + (size_t)(base + i) * k_nn;
      const int mmax = __reduce_max_sync(0xffffffffu, m);
      int n_e = 0, n_o = 0;
      for (int c0 = 0; c0 < mmax; c0 += kMutualLanes) {
        const int c = c0 + gl;
        bool in = false;
        int j = 0;
        float w = 0.f;
        if (c < m) {
          j = row[c];
          const float4 q = kept[base + j];
          const double d2 = sqdist_f64_seq(q.x, q.y, q.z, p.x, p.y, p.z);   // same expression as row j used
          w = fabsf(__fsub_rn(p.w, q.w));
          in = !(d2 > rk2[base + j]);
          if (in && (knn_cnt[base + j] & kTruncatedRow)) {    // row j kept k of its tied neighbours: is i one of them?
            const int32_t* rj = knn + (size_t)(base + j) * k_nn;
            bool found = false;
            for (int t = 0; t < k_nn; ++t) found |= rj[t] == i;
            in = found;
          }
        }
        const bool e = in && (double)w <= eps_first;
        const unsigned be = __ballot_sync(0xffffffffu, e) & gmask, bo = __ballot_sync(0xffffffffu, in && !e) & gmask;
        if (e) orow[n_e + __popc(be & lt)] = j;
        n_e += __popc(be);
        n_o += __popc(bo);
      }
      if (gl == 0 && act) { nbr_cnt[base + i] = n_e + n_o; nbr_eps_cnt[base + i] = n_e; }
    }
    return;
  }
  for (int i = warp; i < n; i += nwarps) {
    const float4 p = kept[base + i];
    const int m = knn_cnt[base + i] & ~kTruncatedRow;
    const int32_t* row = knn + (size_t)(base + i) * k_nn;
    int32_t* orow = nbr + (size_t)(base + i) * k_nn;
    float* wrow = nbr_w + (size_t)(base + i) * k_nn;
    auto candidate = [&](int c, int& j, float& w) {
      j = 0; w = 0.f;
      if (c >= m) return false;
      j = row[c];
      const float4 q = kept[base + j];
      const double d2 = sqdist_f64_seq(q.x, q.y, q.z, p.x, p.y, p.z);   // same expression as row j used
      w = fabsf(__fsub_rn(p.w, q.w));
      if (d2 > rk2[base + j]) return false;
      if (knn_cnt[base + j] & kTruncatedRow) {              // row j kept k of its tied neighbours: is i one of them?
        const int32_t* rj = knn + (size_t)(base + j) * k_nn;
        bool found = false;
        for (int t = 0; t < k_nn; ++t) found |= rj[t] == i;
        return found;
      }
      return true;
    };
    if (partition) {
      int jj[kMutualChunks];
      float ww[kMutualChunks];
      unsigned bal_e[kMutualChunks], bal_o[kMutualChunks];
      int n_e = 0, n_o = 0;
#pragma unroll
      for (int t = 0; t < kMutualChunks; ++t) {
        bal_e[t] = bal_o[t] = 0u;
        if (32 * t < m) {                                   // warp-uniform
          const bool in = candidate(32 * t + lane, jj[t], ww[t]);
          const bool e = in && (double)ww[t] <= eps_first;
          bal_e[t] = __ballot_sync(0xffffffffu, e);
          bal_o[t] = __ballot_sync(0xffffffffu, in && !e);
          n_e += __popc(bal_e[t]);
          n_o += __popc(bal_o[t]);
        }
      }
      int have_e = 0, have_o = n_e;
#pragma unroll
      for (int t = 0; t < kMutualChunks; ++t) {
        const unsigned me = 1u << lane;
        int slot = -1;
        if (bal_e[t] & me) slot = have_e + __popc(bal_e[t] & lt);
        else if ((bal_o[t] & me) && nbr_w) slot = have_o + __popc(bal_o[t] & lt);   // no weights wanted: eps-edges only
        if (slot >= 0) {
          orow[slot] = jj[t];
          if (nbr_w) wrow[slot] = ww[t];
        }
        have_e += __popc(bal_e[t]);
        have_o += __popc(bal_o[t]);
      }
      if (lane == 0) { nbr_cnt[base + i] = n_e + n_o; nbr_eps_cnt[base + i] = n_e; }
      continue;
    }
    int have = 0;
    for (int c0 = 0; c0 < m; c0 += 32) {
      int j;
      float w;
      const bool in = candidate(c0 + lane, j, w);
      const unsigned bal = __ballot_sync(0xffffffffu, in);
      if (in) {
        const int slot = have + __popc(bal & lt);
        orow[slot] = j;
        wrow[slot] = w;
      }
      have += __popc(bal);
    }
    if (lane == 0) {
      nbr_cnt[base + i] = have;
      if (nbr_eps_cnt) nbr_eps_cnt[base + i] = -1;
    }
  }
}

// ---- I: DBSCAN ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dbscan_core_kernel(
    const int64_t* __restrict__ off, const int32_t* __restrict__ n_kept, int k_nn, const float* __restrict__ nbr_w,
    const int32_t* __restrict__ nbr_cnt, const int32_t* __restrict__ eps_cnt, double eps, int min_samples,
    uint8_t* __restrict__ core, int32_t* __restrict__ parent) {
  const int s = blockIdx.y;
  const int n = n_kept[s];
  const int64_t base = off[s];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int deg = 1;                                   // the point itself (sklearn adds the diagonal)
    const int pre = eps_cnt ? eps_cnt[base + i] : -1;
    if (pre >= 0) deg += pre;                      // rows partitioned by mutual_edges_kernel
    else if (nbr_w) {
      const int m = nbr_cnt[base + i];
      const float* w = nbr_w + (size_t)(base + i) * k_nn;
      for (int c = 0; c < m; ++c) deg += ((double)w[c] <= eps);
    }
    core[base + i] = deg >= min_samples;
    parent[base + i] = i;
  }
}

__device__ __forceinline__ int uf_find(int32_t* parent, int a) {
  int p = parent[a];
  while (p != a) {
    const int gp = parent[p];
    if (gp != p) parent[a] = gp;    // path halving (benign race: parents only ever decrease)
    a = p;
    p = parent[a];
  }
  return a;
}
__device__ __forceinline__ void uf_union(int32_t* parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }       // a > b: hang the larger root under the smaller
    const int old = atomicCAS(&parent[a], a, b);
    if (old == a) return;
  }
}

// lanes that share one graph row in the DBSCAN forest passes.  Both passes are chains of dependent
// gathers (row -> neighbour -> its core flag / parent), so what counts is how many rows a warp has in
// flight, not how its lanes line up: measured per 24 scans with 32 / 16 / 8 / 4 / 2 / 1 lanes per row --
// initial forest 179 / 167 / 126 / 105 / 105 / 79 us, union sweep 343 / 290 / 248 / 206 / 202 / 219 us
constexpr int kInitLanes = 1, kUnionLanes = 2;

// Initial forest (ECL-CC style): every core point hangs under its smallest qualifying core
// neighbour with a smaller index, then a few pointer-doubling sweeps shorten the chains, so the
// union sweep below mostly finds "already together".
__global__ void __launch_bounds__(256) dbscan_init_kernel(
    const int64_t* __restrict__ off, const int32_t* __restrict__ n_kept, int k_nn, const int32_t* __restrict__ nbr,
    const float* __restrict__ nbr_w, const int32_t* __restrict__ nbr_cnt, const int32_t* __restrict__ eps_cnt, double eps,
    const uint8_t* __restrict__ core, int32_t* __restrict__ parent) {
  const int s = blockIdx.y;
  const int n = n_kept[s];
  const int64_t base = off[s];
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const int gl = lane & (kInitLanes - 1), sub = lane / kInitLanes;
  for (int iw = warp * (32 / kInitLanes); iw < n; iw += nwarps * (32 / kInitLanes)) {      // warp-uniform
    const int i = iw + sub;
    const bool active = i < n && core[base + i];
    int pre = -1, m = 0;
    if (active) {
      pre = eps_cnt ? eps_cnt[base + i] : -1;      // >= 0: only this prefix holds eps-edges
      m = pre >= 0 ? pre : nbr_cnt[base + i];
    }
    const size_t row = (size_t)(base + i) * k_nn;
    int best = i;
    for (int c = gl; c < m; c += kInitLanes) {
      const int j = nbr[row + c];
      if (j < best && (pre >= 0 || (nbr_w && (double)nbr_w[row + c] <= eps)) && core[base + j]) best = j;
    }
#pragma unroll
    for (int o = kInitLanes / 2; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (gl == 0 && active) parent[base + i] = best;
  }
}

__global__ void __launch_bounds__(256) dbscan_jump_kernel(const int64_t* __restrict__ off, const int32_t* __restrict__ n_kept,
                                                          int32_t* __restrict__ parent) {
  const int s = blockIdx.y;
  const int n = n_kept[s];
  int32_t* par = parent + off[s];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int p = par[i];
    int gp = par[p];
    // parents only ever decrease, so a racing update can only shorten the path further
    if (gp != p) { const int ggp = par[gp]; par[i] = ggp; }
  }
}

// one warp per core row: lanes sweep the row's (<= k) edges, each undirected core-core edge is
// united once (from its larger endpoint)
__global__ void __launch_bounds__(256) dbscan_union_kernel(
    const int64_t* __restrict__ off, const int32_t* __restrict__ n_kept, int k_nn, const int32_t* __restrict__ nbr,
    const float* __restrict__ nbr_w, const int32_t* __restrict__ nbr_cnt, const int32_t* __restrict__ eps_cnt, double eps,
    const uint8_t* __restrict__ core, int32_t* __restrict__ parent) {
  const int s = blockIdx.y;
  const int n = n_kept[s];
  const int64_t base = off[s];
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const int gl = lane & (kUnionLanes - 1), sub = lane / kUnionLanes;
  for (int iw = warp * (32 / kUnionLanes); iw < n; iw += nwarps * (32 / kUnionLanes)) {
    const int i = iw + sub;
    if (i >= n || !core[base + i]) continue;
    const int pre = eps_cnt ? eps_cnt[base + i] : -1;
    const int m = pre >= 0 ? pre : nbr_cnt[base + i];
    const size_t row = (size_t)(base + i) * k_nn;
    // after the initial forest and the pointer-doubling sweeps most points hang directly under
    // their root: two points with the same parent are already together (parents never leave
    // their tree), which spares the two dependent find() walks for almost every edge
    const int pi = parent[base + i];
    for (int c = gl; c < m; c += kUnionLanes) {
      const int j = nbr[row + c];
      if (j < i && (pre >= 0 || (nbr_w && (double)nbr_w[row + c] <= eps)) && core[base + j] && parent[base + j] != pi)
        uf_union(parent + base, i, j);
    }
  }
}

// ---- labelling: (a) rank the roots in index order (one CTA per scan, ordered scan),
//                 (b) cores take the rank of their root, (c) borders take the smallest cluster id
//                 among their core neighbours; (b),(c) run grid-wide.
__global__ void __launch_bounds__(1024) dbscan_rank_roots_kernel(
    const int64_t* __restrict__ off, const int32_t* __restrict__ n_kept, const uint8_t* __restrict__ core,
    const int32_t* __restrict__ parent, int32_t* __restrict__ root_rank, int32_t* __restrict__ n_clusters) {
  const int s = blockIdx.x;
  const int n = n_kept[s];
  const int64_t base = off[s];
  __shared__ int warp_cnt[32];
  __shared__ int tile_total;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int running = 0;
  for (int t0 = 0; t0 < n; t0 += 1024) {
    const int i = t0 + threadIdx.x;
    const bool is_root = i < n && core[base + i] && parent[base + i] == i;
    const unsigned bal = __ballot_sync(0xffffffffu, is_root);
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
    if (is_root) root_rank[base + i] = running + warp_cnt[w] + __popc(bal & ((1u << lane) - 1u));
    running += tile_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) n_clusters[s] = running;
}

__global__ void __launch_bounds__(256) dbscan_label_cores_kernel(
    const int64_t* __restrict__ off, const int32_t* __restrict__ n_kept, const uint8_t* __restrict__ core,
    const int32_t* __restrict__ parent, const int32_t* __restrict__ root_rank, int32_t* __restrict__ labels_kept,
    int32_t* __restrict__ labels_full) {
  const int s = blockIdx.y;
  const int n = n_kept[s];
  const int64_t base = off[s];
  const int n_full = (int)(off[s + 1] - base);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_full; i += gridDim.x * blockDim.x) {
    labels_full[base + i] = -1;
    if (i < n) {
      int lab = -1;
      if (core[base + i]) {
        int r = i;
        while (parent[base + r] != r) r = parent[base + r];
        lab = root_rank[base + r];
      }
      labels_kept[base + i] = lab;
    }
  }
}

__global__ void __launch_bounds__(256) dbscan_label_borders_kernel(
    const int64_t* __restrict__ off, const int32_t* __restrict__ n_kept, const int32_t* __restrict__ kept_idx, int k_nn,
    const int32_t* __restrict__ nbr, const float* __restrict__ nbr_w, const int32_t* __restrict__ nbr_cnt,
    const int32_t* __restrict__ eps_cnt, double eps, const uint8_t* __restrict__ core, const int32_t* __restrict__ labels_kept,
    int32_t* __restrict__ border_lab, int32_t* __restrict__ labels_full) {
  const int s = blockIdx.y;
  const int n = n_kept[s];
  const int64_t base = off[s];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int lab;
    if (core[base + i]) {
      lab = labels_kept[base + i];
    } else {
      lab = 0x7fffffff;
      const int pre = eps_cnt ? eps_cnt[base + i] : -1;
      const int m = pre >= 0 ? pre : nbr_cnt[base + i];
      const size_t row = (size_t)(base + i) * k_nn;
      for (int c = 0; c < m; ++c) {
        const int j = nbr[row + c];
        if ((pre >= 0 || (nbr_w && (double)nbr_w[row + c] <= eps)) && core[base + j]) lab = min(lab, labels_kept[base + j]);
      }
      if (lab == 0x7fffffff) lab = -1;
      border_lab[base + i] = lab;       // labels_kept of cores is still being read by other threads
    }
    labels_full[base + kept_idx[base + i]] = lab;
  }
}

__global__ void __launch_bounds__(256) dbscan_merge_borders_kernel(
    const int64_t* __restrict__ off, const int32_t* __restrict__ n_kept, const uint8_t* __restrict__ core,
    const int32_t* __restrict__ border_lab, int32_t* __restrict__ labels_kept) {
  const int s = blockIdx.y;
  const int n = n_kept[s];
  const int64_t base = off[s];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (!core[base + i]) labels_kept[base + i] = border_lab[base + i];
}

// ---- f-4: non-default graph types of precompute_affinity_matrix (clustering_utils.py:16-31) -----
// 'knn' / 'sym_knn' / 'mutual_knn' need the exact k nearest neighbours at ANY distance and 'radius'
// every pair within a radius; neither is on the seed-label path (the configs use
// radius_mutual_knn), so they get exact brute-force kernels: one warp per point, every other
// point of the scan evaluated in sklearn's sequential f64 arithmetic, the k-th distance by the
// nested-histogram selection used above.
__global__ void __launch_bounds__(kKnnWarps * 32) knn_bruteforce_kernel(
    const float* __restrict__ pts, int stride, int n, int k_nn, double d2_max, int32_t* __restrict__ knn,
    int32_t* __restrict__ knn_cnt, double* __restrict__ rk2_out, int32_t* __restrict__ flags) {
  __shared__ int hist_sh[kKnnWarps][kBins];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int* hist = hist_sh[wib];
  const int kk = min(k_nn, n - 1);
  for (int i = blockIdx.x * kKnnWarps + wib; i < n; i += gridDim.x * kKnnWarps) {
    const float px = pts[(size_t)stride * i], py = pts[(size_t)stride * i + 1], pz = pts[(size_t)stride * i + 2];
    auto scan_all = [&](auto f) {
      for (int j0 = 0; j0 < n; j0 += 32) {
        const int j = j0 + lane;
        const bool live = j < n && j != i;
        double d2 = 0.0;
        if (live) d2 = sqdist_f64_seq(px, py, pz, pts[(size_t)stride * j], pts[(size_t)stride * j + 1], pts[(size_t)stride * j + 2]);
        f(live, d2, j);
      }
    };
    int emitted = 0;
    double rk2 = 0.0;
    if (kk > 0) {
      kth_by_histogram<double>(scan_all, kk, d2_max, hist, lane, &rk2, flags);
      int32_t* out = knn + (size_t)i * k_nn;
      scan_all([&](bool live, double d2, int j) {
        const bool in = live && d2 <= rk2;
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        const int slot = emitted + __popc(bal & ((1u << lane) - 1u));
        if (in && slot < k_nn) out[slot] = j;
        emitted += __popc(bal);
      });
    }
    if (lane == 0) {
      if (emitted > k_nn) { atomicOr(flags, 2); emitted = k_nn; }     // ties at the k-th distance: first k by index
      knn_cnt[i] = emitted;
      rk2_out[i] = rk2;
    }
    __syncwarp();
  }
}

// FILL = false: counts[i] = neighbours of i within r2 (self excluded); FILL = true: their indices, ascending
template <bool FILL>
__global__ void __launch_bounds__(256) radius_bruteforce_kernel(const float* __restrict__ pts, int stride, int n, double r2,
                                                                 int64_t* __restrict__ counts, const int64_t* __restrict__ indptr,
                                                                 int32_t* __restrict__ indices) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp; i < n; i += nwarps) {
    const float px = pts[(size_t)stride * i], py = pts[(size_t)stride * i + 1], pz = pts[(size_t)stride * i + 2];
    long long have = 0;
    for (int j0 = 0; j0 < n; j0 += 32) {
      const int j = j0 + lane;
      const bool in = j < n && j != i &&
                      sqdist_f64_seq(px, py, pz, pts[(size_t)stride * j], pts[(size_t)stride * j + 1], pts[(size_t)stride * j + 2]) <= r2;
      const unsigned bal = __ballot_sync(0xffffffffu, in);
      if (FILL && in) indices[indptr[i] + have + __popc(bal & ((1u << lane) - 1u))] = j;
      have += __popc(bal);
    }
    if (!FILL && lane == 0) counts[i] = have;
  }
}

// edge weights of clustering_utils.py:42-56 for a CSR pattern, float32 arithmetic like numpy's:
// kind 0 'l1' |pp_r - pp_j|, 1 'exp' exp((pp_r - pp_j)^2), 2 '3d_l2_distance' ||row_r - row_j|| over ALL
// `width` columns of the point rows (the reference subtracts whole rows), summed left to right
__global__ void __launch_bounds__(256) edge_affinity_kernel(const float* __restrict__ pts, int stride, int width,
                                                            const float* __restrict__ pp, const int64_t* __restrict__ indptr,
                                                            const int32_t* __restrict__ indices, int n, int kind,
                                                            double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp; i < n; i += nwarps) {
    for (int64_t e = indptr[i] + lane; e < indptr[i + 1]; e += 32) {
      const int j = indices[e];
      float w;
      if (kind == 2) {
        float acc = 0.f;
        for (int c = 0; c < width; ++c) {
          const float d = __fsub_rn(pts[(size_t)stride * i + c], pts[(size_t)stride * j + c]);
          acc = c == 0 ? __fmul_rn(d, d) : __fadd_rn(acc, __fmul_rn(d, d));
        }
        w = __fsqrt_rn(acc);
      } else {
        const float d = __fsub_rn(pp[i], pp[j]);
        w = kind == 0 ? fabsf(d) : expf(__fmul_rn(d, d));
      }
      out[e] = (double)w;
    }
  }
}

}  // namespace modest

using namespace modest;

extern "C" int modest_ground_mask_batch(const float* d_ptc, int point_stride, const int64_t* d_off,
                                        const float* d_pp, const double* d_planes, int n_scans, double offset,
                                        const float* h_only_range, const float* h_limit_range, float* d_kept,
                                        int32_t* d_kept_idx, int32_t* d_n_kept, uint8_t* d_mask, void* stream_) {
  modest::StageRange nvtx_("modest:F,G masks");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_scans <= 0) return MODEST_OK;
  MODEST_REQUIRE(d_ptc && d_off && d_pp && d_planes && h_limit_range && d_kept && d_kept_idx && d_n_kept,
                 "ground_mask: null pointer argument");
  MODEST_REQUIRE(point_stride >= 3, "ground_mask: point_stride %d < 3", point_stride);
  MaskCfg c;
  c.offset = offset;
  c.use_only_range = h_only_range != nullptr;
  if (h_only_range) { c.only_x_lo = h_only_range[0]; c.only_x_hi = h_only_range[1]; c.only_y_lo = h_only_range[2]; c.only_y_hi = h_only_range[3]; }
  else { c.only_x_lo = c.only_x_hi = c.only_y_lo = c.only_y_hi = 0.f; }
  c.lim_x_lo = h_limit_range[0]; c.lim_x_hi = h_limit_range[1]; c.lim_y_lo = h_limit_range[2]; c.lim_y_hi = h_limit_range[3];
  ground_mask_compact_kernel<<<n_scans, 1024, 0, stream>>>(d_ptc, point_stride, d_off, d_pp, d_planes, c,
                                                           reinterpret_cast<float4*>(d_kept), d_kept_idx, d_n_kept, d_mask);
  MODEST_LAUNCH_CHECK("ground_mask_compact_kernel");
  note_launch(1);
  return MODEST_OK;
}

static const float kGraphCell = 0.5f;        // fine cell edge of the kNN grid [m]
static const float kGraphSlack = 1.001f;


extern "C" size_t modest_graph_workspace_bytes(int n_scans, int64_t n_points_total, int n_neighbors, int grid_dim) {
  if (grid_dim <= 0) grid_dim = 288;
  size_t b = 0;
  auto add = [&](size_t bytes) { b = align_up(b, 256) + bytes; };
  add(sizeof(GridMeta) * (size_t)n_scans);
  add(sizeof(int) * (size_t)n_scans * cell_stride(grid_dim));
  add(sizeof(float4) * (size_t)n_points_total);
  add(sizeof(double) * (size_t)n_points_total);                    // rk2
  add(sizeof(int32_t) * (size_t)n_points_total * n_neighbors);     // knn
  add(sizeof(int32_t) * (size_t)n_points_total);                   // knn_cnt
  add(sizeof(int32_t) * (size_t)n_points_total);                   // points deferred to the general kNN kernel
  add(sizeof(int32_t) * (size_t)n_scans);                          // ... and their number per scan
  return b + 256;
}

extern "C" int modest_affinity_graph_batch(const float* d_kept, const int64_t* d_off, const int32_t* d_n_kept,
                                           int n_scans, int64_t n_points_total, int64_t max_points,
                                           int n_neighbors, double radius, int grid_dim, int32_t* d_nbr,
                                           float* d_nbr_w, int32_t* d_nbr_cnt, double partition_eps, int32_t* d_nbr_eps_cnt, int32_t* d_flags, void* d_ws,
                                           size_t ws_bytes, void* stream_) {
  modest::StageRange nvtx_("modest:H affinity graph");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_scans <= 0 || n_points_total <= 0) return MODEST_OK;
  if (grid_dim <= 0) grid_dim = 288;
  MODEST_REQUIRE(d_kept && d_off && d_n_kept && d_nbr && d_nbr_cnt && d_flags && d_ws,
                 "affinity_graph: null pointer argument");
  // without a weights buffer only the eps-edges are produced, which needs the partitioned rows
  MODEST_REQUIRE(d_nbr_w || (d_nbr_eps_cnt && partition_eps >= 0.0 && n_neighbors <= 32 * kMutualChunks),
                 "affinity_graph: d_nbr_w may only be NULL together with partition_eps >= 0, d_nbr_eps_cnt and n_neighbors <= %d",
                 32 * kMutualChunks);
  MODEST_REQUIRE(n_neighbors >= 1 && n_neighbors <= 1024, "affinity_graph: n_neighbors %d out of range", n_neighbors);
  MODEST_REQUIRE(radius > 0.0 && radius <= 64.0, "affinity_graph: radius %g out of range", radius);
  MODEST_REQUIRE(ws_bytes >= modest_graph_workspace_bytes(n_scans, n_points_total, n_neighbors, grid_dim),
                 "affinity_graph: workspace too small");
  MODEST_REQUIRE(n_scans <= 65535, "affinity_graph: more than 65535 scans in one launch");
  const int G = grid_dim;
  Arena ar(d_ws, ws_bytes);
  GridMeta* meta = ar.take<GridMeta>(n_scans);
  int* cells = ar.take<int>((size_t)n_scans * cell_stride(G));
  float4* sorted = ar.take<float4>(n_points_total);
  double* rk2 = ar.take<double>(n_points_total);
  int32_t* knn = ar.take<int32_t>((size_t)n_points_total * n_neighbors);
  int32_t* knn_cnt = ar.take<int32_t>(n_points_total);
  int32_t* queue = ar.take<int32_t>(n_points_total);
  int32_t* queue_cnt = ar.take<int32_t>(n_scans);
  MODEST_REQUIRE(ar.ok(), "workspace too small for the requested sizes");

  const float cell = kGraphCell * kGraphSlack;
  int rc = grid2d_build(d_kept, 4, d_off, d_n_kept, n_scans, max_points, cell, G, meta, cells, sorted, stream);
  if (rc != MODEST_OK) return rc;
  // search windows: +-1, +-2, ... cells (radius = window reach, capped at the graph radius); the
  // last level always covers the full radius
  KnnLevels lv;
  lv.n = 0;
  const double r2_max = radius * radius;
  int L_full = (int)ceil(radius / (double)kGraphCell - 1e-9);
  if (L_full < 1) L_full = 1;
  MODEST_REQUIRE(L_full <= 15, "affinity_graph: radius %g spans more than 15 cells", radius);
  // fine levels (windows of +-L cells, tried in order before the full radius) and the window
  // population, in multiples of k, below which a fine level is not even tried.  One level of
  // +-2 cells tried from 2k points up measured best on the Lyft shape (4.14 -> 3.97 ms per
  // 12-scan step against {+-1 cell, from k points}; more levels gain nothing).
  static const int kFineLevels[] = {2};
  const int fine_n = (int)(sizeof(kFineLevels) / sizeof(kFineLevels[0]));
  const int* fine = kFineLevels;
  const double try_factor = 2.0;
  for (int i = 0; i < fine_n; ++i) {
    const int L = fine[i];
    if (L < 1 || L >= L_full) continue;
    lv.L[lv.n] = L;
    lv.r2[lv.n] = ((double)kGraphCell * L) * ((double)kGraphCell * L);
    ++lv.n;
  }
  lv.L[lv.n] = L_full;
  lv.r2[lv.n] = r2_max;
  ++lv.n;
  lv.min_pop = (int)ceil(try_factor * n_neighbors);
  int wblocks = (int)((max_points + 3) / 4);
  if (wblocks < 1) wblocks = 1;
  if (wblocks > 148 * 16) wblocks = 148 * 16;
  MODEST_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int32_t), stream));
  MODEST_CUDA(cudaMemsetAsync(queue_cnt, 0, sizeof(int32_t) * (size_t)n_scans, stream));
  knn_select_fast_kernel<<<dim3(wblocks, n_scans), kKnnWarps * 32, 0, stream>>>(sorted, cells, meta, d_off, G, n_neighbors, r2_max,
                                                                    lv, rk2, knn, knn_cnt, d_flags, queue, queue_cnt);
  MODEST_LAUNCH_CHECK("knn_select_fast_kernel");
  // the deferred points (dense neighbourhoods; normally a few per mille): a narrow grid, it idles when the queues are empty
  knn_select_kernel<<<dim3(64, n_scans), kKnnWarps * 32, 0, stream>>>(sorted, cells, meta, d_off, G, n_neighbors, r2_max, lv, rk2,
                                                                    knn, knn_cnt, d_flags, queue, queue_cnt);
  MODEST_LAUNCH_CHECK("knn_select_kernel");
  int mblocks = (int)((max_points * 32 + 255) / 256);
  if (mblocks < 1) mblocks = 1;
  if (mblocks > 148 * 8) mblocks = 148 * 8;
  mutual_edges_kernel<<<dim3(mblocks, n_scans), 256, 0, stream>>>(reinterpret_cast<const float4*>(d_kept), d_off,
                                                                 d_n_kept, n_neighbors, rk2, knn, knn_cnt, d_nbr,
                                                                 d_nbr_w, d_nbr_cnt, partition_eps, d_nbr_eps_cnt);
  MODEST_LAUNCH_CHECK("mutual_edges_kernel");
  note_launch(3);
  return MODEST_OK;
}

extern "C" size_t modest_dbscan_workspace_bytes(int64_t n_points_total) {
  return align_up((size_t)n_points_total, 256) + 2 * align_up(sizeof(int32_t) * (size_t)n_points_total, 256) + 512;
}

extern "C" int modest_dbscan_batch(const int64_t* d_off, const int32_t* d_n_kept, const int32_t* d_kept_idx,
                                   int n_scans, int64_t n_points_total, int64_t max_points, int n_neighbors,
                                   const int32_t* d_nbr, const float* d_nbr_w, const int32_t* d_nbr_cnt, const int32_t* d_nbr_eps_cnt, double eps,
                                   int min_samples, int32_t* d_labels_kept, int32_t* d_labels_full,
                                   int32_t* d_n_clusters, void* d_ws, size_t ws_bytes, void* stream_) {
  modest::StageRange nvtx_("modest:I DBSCAN");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_scans <= 0 || n_points_total <= 0) return MODEST_OK;
  MODEST_REQUIRE(d_off && d_n_kept && d_kept_idx && d_nbr && (d_nbr_w || d_nbr_eps_cnt) && d_nbr_cnt && d_labels_kept &&
                     d_labels_full && d_n_clusters && d_ws, "dbscan: null pointer argument");
  MODEST_REQUIRE(ws_bytes >= modest_dbscan_workspace_bytes(n_points_total), "dbscan: workspace too small");
  MODEST_REQUIRE(n_scans <= 65535, "dbscan: more than 65535 scans in one launch");
  Arena ar(d_ws, ws_bytes);
  uint8_t* core = ar.take<uint8_t>(n_points_total);
  int32_t* parent = ar.take<int32_t>(n_points_total);
  int32_t* root_rank = ar.take<int32_t>(n_points_total);
  MODEST_REQUIRE(ar.ok(), "workspace too small for the requested sizes");
  int32_t* border_lab = root_rank;
  int pblocks = (int)((max_points + 255) / 256);
  if (pblocks < 1) pblocks = 1;
  dbscan_core_kernel<<<dim3(pblocks, n_scans), 256, 0, stream>>>(d_off, d_n_kept, n_neighbors, d_nbr_w, d_nbr_cnt,
                                                                d_nbr_eps_cnt, eps, min_samples, core, parent);
  MODEST_LAUNCH_CHECK("dbscan_core_kernel");
  int64_t eb = (max_points * 32 + 255) / 256;          // one warp per row
  if (eb < 1) eb = 1;
  if (eb > 148 * 16) eb = 148 * 16;
  dbscan_init_kernel<<<dim3((unsigned)eb, n_scans), 256, 0, stream>>>(d_off, d_n_kept, n_neighbors, d_nbr, d_nbr_w,
                                                                     d_nbr_cnt, d_nbr_eps_cnt, eps, core, parent);
  MODEST_LAUNCH_CHECK("dbscan_init_kernel");
  for (int r = 0; r < 4; ++r) {
    dbscan_jump_kernel<<<dim3(pblocks, n_scans), 256, 0, stream>>>(d_off, d_n_kept, parent);
    MODEST_LAUNCH_CHECK("dbscan_jump_kernel");
  }
  dbscan_union_kernel<<<dim3((unsigned)eb, n_scans), 256, 0, stream>>>(d_off, d_n_kept, n_neighbors, d_nbr, d_nbr_w,
                                                                      d_nbr_cnt, d_nbr_eps_cnt, eps, core, parent);
  MODEST_LAUNCH_CHECK("dbscan_union_kernel");
  dbscan_rank_roots_kernel<<<n_scans, 1024, 0, stream>>>(d_off, d_n_kept, core, parent, root_rank, d_n_clusters);
  MODEST_LAUNCH_CHECK("dbscan_rank_roots_kernel");
  const dim3 pgrid(pblocks, n_scans);
  dbscan_label_cores_kernel<<<pgrid, 256, 0, stream>>>(d_off, d_n_kept, core, parent, root_rank, d_labels_kept, d_labels_full);
  MODEST_LAUNCH_CHECK("dbscan_label_cores_kernel");
  // root_rank entries of non-root points are free: reuse the array for the border labels
  dbscan_label_borders_kernel<<<pgrid, 256, 0, stream>>>(d_off, d_n_kept, d_kept_idx, n_neighbors, d_nbr, d_nbr_w, d_nbr_cnt,
                                                        d_nbr_eps_cnt, eps, core, d_labels_kept, border_lab, d_labels_full);
  MODEST_LAUNCH_CHECK("dbscan_label_borders_kernel");
  dbscan_merge_borders_kernel<<<pgrid, 256, 0, stream>>>(d_off, d_n_kept, core, border_lab, d_labels_kept);
  MODEST_LAUNCH_CHECK("dbscan_merge_borders_kernel");
  note_launch(11);
  return MODEST_OK;
}

extern "C" int modest_knn_bruteforce(const float* d_pts, int point_stride, int n, int n_neighbors, double d2_max,
                                     int32_t* d_knn, int32_t* d_knn_cnt, double* d_rk2, int32_t* d_flags, void* stream_) {
  modest::StageRange nvtx_("modest:f-4 exact kNN lists");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n <= 0) return MODEST_OK;
  MODEST_REQUIRE(d_pts && d_knn && d_knn_cnt && d_rk2 && d_flags, "knn_bruteforce: null pointer argument");
  MODEST_REQUIRE(point_stride >= 3 && n_neighbors >= 1 && d2_max > 0.0, "knn_bruteforce: bad arguments");
  MODEST_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int32_t), stream));
  int blocks = (n + kKnnWarps - 1) / kKnnWarps;
  if (blocks > 148 * 16) blocks = 148 * 16;
  knn_bruteforce_kernel<<<blocks, kKnnWarps * 32, 0, stream>>>(d_pts, point_stride, n, n_neighbors, d2_max, d_knn, d_knn_cnt,
                                                              d_rk2, d_flags);
  MODEST_LAUNCH_CHECK("knn_bruteforce_kernel");
  note_launch(1);
  return MODEST_OK;
}

extern "C" int modest_radius_graph(const float* d_pts, int point_stride, int n, double radius, int64_t* d_counts,
                                   const int64_t* d_indptr, int32_t* d_indices, void* stream_) {
  modest::StageRange nvtx_("modest:f-4 radius graph");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n <= 0) return MODEST_OK;
  MODEST_REQUIRE(d_pts && point_stride >= 3 && radius >= 0.0, "radius_graph: bad arguments");
  MODEST_REQUIRE((d_counts && !d_indices) || (d_indptr && d_indices), "radius_graph: pass d_counts (count call) or d_indptr + d_indices (fill call)");
  int blocks = (n * 32 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  const double r2 = radius * radius;
  if (d_indices) radius_bruteforce_kernel<true><<<blocks, 256, 0, stream>>>(d_pts, point_stride, n, r2, nullptr, d_indptr, d_indices);
  else radius_bruteforce_kernel<false><<<blocks, 256, 0, stream>>>(d_pts, point_stride, n, r2, d_counts, nullptr, nullptr);
  MODEST_LAUNCH_CHECK("radius_bruteforce_kernel");
  note_launch(1);
  return MODEST_OK;
}

extern "C" int modest_edge_affinity(const float* d_pts, int point_stride, int width, const float* d_pp, const int64_t* d_indptr,
                                    const int32_t* d_indices, int n, int kind, double* d_out, void* stream_) {
  modest::StageRange nvtx_("modest:f-4 edge affinity");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n <= 0) return MODEST_OK;
  MODEST_REQUIRE(d_indptr && d_indices && d_out && kind >= 0 && kind <= 2, "edge_affinity: bad arguments");
  MODEST_REQUIRE(kind == 2 ? (d_pts && width >= 1 && width <= point_stride) : d_pp != nullptr, "edge_affinity: missing input for kind %d", kind);
  int blocks = (n * 32 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  edge_affinity_kernel<<<blocks, 256, 0, stream>>>(d_pts, point_stride, width, d_pp, d_indptr, d_indices, n, kind, d_out);
  MODEST_LAUNCH_CHECK("edge_affinity_kernel");
  note_launch(1);
  return MODEST_OK;
}
