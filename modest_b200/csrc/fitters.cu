// the non-default rectangle fitters of get_obj() (utils/pointcloud_utils.py:88-165,
// 219-275) and get_lowest_point_rect() (:278-290) as operator-level kernels -- one cluster per
// call, not on the seed-label path (the configs use closeness_to_edge, csrc/boxes.cu).
//
//   variance_to_edge  901 headings; per heading the points nearer an x-edge / a y-edge of the
//                     bounding rectangle, score -var(Ex) - var(Ey) with numpy's np.var arithmetic
//                     (two pairwise-summed passes), first strict maximum; bit-faithful to numpy.
//   PCA               principal axes of the 2x2 covariance (closed-form symmetric eigensolver,
//                     sklearn's sign convention); agrees with sklearn to rounding, not to the bit
//                     (sklearn accumulates the covariance through BLAS).
//   min_zx_area_fit   convex hull by gift wrapping, then the smallest bounding rectangle with a
//                     side on a hull edge.  The reference skips ONE hull edge -- the one that
//                     closes scipy's vertex list, whose start is a qhull internal -- so its
//                     result can be a slightly larger rectangle when that edge is the optimal
//                     one; every edge is tried here (documented deviation).
#include "common.cuh"

namespace modest {
extern void note_launch(int n);

__device__ __forceinline__ void project_xz(double x, double z, double c, double s, double* px, double* py) {
  // cluster_ptc @ [[c, s], [-s, c]].T  -- dgemm rounding: fma(second term, first product) (see boxes.cu)
  *px = __fma_rn(z, s, __dmul_rn(x, c));
  *py = __fma_rn(z, c, __dmul_rn(x, -s));
}

// numpy's pairwise summation of a contiguous double array (numpy/_core/src/umath/loops_utils.h.src:
// pairwise_sum): n < 8 sequential from the first element; n <= 128 eight running sums combined as
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) then the tail; longer arrays split at n/2 rounded down to a
// multiple of 8.  Iterative over the 128-blocks in the order the recursion visits them would change
// the association, so the recursion is kept (depth log2(n/128)).
__device__ double np_pairwise_sum(const double* a, int n) {
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
  }
  if (n <= 128) {
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return __dadd_rn(np_pairwise_sum(a, n2), np_pairwise_sum(a + n2, n - n2));
}

// np.var(a) for a contiguous f64 array of n >= 1 elements, a is overwritten (scratch)
__device__ double np_var_inplace(double* a, int n) {
  const double mean = __ddiv_rn(np_pairwise_sum(a, n), (double)n);
  for (int i = 0; i < n; ++i) {
    const double d = __dsub_rn(a[i], mean);
    a[i] = __dmul_rn(d, d);
  }
  return __ddiv_rn(np_pairwise_sum(a, n), (double)n);
}

// one THREAD per heading: the numpy reductions are sequential recipes, and a cluster is small
__global__ void __launch_bounds__(128) variance_score_kernel(const double* __restrict__ xz, int n, const double* __restrict__ trig,
                                                             int n_angles, double* __restrict__ scratch /* (n_angles, 2n) */,
                                                             double* __restrict__ score) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_angles) return;
  const double cs = trig[a], sn = trig[n_angles + a];
  double lox = 1e300, hix = -1e300, loy = 1e300, hiy = -1e300;
  for (int i = 0; i < n; ++i) {
    double px, py;
    project_xz(xz[2 * i], xz[2 * i + 1], cs, sn, &px, &py);
    lox = fmin(lox, px); hix = fmax(hix, px); loy = fmin(loy, py); hiy = fmax(hiy, py);
  }
  double* ex = scratch + (size_t)a * 2 * n;
  double* ey = ex + n;
  int nx = 0, ny = 0;
  for (int i = 0; i < n; ++i) {
    double px, py;
    project_xz(xz[2 * i], xz[2 * i + 1], cs, sn, &px, &py);
    const double dx = fmin(__dsub_rn(px, lox), __dsub_rn(hix, px));
    const double dy = fmin(__dsub_rn(py, loy), __dsub_rn(hiy, py));
    if (dx < dy) ex[nx++] = dx;
    if (dy < dx) ey[ny++] = dy;
  }
  double var = 0.0;                                   // pointcloud_utils.py:239-243
  if (nx > 0) var = __dadd_rn(var, -np_var_inplace(ex, nx));
  if (ny > 0) var = __dadd_rn(var, -np_var_inplace(ey, ny));
  score[a] = var;
}

struct RectOut { double corners[8]; double angle, area, status; };

// extents of the cluster at heading (cs, sn); block-wide (256 threads)
__device__ void block_extents(const double* xz, int n, double cs, double sn, double* ext /*lox,hix,loy,hiy*/, double (*sh)[8]) {
  double lox = 1e300, hix = -1e300, loy = 1e300, hiy = -1e300;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double px, py;
    project_xz(xz[2 * i], xz[2 * i + 1], cs, sn, &px, &py);
    lox = fmin(lox, px); hix = fmax(hix, px); loy = fmin(loy, py); hiy = fmax(hiy, py);
  }
  lox = warp_min(lox); loy = warp_min(loy); hix = warp_max(hix); hiy = warp_max(hiy);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) { sh[0][w] = lox; sh[1][w] = hix; sh[2][w] = loy; sh[3][w] = hiy; }
  __syncthreads();
  lox = sh[0][0]; hix = sh[1][0]; loy = sh[2][0]; hiy = sh[3][0];
  for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
    lox = fmin(lox, sh[0][k]); hix = fmax(hix, sh[1][k]); loy = fmin(loy, sh[2][k]); hiy = fmax(hiy, sh[3][k]);
  }
  ext[0] = lox; ext[1] = hix; ext[2] = loy; ext[3] = hiy;
}

// rval @ components with components = [[c, s], [-s, c]] (the layout of :196-201 / :258-263)
__device__ void write_corners(const double* ext, double cs, double sn, RectOut* o) {
  auto corner = [&](double a, double b, double* ox, double* oy) {
    *ox = __fma_rn(b, -sn, __dmul_rn(a, cs));
    *oy = __fma_rn(b, cs, __dmul_rn(a, sn));
  };
  corner(ext[1], ext[2], &o->corners[0], &o->corners[1]);
  corner(ext[0], ext[2], &o->corners[2], &o->corners[3]);
  corner(ext[0], ext[3], &o->corners[4], &o->corners[5]);
  corner(ext[1], ext[3], &o->corners[6], &o->corners[7]);
  o->area = __dmul_rn(__dsub_rn(ext[1], ext[0]), __dsub_rn(ext[3], ext[2]));
}

// first strict maximum over the headings, then the rectangle at that heading (or heading + pi/2
// when the box is taller than wide), pointcloud_utils.py:244-275
__global__ void __launch_bounds__(256) variance_finish_kernel(const double* __restrict__ xz, int n, const double* __restrict__ trig,
                                                              const double* __restrict__ angles, int n_angles,
                                                              const double* __restrict__ score, RectOut* __restrict__ out) {
  __shared__ double sh[4][8];
  __shared__ int s_sel;
  if (threadIdx.x == 0) {
    double best = -1e308 * 10.0;                       // -inf
    int sel = 0;
    for (int a = 0; a < n_angles; ++a)
      if (score[a] > best) { best = score[a]; sel = a; }
    s_sel = sel;
  }
  __syncthreads();
  const int sel = s_sel;
  double ext[4];
  int variant = 0;
  block_extents(xz, n, trig[sel], trig[n_angles + sel], ext, sh);
  if (__dsub_rn(ext[1], ext[0]) < __dsub_rn(ext[3], ext[2])) {
    variant = 1;
    block_extents(xz, n, trig[2 * n_angles + sel], trig[3 * n_angles + sel], ext, sh);
  }
  if (threadIdx.x == 0) {
    write_corners(ext, trig[2 * variant * n_angles + sel], trig[(2 * variant + 1) * n_angles + sel], out);
    out->angle = angles[variant * n_angles + sel];
    out->status = 0.0;
  }
}

// PCA_rectangle (:148-165): components = principal axes (largest variance first), rows made
// sign-deterministic like sklearn.utils.extmath.svd_flip(u_based_decision=False): the entry of
// largest magnitude in each row is positive.
__global__ void __launch_bounds__(256) pca_rect_kernel(const double* __restrict__ xz, int n, RectOut* __restrict__ out) {
  __shared__ double sh[4][8];
  __shared__ double s_c[4];
  double sx = 0, sz = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { sx += xz[2 * i]; sz += xz[2 * i + 1]; }
  sx = warp_sum(sx); sz = warp_sum(sz);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sh[0][w] = sx; sh[1][w] = sz; }
  __syncthreads();
  double mx = 0, mz = 0;
  for (int k = 0; k < 8; ++k) { mx += sh[0][k]; mz += sh[1][k]; }
  mx /= n; mz /= n;
  __syncthreads();
  double cxx = 0, cxz = 0, czz = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double dx = xz[2 * i] - mx, dz = xz[2 * i + 1] - mz;
    cxx += dx * dx; cxz += dx * dz; czz += dz * dz;
  }
  cxx = warp_sum(cxx); cxz = warp_sum(cxz); czz = warp_sum(czz);
  if (lane == 0) { sh[0][w] = cxx; sh[1][w] = cxz; sh[2][w] = czz; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0, c = 0;
    for (int k = 0; k < 8; ++k) { a += sh[0][k]; b += sh[1][k]; c += sh[2][k]; }
    // symmetric 2x2 [[a, b], [b, c]]: eigenvector of the larger eigenvalue (LAPACK dlaev2's recipe)
    const double sm = a + c, df = a - c, adf = fabs(df), tb = b + b, ab = fabs(tb);
    double rt;
    if (adf > ab) rt = adf * sqrt(1.0 + (ab / adf) * (ab / adf));
    else if (adf < ab) rt = ab * sqrt(1.0 + (adf / ab) * (adf / ab));
    else rt = ab * sqrt(2.0);
    (void)sm;
    double cs, ct, tn, cs1, sn1;
    int sgn2;
    if (df >= 0.0) { cs = df + rt; sgn2 = 1; } else { cs = df - rt; sgn2 = -1; }
    if (fabs(cs) > ab) { ct = -tb / cs; sn1 = 1.0 / sqrt(1.0 + ct * ct); cs1 = ct * sn1; }
    else if (ab == 0.0) { cs1 = 1.0; sn1 = 0.0; }
    else { tn = -cs / tb; cs1 = 1.0 / sqrt(1.0 + tn * tn); sn1 = tn * cs1; }
    const int sgn1 = sm < 0.0 ? -1 : 1;
    if (sgn1 == sgn2) { tn = cs1; cs1 = -sn1; sn1 = tn; }
    // (cs1, sn1) is the unit eigenvector of the eigenvalue of larger ABSOLUTE value = the larger one (PSD matrix)
    double v0x = cs1, v0z = sn1, v1x = -sn1, v1z = cs1;
    if (fabs(v0x) >= fabs(v0z) ? v0x < 0 : v0z < 0) { v0x = -v0x; v0z = -v0z; }
    if (fabs(v1x) >= fabs(v1z) ? v1x < 0 : v1z < 0) { v1x = -v1x; v1z = -v1z; }
    s_c[0] = v0x; s_c[1] = v0z; s_c[2] = v1x; s_c[3] = v1z;
  }
  __syncthreads();
  const double c00 = s_c[0], c01 = s_c[1], c10 = s_c[2], c11 = s_c[3];
  // on_component_ptc = cluster_ptc @ components.T
  double lox = 1e300, hix = -1e300, loy = 1e300, hiy = -1e300;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double px = __fma_rn(xz[2 * i + 1], c01, __dmul_rn(xz[2 * i], c00));
    const double py = __fma_rn(xz[2 * i + 1], c11, __dmul_rn(xz[2 * i], c10));
    lox = fmin(lox, px); hix = fmax(hix, px); loy = fmin(loy, py); hiy = fmax(hiy, py);
  }
  lox = warp_min(lox); loy = warp_min(loy); hix = warp_max(hix); hiy = warp_max(hiy);
  __syncthreads();
  if (lane == 0) { sh[0][w] = lox; sh[1][w] = hix; sh[2][w] = loy; sh[3][w] = hiy; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) { lox = fmin(lox, sh[0][k]); hix = fmax(hix, sh[1][k]); loy = fmin(loy, sh[2][k]); hiy = fmax(hiy, sh[3][k]); }
    // rval = [[max_x,min_y],[min_x,min_y],[min_x,max_y],[max_x,max_y]] @ components
    auto corner = [&](double a, double b, double* ox, double* oy) {
      *ox = __fma_rn(b, c10, __dmul_rn(a, c00));
      *oy = __fma_rn(b, c11, __dmul_rn(a, c01));
    };
    corner(hix, loy, &out->corners[0], &out->corners[1]);
    corner(lox, loy, &out->corners[2], &out->corners[3]);
    corner(lox, hiy, &out->corners[4], &out->corners[5]);
    corner(hix, hiy, &out->corners[6], &out->corners[7]);
    out->area = (hix - lox) * (hiy - loy);
    out->angle = atan2(c01, c00);
    out->status = 0.0;
  }
}

// minimum_bounding_rectangle (:88-146).  Hull by gift wrapping (counterclockwise from the lowest of
// the leftmost points), candidate headings |atan2(edge) mod pi/2| in ascending order, first minimum.
constexpr int kHullCap = 1024;
__global__ void __launch_bounds__(256) min_area_rect_kernel(const double* __restrict__ xz, int n, RectOut* __restrict__ out) {
  __shared__ int hull[kHullCap];
  __shared__ double ang[kHullCap];
  __shared__ double s_v[8];
  __shared__ int s_i[8];
  __shared__ int s_h;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // leftmost (then lowest) point
  auto block_best = [&](double key, double key2, int idx) {      // lexicographic minimum of (key, key2, idx)
    for (int o = 16; o > 0; o >>= 1) {
      const double k1 = __shfl_xor_sync(0xffffffffu, key, o), k2 = __shfl_xor_sync(0xffffffffu, key2, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (k1 < key || (k1 == key && (k2 < key2 || (k2 == key2 && oi < idx)))) { key = k1; key2 = k2; idx = oi; }
    }
    __syncthreads();
    if (lane == 0) { s_v[w] = key; ang[w] = key2; s_i[w] = idx; }   // ang[0..7] borrowed as scratch before the hull exists
    __syncthreads();
    double bk = s_v[0], bk2 = ang[0];
    int bi = s_i[0];
    for (int k = 1; k < 8; ++k)
      if (s_v[k] < bk || (s_v[k] == bk && (ang[k] < bk2 || (ang[k] == bk2 && s_i[k] < bi)))) { bk = s_v[k]; bk2 = ang[k]; bi = s_i[k]; }
    __syncthreads();
    return bi;
  };
  double k1 = 1e300, k2 = 1e300;
  int ki = 0x7fffffff;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    if (xz[2 * i] < k1 || (xz[2 * i] == k1 && xz[2 * i + 1] < k2)) { k1 = xz[2 * i]; k2 = xz[2 * i + 1]; ki = i; }
  const int start = block_best(k1, k2, ki);
  int cur = start, h = 0;
  while (h < kHullCap) {
    if (threadIdx.x == 0) hull[h] = cur;
    ++h;
    // next vertex: the point q such that every other point is to the left of cur -> q (most clockwise turn);
    // among collinear candidates the farthest.  Each thread keeps its best, then a tournament.
    const double cx = xz[2 * cur], cz = xz[2 * cur + 1];
    int best = -1;
    double bx = 0, bz = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      if (i == cur) continue;
      const double qx = xz[2 * i] - cx, qz = xz[2 * i + 1] - cz;
      if (qx == 0.0 && qz == 0.0) continue;                         // duplicate of the current vertex
      if (best < 0) { best = i; bx = qx; bz = qz; continue; }
      const double cr = bx * qz - bz * qx;                         // > 0: q is to the left of cur->best
      if (cr < 0.0 || (cr == 0.0 && qx * qx + qz * qz > bx * bx + bz * bz)) { best = i; bx = qx; bz = qz; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const int ob = __shfl_xor_sync(0xffffffffu, best, o);
      const double ox = __shfl_xor_sync(0xffffffffu, bx, o), oz = __shfl_xor_sync(0xffffffffu, bz, o);
      if (ob >= 0) {
        const double cr = bx * oz - bz * ox;
        if (best < 0 || cr < 0.0 || (cr == 0.0 && ox * ox + oz * oz > bx * bx + bz * bz)) { best = ob; bx = ox; bz = oz; }
      }
    }
    __syncthreads();
    if (lane == 0) { s_i[w] = best; s_v[w] = bx; ang[kHullCap - 8 + w] = bz; }
    __syncthreads();
    best = s_i[0]; bx = s_v[0]; bz = ang[kHullCap - 8];
    for (int k = 1; k < 8; ++k) {
      const int ob = s_i[k];
      if (ob < 0) continue;
      const double ox = s_v[k], oz = ang[kHullCap - 8 + k];
      const double cr = bx * oz - bz * ox;
      if (best < 0 || cr < 0.0 || (cr == 0.0 && ox * ox + oz * oz > bx * bx + bz * bz)) { best = ob; bx = ox; bz = oz; }
    }
    __syncthreads();
    if (best < 0 || best == start) break;
    cur = best;
  }
  if (threadIdx.x == 0) s_h = h;
  __syncthreads();
  h = s_h;
  if (h >= kHullCap - 8) { if (threadIdx.x == 0) out->status = 2.0; return; }   // hull too large for the scratch
  const double pi2 = 1.5707963267948966;
  // the walk above turns clockwise; edges as in the reference: hull[k+1] - hull[k] (the closing edge included here)
  for (int k = threadIdx.x; k < h; k += blockDim.x) {
    const int a = hull[k], b = hull[(k + 1) % h];
    const double a0 = atan2(xz[2 * b + 1] - xz[2 * a + 1], xz[2 * b] - xz[2 * a]);
    double r = fmod(a0, pi2);                                         // np.mod: C fmod, then the sign of the divisor
    if (r != 0.0 && r < 0.0) r += pi2;
    ang[k] = fabs(r);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // np.unique: ascending, duplicates dropped (insertion sort: a hull has tens of vertices)
    for (int i = 1; i < h; ++i) {
      const double v = ang[i];
      int j = i - 1;
      while (j >= 0 && ang[j] > v) { ang[j + 1] = ang[j]; --j; }
      ang[j + 1] = v;
    }
    int m = 0;
    for (int i = 0; i < h; ++i)
      if (m == 0 || ang[i] != ang[m - 1]) ang[m++] = ang[i];
    double best_area = 1e300, bx1 = 0, bx2 = 0, by1 = 0, by2 = 0, br[4] = {1, 0, 0, 1}, bang = 0;
    for (int t = 0; t < m; ++t) {
      const double a = ang[t];
      const double r00 = cos(a), r01 = cos(a - pi2), r10 = cos(a + pi2), r11 = cos(a);
      double lox = 1e300, hix = -1e300, loy = 1e300, hiy = -1e300;
      for (int k = 0; k < h; ++k) {
        const double x = xz[2 * hull[k]], z = xz[2 * hull[k] + 1];
        const double px = __fma_rn(z, r01, __dmul_rn(x, r00)), py = __fma_rn(z, r11, __dmul_rn(x, r10));
        lox = fmin(lox, px); hix = fmax(hix, px); loy = fmin(loy, py); hiy = fmax(hiy, py);
      }
      const double area = __dmul_rn(__dsub_rn(hix, lox), __dsub_rn(hiy, loy));
      if (area < best_area) { best_area = area; bx1 = hix; bx2 = lox; by1 = hiy; by2 = loy; br[0] = r00; br[1] = r01; br[2] = r10; br[3] = r11; bang = a; }
    }
    // rval[k] = np.dot([x, y], r)
    auto corner = [&](double x, double y, double* ox, double* oy) {
      *ox = __fma_rn(y, br[2], __dmul_rn(x, br[0]));
      *oy = __fma_rn(y, br[3], __dmul_rn(x, br[1]));
    };
    corner(bx1, by2, &out->corners[0], &out->corners[1]);
    corner(bx2, by2, &out->corners[2], &out->corners[3]);
    corner(bx2, by1, &out->corners[4], &out->corners[5]);
    corner(bx1, by1, &out->corners[6], &out->corners[7]);
    out->angle = bang;
    out->area = best_area;
    out->status = 0.0;
  }
}

// get_lowest_point_rect (:278-290): max rect-y of the points strictly inside the footprint
__global__ void __launch_bounds__(256) lowest_point_kernel(const double* __restrict__ rect, int n, double cx, double cz, double c, double s,
                                                           double hl, double hw, unsigned long long* __restrict__ bottom) {
  unsigned long long best = 0ull;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double dx = __dsub_rn(rect[3 * i], cx), dz = __dsub_rn(rect[3 * i + 2], cz);
    // (ptc_xz - c) @ [[cos, -sin], [sin, cos]].T
    const double lx = __fma_rn(dz, -s, __dmul_rn(dx, c));
    const double ly = __fma_rn(dz, c, __dmul_rn(dx, s));
    if (lx > -hl && lx < hl && ly > -hw && ly < hw) {
      const unsigned long long v = f64_ordered(rect[3 * i + 1]);
      best = v > best ? v : best;
    }
  }
  if (best) atomicMax(bottom, best);
}

// ordered-u64 image -> the double it stands for; NaN when no point was inside (ys.max() of an empty
// array raises in the reference)
__global__ void lowest_point_decode_kernel(unsigned long long* bottom) {
  const unsigned long long u = *bottom;
  const double v = u ? f64_from_ordered(u) : __longlong_as_double(0x7ff8000000000000ll);
  *reinterpret_cast<double*>(bottom) = v;
}

}  // namespace modest

using namespace modest;

extern "C" size_t modest_fit_rectangle_workspace_bytes(int n, int n_angles) {
  return sizeof(double) * ((size_t)n_angles * 2 * (size_t)(n > 0 ? n : 1) + (size_t)n_angles) + 256;
}

extern "C" int modest_fit_rectangle(const double* d_xz, int n, int method, const double* d_trig, const double* d_angles,
                                    int n_angles, double* d_out, void* d_ws, size_t ws_bytes, void* stream_) {
  modest::StageRange nvtx_("modest:f-4 rectangle fit");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MODEST_REQUIRE(d_xz && d_out && n >= 1, "fit_rectangle: bad arguments");
  MODEST_REQUIRE(method >= 0 && method <= 2, "fit_rectangle: method %d (0 min area, 1 PCA, 2 variance_to_edge)", method);
  static_assert(sizeof(RectOut) == 11 * sizeof(double), "RectOut is (corners 8, angle, area, status)");
  RectOut* out = reinterpret_cast<RectOut*>(d_out);
  if (method == 0) {
    min_area_rect_kernel<<<1, 256, 0, stream>>>(d_xz, n, out);
    MODEST_LAUNCH_CHECK("min_area_rect_kernel");
    note_launch(1);
  } else if (method == 1) {
    pca_rect_kernel<<<1, 256, 0, stream>>>(d_xz, n, out);
    MODEST_LAUNCH_CHECK("pca_rect_kernel");
    note_launch(1);
  } else {
    MODEST_REQUIRE(d_trig && d_angles && n_angles >= 1 && d_ws, "fit_rectangle: variance_to_edge needs the angle tables and a workspace");
    MODEST_REQUIRE(ws_bytes >= modest_fit_rectangle_workspace_bytes(n, n_angles), "fit_rectangle: workspace too small");
    double* scratch = static_cast<double*>(d_ws);
    double* score = scratch + (size_t)n_angles * 2 * n;
    MODEST_CUDA(cudaDeviceSetLimit(cudaLimitStackSize, 4096));         // np_pairwise_sum recurses log2(n/128) deep
    variance_score_kernel<<<(n_angles + 127) / 128, 128, 0, stream>>>(d_xz, n, d_trig, n_angles, scratch, score);
    MODEST_LAUNCH_CHECK("variance_score_kernel");
    variance_finish_kernel<<<1, 256, 0, stream>>>(d_xz, n, d_trig, d_angles, n_angles, score, out);
    MODEST_LAUNCH_CHECK("variance_finish_kernel");
    note_launch(2);
  }
  return MODEST_OK;
}

extern "C" int modest_lowest_point_rect(const double* d_rect, int n, double cx, double cz, double cos_ry, double sin_ry,
                                        double l, double w, double* d_bottom, void* stream_) {
  modest::StageRange nvtx_("modest:L lowest point in footprint");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MODEST_REQUIRE(d_rect && d_bottom && n >= 0, "lowest_point_rect: bad arguments");
  MODEST_CUDA(cudaMemsetAsync(d_bottom, 0, sizeof(double), stream));   // ordered-u64 image 0 = "no point inside"
  if (n > 0) {
    int blocks = (n + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    lowest_point_kernel<<<blocks, 256, 0, stream>>>(d_rect, n, cx, cz, cos_ry, sin_ry, l / 2.0, w / 2.0,
                                                    reinterpret_cast<unsigned long long*>(d_bottom));
    MODEST_LAUNCH_CHECK("lowest_point_kernel");
    note_launch(1);
  }
  lowest_point_decode_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<unsigned long long*>(d_bottom));
  MODEST_LAUNCH_CHECK("lowest_point_decode_kernel");
  return MODEST_OK;
}
