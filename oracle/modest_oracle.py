"""TEST INFRASTRUCTURE ONLY -- CPU restatement of MODEST's seed-label hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs
may import this module, and only as the checker or the timed CPU baseline.  The product
(`modest_b200/`) never imports it.

Parity status: PINNED.  The reference ships no tests or golden vectors, so
the pin is the reference itself: `oracle/make_golden.py` runs the unmodified reference modules
(via `oracle/reference_harness.py`, build container only) on seeded synthetic scans, checks
every function below against them, and commits the outputs under tests/golden/.

Two kinds of function live here:

* `*_lib` / default functions call the same third-party routines the reference calls
  (scipy cKDTree, sklearn RANSACRegressor / kneighbors_graph / radius_neighbors_graph /
  DBSCAN, numpy percentile) with the same arguments -- those libraries are un-vendored
  dependencies of the reference (README.md:32-39 pins nothing; installed here: numpy 2.3.5,
  scipy 1.18.1, scikit-learn 1.9.0).
* `*_bruteforce` / `*_restated` functions spell the published semantics of those routines out
  in plain numpy (small inputs only).  They are what the CUDA kernels implement, and the
  tests show both kinds agree.

Every function cites the reference lines it follows (paths relative to
/root/reference/generate_cluster_mask/).
"""
from __future__ import annotations

import math
import types

import numpy as np

# ------------------------------------------------------------------------------------------
# A-B  scan IO and pose chain
# ------------------------------------------------------------------------------------------


def read_velodyne(path):
    """utils/pointcloud_utils.py:22-25 -- flat f32 file viewed as rows of 4."""
    return np.fromfile(path, dtype=np.float32).reshape(-1, 4)


def relative_pose(fixed_l2e, fixed_ego, query_l2e, query_ego, k2n):
    """pre_compute_pp_score.py:27-28 -- K^-1 * l2e_f^-1 * ego_f^-1 * ego_q * l2e_q * K, three
    nested solves in the promoted dtype, result cast to f32."""
    inner = query_ego @ query_l2e @ k2n
    for lhs in (fixed_ego, fixed_l2e, k2n):
        inner = np.linalg.solve(lhs, inner)
    return inner.astype(np.float32)


def apply_pose(xyz, tr):
    """utils/pointcloud_utils.py:11-19 -- homogeneous row-vector product, f32 ones column."""
    hom = np.hstack((xyz, np.ones((xyz.shape[0], 1), dtype=np.float32)))
    return np.dot(hom, tr.T).reshape(-1, 4)[:, :3]


def drop_ego_box(xyz, xr=(-1.15, 1.75), yr=(-0.65, 0.65)):
    """pre_compute_pp_score.py:48-52 (nuScenes history frames only, :141-142)."""
    inside = (xyz[:, 0] >= xr[0]) & (xyz[:, 0] < xr[1]) & (xyz[:, 1] >= yr[0]) & (xyz[:, 1] < yr[1])
    return xyz[~inside]


# ------------------------------------------------------------------------------------------
# C-D  persistence-point score
# ------------------------------------------------------------------------------------------
def neighbor_counts(query_xyz, history, radius=0.3):
    """pre_compute_pp_score.py:54-60,188-193 -- one cKDTree per traversal,
    query_ball_point(..., return_length=True); result (N,T) int64."""
    from scipy.spatial import cKDTree
    cols = [cKDTree(h).query_ball_point(query_xyz[:, :3], r=radius, return_length=True)
            for h in history]
    return np.stack(cols).T


def neighbor_counts_bruteforce(query_xyz, history, radius=0.3, block=512):
    """What cKDTree's p=2 ball query decides, spelt out: coordinates widened f32->f64,
    d2 = dx*dx; d2 += dy*dy; d2 += dz*dz (sequential f64), hit iff d2 <= radius*radius."""
    q = np.asarray(query_xyz[:, :3], dtype=np.float64)
    r2 = float(radius) * float(radius)
    out = np.zeros((q.shape[0], len(history)), dtype=np.int64)
    for t, h in enumerate(history):
        h64 = np.asarray(h, dtype=np.float64)
        for s in range(0, q.shape[0], block):
            qb = q[s:s + block]
            d2 = (qb[:, None, 0] - h64[None, :, 0]) ** 2
            d2 = d2 + (qb[:, None, 1] - h64[None, :, 1]) ** 2
            d2 = d2 + (qb[:, None, 2] - h64[None, :, 2]) ** 2
            out[s:s + block, t] = (d2 <= r2).sum(axis=1)
    return out


def persistence_entropy(counts):
    """pre_compute_pp_score.py:68-75 -- P = c / (sum_t c + 1e-8); H = -sum P ln(P+1e-8) / ln T,
    all f64.  The CLI stores H.astype(f32) (:195-196)."""
    c = np.asarray(counts)
    n_trav = c.shape[1]
    p = c / (c.sum(axis=1)[:, None] + 1e-8)
    return (-p * np.log(p + 1e-8)).sum(axis=1) / np.log(n_trav)


def pp_score(query_xyz, history, radius=0.3):
    return persistence_entropy(neighbor_counts(query_xyz, history, radius)).astype(np.float32)


# ------------------------------------------------------------------------------------------
# E  RANSAC ground plane
# ------------------------------------------------------------------------------------------
def plane_candidates_mask(xyz, max_hs, ptc_range):
    """utils/pointcloud_utils.py:45-49 -- z below max_hs, x/y strictly inside the range."""
    (x0, x1), (y0, y1) = ptc_range
    return ((xyz[:, 2] < max_hs) & (xyz[:, 0] > x0) & (xyz[:, 0] < x1)
            & (xyz[:, 1] > y0) & (xyz[:, 1] < y1))


def _plane_from_linear_model(coef, intercept):
    """utils/pointcloud_utils.py:53-62 -- (a,b,-1,h)/||(a,b,-1)|| negated so that c > 0."""
    w = np.array([coef[0], coef[1], -1.0], dtype=np.float64)
    nrm = np.linalg.norm(w)
    return -np.array([w[0] / nrm, w[1] / nrm, w[2] / nrm, intercept / nrm])


class injected_minimal_sets:
    """Context manager: while active, sklearn's RANSAC trial loop (_ransac.py:485-487) takes its
    minimal sets from `triples` ((n,3) indices into the candidate list, consumed in order)
    instead of drawing them from numpy's global RandomState.  The reference never seeds that
    stream, so ANY sequence of draws is a valid run of it; this is how the
    tests check the GPU's throughput mode (device-drawn minimal sets) against the unmodified
    library: same draws in, same n_trials_ / consensus set / labels out.  `used` counts draws."""

    def __init__(self, triples):
        self.triples = np.asarray(triples, dtype=np.int64).reshape(-1, 3)
        self.used = 0

    def _draw(self, n_population, n_samples, random_state=None, method="auto"):
        assert n_samples == 3 and self.used < len(self.triples), "injected minimal sets exhausted"
        t = self.triples[self.used]
        assert t.min() >= 0 and t.max() < n_population
        self.used += 1
        return t.copy()

    def __enter__(self):
        import sklearn.linear_model._ransac as mod
        self._mod, self._orig = mod, mod.sample_without_replacement
        mod.sample_without_replacement = self._draw
        return self

    def __exit__(self, *exc):
        self._mod.sample_without_replacement = self._orig
        return False


def fit_ground_plane(xyz, max_hs=-1.5, ptc_range=((-20, 70), (-20, 20)), return_model=False, draws=None):
    """utils/pointcloud_utils.py:44-65 with it=1 -- sklearn RANSACRegressor() defaults on
    (x,y)->z of the candidate points; consumes the *global* numpy RNG like the reference
    (or, for the tests of the device-drawn mode, the minimal sets in `draws`)."""
    from sklearn.linear_model import RANSACRegressor
    sel = xyz[plane_candidates_mask(xyz, max_hs, ptc_range)]
    if draws is not None:
        with injected_minimal_sets(draws):
            model = RANSACRegressor().fit(sel[:, [0, 1]], sel[:, 2])
    else:
        model = RANSACRegressor().fit(sel[:, [0, 1]], sel[:, 2])
    plane = _plane_from_linear_model(model.estimator_.coef_, model.estimator_.intercept_)
    return (plane, model) if return_model else plane


def draw_minimal_subset(n_population, n_samples=3, rng=None):
    """sklearn.utils.random.sample_without_replacement(method='auto') restated
    (sklearn/utils/_random.pyx:222-265): ratio < 0.01 -> tracking selection (rejection on
    rng.randint), 0.01..0.99 -> rng.permutation(n)[:k], > 0.99 -> reservoir sampling."""
    rng = np.random.mtrand._rand if rng is None else rng
    ratio = n_samples / n_population if n_population else 1.0
    if 0.01 < ratio < 0.99:
        return rng.permutation(n_population)[:n_samples].astype(np.int64)
    if ratio < 0.2:
        seen, out = set(), np.empty(n_samples, np.int64)
        for i in range(n_samples):
            j = rng.randint(n_population)
            while j in seen:
                j = rng.randint(n_population)
            seen.add(j)
            out[i] = j
        return out
    out = np.arange(n_samples, dtype=np.int64)
    for i in range(n_samples, n_population):
        j = rng.randint(0, i + 1)
        if j < n_samples:
            out[j] = i
    return out


def mad_threshold(z):
    """sklearn/linear_model/_ransac.py:396-398 -- median(|z - median(z)|) in z's dtype."""
    return np.median(np.abs(z - np.median(z)))


def dynamic_max_trials(n_inliers, n_samples, min_samples=3, probability=0.99):
    """sklearn/linear_model/_ransac.py:47-78."""
    eps = np.spacing(1)
    ratio = n_inliers / float(n_samples)
    nom = max(eps, 1 - probability)
    denom = max(eps, 1 - ratio ** min_samples)
    if nom == 1:
        return 0
    if denom == 1:
        return float("inf")
    return abs(float(np.ceil(np.log(nom) / np.log(denom))))


def ransac_restated(xy, z, rng=None, max_trials=100):
    """The RANSACRegressor(LinearRegression) trial loop of sklearn 1.9.0
    (_ransac.py:447-560) with every default the reference relies on: 3-point minimal sets from
    the global RNG, f32 |residual| <= MAD, more-inliers-wins then R^2 tiebreak, dynamic early
    stop at p=0.99, final least-squares refit on the consensus set.

    The minimal-set model is the exact plane through the three points evaluated in f64 and
    rounded to f32 (sklearn solves the same 3x3 system with f32 gelsd), so inlier masks can
    differ from sklearn's only for points within an ulp of the threshold.
    Returns dict(coef, intercept, inlier_mask, n_trials, draws)."""
    rng = np.random.mtrand._rand if rng is None else rng
    n = xy.shape[0]
    thr = mad_threshold(z)
    best_n, best_score, best_mask = 1, -np.inf, None
    trials, draws = 0, []
    while trials < max_trials:
        trials += 1
        idx = draw_minimal_subset(n, 3, rng)
        draws.append(idx)
        coef, icpt = exact_plane_through(xy[idx].astype(np.float64), z[idx].astype(np.float64))
        pred = (xy[:, 0] * np.float32(coef[0]) + xy[:, 1] * np.float32(coef[1])
                + np.float32(icpt)).astype(np.float32)
        mask = np.abs(z - pred) <= thr
        n_in = int(mask.sum())
        if n_in < best_n:
            continue
        zi, pi = z[mask].astype(np.float64), pred[mask].astype(np.float64)
        ss_res = ((zi - pi) ** 2).sum()
        ss_tot = ((zi - zi.mean()) ** 2).sum()
        score = 1.0 - ss_res / ss_tot if ss_tot > 0 else (1.0 if ss_res == 0 else 0.0)
        if n_in == best_n and score < best_score:
            continue
        best_n, best_score, best_mask = n_in, score, mask
        max_trials = min(max_trials, dynamic_max_trials(best_n, n))
    coef, icpt = least_squares_plane(xy[best_mask], z[best_mask])
    return dict(coef=coef, intercept=icpt, inlier_mask=best_mask, n_trials=trials, draws=draws,
                threshold=thr)


def exact_plane_through(xy3, z3):
    """z = a x + b y + c through three points (f64), centred like LinearRegression does."""
    mx, mz = xy3.mean(axis=0), z3.mean()
    a = xy3 - mx
    sol, *_ = np.linalg.lstsq(a, z3 - mz, rcond=None)
    return sol, mz - mx @ sol


def least_squares_plane(xy, z):
    """LinearRegression(fit_intercept=True) on the consensus set: centre, solve the 2x2
    normal equations in f64 (sklearn: f32 gelsd; agreement ~1e-8)."""
    xy64, z64 = xy.astype(np.float64), z.astype(np.float64)
    mx, mz = xy64.mean(axis=0), z64.mean()
    a = xy64 - mx
    g = a.T @ a
    rhs = a.T @ (z64 - mz)
    sol = np.linalg.solve(g, rhs)
    return sol, mz - mx @ sol


# ------------------------------------------------------------------------------------------
# F-G  masks
# ------------------------------------------------------------------------------------------
def signed_plane_distance(xyz, plane):
    """utils/pointcloud_utils.py:76-81 -- (p . n + d)/||n||, f32 points against an f64 plane."""
    return (xyz @ plane[:3] + plane[3]) / np.sqrt((plane[:3] ** 2).sum())


def keep_above_plane(xyz, plane, offset=0.05, only_range=((-30, 30), (-30, 30))):
    """utils/pointcloud_utils.py:68-74 -- a point is dropped only if it is below plane+offset
    AND strictly inside `only_range`."""
    drop = signed_plane_distance(xyz, plane) < offset
    if only_range is not None:
        (x0, x1), (y0, y1) = only_range
        drop &= (xyz[:, 0] > x0) & (xyz[:, 0] < x1) & (xyz[:, 1] > y0) & (xyz[:, 1] < y1)
    return ~drop


def limit_range_mask(ptc, limit_range):
    """generate_mask.py:61-64 -- half-open on the low side, closed on the high side."""
    (x0, x1), (y0, y1) = limit_range
    return (ptc[:, 0] > x0) & (ptc[:, 0] <= x1) & (ptc[:, 1] > y0) & (ptc[:, 1] <= y1)


# ------------------------------------------------------------------------------------------
# H  mutual-kNN / radius graph with |delta pp| weights
# ------------------------------------------------------------------------------------------
def affinity_graph(ptc, pp, n_neighbors=70, radius=2.0):
    """utils/clustering_utils.py:32-60 for neighbor_type='radius_mutual_knn',
    affinity_type='l1': kNN connectivity graph AND its transpose AND the radius graph, then
    data[k] = |pp[row] - pp[col]| evaluated in f32 and stored in the f64 CSR."""
    import scipy.sparse
    import sklearn.neighbors as skn
    knn = skn.kneighbors_graph(ptc[:, :3], n_neighbors=n_neighbors, n_jobs=-1)
    g = knn.multiply(knn.T)
    g = g.multiply(skn.radius_neighbors_graph(ptc[:, :3], radius=radius, n_jobs=-1))
    g = scipy.sparse.csr_matrix(g)
    g.eliminate_zeros()
    rows = np.repeat(np.arange(g.shape[0]), np.diff(g.indptr))
    w = np.abs(pp[rows] - pp[g.indices])            # f32 arithmetic
    return scipy.sparse.csr_matrix((w.astype(np.float64), g.indices, g.indptr), shape=g.shape)


def affinity_graph_variant(ptc, pp, neighbor_type, affinity_type, n_neighbors=70, radius=2.0):
    """utils/clustering_utils.py:7-60 with every branch (the non-default graph / affinity types of
    SURVEY 8(f-4)), the same sklearn / scipy / numpy calls as the reference."""
    from sklearn import neighbors
    import scipy.sparse
    if neighbor_type == "knn":
        graph = neighbors.kneighbors_graph(ptc[:, :3], n_neighbors=n_neighbors)
    elif neighbor_type == "sym_knn":
        graph = neighbors.kneighbors_graph(ptc[:, :3], n_neighbors=n_neighbors)
        graph = graph + graph.T
        graph.eliminate_zeros()
    elif neighbor_type == "mutual_knn":
        graph = neighbors.kneighbors_graph(ptc[:, :3], n_neighbors=n_neighbors)
        graph = graph.multiply(graph.T)
        graph.eliminate_zeros()
    elif neighbor_type == "radius":
        graph = neighbors.radius_neighbors_graph(ptc[:, :3], radius=radius)
    elif neighbor_type == "radius_mutual_knn":
        graph = neighbors.kneighbors_graph(ptc[:, :3], n_neighbors=n_neighbors)
        graph = graph.multiply(graph.T)
        graph = graph.multiply(neighbors.radius_neighbors_graph(ptc[:, :3], radius=radius))
        graph.eliminate_zeros()
    else:
        raise NotImplementedError(neighbor_type)
    graph = scipy.sparse.csr_matrix(graph)
    data = graph.data.copy()
    for r in range(graph.indptr.shape[0] - 1):
        sl = slice(graph.indptr[r], graph.indptr[r + 1])
        if affinity_type == "l1":
            data[sl] = np.abs(pp[r] - pp[graph.indices[sl]])
        elif affinity_type == "exp":
            data[sl] = np.exp((pp[r] - pp[graph.indices[sl]]) ** 2)
        elif affinity_type == "3d_l2_distance":
            data[sl] = np.linalg.norm(ptc[r].reshape(1, -1) - ptc[graph.indices[sl]], axis=1)
        else:
            raise NotImplementedError(affinity_type)
    return scipy.sparse.csr_matrix((data, graph.indices, graph.indptr), shape=graph.shape)


def affinity_edges_bruteforce(ptc, pp, n_neighbors=70, radius=2.0):
    """Same edge set from first principles (small inputs): f64 squared distances
    (dx*dx + dy*dy) + dz*dz, j in kNN_k(i) with self excluded, mutual, d2 <= radius^2.
    Returns a dense boolean adjacency and the f32 weight matrix."""
    x = ptc[:, :3].astype(np.float64)
    n = x.shape[0]
    d2 = (x[:, None, 0] - x[None, :, 0]) ** 2
    d2 = d2 + (x[:, None, 1] - x[None, :, 1]) ** 2
    d2 = d2 + (x[:, None, 2] - x[None, :, 2]) ** 2
    np.fill_diagonal(d2, np.inf)
    k = min(n_neighbors, n - 1)
    kth = np.sort(d2, axis=1)[:, k - 1]
    in_knn = d2 <= kth[:, None]
    adj = in_knn & in_knn.T & (d2 <= radius * radius)
    w = np.abs(pp[:, None] - pp[None, :])
    return adj, w


# ------------------------------------------------------------------------------------------
# I  DBSCAN on the precomputed sparse graph
# ------------------------------------------------------------------------------------------
def dbscan_labels(graph, eps=0.1, min_samples=10):
    """generate_mask.py:77-81."""
    from sklearn import cluster
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return cluster.DBSCAN(metric="precomputed", eps=eps, min_samples=min_samples,
                              n_jobs=-1).fit(graph).labels_


def dbscan_restated(indptr, indices, weights, eps=0.1, min_samples=10):
    """sklearn/cluster/_dbscan.py:427-463 + _dbscan_inner.pyx spelt out: neighbourhood(i) =
    {j : stored edge with weight <= eps} + {i}; core iff |neighbourhood| >= min_samples;
    clusters are the components of core-core edges numbered by their smallest core index;
    a border point takes the smallest cluster id among its core neighbours; the rest -1."""
    n = len(indptr) - 1
    nbrs = []
    for i in range(n):
        sl = slice(indptr[i], indptr[i + 1])
        nbrs.append(indices[sl][weights[sl] <= eps])
    core = np.array([len(nb) + 1 >= min_samples for nb in nbrs])
    parent = np.arange(n)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for i in range(n):
        if core[i]:
            for j in nbrs[i]:
                if core[j]:
                    ra, rb = find(i), find(j)
                    if ra != rb:
                        parent[max(ra, rb)] = min(ra, rb)
    labels = np.full(n, -1, dtype=np.int64)
    roots = sorted({find(i) for i in range(n) if core[i]})
    rank = {r: k for k, r in enumerate(roots)}
    for i in range(n):
        if core[i]:
            labels[i] = rank[find(i)]
    for i in range(n):
        if not core[i]:
            ids = [labels[j] for j in nbrs[i] if core[j]]
            if ids:
                labels[i] = min(ids)
    return labels


# ------------------------------------------------------------------------------------------
# J  cluster filtering
# ------------------------------------------------------------------------------------------
def percentile_f32_restated(values, q):
    """numpy 2.3 np.percentile(values_f32, q) 'linear': the quantile, the virtual index
    (n-1)*q, gamma and the lerp are all evaluated in float32
    (numpy/lib/_function_base_impl.py:4277,4633-4678)."""
    v = np.sort(np.asarray(values, dtype=np.float32))
    qf = np.float32(q) / np.float32(100)
    vi = np.float32(v.size - 1) * qf
    lo = int(np.floor(vi))
    hi = min(lo + 1, v.size - 1)
    g = np.float32(vi - np.float32(lo))
    a, b = v[lo], v[hi]
    d = np.float32(b - a)
    if g >= np.float32(0.5):
        return np.float32(b - np.float32(d * np.float32(np.float32(1) - g)))
    return np.float32(a + np.float32(d * g))


def cluster_is_valid(xyz, pp, plane, min_points=10, max_volume=40, min_volume=0.5,
                     max_min_height=4, min_max_height=0, percentile=10,
                     min_percentile_pp_score=0.7):
    """utils/clustering_utils.py:94-117 (the volume gates are commented out there)."""
    if xyz.shape[0] < min_points:
        return False
    h = signed_plane_distance(xyz, plane)
    if h.min() > max_min_height or h.max() < min_max_height:
        return False
    return not (np.percentile(pp, percentile) > min_percentile_pp_score)


def filter_cluster_labels(ptc, pp, labels, plane_draws=None, info=None, **gates):
    """utils/clustering_utils.py:119-135 -- a SECOND RANSAC plane with hard-coded
    max_hs=-1.5 / range ((-70,70),(-50,50)), invalid clusters -> -1, then ids are
    re-numbered by sorted(set(labels)) (so noise becomes 0 when any noise exists)."""
    out = labels.copy()
    plane, model = fit_ground_plane(ptc, max_hs=-1.5, ptc_range=((-70, 70), (-50, 50)), draws=plane_draws,
                                    return_model=True)
    if info is not None:
        info["n_trials2"] = int(model.n_trials_)
    for cid in range(out.max() + 1):
        member = out == cid
        if not cluster_is_valid(ptc[member, :3], pp[member], plane, **gates):
            out[member] = -1
    uniq = np.unique(out)
    return np.searchsorted(uniq, out).astype(out.dtype), plane


# ------------------------------------------------------------------------------------------
# K  calibration
# ------------------------------------------------------------------------------------------
class Calib:
    """utils/kitti_util.py:232-257,264-279 -- P2, R0_rect, Tr_velo_to_cam from a KITTI txt."""

    def __init__(self, path=None, table=None):
        if table is None:
            table = {}
            with open(path) as fh:
                for line in fh:
                    line = line.rstrip()
                    if not line:
                        continue
                    key, val = line.split(":", 1)
                    try:
                        table[key] = np.array([float(t) for t in val.split()])
                    except ValueError:
                        pass
        self.P = np.asarray(table["P2"], dtype=np.float64).reshape(3, 4)
        self.V2C = np.asarray(table["Tr_velo_to_cam"], dtype=np.float64).reshape(3, 4)
        self.R0 = np.asarray(table["R0_rect"], dtype=np.float64).reshape(3, 3)

    def velo_to_rect(self, xyz):
        """kitti_util.py:293-329 -- R0 (V2C [p,1]) with f64 ones."""
        hom = np.hstack((xyz, np.ones((xyz.shape[0], 1))))
        return (self.R0 @ (hom @ self.V2C.T).T).T

    def rect_to_image(self, xyz):
        """kitti_util.py:334-342."""
        hom = np.hstack((xyz, np.ones((xyz.shape[0], 1))))
        uvw = hom @ self.P.T
        return uvw[:, :2] / uvw[:, 2:3]


# ------------------------------------------------------------------------------------------
# L  closeness-to-edge box fit
# ------------------------------------------------------------------------------------------
SEARCH_ANGLES_DEG = np.arange(0, 90 + 0.1, 0.1)     # pointcloud_utils.py:170 -> 901 values


def _axes(theta):
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[c, s], [-s, c]])


def closeness_fit(xz, d0=1e-2):
    """utils/pointcloud_utils.py:167-216 -- scan 901 headings, score sum 1/max(min(Dx,Dy),d0),
    first strict maximum wins; swap axes if the box is taller than wide; corners back in the
    input frame.  Returns (corners (4,2), angle, area)."""
    best, theta = -math.inf, None
    for deg in SEARCH_ANGLES_DEG:
        ang = deg / 180.0 * np.pi
        pr = xz @ _axes(ang).T
        lo, hi = pr.min(axis=0), pr.max(axis=0)
        edge = np.minimum(pr - lo, hi - pr).min(axis=1)
        score = (1.0 / np.maximum(edge, d0)).sum()
        if score > best:
            best, theta = score, ang
    pr = xz @ _axes(theta).T
    lo, hi = pr.min(axis=0), pr.max(axis=0)
    if (hi[0] - lo[0]) < (hi[1] - lo[1]):
        theta = theta + np.pi / 2
        pr = xz @ _axes(theta).T
        lo, hi = pr.min(axis=0), pr.max(axis=0)
    area = (hi[0] - lo[0]) * (hi[1] - lo[1])
    corners = np.array([[hi[0], lo[1]], [lo[0], lo[1]], [lo[0], hi[1]], [hi[0], hi[1]]]) @ _axes(theta)
    return corners, theta, area


def min_area_fit(points):
    """utils/pointcloud_utils.py:88-146 (minimum_bounding_rectangle), literally: scipy ConvexHull,
    the h-1 edges between consecutive vertices of its list, np.unique of |angle mod pi/2|."""
    from scipy.spatial import ConvexHull
    pi2 = np.pi / 2.
    hull_points = points[ConvexHull(points).vertices]
    edges = hull_points[1:] - hull_points[:-1]
    angles = np.unique(np.abs(np.mod(np.arctan2(edges[:, 1], edges[:, 0]), pi2)))
    rotations = np.vstack([np.cos(angles), np.cos(angles - pi2), np.cos(angles + pi2), np.cos(angles)]).T.reshape((-1, 2, 2))
    rot_points = np.dot(rotations, hull_points.T)
    min_x, max_x = np.nanmin(rot_points[:, 0], axis=1), np.nanmax(rot_points[:, 0], axis=1)
    min_y, max_y = np.nanmin(rot_points[:, 1], axis=1), np.nanmax(rot_points[:, 1], axis=1)
    areas = (max_x - min_x) * (max_y - min_y)
    k = np.argmin(areas)
    x1, x2, y1, y2, r = max_x[k], min_x[k], max_y[k], min_y[k], rotations[k]
    rval = np.array([np.dot([x1, y2], r), np.dot([x2, y2], r), np.dot([x2, y1], r), np.dot([x1, y1], r)])
    return rval, angles[k], areas[k]


def pca_fit(cluster_ptc):
    """utils/pointcloud_utils.py:148-165 (PCA_rectangle) with sklearn's PCA."""
    import sklearn.decomposition
    comp = sklearn.decomposition.PCA(n_components=2).fit(cluster_ptc).components_
    on = cluster_ptc @ comp.T
    min_x, max_x, min_y, max_y = on[:, 0].min(), on[:, 0].max(), on[:, 1].min(), on[:, 1].max()
    rval = np.array([[max_x, min_y], [min_x, min_y], [min_x, max_y], [max_x, max_y]]) @ comp
    return rval, np.arctan2(comp[0, 1], comp[0, 0]), (max_x - min_x) * (max_y - min_y)


def variance_fit(cluster_ptc, delta=0.1):
    """utils/pointcloud_utils.py:219-275 (variance_rectangle)."""
    max_var, choose = -float("inf"), None

    def comps(a):
        return np.array([[np.cos(a), np.sin(a)], [-np.sin(a), np.cos(a)]])
    for angle in np.arange(0, 90 + delta, delta):
        angle = angle / 180. * np.pi
        pr = cluster_ptc @ comps(angle).T
        min_x, max_x, min_y, max_y = pr[:, 0].min(), pr[:, 0].max(), pr[:, 1].min(), pr[:, 1].max()
        dx = np.vstack((pr[:, 0] - min_x, max_x - pr[:, 0])).min(axis=0)
        dy = np.vstack((pr[:, 1] - min_y, max_y - pr[:, 1])).min(axis=0)
        var = 0
        if (dx < dy).sum() > 0:
            var += -np.var(dx[dx < dy])
        if (dy < dx).sum() > 0:
            var += -np.var(dy[dy < dx])
        if var > max_var:
            max_var, choose = var, angle
    angle = choose
    pr = cluster_ptc @ comps(angle).T
    min_x, max_x, min_y, max_y = pr[:, 0].min(), pr[:, 0].max(), pr[:, 1].min(), pr[:, 1].max()
    if (max_x - min_x) < (max_y - min_y):
        angle = choose + np.pi / 2
        pr = cluster_ptc @ comps(angle).T
        min_x, max_x, min_y, max_y = pr[:, 0].min(), pr[:, 0].max(), pr[:, 1].min(), pr[:, 1].max()
    rval = np.array([[max_x, min_y], [min_x, min_y], [min_x, max_y], [max_x, max_y]]) @ comps(angle)
    return rval, angle, (max_x - min_x) * (max_y - min_y)


def fit_box_variant(cluster_rect, all_rect, fit_method):
    """utils/pointcloud_utils.py:292-317 (get_obj) for the three non-default fit methods."""
    xz = cluster_rect[:, [0, 2]]
    corners, ry, area = {"min_zx_area_fit": min_area_fit, "PCA": pca_fit, "variance_to_edge": variance_fit}[fit_method](xz)
    ry = ry * -1
    l = np.linalg.norm(corners[0] - corners[1])
    w = np.linalg.norm(corners[0] - corners[-1])
    c = (corners[0] + corners[2]) / 2
    bottom = lowest_point_in_footprint(all_rect, c, l, w, ry)
    h = bottom - cluster_rect[:, 1].min()
    return types.SimpleNamespace(t=np.array([c[0], bottom, c[1]]), l=l, w=w, h=h, ry=ry, volume=area * h)


def lowest_point_in_footprint(all_rect, centre_xz, length, width, ry):
    """utils/pointcloud_utils.py:278-290 -- max rect-y of ALL scan points strictly inside the
    rotated footprint."""
    c, s = np.cos(ry), np.sin(ry)
    local = (all_rect[:, [0, 2]] - centre_xz) @ np.array([[c, -s], [s, c]]).T
    inside = (np.abs(local[:, 0]) < length / 2) & (np.abs(local[:, 1]) < width / 2)
    # NB: the reference writes the four strict comparisons separately; |v| < a/2 is the same set
    return all_rect[inside, 1].max()


def fit_box(cluster_rect, all_rect):
    """utils/pointcloud_utils.py:292-317 for fit_method='closeness_to_edge'."""
    corners, ang, area = closeness_fit(cluster_rect[:, [0, 2]])
    ry = -ang
    length = np.linalg.norm(corners[0] - corners[1])
    width = np.linalg.norm(corners[0] - corners[3])
    centre = (corners[0] + corners[2]) / 2
    bottom = lowest_point_in_footprint(all_rect, centre, length, width, ry)
    height = bottom - cluster_rect[:, 1].min()
    return types.SimpleNamespace(t=np.array([centre[0], bottom, centre[1]]), l=length, w=width,
                                 h=height, ry=ry, volume=area * height)


# ------------------------------------------------------------------------------------------
# N  rotated BEV IoU + greedy suppression
# ------------------------------------------------------------------------------------------
def boxes_for_nms(objs):
    """utils/pointcloud_utils.py:322-324 -- [t.x, t.z, 0, l, w, h, -ry] rounded to f32."""
    return np.array([[o.t[0], o.t[2], 0, o.l, o.w, o.h, -o.ry] for o in objs]).astype(np.float32)


def greedy_suppress(iou, order, thr):
    """utils/pointcloud_utils.py:330-343 -- visit boxes in `order`; a kept box clears every
    box whose IoU with it exceeds thr (then re-keeps itself)."""
    keep = np.ones(iou.shape[0], dtype=bool)
    for i in order:
        if keep[i]:
            keep[iou[i] > thr] = False
            keep[i] = True
    return keep


def nms_order_from_self_iou(iou):
    """utils/pointcloud_utils.py:335-336 -- descending argsort of the diagonal."""
    return np.diag(iou).argsort()[::-1]


def _f32(x):
    return np.float32(x)


def bev_overlap_f32(a, b):
    """utils/iou3d_nms/src/iou3d_nms_kernel.cu:104-225 (== iou3d_cpu.cpp:97-213) restated in
    numpy float32 scalars without FMA contraction: edge-edge intersections, corners inside
    the other box (1e-2 margin), bubble sort by atan2 around the centroid, shoelace."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)

    def corners(bx):
        hx, hy = bx[3] / _f32(2), bx[4] / _f32(2)
        base = [(bx[0] - hx, bx[1] - hy), (bx[0] + hx, bx[1] - hy),
                (bx[0] + hx, bx[1] + hy), (bx[0] - hx, bx[1] + hy)]
        c, s = _f32(np.cos(bx[6])), _f32(np.sin(bx[6]))
        out = []
        for (px, py) in base:
            dx, dy = px - bx[0], py - bx[1]
            out.append((_f32(dx * c + dy * (-s) + bx[0]), _f32(dx * s + dy * c + bx[1])))
        out.append(out[0])
        return out

    def cross3(p1, p2, p0):
        return _f32((p1[0] - p0[0]) * (p2[1] - p0[1]) - (p2[0] - p0[0]) * (p1[1] - p0[1]))

    def seg_hit(p1, p0, q1, q0):
        if not (min(p0[0], p1[0]) <= max(q0[0], q1[0]) and min(q0[0], q1[0]) <= max(p0[0], p1[0])
                and min(p0[1], p1[1]) <= max(q0[1], q1[1])
                and min(q0[1], q1[1]) <= max(p0[1], p1[1])):
            return None
        s1, s2 = cross3(q0, p1, p0), cross3(p1, q1, p0)
        s3, s4 = cross3(p0, q1, q0), cross3(q1, p1, q0)
        if not (s1 * s2 > 0 and s3 * s4 > 0):
            return None
        s5 = cross3(q1, p1, p0)
        if abs(_f32(s5 - s1)) > _f32(1e-8):
            return (_f32((s5 * q0[0] - s1 * q1[0]) / (s5 - s1)),
                    _f32((s5 * q0[1] - s1 * q1[1]) / (s5 - s1)))
        a0, b0, c0 = p0[1] - p1[1], p1[0] - p0[0], p0[0] * p1[1] - p1[0] * p0[1]
        a1, b1, c1 = q0[1] - q1[1], q1[0] - q0[0], q0[0] * q1[1] - q1[0] * q0[1]
        det = a0 * b1 - a1 * b0
        return (_f32((b0 * c1 - b1 * c0) / det), _f32((a1 * c0 - a0 * c1) / det))

    def inside(bx, p):
        c, s = _f32(np.cos(-bx[6])), _f32(np.sin(-bx[6]))
        rx = (p[0] - bx[0]) * c + (p[1] - bx[1]) * (-s)
        ry = (p[0] - bx[0]) * s + (p[1] - bx[1]) * c
        return abs(rx) < bx[3] / _f32(2) + _f32(1e-2) and abs(ry) < bx[4] / _f32(2) + _f32(1e-2)

    with np.errstate(all="ignore"):
        ca, cb = corners(a), corners(b)
        poly = []
        for i in range(4):
            for j in range(4):
                hit = seg_hit(ca[i + 1], ca[i], cb[j + 1], cb[j])
                if hit is not None:
                    poly.append(hit)
        for k in range(4):
            if inside(a, cb[k]):
                poly.append(cb[k])
            if inside(b, ca[k]):
                poly.append(ca[k])
        n = len(poly)
        if n == 0:
            return _f32(0.0)
        cx = _f32(0)
        cy = _f32(0)
        for p in poly:
            cx, cy = _f32(cx + p[0]), _f32(cy + p[1])
        cx, cy = _f32(cx / _f32(n)), _f32(cy / _f32(n))
        ang = [_f32(np.arctan2(_f32(p[1] - cy), _f32(p[0] - cx))) for p in poly]
        for j in range(n - 1):
            for i in range(n - j - 1):
                if ang[i] > ang[i + 1]:
                    poly[i], poly[i + 1] = poly[i + 1], poly[i]
                    ang[i], ang[i + 1] = ang[i + 1], ang[i]
        area = _f32(0)
        for k in range(n - 1):
            ux, uy = poly[k][0] - poly[0][0], poly[k][1] - poly[0][1]
            vx, vy = poly[k + 1][0] - poly[0][0], poly[k + 1][1] - poly[0][1]
            area = _f32(area + _f32(ux * vy - uy * vx))
        return _f32(abs(area) / _f32(2))


def bev_iou_matrix_f32(boxes_a, boxes_b):
    """iou3d_nms_kernel.cu:227-234,251-265 -- overlap / max(Sa + Sb - overlap, 1e-8)."""
    out = np.zeros((len(boxes_a), len(boxes_b)), np.float32)
    for i, a in enumerate(boxes_a):
        for j, b in enumerate(boxes_b):
            ov = bev_overlap_f32(a, b)
            sa, sb = _f32(a[3] * a[4]), _f32(b[3] * b[4])
            out[i, j] = ov / max(_f32(sa + sb - ov), _f32(1e-8))
    return out


# ------------------------------------------------------------------------------------------
# O  FOV gate and KITTI label text
# ------------------------------------------------------------------------------------------
def box_in_fov(obj, calib, image_shape):
    """utils/pointcloud_utils.py:373-379 -- box centre (half a height above its bottom)
    projects inside the image and lies in front of the camera."""
    centre = obj.t.copy()
    centre[1] -= obj.h / 2
    u, v = calib.rect_to_image(centre.reshape(1, 3))[0]
    return bool(0 <= u < image_shape[1] and 0 <= v < image_shape[0] and centre[2] > 0)


def box_image_extent(obj, P):
    """utils/kitti_util.py:430-478 -- min/max over the 8 projected corners (not clipped)."""
    c, s = np.cos(obj.ry), np.sin(obj.ry)
    rot = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    hl, hw, h = obj.l / 2, obj.w / 2, obj.h
    local = np.array([[hl, hl, -hl, -hl, hl, hl, -hl, -hl],
                      [0, 0, 0, 0, -h, -h, -h, -h],
                      [hw, -hw, -hw, hw, hw, -hw, -hw, hw]])
    pts = rot @ local
    pts[0] += obj.t[0]
    pts[1] += obj.t[1]
    pts[2] += obj.t[2]
    hom = np.hstack((pts.T, np.ones((8, 1))))
    uvw = hom @ P.T
    uv = uvw[:, :2] / uvw[:, 2:3]
    return np.concatenate([uv.min(axis=0), uv.max(axis=0)])


def kitti_label_text(objs, calib, obj_type="Dynamic", with_score=False):
    """utils/pointcloud_utils.py:347-370 -- one '%.4f' line per box, '\\n'-joined, no trailing
    newline."""
    lines = []
    for o in objs:
        alpha = -np.arctan2(o.t[0], o.t[2]) + o.ry
        ext = box_image_extent(o, calib.P)
        vals = [alpha, *ext, o.h, o.w, o.l, o.t[0], o.t[1], o.t[2], o.ry]
        if with_score:
            vals.append(getattr(o, "score", -1))
        lines.append(f"{obj_type} -1 -1 " + " ".join(f"{v:.4f}" for v in vals))
    return "\n".join(lines)


# ------------------------------------------------------------------------------------------
# f-3  road planes (data_preprocessing/RANSAC.py)
# ------------------------------------------------------------------------------------------
def road_plane_for_scan(pc_velo, calib, min_h=1.5, max_h=2.0):
    """data_preprocessing/RANSAC.py:28-52 -- RANSAC on rect (x,z)->y of the points with
    min_h < y < max_h, -10 < z < 70, -20 < x < 20; returns (w (3,), h) as written to the file."""
    from sklearn.linear_model import RANSACRegressor
    rect = calib.velo_to_rect(pc_velo[:, :3])
    ok = ((rect[:, 1] > min_h) & (rect[:, 1] < max_h) & (rect[:, 2] > -10) & (rect[:, 2] < 70)
          & (rect[:, 0] > -20) & (rect[:, 0] < 20))
    rect = rect[ok]
    if len(rect) < 5:
        return np.array([0.0, -1.0, 0.0]), 1.65
    reg = RANSACRegressor().fit(rect[:, [0, 2]], rect[:, 1])
    w = np.array([reg.estimator_.coef_[0], -1.0, reg.estimator_.coef_[1]])
    nrm = np.linalg.norm(w)
    return w / nrm, reg.estimator_.intercept_ / nrm


def road_plane_text(w, h):
    """data_preprocessing/RANSAC.py:60-67."""
    return "\n".join(["# Plane", "Width 4", "Height 1", "{:e} {:e} {:e} {:e}".format(w[0], w[1], w[2], h)])


# ------------------------------------------------------------------------------------------
# f-1  self-training merge (combine_labels.py)
# ------------------------------------------------------------------------------------------
def detections_to_boxes(preds):
    """combine_labels.py:23-34 -- dimensions are stored (l, h, w)."""
    out = []
    for i in range(preds["location"].shape[0]):
        d = preds["dimensions"][i]
        out.append(types.SimpleNamespace(t=preds["location"][i], l=d[0], h=d[1], w=d[2],
                                         ry=preds["rotation_y"][i], score=preds["score"][i]))
    return out


def detection_passes_pp_gate(rect, pp, obj, percentile=50, threshold=0.5):
    """combine_labels.py:42-60 -- PP percentile of the points inside the box (rotated footprint,
    t.y - h < y <= t.y) must not exceed the threshold; an empty box fails."""
    c, s = np.cos(obj.ry), np.sin(obj.ry)
    local = (rect[:, [0, 2]] - obj.t[[0, 2]]) @ np.array([[c, -s], [s, c]]).T
    inside = (np.abs(local[:, 0]) < obj.l / 2) & (np.abs(local[:, 1]) < obj.w / 2)
    inside &= (rect[:, 1] > obj.t[1] - obj.h) & (rect[:, 1] <= obj.t[1])
    return bool(inside.sum() > 0 and not (np.percentile(pp[inside], percentile) > threshold))


def merge_labels_for_scan(preds, seed_objs, rect, pp, calib, iou_fn, image_shape=(1024, 1224), percentile=50,
                          threshold=0.5, score_filtering=-1, nms_threshold=0.1, fov_only=True, with_score=False):
    """combine_labels.py:101-121 for one frame."""
    dets = [o for o in detections_to_boxes(preds)
            if detection_passes_pp_gate(rect, pp, o, percentile, threshold) & (o.score > score_filtering)]
    for o in seed_objs:
        o.score = -999 + o.w * o.l                                   # add_area_score, :37-39
    objs = dets + list(seed_objs)
    if objs:
        iou = iou_fn(boxes_for_nms(objs))
        order = np.argsort([o.score for o in objs])[::-1]
        keep = greedy_suppress(iou, order, nms_threshold)
        objs = [o for o, k in zip(objs, keep) if k]
    if fov_only:
        objs = [o for o in objs if box_in_fov(o, calib, image_shape)]
    return kitti_label_text(objs, calib, with_score=with_score), objs


# ------------------------------------------------------------------------------------------
# Whole-scan drivers (the bodies of the three CLI loops)
# ------------------------------------------------------------------------------------------
DEFAULT_MASK_CFG = dict(
    plane_estimate=dict(range=[[-70, 70], [-20, 20]], max_hs=-1.5, offset=0.05),
    limit_range=[[-70, 70], [-40, 40]],
    graph=dict(neighbor_type="radius_mutual_knn", affinity_type="l1", n_neighbors=70, radius=2.0),
    clustering=dict(method="DBSCAN", DBSCAN=dict(eps=0.1, min_samples=10)),
    filtering=dict(min_points=10, max_volume=120, min_volume=0.5, min_max_height=0.5,
                   max_min_height=1.0, percentile=20, min_percentile_pp_score=0.7),
    bbox_gen=dict(fit_method="closeness_to_edge"),
)


def seed_mask_for_scan(ptc, pp, calib, cfg=None, seed=None, return_stages=False, draws=None):
    """generate_mask.py:52-103 for one scan.  `seed` re-seeds the global numpy RNG first (the
    reference never seeds; SURVEY.md 8(a)-R / 8(d) define seed = 1024 + scan id for parity);
    `draws` = (minimal sets of the first fit, of filter_labels' fit) replaces the stream."""
    cfg = DEFAULT_MASK_CFG if cfg is None else cfg
    if seed is not None:
        np.random.seed(seed)
    pe = cfg["plane_estimate"]
    plane, model1 = fit_ground_plane(ptc[:, :3], max_hs=pe["max_hs"], ptc_range=pe["range"],
                                     draws=None if draws is None else draws[0], return_model=True)
    info = dict(n_trials1=int(model1.n_trials_))
    keep = keep_above_plane(ptc[:, :3], plane, offset=pe["offset"], only_range=pe["range"])
    keep &= limit_range_mask(ptc, cfg["limit_range"])
    g = cfg["graph"]
    graph = affinity_graph(ptc[keep], pp[keep], g["n_neighbors"], g["radius"])
    db = cfg["clustering"]["DBSCAN"]
    raw = np.full(ptc.shape[0], -1, dtype=np.int64)
    raw[keep] = dbscan_labels(graph, db["eps"], db["min_samples"])
    labels, plane2 = filter_cluster_labels(ptc, pp, raw, plane_draws=None if draws is None else draws[1],
                                           info=info, **cfg["filtering"])
    rect = calib.velo_to_rect(ptc[:, :3])
    objs = []
    f = cfg["filtering"]
    for cid in range(1, labels.max() + 1):
        box = fit_box(rect[labels == cid], rect)
        if f["min_volume"] < box.volume < f["max_volume"]:
            objs.append(box)
        else:
            labels[labels == cid] = 0
    uniq = np.unique(labels)
    labels = np.searchsorted(uniq, labels).astype(labels.dtype)
    if return_stages:
        return labels, objs, dict(plane=plane, keep=keep, raw=raw, plane2=plane2, graph=graph, **info)
    return labels, objs


def labels_for_scan(objs, calib, iou_fn, image_shape=(1024, 1224), nms_threshold=0.1,
                    nms=True, fov_only=True):
    """gen_label_files.py:41-52 for one scan; `iou_fn(boxes_f32) -> (K,K) f32` is the BEV IoU
    (the reference's CUDA op on a GPU box, `bev_iou_matrix_f32` elsewhere)."""
    if nms and len(objs) > 0:
        iou = iou_fn(boxes_for_nms(objs))
        keep = greedy_suppress(iou, nms_order_from_self_iou(iou), nms_threshold)
        objs = [o for o, k in zip(objs, keep) if k]
    if fov_only:
        objs = [o for o in objs if box_in_fov(o, calib, image_shape)]
    return kitti_label_text(objs, calib), objs
