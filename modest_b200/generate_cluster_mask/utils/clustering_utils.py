"""Drop-in for the reference's utils/clustering_utils.py (same names / arguments / return
types) on top of libmodest_b200.  The configured default (neighbor_type='radius_mutual_knn',
affinity_type='l1') takes the grid kernels of the seed-label path; the other graph types of
clustering_utils.py:16-31 ('knn', 'sym_knn', 'mutual_knn', 'radius') and affinities of :49-56
('exp', '3d_l2_distance') take exact brute-force kernels.  There is no CPU
fallback: distances, neighbour selection and edge weights are computed on the GPU; only the
sparse-pattern bookkeeping (transpose / union / intersection of index sets, which the reference
does with scipy as well) happens on the host."""
import ctypes as C

import numpy as np
import scipy.sparse
import torch

from modest_b200 import _lib
from modest_b200 import pipeline as _pl
from .pointcloud_utils import _as_batch, _pipe, distance_to_plane, estimate_plane  # noqa: F401

_NEIGHBOR_TYPES = ('knn', 'sym_knn', 'mutual_knn', 'radius', 'radius_mutual_knn')
_AFFINITY_KINDS = {'l1': 0, 'exp': 1, '3d_l2_distance': 2}


def _grid_mutual_graph(ptc, pp_score, n_neighbors, radius):
    """(indptr, indices, l1 weights) of the radius_mutual_knn graph from the seed-label kernels."""
    n = ptc.shape[0]
    pipe = _pl.SeedLabelPipeline(dict(graph=dict(neighbor_type='radius_mutual_knn', affinity_type='l1',
                                                 n_neighbors=int(n_neighbors), radius=float(radius))))
    kept = np.zeros((n, 4), dtype=np.float32)
    kept[:, :3] = ptc[:, :3]
    kept[:, 3] = pp_score
    kept_d = torch.from_numpy(kept).cuda()
    off = torch.tensor([0, n], dtype=torch.int64, device="cuda")
    n_kept = torch.tensor([n], dtype=torch.int32, device="cuda")
    nbr, nbr_w, nbr_cnt, flags = pipe.affinity_graph(kept_d, off, n_kept, 1, n, n)
    k = int(n_neighbors)
    cnt = nbr_cnt.cpu().numpy()[:n]
    idx = nbr.cpu().numpy()[:n * k].reshape(n, k)
    w = nbr_w.cpu().numpy()[:n * k].reshape(n, k)
    take = np.arange(k)[None, :] < cnt[:, None]
    order = np.argsort(np.where(take, idx, np.iinfo(np.int32).max), axis=1, kind="stable")
    idx_s, w_s = np.take_along_axis(idx, order, 1), np.take_along_axis(w, order, 1)
    take_s = np.arange(k)[None, :] < cnt[:, None]
    indptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    return indptr, idx_s[take_s].astype(np.int64), w_s[take_s].astype(np.float64)


def _knn_pattern(xyz_d, n, k):
    """kneighbors_graph(X, k) as a 0/1 CSR (columns sorted): exact k nearest, self excluded."""
    lib = _lib.lib()
    h = xyz_d[:, :3]
    ext = (h.max(dim=0).values - h.min(dim=0).values).double()
    d2_max = float((ext * ext).sum().item()) * (1.0 + 1e-9) + 1e-12
    knn = torch.empty((n, k), dtype=torch.int32, device="cuda")
    cnt = torch.empty(n, dtype=torch.int32, device="cuda")
    rk2 = torch.empty(n, dtype=torch.float64, device="cuda")
    flags = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(lib.modest_knn_bruteforce(_lib.ptr(xyz_d), int(xyz_d.shape[1]), n, k, d2_max, _lib.ptr(knn), _lib.ptr(cnt),
                                         _lib.ptr(rk2), _lib.ptr(flags), _lib.stream_ptr()), "modest_knn_bruteforce")
    if int(flags.item()) & 2:
        import warnings
        warnings.warn("kNN distance ties at the k-th neighbour: the first k by index were kept")
    c = cnt.cpu().numpy().astype(np.int64)
    idx = knn.cpu().numpy()
    take = np.arange(k)[None, :] < c[:, None]
    rows = np.repeat(np.arange(n), c)
    g = scipy.sparse.csr_matrix((np.ones(int(c.sum())), (rows, idx[take])), shape=(n, n))
    g.sort_indices()
    return g


def _radius_pattern(xyz_d, n, radius):
    lib = _lib.lib()
    counts = torch.zeros(n, dtype=torch.int64, device="cuda")
    _lib.check(lib.modest_radius_graph(_lib.ptr(xyz_d), int(xyz_d.shape[1]), n, float(radius), _lib.ptr(counts), None, None,
                                       _lib.stream_ptr()), "modest_radius_graph")
    indptr = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    indptr[1:] = torch.cumsum(counts, 0)
    nnz = int(indptr[-1].item())
    indices = torch.empty(max(nnz, 1), dtype=torch.int32, device="cuda")
    _lib.check(lib.modest_radius_graph(_lib.ptr(xyz_d), int(xyz_d.shape[1]), n, float(radius), None, _lib.ptr(indptr),
                                       _lib.ptr(indices), _lib.stream_ptr()), "modest_radius_graph")
    return scipy.sparse.csr_matrix((np.ones(nnz), indices.cpu().numpy()[:nnz].astype(np.int64), indptr.cpu().numpy()),
                                   shape=(n, n))


def precompute_affinity_matrix(ptc, pp_score, neighbor_type='mutual_knn', affinity_type='l1', n_neighbors=50,
                               radius=1.):
    """clustering_utils.py:7-60 -- CSR (N,N) f64: the edges of the chosen graph type with the chosen
    affinity as data (evaluated in float32 like numpy does, explicit zeros kept).  Column indices
    are sorted within a row (the reference returns sklearn's traversal order for 'knn')."""
    assert ptc.shape[0] == pp_score.shape[0]
    if neighbor_type not in _NEIGHBOR_TYPES:
        raise NotImplementedError(neighbor_type)
    if affinity_type not in _AFFINITY_KINDS:
        raise NotImplementedError(affinity_type)
    n = int(ptc.shape[0])
    ptc = np.ascontiguousarray(ptc, dtype=np.float32)
    pp32 = np.ascontiguousarray(pp_score, dtype=np.float32)
    if neighbor_type == 'radius_mutual_knn':
        indptr, indices, w = _grid_mutual_graph(ptc, pp32, n_neighbors, radius)
        if affinity_type == 'l1':
            return scipy.sparse.csr_matrix((w, indices, indptr), shape=(n, n))
        pattern = scipy.sparse.csr_matrix((np.ones(len(indices)), indices, indptr), shape=(n, n))
    else:
        ptc_d = torch.from_numpy(ptc).cuda()
        if neighbor_type == 'radius':
            pattern = _radius_pattern(ptc_d, n, radius)
        else:
            g = _knn_pattern(ptc_d, n, int(n_neighbors))
            if neighbor_type == 'sym_knn':
                pattern = (g + g.T).tocsr()              # :19-22, the union of the two index sets
            elif neighbor_type == 'mutual_knn':
                pattern = g.multiply(g.T).tocsr()        # :23-27, their intersection
                pattern.eliminate_zeros()
            else:
                pattern = g
    pattern.sort_indices()
    indptr = np.ascontiguousarray(pattern.indptr, dtype=np.int64)
    indices = np.ascontiguousarray(pattern.indices, dtype=np.int32)
    # edge weights on the GPU (:42-56)
    ptc_d = torch.from_numpy(ptc).cuda()
    pp_d = torch.from_numpy(pp32).cuda()
    ip_d, ix_d = torch.from_numpy(indptr).cuda(), torch.from_numpy(indices).cuda()
    out = torch.empty(max(len(indices), 1), dtype=torch.float64, device="cuda")
    _lib.check(_lib.lib().modest_edge_affinity(_lib.ptr(ptc_d), int(ptc.shape[1]), int(ptc.shape[1]), _lib.ptr(pp_d), _lib.ptr(ip_d),
                                               _lib.ptr(ix_d), n, _AFFINITY_KINDS[affinity_type], _lib.ptr(out),
                                               _lib.stream_ptr()), "modest_edge_affinity")
    return scipy.sparse.csr_matrix((out.cpu().numpy()[:len(indices)], indices.astype(np.int64), indptr), shape=(n, n))


def smoothing(*args, **kwargs):
    raise NotImplementedError("smoothing() is dead code in the reference (clustering_utils.py:63-92)")


def _filter(ptc, pp_score, labels, plane, **gates):
    pipe = _pl.SeedLabelPipeline(dict(filtering=gates)) if gates else _pipe()
    b = _as_batch(ptc, pp_score)
    lab = torch.from_numpy(np.ascontiguousarray(labels, dtype=np.int32)).cuda()
    ncl = torch.tensor([int(labels.max()) + 1 if labels.size else 0], dtype=torch.int32, device="cuda")
    pl = torch.from_numpy(np.asarray(plane, dtype=np.float64).reshape(1, 4).copy()).cuda()
    return pipe.filter_and_fit(b, lab, ncl, pl)


def is_valid_cluster(ptc, pp_score, plane, min_points=10, max_volume=40, min_volume=0.5, max_min_height=4,
                     min_max_height=0, percentile=10, min_percentile_pp_score=0.7):
    """clustering_utils.py:94-117"""
    gates = dict(min_points=min_points, max_volume=1e300, min_volume=-1e300, max_min_height=max_min_height,
                 min_max_height=min_max_height, percentile=percentile,
                 min_percentile_pp_score=min_percentile_pp_score)
    labels = np.zeros(ptc.shape[0], dtype=np.int32)
    *_, n_valid, _ = _filter(ptc, pp_score, labels, plane, **gates)
    return bool(int(n_valid.cpu()[0]) == 1)


def filter_labels(ptc, pp_score, labels, **kwargs):
    """clustering_utils.py:119-135 -- second RANSAC plane (global numpy RNG), per-cluster
    gates, ids re-numbered by sorted(set(labels))."""
    plane = estimate_plane(ptc, max_hs=-1.5, ptc_range=((-70, 70), (-50, 50)))
    gates = dict(_pl.DEFAULT_CFG["filtering"])
    gates.update(kwargs)
    lf, *_ = _filter(ptc, pp_score, labels, plane, **gates)
    return lf.cpu().numpy().astype(labels.dtype)
