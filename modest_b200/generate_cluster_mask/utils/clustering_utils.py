"""Drop-in for the reference's utils/clustering_utils.py (same names / arguments / return
types) on top of libmodest_b200.  Only the configured defaults have CUDA implementations:
neighbor_type='radius_mutual_knn' with affinity_type='l1'; the other branches of
clustering_utils.py:16-31,49-56 raise NotImplementedError (there is no CPU fallback)."""
import numpy as np
import scipy.sparse
import torch

from modest_b200 import pipeline as _pl
from .pointcloud_utils import _as_batch, _pipe, distance_to_plane, estimate_plane  # noqa: F401


def precompute_affinity_matrix(ptc, pp_score, neighbor_type='mutual_knn', affinity_type='l1', n_neighbors=50,
                               radius=1.):
    """clustering_utils.py:7-60 -- CSR (N,N) f64 whose stored entries are the mutual-kNN AND
    radius edges with data = |pp_i - pp_j| (evaluated in float32)."""
    assert ptc.shape[0] == pp_score.shape[0]
    if neighbor_type != 'radius_mutual_knn':
        raise NotImplementedError(neighbor_type)
    if affinity_type != 'l1':
        raise NotImplementedError(affinity_type)
    n = ptc.shape[0]
    pipe = _pl.SeedLabelPipeline(dict(graph=dict(neighbor_type=neighbor_type, affinity_type=affinity_type,
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
    return scipy.sparse.csr_matrix((w_s[take_s].astype(np.float64), idx_s[take_s].astype(np.int64), indptr),
                                   shape=(n, n))


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
