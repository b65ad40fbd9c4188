"""Host side of the PP-score stage: packs ragged scans and calls modest_pp_score_batch.

Mirrors count_neighbors()/compute_ephe_score() of the reference
(generate_cluster_mask/pre_compute_pp_score.py:54-75): the inputs are the query scan and the
per-traversal accumulated history clouds, all already expressed in the fixed frame.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _lib


@dataclass
class PPBatch:
    """A ragged batch of scans resident on one GPU."""
    query_xyz: torch.Tensor     # (NQ,3) f32
    q_off: torch.Tensor         # (S+1) i64
    hist_xyz: torch.Tensor      # (NH,3) f32
    h_off: torch.Tensor         # (G+1) i64, G = total traversals
    trav_off: torch.Tensor      # (S+1) i32
    count_off: torch.Tensor     # (S+1) i64
    n_scans: int
    n_trav_total: int
    n_query_total: int
    n_count_total: int
    max_query_points: int
    max_trav_points: int
    h_q_off: np.ndarray         # host copies (the tiled history pass cuts the batch into groups with them)
    h_trav_off: np.ndarray
    h_count_off: np.ndarray
    h_h_off: np.ndarray = None

    @property
    def algorithmic_bytes(self) -> int:
        """12 B per query point + 12 B per history point + 4 B per score."""
        return 12 * self.n_query_total + 12 * int(self.hist_xyz.shape[0]) + 4 * self.n_query_total


def pack_batch(queries, histories, device="cuda") -> PPBatch:
    """queries: list of (N_s,>=3) arrays/tensors; histories: list (per scan) of lists (per
    traversal) of (M,3) arrays/tensors."""
    def to_dev(a):
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
        t = t[:, :3].to(torch.float32)
        return t.to(device, non_blocking=True).contiguous()

    q_sizes = [int(q.shape[0]) for q in queries]
    trav_counts = [len(h) for h in histories]
    h_sizes = [int(t.shape[0]) for h in histories for t in h]
    q_off = np.concatenate([[0], np.cumsum(q_sizes)]).astype(np.int64)
    h_off = np.concatenate([[0], np.cumsum(h_sizes)]).astype(np.int64)
    trav_off = np.concatenate([[0], np.cumsum(trav_counts)]).astype(np.int32)
    count_off = np.concatenate([[0], np.cumsum(np.array(q_sizes, np.int64) * np.array(trav_counts, np.int64))]).astype(np.int64)
    qx = torch.cat([to_dev(q) for q in queries]) if queries else torch.zeros((0, 3), device=device)
    flat = [to_dev(t) for h in histories for t in h]
    hx = torch.cat(flat) if flat else torch.zeros((0, 3), dtype=torch.float32, device=device)
    dev = lambda a: torch.from_numpy(a).to(device)
    return PPBatch(qx, dev(q_off), hx, dev(h_off), dev(trav_off), dev(count_off),
                   n_scans=len(queries), n_trav_total=int(trav_off[-1]), n_query_total=int(q_off[-1]),
                   n_count_total=int(count_off[-1]), max_query_points=max(q_sizes, default=0),
                   max_trav_points=max(h_sizes, default=0), h_q_off=q_off, h_trav_off=trav_off,
                   h_count_off=count_off, h_h_off=h_off)


class PPScorer:
    """Reusable workspace + launch wrapper.

    history_pass="hash" (default): the global-hash pass -- every history point probes the query's
    hash grid; any number of traversals per scan.
    history_pass="tiled": coarse spatial partition + per-tile shared-memory join, the batch cut into
    groups of about `group_points` history points (0 = library default).  Same bits; measured
    slower than the hash pass on the synthetic Lyft shape, kept selectable."""

    def __init__(self, radius: float = 0.3, grid_dim: int = 512, group_points: int = 0, history_pass: str = "hash"):
        self.radius = float(radius)
        self.grid_dim = int(grid_dim)
        self.group_points = int(group_points)
        if history_pass not in ("tiled", "hash"):
            raise ValueError(history_pass)
        self.history_pass = history_pass
        self._ws = None

    def _host_tables(self, b: PPBatch):
        if self.history_pass != "tiled" or b.h_h_off is None:
            return None, None, None, 0
        q = np.ascontiguousarray(b.h_q_off, dtype=np.int64)
        h = np.ascontiguousarray(b.h_h_off, dtype=np.int64)
        t = np.ascontiguousarray(b.h_trav_off, dtype=np.int32)
        rec = _lib.lib().modest_pp_bin_records(h.ctypes.data, t.ctypes.data, b.n_scans, self.group_points)
        return q, h, t, int(rec)

    def _workspace(self, b: PPBatch, bin_records: int):
        need = _lib.lib().modest_pp_workspace_bytes(b.n_scans, b.n_query_total, b.n_count_total, self.grid_dim, bin_records)
        if self._ws is None or self._ws.numel() < need or self._ws.device != b.query_xyz.device:
            self._ws = torch.empty(int(need), dtype=torch.uint8, device=b.query_xyz.device)
        return self._ws

    def __call__(self, b: PPBatch, out: torch.Tensor | None = None, counts: torch.Tensor | None = None,
                 stream=None) -> torch.Tensor:
        if out is None:
            out = torch.empty(b.n_query_total, dtype=torch.float32, device=b.query_xyz.device)
        hq, hh, ht, rec = self._host_tables(b)
        ws = self._workspace(b, rec)
        hp = lambda a: None if a is None else a.ctypes.data
        rc = _lib.lib().modest_pp_score_batch(
            _lib.ptr(b.query_xyz), _lib.ptr(b.q_off), _lib.ptr(b.hist_xyz), _lib.ptr(b.h_off),
            _lib.ptr(b.trav_off), b.n_scans, b.n_trav_total, b.n_query_total, b.n_count_total,
            b.max_query_points, b.max_trav_points, self.radius, self.grid_dim,
            _lib.ptr(counts), _lib.ptr(b.count_off), _lib.ptr(out), hp(hq), hp(hh), hp(ht),
            self.group_points if self.history_pass == "tiled" else -1, rec, _lib.ptr(ws), ws.numel(),
            _lib.stream_ptr(stream))
        _lib.check(rc, "modest_pp_score_batch")
        return out


def count_neighbors_and_score(query_xyz, history, radius=0.3, grid_dim=512, return_counts=False,
                              history_pass="hash", group_points=0):
    """One scan: query (N,>=3), history = list of (M_t,3).  Returns pp (N,) f32 numpy
    [and counts (N,T) int64 like the reference's count_neighbors]."""
    b = pack_batch([query_xyz], [history])
    counts = torch.zeros(b.n_count_total, dtype=torch.int32, device="cuda") if return_counts else None
    pp = PPScorer(radius, grid_dim, group_points=group_points, history_pass=history_pass)(b, counts=counts)
    torch.cuda.synchronize()
    if return_counts:
        return pp.cpu().numpy(), counts.cpu().numpy().reshape(b.n_query_total, len(history)).astype(np.int64)
    return pp.cpu().numpy()
