"""Batched seed-label pipeline: host orchestration of the CUDA stages behind the C ABI.

One `ScanBatch` holds S scans resident on a GPU; `SeedLabelPipeline.run` takes it through the
body of the reference's generate_mask.py:52-103 and gen_label_files.py:41-52 loops:

    RANSAC plane -> ground/range masks -> mutual-kNN graph -> DBSCAN -> second plane ->
    cluster gates -> closeness-to-edge boxes -> volume gate -> BEV NMS -> FOV gate -> label text

Stage methods are public so the drop-in operator modules
(modest_b200/generate_cluster_mask/utils/*.py) can call them one at a time with S = 1.
Nothing here computes on the CPU except RNG bookkeeping (parity mode) and text assembly.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib, ransac_host

DEFAULT_CFG = dict(
    plane_estimate=dict(range=[[-70, 70], [-20, 20]], max_hs=-1.5, offset=0.05),
    limit_range=[[-70, 70], [-40, 40]],
    graph=dict(neighbor_type="radius_mutual_knn", affinity_type="l1", n_neighbors=70, radius=2.0),
    clustering=dict(method="DBSCAN", DBSCAN=dict(eps=0.1, min_samples=10)),
    filtering=dict(min_points=10, max_volume=120, min_volume=0.5, min_max_height=0.5,
                   max_min_height=1.0, percentile=20, min_percentile_pp_score=0.7),
    bbox_gen=dict(fit_method="closeness_to_edge"),
    # gen_label_files.py side
    image_shape=[1024, 1224], fov_only=True, nms=dict(enable=True, threshold=0.1),
)

# filter_labels() hard-codes its own plane arguments (utils/clustering_utils.py:126)
FILTER_PLANE = dict(max_hs=-1.5, range=((-70, 70), (-50, 50)))
CLOSENESS_D0 = 1e-2          # utils/pointcloud_utils.py:167
MAX_TRIALS = 100             # sklearn RANSACRegressor default


def _plain(cfg):
    """OmegaConf-like / AttrDict -> plain nested dict."""
    if hasattr(cfg, "items"):
        return {k: _plain(v) for k, v in cfg.items()}
    if isinstance(cfg, (list, tuple)):
        return [_plain(v) for v in cfg]
    return cfg


def search_angle_tables():
    """cos/sin of the 901 search headings and of heading + pi/2, evaluated with numpy exactly
    the way closeness_rectangle does (utils/pointcloud_utils.py:170-176,196-201):
    angle = np.arange(0, 90.1, 0.1)[k] / 180. * np.pi.  A constant table, built once."""
    deg = np.arange(0, 90 + 0.1, 0.1)
    ang = np.array([d / 180. * np.pi for d in deg])
    ang2 = np.array([a + np.pi / 2 for a in ang])
    trig = np.stack([np.array([np.cos(a) for a in ang]), np.array([np.sin(a) for a in ang]),
                     np.array([np.cos(a) for a in ang2]), np.array([np.sin(a) for a in ang2])])
    return np.ascontiguousarray(trig, dtype=np.float64), np.ascontiguousarray(np.stack([ang, ang2]))


@dataclass
class ScanBatch:
    ptc: torch.Tensor            # (NP,4) f32 cuda  [x,y,z,intensity]
    off: torch.Tensor            # (S+1) i64 cuda
    pp: torch.Tensor             # (NP) f32 cuda
    calib: torch.Tensor          # (S,21) f64 cuda: Tr_velo_to_cam (12) + R0_rect (9)
    P2: np.ndarray               # (S,3,4) f64 host
    h_off: np.ndarray            # host copy of off
    scan_ids: list = field(default_factory=list)
    scan_keys: torch.Tensor = None   # (S) i64 cuda: per-scan key of the device RANSAC draws (the scan id)

    @property
    def n_scans(self):
        return len(self.h_off) - 1

    @property
    def n_points(self):
        return int(self.h_off[-1])

    @property
    def max_points(self):
        return int(np.diff(self.h_off).max()) if self.n_scans else 0


def calib_row(calib) -> np.ndarray:
    """(21,) f64 from an object with V2C (3,4) and R0 (3,3) or a dict of KITTI matrices."""
    if isinstance(calib, dict):
        v2c, r0 = np.asarray(calib["Tr_velo_to_cam"], np.float64), np.asarray(calib["R0_rect"], np.float64)
    else:
        v2c, r0 = np.asarray(calib.V2C, np.float64), np.asarray(calib.R0, np.float64)
    return np.concatenate([v2c.reshape(12), r0.reshape(9)])


def calib_P2(calib) -> np.ndarray:
    P = calib["P2"] if isinstance(calib, dict) else calib.P
    return np.asarray(P, np.float64).reshape(3, 4)


def scan_key(scan_id) -> int:
    """i64 key of a scan for the device RANSAC draws: the integer scan id ("000123" -> 123);
    any other identifier is hashed (stable across runs and ranks)."""
    try:
        return int(scan_id) & (2**63 - 1)
    except (TypeError, ValueError):
        import zlib
        return zlib.crc32(str(scan_id).encode())


def make_batch(ptcs, pps, calibs, scan_ids=None, device="cuda") -> ScanBatch:
    sizes = [int(p.shape[0]) for p in ptcs]
    h_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)

    def dev(a, dt):
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
        return t.to(device=device, dtype=dt, non_blocking=True)
    ptc = torch.cat([dev(p, torch.float32).reshape(-1, 4) for p in ptcs]).contiguous() if ptcs else \
        torch.zeros((0, 4), device=device)
    pp = torch.cat([dev(p, torch.float32).reshape(-1) for p in pps]).contiguous() if pps else \
        torch.zeros((0,), device=device)
    crow = np.stack([calib_row(c) for c in calibs]) if calibs else np.zeros((0, 21))
    P2 = np.stack([calib_P2(c) for c in calibs]) if calibs else np.zeros((0, 3, 4))
    ids = list(scan_ids) if scan_ids is not None else list(range(len(ptcs)))
    return ScanBatch(ptc=ptc, off=torch.from_numpy(h_off).to(device), pp=pp,
                     calib=torch.from_numpy(crow).to(device), P2=P2, h_off=h_off, scan_ids=ids,
                     scan_keys=torch.tensor([scan_key(i) for i in ids], dtype=torch.int64).to(device))


@dataclass
class BatchResult:
    plane: torch.Tensor = None           # (S,4) f64
    plane2: torch.Tensor = None          # (S,4) f64
    ransac_info: torch.Tensor = None     # (S,4) i32 (first fit)
    ransac_info2: torch.Tensor = None
    triples: torch.Tensor = None         # (S,100,3) i32 minimal sets of the first / second fit (want_debug)
    triples2: torch.Tensor = None
    n_kept: torch.Tensor = None          # (S) i32
    kept_idx: torch.Tensor = None        # (NP) i32
    mask: torch.Tensor = None            # (NP) u8 final_mask
    labels_raw: torch.Tensor = None      # (NP) i32 DBSCAN labels on the full scan
    n_clusters: torch.Tensor = None
    labels_filtered: torch.Tensor = None
    labels: torch.Tensor = None          # (NP) i32 final seg labels
    boxes: torch.Tensor = None           # (S,max_boxes,8) f64
    n_boxes: torch.Tensor = None
    keep: torch.Tensor = None            # (S,max_boxes) u8 after NMS
    iou: torch.Tensor = None
    flags: dict = field(default_factory=dict)


class _Scratch:
    """Grow-only device buffers keyed by name (the library itself never allocates)."""

    def __init__(self):
        self._bufs = {}
        self.generation = 0       # (re)allocations so far: a captured CUDA graph is stale when this moves

    def get(self, name, numel, dtype, device):
        t = self._bufs.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype or t.device != torch.device(device):
            t = torch.empty(max(int(numel), 1), dtype=dtype, device=device)
            self._bufs[name] = t
            self.generation += 1
        return t


class SeedLabelPipeline:
    def __init__(self, cfg=None, max_clusters=2048, max_boxes=128, graph_grid=288):
        base = _plain(DEFAULT_CFG)
        if cfg is not None:
            for k, v in _plain(cfg).items():
                if isinstance(v, dict) and isinstance(base.get(k), dict):
                    base[k] = {**base[k], **v}
                else:
                    base[k] = v
        self.cfg = base
        g = self.cfg["graph"]
        if g["neighbor_type"] != "radius_mutual_knn" or g["affinity_type"] != "l1":
            # utils/clustering_utils.py:16-31,49-56 has other branches: the operator module
            # (generate_cluster_mask/utils/clustering_utils.py) implements them, the fused batched
            # pipeline only the configured default
            raise NotImplementedError(f"{g['neighbor_type']}/{g['affinity_type']} in the fused pipeline")
        if self.cfg["clustering"]["method"] != "DBSCAN":
            raise NotImplementedError(self.cfg["clustering"]["method"])    # generate_mask.py:82-83
        if self.cfg["bbox_gen"]["fit_method"] != "closeness_to_edge":
            raise NotImplementedError(self.cfg["bbox_gen"]["fit_method"])   # pointcloud_utils.py:301-302
        self.max_clusters, self.max_boxes, self.graph_grid = int(max_clusters), int(max_boxes), int(graph_grid)
        self._scr = _Scratch()
        self._tables = None
        self.lib = _lib.lib()

    # ------------------------------------------------------------------ helpers
    def _angle_tables(self, device):
        if self._tables is None or self._tables[0].device != torch.device(device):
            trig, ang = search_angle_tables()
            self._tables = (torch.from_numpy(trig).to(device), torch.from_numpy(ang).to(device), trig.shape[1])
        return self._tables

    @staticmethod
    def _range4(r):
        return np.array([r[0][0], r[0][1], r[1][0], r[1][1]], dtype=np.float32)

    # ------------------------------------------------------------------ stage E
    def fit_planes(self, b: ScanBatch, max_hs, ptc_range, rng="device", seed=0, stream=None,
                   return_debug=False):
        """estimate_plane() for every scan of the batch.

        rng="device": minimal sets drawn on the GPU from (`seed`, b.scan_keys[s], trial) -- no host
                      sync, and a scan's draws do not depend on its slot in the batch.
        rng="numpy":  minimal sets taken from numpy's global RandomState exactly as sklearn
                      would; scans are processed in order, the stream advances by n_trials_ draws
                      per scan (two host syncs for the whole batch when `seed` is a list of
                      per-scan seeds handled by the caller).
        Returns (plane (S,4) f64 cuda, info (S,4) i32 cuda[, debug dict])."""
        S, dev = b.n_scans, b.ptc.device
        sp = _lib.stream_ptr(stream)
        cand = self._scr.get("cand", 3 * b.n_points, torch.float32, dev)
        n_cand = torch.empty(S, dtype=torch.int32, device=dev)
        thr = torch.empty(S, dtype=torch.float32, device=dev)
        r4 = self._range4(ptc_range)
        _lib.check(self.lib.modest_plane_candidates_batch(
            _lib.ptr(b.ptc), 4, _lib.ptr(b.off), S, float(max_hs), float(r4[0]), float(r4[1]), float(r4[2]),
            float(r4[3]), _lib.ptr(cand), _lib.ptr(n_cand), _lib.ptr(thr), sp), "modest_plane_candidates_batch")
        triples = None
        h_ncand = None
        if rng == "numpy":
            h_ncand = n_cand.cpu().numpy()
            if S != 1:
                raise ValueError("rng='numpy' consumes one global stream: call with one scan at a time")
            triples = torch.from_numpy(ransac_host.peek_triples(int(h_ncand[0]), MAX_TRIALS)[None]).to(dev)
        elif isinstance(rng, torch.Tensor):
            triples = rng
        plane = torch.empty((S, 4), dtype=torch.float64, device=dev)
        model = torch.empty((S, 3), dtype=torch.float64, device=dev)
        info = torch.empty((S, 4), dtype=torch.int32, device=dev)
        need = self.lib.modest_ransac_workspace_bytes(S, MAX_TRIALS)
        ws = self._scr.get("ransac_ws", need, torch.uint8, dev)
        inl = self._scr.get("inlier", b.n_points, torch.uint8, dev) if return_debug else None
        tri_out = torch.empty((S, MAX_TRIALS, 3), dtype=torch.int32, device=dev) if return_debug else None
        _lib.check(self.lib.modest_ransac_fit_batch(
            _lib.ptr(cand), _lib.ptr(b.off), _lib.ptr(n_cand), _lib.ptr(thr), S, b.max_points,
            _lib.ptr(triples), C.c_uint64(int(seed) & (2**64 - 1)), _lib.ptr(b.scan_keys), MAX_TRIALS, _lib.ptr(plane),
            _lib.ptr(model),
            _lib.ptr(info), _lib.ptr(tri_out), _lib.ptr(inl), _lib.ptr(ws), ws.numel(), sp), "modest_ransac_fit_batch")
        if rng == "numpy":
            h_info = info.cpu().numpy()
            ransac_host.consume_trials(int(h_ncand[0]), int(h_info[0, 1]))
        if return_debug:
            return plane, info, dict(n_cand=n_cand, thr=thr, model=model, inlier=inl, cand=cand, triples=tri_out)
        return plane, info

    # ------------------------------------------------------------------ stages F+G
    def ground_masks(self, b: ScanBatch, plane, offset, only_range, limit_range, want_mask=True, stream=None):
        S, dev = b.n_scans, b.ptc.device
        kept = self._scr.get("kept", 4 * b.n_points, torch.float32, dev)
        kept_idx = torch.empty(b.n_points, dtype=torch.int32, device=dev)
        n_kept = torch.empty(S, dtype=torch.int32, device=dev)
        mask = torch.empty(b.n_points, dtype=torch.uint8, device=dev) if want_mask else None
        only = self._range4(only_range) if only_range is not None else None
        lim = self._range4(limit_range)
        _lib.check(self.lib.modest_ground_mask_batch(
            _lib.ptr(b.ptc), 4, _lib.ptr(b.off), _lib.ptr(b.pp), _lib.ptr(plane), S, float(offset),
            None if only is None else only.ctypes.data_as(C.c_void_p), lim.ctypes.data_as(C.c_void_p),
            _lib.ptr(kept), _lib.ptr(kept_idx), _lib.ptr(n_kept), _lib.ptr(mask), _lib.stream_ptr(stream)),
            "modest_ground_mask_batch")
        return kept, kept_idx, n_kept, mask

    # ------------------------------------------------------------------ stage H
    def affinity_graph(self, kept, off, n_kept, n_scans, n_points, max_points, stream=None, partition_eps=None,
                       eps_edges_only=False):
        """Returns nbr, nbr_w, nbr_cnt, flags.  With `partition_eps` (the DBSCAN radius the caller
        will use) every row lists its eps-edges first and `self.nbr_eps_cnt` holds their number;
        with `eps_edges_only` nothing else is written (nbr_w is None): all DBSCAN needs."""
        dev = kept.device
        g = self.cfg["graph"]
        k = int(g["n_neighbors"])
        eps_edges_only = bool(eps_edges_only and partition_eps is not None and k <= 96)
        nbr = self._scr.get("nbr", n_points * k, torch.int32, dev)
        nbr_w = None if eps_edges_only else self._scr.get("nbr_w", n_points * k, torch.float32, dev)
        nbr_cnt = self._scr.get("nbr_cnt", n_points, torch.int32, dev)
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
        self.nbr_eps_cnt = self._scr.get("nbr_eps_cnt", n_points, torch.int32, dev) if partition_eps is not None else None
        need = self.lib.modest_graph_workspace_bytes(n_scans, n_points, k, self.graph_grid)
        ws = self._scr.get("graph_ws", need, torch.uint8, dev)
        _lib.check(self.lib.modest_affinity_graph_batch(
            _lib.ptr(kept), _lib.ptr(off), _lib.ptr(n_kept), n_scans, n_points, max_points, k, float(g["radius"]),
            self.graph_grid, _lib.ptr(nbr), _lib.ptr(nbr_w), _lib.ptr(nbr_cnt),
            -1.0 if partition_eps is None else float(partition_eps), _lib.ptr(self.nbr_eps_cnt), _lib.ptr(flags), _lib.ptr(ws),
            ws.numel(), _lib.stream_ptr(stream)), "modest_affinity_graph_batch")
        return nbr, nbr_w, nbr_cnt, flags

    # ------------------------------------------------------------------ stage I
    def dbscan(self, off, n_kept, kept_idx, n_scans, n_points, max_points, nbr, nbr_w, nbr_cnt, stream=None,
               nbr_eps_cnt=None):
        dev = off.device
        d = self.cfg["clustering"]["DBSCAN"]
        k = int(self.cfg["graph"]["n_neighbors"])
        labels_kept = torch.empty(n_points, dtype=torch.int32, device=dev)
        labels_full = torch.empty(n_points, dtype=torch.int32, device=dev)
        n_clusters = torch.empty(n_scans, dtype=torch.int32, device=dev)
        need = self.lib.modest_dbscan_workspace_bytes(n_points)
        ws = self._scr.get("dbscan_ws", need, torch.uint8, dev)
        _lib.check(self.lib.modest_dbscan_batch(
            _lib.ptr(off), _lib.ptr(n_kept), _lib.ptr(kept_idx), n_scans, n_points, max_points, k, _lib.ptr(nbr),
            _lib.ptr(nbr_w), _lib.ptr(nbr_cnt), _lib.ptr(nbr_eps_cnt), float(d["eps"]), int(d["min_samples"]), _lib.ptr(labels_kept),
            _lib.ptr(labels_full), _lib.ptr(n_clusters), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(stream)),
            "modest_dbscan_batch")
        return labels_kept, labels_full, n_clusters

    # ------------------------------------------------------------------ stages J-M
    def gates_array(self):
        f = self.cfg["filtering"]
        q32 = np.float32(f["percentile"]) / np.float32(100)     # numpy: q / arr.dtype.type(100)
        return np.array([f["min_points"], f["max_min_height"], f["min_max_height"], float(q32),
                         f["min_percentile_pp_score"], f["min_volume"], f["max_volume"], CLOSENESS_D0],
                        dtype=np.float64)

    def filter_and_fit(self, b: ScanBatch, labels_full, n_clusters, plane2, stream=None, rect=None, gates=None):
        S, dev = b.n_scans, b.ptc.device
        trig, ang, n_ang = self._angle_tables(dev)
        gates = self.gates_array() if gates is None else np.ascontiguousarray(gates, dtype=np.float64)
        labels_filtered = torch.empty(b.n_points, dtype=torch.int32, device=dev)
        labels_final = torch.empty(b.n_points, dtype=torch.int32, device=dev)
        boxes = torch.zeros((S, self.max_boxes, 8), dtype=torch.float64, device=dev)
        n_boxes = torch.empty(S, dtype=torch.int32, device=dev)
        n_valid = torch.empty(S, dtype=torch.int32, device=dev)
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
        need = self.lib.modest_filter_workspace_bytes(S, b.n_points, self.max_clusters)
        ws = self._scr.get("filter_ws", need, torch.uint8, dev)
        _lib.check(self.lib.modest_filter_and_fit_batch(
            _lib.ptr(b.ptc), 4, _lib.ptr(b.off), _lib.ptr(b.pp), _lib.ptr(labels_full), _lib.ptr(n_clusters),
            _lib.ptr(plane2), _lib.ptr(b.calib), _lib.ptr(rect), S, b.n_points, b.max_points, self.max_clusters, self.max_boxes,
            gates.ctypes.data_as(C.c_void_p), _lib.ptr(trig), _lib.ptr(ang), n_ang, _lib.ptr(labels_filtered),
            _lib.ptr(labels_final), _lib.ptr(boxes), _lib.ptr(n_boxes), _lib.ptr(n_valid), _lib.ptr(flags),
            _lib.ptr(ws), ws.numel(), _lib.stream_ptr(stream)), "modest_filter_and_fit_batch")
        return labels_filtered, labels_final, boxes, n_boxes, n_valid, flags

    # ------------------------------------------------------------------ stage N
    def seed_nms(self, boxes, n_boxes, want_iou=False, stream=None):
        S, dev = boxes.shape[0], boxes.device
        mb = self.max_boxes
        iou_ws = self._scr.get("iou_ws", S * mb * mb, torch.float32, dev)
        iou = torch.zeros((S, mb, mb), dtype=torch.float32, device=dev) if want_iou else None
        keep = torch.empty((S, mb), dtype=torch.uint8, device=dev)
        _lib.check(self.lib.modest_seed_nms_batch(
            _lib.ptr(boxes), _lib.ptr(n_boxes), S, mb, float(self.cfg["nms"]["threshold"]), _lib.ptr(iou),
            _lib.ptr(iou_ws), _lib.ptr(keep), _lib.stream_ptr(stream)), "modest_seed_nms_batch")
        return keep, iou

    # ------------------------------------------------------------------ whole pipeline
    def run(self, b: ScanBatch, rng="device", seed=0, want_debug=False, stream=None) -> BatchResult:
        """generate_mask.py:52-103 + the NMS of gen_label_files.py:44-45 for every scan.

        rng="device": throughput mode, no host synchronisation anywhere.
        rng="numpy":  parity mode (one scan per batch): both RANSAC fits draw from numpy's
                      global RandomState in the reference's order (first estimate_plane, then
                      filter_labels' own)."""
        if stream is not None and stream != torch.cuda.current_stream():
            # temporaries are allocated (and later freed) by torch's caching allocator on the CURRENT
            # stream: make `stream` current for the whole call so kernels and allocations agree
            with torch.cuda.stream(stream):
                return self.run(b, rng=rng, seed=seed, want_debug=want_debug, stream=stream)
        cfg = self.cfg
        pe = cfg["plane_estimate"]
        r = BatchResult()
        if want_debug:     # also hand back the minimal sets that were used (tests replay them in sklearn)
            r.plane, r.ransac_info, dbg = self.fit_planes(b, pe["max_hs"], pe["range"], rng=rng, seed=seed,
                                                          stream=stream, return_debug=True)
            r.triples = dbg["triples"].clone()
        else:
            r.plane, r.ransac_info = self.fit_planes(b, pe["max_hs"], pe["range"], rng=rng, seed=seed, stream=stream)
        kept, r.kept_idx, r.n_kept, r.mask = self.ground_masks(
            b, r.plane, pe["offset"], pe["range"], cfg["limit_range"], want_mask=want_debug, stream=stream)
        eps = float(cfg["clustering"]["DBSCAN"]["eps"])
        nbr, nbr_w, nbr_cnt, gflags = self.affinity_graph(kept, b.off, r.n_kept, b.n_scans, b.n_points,
                                                          b.max_points, stream=stream, partition_eps=eps,
                                                          eps_edges_only=not want_debug)
        _, r.labels_raw, r.n_clusters = self.dbscan(b.off, r.n_kept, r.kept_idx, b.n_scans, b.n_points,
                                                    b.max_points, nbr, nbr_w, nbr_cnt, stream=stream,
                                                    nbr_eps_cnt=self.nbr_eps_cnt)
        if want_debug:
            r.plane2, r.ransac_info2, dbg = self.fit_planes(b, FILTER_PLANE["max_hs"], FILTER_PLANE["range"], rng=rng,
                                                            seed=seed + 0x9E3779B9, stream=stream, return_debug=True)
            r.triples2 = dbg["triples"].clone()
        else:
            r.plane2, r.ransac_info2 = self.fit_planes(b, FILTER_PLANE["max_hs"], FILTER_PLANE["range"], rng=rng,
                                                       seed=seed + 0x9E3779B9, stream=stream)
        r.labels_filtered, r.labels, r.boxes, r.n_boxes, _, fflags = self.filter_and_fit(
            b, r.labels_raw, r.n_clusters, r.plane2, stream=stream)
        if cfg["nms"]["enable"]:
            r.keep, r.iou = self.seed_nms(r.boxes, r.n_boxes, want_iou=want_debug, stream=stream)
        else:
            r.keep = (torch.arange(self.max_boxes, device=b.ptc.device)[None, :] < r.n_boxes[:, None]).to(torch.uint8)
        r.flags = dict(graph=gflags, fit=fflags)
        return r

    @staticmethod
    def check_flags(result: "BatchResult"):
        """Raise when a fixed-capacity device buffer overflowed (results would be silently
        truncated); warn about distance ties the reference resolves by traversal order."""
        SeedLabelPipeline.check_flag_values(int(result.flags["graph"].item()), int(result.flags["fit"].item()))

    @staticmethod
    def check_flag_values(g: int, f: int):
        if f & 4:
            raise RuntimeError("more DBSCAN clusters than max_clusters: rebuild the pipeline with a larger capacity")
        if f & 16:
            raise RuntimeError("more boxes than max_boxes: rebuild the pipeline with a larger capacity")
        if f & 8:
            raise ValueError("a fitted box contains no scan point (the reference raises on ys.max() of an empty array)")
        if g & 3:
            import warnings
            warnings.warn("kNN distance ties" + (" beyond n_neighbors (duplicate points?)" if g & 2 else "") +
                          ": neighbour sets may differ from scikit-learn's traversal order")

    # ------------------------------------------------------------------ stage O (host)
    def label_texts(self, b: ScanBatch, boxes, n_boxes, keep):
        """KITTI label text per scan (gen_label_files.py:46-52); one D2H copy of the small box
        tables, formatting in the library's host routine."""
        h_boxes = boxes.cpu().numpy()
        h_n = n_boxes.cpu().numpy()
        h_keep = keep.cpu().numpy()
        return self.format_labels_batch(h_boxes, h_n, h_keep, b.P2)

    def format_labels_batch(self, h_boxes, h_n, h_keep, P2, obj_type="Dynamic"):
        """All scans of a batch in one library call: h_boxes (S,max_boxes,8) f64, h_n (S) i32,
        h_keep (S,max_boxes) u8, P2 (S,3,4) f64 -- host arrays -> list of S label texts."""
        S, mb = int(h_boxes.shape[0]), int(h_boxes.shape[1])
        h_boxes = np.ascontiguousarray(h_boxes, dtype=np.float64)
        h_n = np.ascontiguousarray(h_n, dtype=np.int32)
        h_keep = np.ascontiguousarray(h_keep, dtype=np.uint8)
        P2 = np.ascontiguousarray(P2, dtype=np.float64)
        cap = 256 * max(int(h_n.sum()), 1) + 16
        buf = C.create_string_buffer(cap)
        offs = np.zeros(S + 1, dtype=np.int64)
        ish = self.cfg["image_shape"]
        _lib.check(self.lib.modest_kitti_labels_batch_host(
            h_boxes.ctypes.data_as(C.c_void_p), h_n.ctypes.data_as(C.c_void_p), h_keep.ctypes.data_as(C.c_void_p), S, mb,
            P2.ctypes.data_as(C.c_void_p), 1 if self.cfg["fov_only"] else 0, int(ish[0]), int(ish[1]), obj_type.encode(),
            buf, cap, offs.ctypes.data_as(C.c_void_p)), "modest_kitti_labels_batch_host")
        raw = buf.raw
        return [raw[offs[s]:offs[s + 1]].decode() for s in range(S)]

    def format_labels(self, boxes8, keep, P2, scores=None, obj_type="Dynamic"):
        n = int(boxes8.shape[0])
        boxes8 = np.ascontiguousarray(boxes8, dtype=np.float64)
        keep = None if keep is None else np.ascontiguousarray(keep, dtype=np.uint8)
        P2 = np.ascontiguousarray(P2, dtype=np.float64)
        cap = 256 * max(n, 1) + 16
        buf = C.create_string_buffer(cap)
        ln, cnt = C.c_size_t(0), C.c_int(0)
        kept = np.zeros(max(n, 1), dtype=np.uint8)
        ish = self.cfg["image_shape"]
        sc = None if scores is None else np.ascontiguousarray(scores, dtype=np.float64)
        _lib.check(self.lib.modest_kitti_labels_host(
            boxes8.ctypes.data_as(C.c_void_p), n, None if keep is None else keep.ctypes.data_as(C.c_void_p),
            P2.ctypes.data_as(C.c_void_p), 1 if self.cfg["fov_only"] else 0, int(ish[0]), int(ish[1]),
            obj_type.encode(), None if sc is None else sc.ctypes.data_as(C.c_void_p), buf, cap,
            C.byref(ln), C.byref(cnt), kept.ctypes.data_as(C.c_void_p)), "modest_kitti_labels_host")
        return buf.raw[:ln.value].decode(), kept[:n].astype(bool)
