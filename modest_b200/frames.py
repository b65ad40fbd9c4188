"""History front end of the streaming engine.

The reference re-reads and re-transforms every history frame for every query scan
(pre_compute_pp_score.py:133-150: F frames x T traversals per scan), although consecutive query
scans share almost all of them.  Here a scan is described by frame ids and 4x4 poses; the raw
velodyne frames cross PCIe once, stay in a device-resident cache, and
`modest_transform_gather_batch` (stage B) brings them into each scan's fixed frame on the GPU.

    JobBatch          what a loader hands to the engine: ids + poses, flat numpy arrays
    DeviceFrameCache  frame id -> (N,4) f32 cuda tensor, least recently used out first
    jobs_from_dataset builds JobBatches for a synth.TrackDataset (tests, bench)
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib

# the 96-byte job record of modest_transform_gather_batch (include/modest_b200.h)
FRAME_JOB = np.dtype([("src", "<u8"), ("dst_row", "<i8"), ("n", "<i4"), ("flags", "<i4"), ("T", "<f4", (16,)),
                      ("pad", "<i8")])
assert FRAME_JOB.itemsize == 96
CENTER_BOX = np.array([-1.15, 1.75, -0.65, 0.65], dtype=np.float32)       # pre_compute_pp_score.py:48


@dataclass
class JobBatch:
    """S query scans.  History frames are listed scan-major, then in the order the reference
    concatenates them (traversal by traversal, pre_compute_pp_score.py:133-150)."""
    scan_ids: list
    query_fid: np.ndarray            # (S,) frame id of every query scan (its raw frame is the scan itself)
    query_T: np.ndarray              # (S,4,4) f32 fixed frame <- query frame
    hist_fid: np.ndarray             # (H,) frame ids
    hist_T: np.ndarray               # (H,4,4) f32 fixed frame <- history frame
    frames_per_trav: list            # per scan: frames in each traversal (sums to that scan's share of H)
    calibs: list
    remove_center: bool = False      # nuScenes: ego returns are dropped from HISTORY frames only (:141-142)

    @property
    def n_scans(self):
        return len(self.scan_ids)


class DeviceFrameCache:
    """Raw frames resident on one GPU.  `source(fid)` must return the frame as a pinned (N,4)
    float32 host tensor (e.g. velodyne/%06d.bin read into pinned memory); uploads are enqueued
    on the stream that is current when `get` is called.

    Frames are packed into slabs of `slab_bytes` (one device allocation per slab, not per frame:
    an allocation per 1 MB frame costs more host time than the whole batch's launches); when
    `device_bytes` is exceeded the oldest slab is dropped with every frame in it."""

    def __init__(self, source, device_bytes=48 << 30, slab_bytes=256 << 20):
        self.source, self.device_bytes, self.slab_bytes = source, int(device_bytes), int(slab_bytes)
        self._d = {}                      # fid -> (N,4) view into a slab
        self._slabs = []                  # [tensor, used floats, [fids]] oldest first
        self.hits = self.misses = 0
        self.h2d_bytes = 0

    def __contains__(self, fid):
        return fid in self._d

    def _room(self, n_floats):
        if self._slabs and self._slabs[-1][1] + n_floats <= self._slabs[-1][0].numel():
            return self._slabs[-1]
        size = max(self.slab_bytes // 4, n_floats)
        while self._slabs and 4 * (sum(sl[0].numel() for sl in self._slabs) + size) > self.device_bytes:
            _, _, fids = self._slabs.pop(0)
            for f in fids:
                self._d.pop(f, None)
        self._slabs.append([torch.empty(size, dtype=torch.float32, device="cuda"), 0, []])
        return self._slabs[-1]

    def get(self, fid):
        t = self._d.get(fid)
        if t is not None:
            self.hits += 1
            return t
        self.misses += 1
        host = self.source(fid)
        n = host.numel()
        slab = self._room(n)
        t = slab[0][slab[1]:slab[1] + n].view(host.shape)
        slab[1] += (n + 3) // 4 * 4                     # frames stay 16-byte aligned
        slab[2].append(fid)
        t.copy_(host, non_blocking=True)
        self.h2d_bytes += n * 4
        self._d[fid] = t
        return t

    def get_many(self, fids, stream=None):
        """`get` for a list of frames with the uploads of all misses enqueued by one library call
        (a batch of a drive can need ~800 frames the cache has not seen: one torch copy per frame
        costs the host 30-80 us each, the library's loop 2-3 us)."""
        out, todo = [], []
        for fid in fids:
            t = self._d.get(fid)
            if t is not None:
                self.hits += 1
            else:
                self.misses += 1
                host = self.source(fid)
                n = host.numel()
                slab = self._room(n)
                t = slab[0][slab[1]:slab[1] + n].view(host.shape)
                slab[1] += (n + 3) // 4 * 4                     # frames stay 16-byte aligned
                slab[2].append(fid)
                self.h2d_bytes += n * 4
                self._d[fid] = t
                todo.append((host, t, 4 * n))
            out.append(t)
        if todo:
            src = np.array([h.data_ptr() for h, _, _ in todo], dtype=np.uint64)
            dst = np.array([t.data_ptr() for _, t, _ in todo], dtype=np.uint64)
            nb = np.array([b for _, _, b in todo], dtype=np.int64)
            _lib.check(_lib.lib().modest_upload_frames(src.ctypes.data, dst.ctypes.data, nb.ctypes.data, len(todo),
                                                       _lib.stream_ptr(stream)), "modest_upload_frames")
            self._keepalive = [h for h, _, _ in todo]           # host buffers stay referenced until the next call
        return out

    def clear(self):
        self._d.clear()
        self._slabs.clear()


def pinned_frame_source(frames: dict):
    """frame id -> pinned tensor, for frames already in host memory (dict of (N,4) arrays)."""
    pinned = {}

    def source(fid):
        t = pinned.get(fid)
        if t is None:
            t = torch.from_numpy(np.ascontiguousarray(frames[fid], dtype=np.float32)).pin_memory()
            pinned[fid] = t
        return t
    return source


def bin_file_source(velodyne_dir: str):
    """frame id -> pinned tensor read from <velodyne_dir>/%06d.bin (utils/pointcloud_utils.py:22-25)."""
    import os

    def source(fid):
        arr = np.fromfile(os.path.join(velodyne_dir, f"{int(fid):06d}.bin"), dtype=np.float32).reshape(-1, 4)
        return torch.from_numpy(arr).pin_memory()
    return source


def jobs_from_dataset(ds, scan_ids, batch_size) -> list:
    """JobBatches over `scan_ids` of a synth.TrackDataset, poses by the reference's pose chain
    (pre_compute_pp_score.py:27-28, three stacked solves per batch: TrackDataset.relative_poses_batch)."""
    out = []
    for i in range(0, len(scan_ids), batch_size):
        ids = list(scan_ids[i:i + batch_size])
        fpt, flat_scan, flat_fid = [], [], []
        for sid in ids:
            groups = ds.history_frames(sid)
            fpt.append([len(g) for g in groups])
            flat_scan.append(sid)                                   # the query itself first, then its history
            flat_fid.append(sid)
            for g in groups:
                flat_scan.extend([sid] * len(g))
                flat_fid.extend(g)
        poses = ds.relative_poses_batch(flat_scan, flat_fid)
        is_q = np.zeros(len(flat_fid), dtype=bool)
        is_q[np.cumsum([0] + [1 + sum(f) for f in fpt[:-1]])] = True
        out.append(JobBatch(scan_ids=ids, query_fid=np.array(ids, dtype=np.int64), query_T=poses[is_q],
                            hist_fid=np.array(flat_fid, dtype=np.int64)[~is_q], hist_T=poses[~is_q],
                            frames_per_trav=fpt, calibs=[ds.calib] * len(ids), remove_center=bool(ds.shape.nusc)))
    return out
