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
    """Raw frames resident on one GPU, bounded by `device_bytes`, LRU eviction.  `source(fid)` must
    return the frame as a pinned (N,4) float32 host tensor (e.g. velodyne/%06d.bin read into pinned
    memory); uploads are enqueued on the stream that is current when `get` is called."""

    def __init__(self, source, device_bytes=48 << 30):
        self.source, self.device_bytes = source, int(device_bytes)
        self._d, self._used = OrderedDict(), 0
        self.hits = self.misses = 0
        self.h2d_bytes = 0

    def __contains__(self, fid):
        return fid in self._d

    def get(self, fid):
        t = self._d.get(fid)
        if t is not None:
            self._d.move_to_end(fid)
            self.hits += 1
            return t
        self.misses += 1
        host = self.source(fid)
        t = torch.empty(host.shape, dtype=torch.float32, device="cuda")
        t.copy_(host, non_blocking=True)
        nbytes = t.numel() * 4
        self.h2d_bytes += nbytes
        while self._d and self._used + nbytes > self.device_bytes:
            _, old = self._d.popitem(last=False)
            self._used -= old.numel() * 4
        self._d[fid] = t
        self._used += nbytes
        return t

    def clear(self):
        self._d.clear()
        self._used = 0


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
    (synth.relative_pose_f32 restates pre_compute_pp_score.py:27-28)."""
    out = []
    for i in range(0, len(scan_ids), batch_size):
        ids = list(scan_ids[i:i + batch_size])
        q_T, h_fid, h_T, fpt = [], [], [], []
        for sid in ids:
            groups = ds.history_frames(sid)
            fpt.append([len(g) for g in groups])
            flat = [f for g in groups for f in g]
            poses = ds.relative_poses(sid, [sid] + flat)          # one stacked solve per scan
            q_T.append(poses[0])
            h_fid.extend(flat)
            h_T.append(poses[1:])
        out.append(JobBatch(scan_ids=ids, query_fid=np.array(ids, dtype=np.int64), query_T=np.stack(q_T).astype(np.float32),
                            hist_fid=np.array(h_fid, dtype=np.int64), hist_T=np.concatenate(h_T).astype(np.float32),
                            frames_per_trav=fpt, calibs=[ds.calib] * len(ids), remove_center=bool(ds.shape.nusc)))
    return out
