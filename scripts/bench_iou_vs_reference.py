"""Stage N, native vs native: libmodest_b200's BEV IoU / NMS against the reference's own
iou3d_nms_cuda extension (built unmodified for sm_100 by oracle/build_ref.py) on the same box.
    python scripts/bench_iou_vs_reference.py > profiles/<name>.md
Calls are timed the way the reference's callers make them (iou3d_nms_utils.py:37-51,54-74):
boxes_iou_bev_gpu is asynchronous (CUDA events), nms_gpu ends with the blocking copy of the keep
list (wall clock around the call, device idle before)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402
from modest_b200.generate_cluster_mask.utils.iou3d_nms import iou3d_nms_cuda as ours  # noqa: E402

ref = build_ref.load()
if ref is None:
    print("oracle/_ref/iou3d_nms_cuda.so not built: nothing to compare with")
    sys.exit(0)


def boxes(n, seed):
    rng = np.random.default_rng(seed)
    b = np.zeros((n, 7), np.float32)
    side = 6.0 * np.sqrt(n)                 # ~ constant box density: a few overlaps per box
    b[:, 0] = rng.uniform(-side, side, n); b[:, 1] = rng.uniform(-side, side, n)
    b[:, 3] = rng.uniform(3.5, 5.0, n); b[:, 4] = rng.uniform(1.6, 2.1, n); b[:, 5] = rng.uniform(1.4, 1.9, n)
    b[:, 6] = rng.uniform(-np.pi, np.pi, n)
    return torch.from_numpy(b).cuda()


def time_events(fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps      # us


def time_wall(fn, reps=100):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return 1e6 * (time.perf_counter() - t0) / reps


print("# Stage N: ours vs the reference's iou3d_nms_cuda (sm_100 build), same B200, same boxes\n")
print(f"device: {torch.cuda.get_device_name(0)}; times in microseconds per call\n")
print("| K boxes | boxes_iou_bev_gpu ref | ours | ratio | nms_gpu (incl. blocking D2H) ref | ours | ratio | equal |")
print("|---|---|---|---|---|---|---|---|")
for n in (16, 64, 300, 2000):
    bt = boxes(n, n)
    o_ref = torch.zeros((n, n), device="cuda"); o_our = torch.zeros((n, n), device="cuda")
    t_ref = time_events(lambda: ref.boxes_iou_bev_gpu(bt, bt, o_ref))
    t_our = time_events(lambda: ours.boxes_iou_bev_gpu(bt, bt, o_our))
    k_ref = torch.zeros(n, dtype=torch.long); k_our = torch.zeros(n, dtype=torch.long)
    n_ref = ref.nms_gpu(bt, k_ref, 0.1); n_our = ours.nms_gpu(bt, k_our, 0.1)
    tn_ref = time_wall(lambda: ref.nms_gpu(bt, k_ref, 0.1))
    tn_our = time_wall(lambda: ours.nms_gpu(bt, k_our, 0.1))
    same = torch.equal(o_ref, o_our) and n_ref == n_our and torch.equal(k_ref[:n_ref], k_our[:n_our])
    print(f"| {n} | {t_ref:.1f} | {t_our:.1f} | {t_ref / t_our:.2f}x | {tn_ref:.1f} | {tn_our:.1f} | {tn_ref / tn_our:.2f}x | {same} |")
print("\nThe seed-label path itself never makes these calls one scan at a time: `modest_seed_nms_batch` runs the K x K IoU,"
      " the ordering and the greedy suppression of every scan of a batch in one launch (one CTA per scan), without the host"
      " round trip objs_nms makes per scan (pointcloud_utils.py:322-341).")
S = 48
from modest_b200 import _lib  # noqa: E402
from modest_b200 import pipeline as pl  # noqa: E402
p = pl.SeedLabelPipeline()
bx = torch.zeros((S, p.max_boxes, 8), dtype=torch.float64, device="cuda")
nb = torch.full((S,), 30, dtype=torch.int32, device="cuda")
for s in range(S):
    b7 = boxes(30, 100 + s).double()
    bx[s, :30, 0] = b7[:, 0]; bx[s, :30, 2] = b7[:, 1]; bx[s, :30, 3] = b7[:, 3]; bx[s, :30, 4] = b7[:, 4]
    bx[s, :30, 5] = b7[:, 5]; bx[s, :30, 6] = -b7[:, 6]
t_batch = time_events(lambda: p.seed_nms(bx, nb))
b30 = boxes(30, 7)
o30 = torch.zeros((30, 30), device="cuda")
t_ref_scan = time_wall(lambda: (ref.boxes_iou_bev_gpu(b30, b30, o30), o30.cpu()))
print(f"\n| 48 scans x 30 boxes | reference way (IoU kernel + .cpu() per scan, host NMS not counted) | modest_seed_nms_batch (one launch) |")
print("|---|---|---|")
print(f"| us per batch | {48 * t_ref_scan:.0f} | {t_batch:.0f} |")
