"""compute-sanitizer target: smoke() (PP + whole pipeline on a 6k-point scan, parity-checked) plus the
tiled PP pass and the engine's frame-job path on small inputs.  Run as
    compute-sanitizer --tool memcheck|racecheck|initcheck|synccheck python scripts/sanitize_smoke.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from modest_b200 import engine as eng, frames as fr, pp_score, synth  # noqa: E402

ge.smoke()
case = synth.make_scan_case(3, synth.LYFT, n_traversals=3, n_points=6000)
a = pp_score.count_neighbors_and_score(case.query_fixed, case.history, return_counts=True, history_pass="tiled")
b = pp_score.count_neighbors_and_score(case.query_fixed, case.history, return_counts=True, history_pass="hash")
assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])
ds = synth.make_track_dataset(synth.NUSC, n_traversals=3, frames_per_traversal=2, history_frames=2, n_points=4000, seed=5)
e = eng.SeedLabelEngine(frame_source=fr.pinned_frame_source(ds.frames))
jobs = fr.jobs_from_dataset(ds, ds.scan_ids, 4)
out = list(e.process(jobs * 5))            # every lane sees a shape three times: eager, CUDA-graph capture, replay
assert e.graph_replays > 0
torch.cuda.synchronize()
# f-4 operator kernels (brute-force graphs, edge affinities, the three other fitters, lowest point)
from modest_b200.generate_cluster_mask.utils import clustering_utils as cu, pointcloud_utils as pu  # noqa: E402
rng = np.random.default_rng(0)
pts = np.concatenate([rng.normal(0, 3, (700, 3)), rng.uniform(0, 1, (700, 1))], 1).astype(np.float32)
ppv = rng.uniform(0, 1, 700).astype(np.float32)
for nt, at in (("knn", "l1"), ("sym_knn", "exp"), ("mutual_knn", "3d_l2_distance"), ("radius", "l1"), ("radius_mutual_knn", "exp")):
    g = cu.precompute_affinity_matrix(pts, ppv, nt, at, 9, 1.2)
    assert g.nnz > 0
rect = rng.normal(0, 4, (3000, 3))
cl = rect[:400] * [0.5, 0.2, 0.2] + [3, 0, 2]
rect = np.concatenate([rect, cl])
for m in ("min_zx_area_fit", "PCA", "variance_to_edge", "closeness_to_edge"):
    o = pu.get_obj(cl, rect, m)
    assert np.isfinite([o.l, o.w, o.h, o.ry, o.volume]).all()
# a scan with a huge cluster (box pre-rejection) and the parallel finalize
big = synth.make_scan_case(9, synth.LYFT, n_traversals=2, n_points=30000)
ppb = pp_score.count_neighbors_and_score(big.query_fixed, big.history)
from modest_b200 import pipeline as pl  # noqa: E402
pipe = pl.SeedLabelPipeline()
r = pipe.run(pl.make_batch([big.query, case.query], [ppb, b[0]], [big.calib, case.calib], scan_ids=[9, 3]), rng="device", seed=2)
pipe.check_flags(r)
torch.cuda.synchronize()
print("sanitize target ok:", sum(len(t) for _, ts in out for t in ts), "label bytes;", int(r.n_boxes.sum()), "boxes")
