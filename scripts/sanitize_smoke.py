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
out = list(e.process(fr.jobs_from_dataset(ds, ds.scan_ids, 4)))
torch.cuda.synchronize()
print("sanitize target ok:", sum(len(t) for _, ts in out for t in ts), "label bytes")
