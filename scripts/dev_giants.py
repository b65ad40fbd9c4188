import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modest_b200 import pipeline as pl, pp_score, synth
ds = synth.make_track_dataset(synth.LYFT, n_traversals=16, frames_per_traversal=2, n_points=60000, seed=1024, workers=0)
cases = [synth.scan_case_from_dataset(ds, sid) for sid in ds.scan_ids[:6]]
b = pp_score.pack_batch([c.query_fixed for c in cases], [c.history for c in cases])
pp = pp_score.PPScorer()(b)
sb = pl.make_batch([c.query for c in cases], [pp[b.h_q_off[s]:b.h_q_off[s + 1]] for s in range(len(cases))], [c.calib for c in cases],
                   scan_ids=[c.scan_id for c in cases])
p = pl.SeedLabelPipeline()
r = p.run(sb, rng="device", seed=1)
lf, lfin = r.labels_filtered.cpu().numpy(), r.labels.cpu().numpy()
for s in range(len(cases)):
    a, f = lf[sb.h_off[s]:sb.h_off[s + 1]], lfin[sb.h_off[s]:sb.h_off[s + 1]]
    sizes = np.bincount(a)[1:]
    kept = np.array([f[a == k + 1].max() > 0 for k in range(len(sizes))])
    print("scan", s, "valid clusters", len(sizes), "kept", int(kept.sum()), "dropped sizes", np.sort(sizes[~kept])[::-1][:8], "kept max", sizes[kept].max() if kept.any() else 0)
