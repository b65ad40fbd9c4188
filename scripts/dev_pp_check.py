"""Developer probe (GPU): PP-score parity vs the oracle and a first timing."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from modest_b200 import synth, pp_score
from oracle import modest_oracle as orc

case = synth.make_scan_case(0, n_traversals=4)
t0 = time.time()
ref_counts = orc.neighbor_counts(case.query_fixed, case.history)
ref_pp = orc.persistence_entropy(ref_counts).astype(np.float32)
print("oracle s", time.time() - t0)
pp, counts = pp_score.count_neighbors_and_score(case.query_fixed, case.history, return_counts=True)
print("counts equal:", np.array_equal(counts, ref_counts), "mismatch", int((counts != ref_counts).sum()),
      "sum", counts.sum(), ref_counts.sum())
print("pp max abs err", np.abs(pp - ref_pp).max(), "bit-equal", np.array_equal(pp, ref_pp))

# timing: B scans x T=16
B = int(os.environ.get("B", 8))
cases = [synth.make_scan_case(100 + i, n_traversals=16) for i in range(B)]
batch = pp_score.pack_batch([c.query_fixed for c in cases], [c.history for c in cases])
scorer = pp_score.PPScorer()
out = scorer(batch)
torch.cuda.synchronize()
for g in (512, 768):
    scorer = pp_score.PPScorer(grid_dim=g)
    for _ in range(3):
        scorer(batch, out=out)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(10):
        scorer(batch, out=out)
    ev[1].record(); torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 10
    gbs = batch.algorithmic_bytes / (ms * 1e-3) / 1e9
    print(f"grid {g}: {ms*1e3/B:.1f} us/scan, {B/ms*1e3:.0f} scans/s, algorithmic {gbs:.1f} GB/s = {gbs/6450.6*100:.2f}% of 6450.6")
