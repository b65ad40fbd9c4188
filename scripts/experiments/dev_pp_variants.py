"""Developer probe (GPU): every variant of the PP history pass against the oracle, then timed.
Usage: python scripts/dev_pp_variants.py [n_scans]"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from modest_b200 import _lib, synth, pp_score
from oracle import modest_oracle as orc

lib = _lib.lib()
def tune(k, v): assert lib.modest_pp_tune(k, v) == 0

ALWAYS_DEAL, NEVER_DEAL = -(1 << 30), 1 << 30
configs = [("v3", 0, 2048, 0, 0), ("v4 lane pf", 1, 2048, 0, 0), ("v4 lane nopf", 3, 2048, 0, 0),
           ("v4 deal pf", 2, 2048, ALWAYS_DEAL, 0), ("v4 adaptive pf", 2, 2048, 640, 40),
           ("v4 adaptive nopf", 4, 2048, 640, 40), ("v5", 5, 2048, 0, 0), ("v6", 6, 2048, 0, 0), ("v6 short 12", 6, 1024, 12, 0)]

# ---- parity
cases = [synth.make_scan_case(3, synth.LYFT, n_traversals=3, n_points=6000),
         synth.make_scan_case(0, n_traversals=4),
         synth.make_scan_case(5, synth.NUSC, n_traversals=3, frames_per_traversal=2)]
for ci, case in enumerate(cases):
    ref = orc.neighbor_counts(case.query_fixed, case.history)
    for name, var, chunk, df, dp in configs:
        tune(0, var); tune(1, chunk); tune(2, df); tune(3, dp)
        pp, counts = pp_score.count_neighbors_and_score(case.query_fixed, case.history, return_counts=True)
        bad = int((counts != ref).sum())
        print(f"case {ci} {name:18s} mismatches {bad} (sum {counts.sum()} ref {ref.sum()})", flush=True)

# ---- timing
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
pool = [synth.make_scan_case(100 + i, n_traversals=16) for i in range(B)]
batch = pp_score.pack_batch([c.query_fixed for c in pool], [c.history for c in pool])
scorer = pp_score.PPScorer()
out = scorer(batch)
torch.cuda.synchronize()
sweep = list(configs)
for sr in (4, 8, 12, 16, 24, 32):
    sweep.append((f"v6 chunk1024 short_row {sr}", 6, 1024, sr, 0))
buf = (ctypes.c_float * 64)()
for name, var, chunk, df, dp in sweep:
    tune(0, var); tune(1, chunk); tune(2, df); tune(3, dp)
    for _ in range(3):
        scorer(batch, out=out)
    torch.cuda.synchronize()
    lib.modest_pp_profile_enable(16)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(10):
        scorer(batch, out=out)
    ev[1].record(); torch.cuda.synchronize()
    n = lib.modest_pp_profile_read(buf, 64)
    kms = float(np.mean([buf[i] for i in range(n)]))
    lib.modest_pp_profile_enable(0)
    ms = ev[0].elapsed_time(ev[1]) / 10
    gbs = batch.algorithmic_bytes / (kms * 1e-3) / 1e9
    print(f"{name:28s} count kernel {kms*1e3/B:6.1f} us/scan ({gbs:7.1f} GB/s = {gbs/6550.1*100:5.2f}%)  whole PP stage {ms*1e3/B:6.1f} us/scan", flush=True)
tune(0, 2); tune(1, 2048); tune(2, 640); tune(3, 40)
