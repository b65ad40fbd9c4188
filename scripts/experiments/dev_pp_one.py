"""Developer probe (GPU): one PP launch of a chosen variant (for ncu).
Usage: python scripts/dev_pp_one.py variant chunk deal_fixed deal_per [n_scans]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from modest_b200 import _lib, synth, pp_score
lib = _lib.lib()
var, chunk, df, dp = (int(a) for a in sys.argv[1:5])
B = int(sys.argv[5]) if len(sys.argv) > 5 else 8
for k, v in enumerate((var, chunk, df, dp)):
    assert lib.modest_pp_tune(k, v) == 0
pool = [synth.make_scan_case(100 + i, n_traversals=16) for i in range(B)]
batch = pp_score.pack_batch([c.query_fixed for c in pool], [c.history for c in pool])
scorer = pp_score.PPScorer()
for _ in range(3):
    out = scorer(batch)
torch.cuda.synchronize()
