"""One tiled PP call on N scans (for ncu launch lists / captures).  python scripts/dev_pp_one.py [n] [group_points] [mode]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modest_b200 import pp_score, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
gp = int(sys.argv[2]) if len(sys.argv) > 2 else 0
mode = sys.argv[3] if len(sys.argv) > 3 else "tiled"
cases = [synth.make_scan_case(500 + i, synth.LYFT, n_traversals=16, n_points=60000) for i in range(n)]
b = pp_score.pack_batch([c.query_fixed for c in cases], [c.history for c in cases])
sc = pp_score.PPScorer(history_pass=mode, group_points=gp)
for _ in range(2):
    sc(b)
torch.cuda.synchronize()
