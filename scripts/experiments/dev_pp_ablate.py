"""Developer probe (GPU): where the history pass spends its time -- v6 with parts switched off
(results are wrong in the ablated runs; timing only)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from modest_b200 import _lib, synth, pp_score
lib = _lib.lib()
def tune(k, v): assert lib.modest_pp_tune(k, v) == 0
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
pool = [synth.make_scan_case(100 + i, n_traversals=16) for i in range(B)]
batch = pp_score.pack_batch([c.query_fixed for c in pool], [c.history for c in pool])
scorer = pp_score.PPScorer()
out = scorer(batch)
torch.cuda.synchronize()
buf = (ctypes.c_float * 64)()
for name, var, abl in (("v6 full", 6, 0), ("v6 no RED", 6, 1), ("v6 no test rounds (phase 1 + prefix)", 6, 2),
                       ("v6 point loads only", 6, 4), ("v5 full", 5, 0), ("v5 no RED", 5, 1), ("v5 no rounds", 5, 2)):
    tune(0, var); tune(1, 1024); tune(4, abl)
    for _ in range(3):
        scorer(batch, out=out)
    torch.cuda.synchronize()
    lib.modest_pp_profile_enable(16)
    for _ in range(10):
        scorer(batch, out=out)
    torch.cuda.synchronize()
    n = lib.modest_pp_profile_read(buf, 64)
    kms = float(np.mean([buf[i] for i in range(n)]))
    lib.modest_pp_profile_enable(0)
    print(f"{name:40s} count kernel {kms*1e3/B:6.1f} us/scan", flush=True)
tune(4, 0)
