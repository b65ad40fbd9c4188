"""Developer probe: time the PP hash pass of alternative builds of the library (exp_libs/lib_<name>.so).
python scripts/experiments/dev_pp_libs.py <lib.so> [n_scans]   -> prints history-pass us/scan and a digest of the counts"""
import ctypes
import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from modest_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.path.abspath(sys.argv[1])
from modest_b200 import pp_score, synth  # noqa: E402

n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
cases = [synth.make_scan_case(500 + i, synth.LYFT, n_traversals=16, frames_per_traversal=1, n_points=60000) for i in range(n)]
b = pp_score.pack_batch([c.query_fixed for c in cases], [c.history for c in cases])
lib = _lib.lib()
sc = pp_score.PPScorer()
counts = torch.zeros(b.n_count_total, dtype=torch.int32, device="cuda")
pp = sc(b, counts=counts)
torch.cuda.synchronize()
digest = hashlib.sha1(counts.cpu().numpy().tobytes() + pp.cpu().numpy().tobytes()).hexdigest()[:12]
for _ in range(3):
    sc(b)
torch.cuda.synchronize()
reps = 10
lib.modest_pp_profile_enable(reps)
for _ in range(reps):
    sc(b)
torch.cuda.synchronize()
buf = (ctypes.c_float * 256)()
k = lib.modest_pp_profile_read(buf, 256)
ms = sorted(buf[i] for i in range(k))
print(f"{os.path.basename(sys.argv[1]):16s} pp_count {1e3 * np.median(ms) / n:6.2f} us/scan (min {1e3 * ms[0] / n:6.2f})   digest {digest}", flush=True)
