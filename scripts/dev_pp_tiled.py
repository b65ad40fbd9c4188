"""Developer probe: tiled vs global-hash PP history pass on N Lyft-shaped scans (parity + CUDA-event timing).
python scripts/dev_pp_tiled.py [n_scans] [group_points ...]"""
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modest_b200 import _lib, pp_score, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
groups = [int(a) for a in sys.argv[2:]] or [0]
F = int(os.environ.get("PP_F", 1))            # history frames per traversal
NPTS = int(os.environ.get("PP_N", 60000))
t0 = time.time()
cases = [synth.make_scan_case(500 + i, synth.LYFT, n_traversals=16, frames_per_traversal=F, n_points=NPTS) for i in range(n)]
print(f"{n} scans generated in {time.time() - t0:.1f}s", flush=True)
b = pp_score.pack_batch([c.query_fixed for c in cases], [c.history for c in cases])
lib = _lib.lib()
ref = None
for mode, gp in [("hash", 0)] + [("tiled", g) for g in groups]:
    sc = pp_score.PPScorer(history_pass=mode, group_points=gp)
    counts = torch.zeros(b.n_count_total, dtype=torch.int32, device="cuda")
    pp = sc(b, counts=counts)
    torch.cuda.synchronize()
    res = (pp.cpu().numpy(), counts.cpu().numpy())
    if ref is None:
        ref = res
    else:
        print(f"  {mode} gp={gp}: count mismatches {(res[1] != ref[1]).sum()}  pp bit-equal {np.array_equal(res[0].view(np.uint32), ref[0].view(np.uint32))}")
    for _ in range(3):
        sc(b)
    torch.cuda.synchronize()
    reps = 10
    lib.modest_pp_profile_enable(reps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sc(b)
    e1.record()
    torch.cuda.synchronize()
    buf = (ctypes.c_float * 256)()
    k = lib.modest_pp_profile_read(buf, 256)
    hist_ms = float(np.mean([buf[i] for i in range(k)]))
    lib.modest_pp_profile_enable(0)
    stage_ms = e0.elapsed_time(e1) / reps
    frac = b.algorithmic_bytes / (hist_ms * 1e-3) / 1e9 / 6550.1
    print(f"{mode:6s} group_points={gp:<9d} history pass {1e3 * hist_ms / n:7.1f} us/scan ({frac * 100:5.2f} % of 6550 GB/s)   "
          f"whole PP stage {1e3 * stage_ms / n:7.1f} us/scan", flush=True)
