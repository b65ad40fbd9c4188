"""Developer probe: sorted-history vs global-hash PP history pass (parity + CUDA-event timing).
python scripts/dev_pp_sorted.py [n_scans]      env PP_F (history frames per traversal), PP_N (points), PP_MODES"""
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modest_b200 import _lib, pp_score, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
F = int(os.environ.get("PP_F", 1))
NPTS = int(os.environ.get("PP_N", 60000))
modes = os.environ.get("PP_MODES", "hash,sorted").split(",")
lib = _lib.lib()

# parity on small ragged cases first (incl. NaN rows, an empty traversal, a scan without history)
small = [synth.make_scan_case(40 + i, synth.LYFT, n_traversals=3 + i, frames_per_traversal=1, n_points=3000 + 1500 * i) for i in range(3)]
hs = [list(c.history) for c in small]
hs[1][0] = hs[1][0].copy(); hs[1][0][::7] = np.nan
hs[2][1] = hs[2][1][:0]
sb = pp_score.pack_batch([c.query_fixed for c in small], hs)
res = {}
for mode in modes:
    counts = torch.zeros(sb.n_count_total, dtype=torch.int32, device="cuda")
    pp = pp_score.PPScorer(history_pass=mode)(sb, counts=counts)
    torch.cuda.synchronize()
    res[mode] = (pp.cpu().numpy(), counts.cpu().numpy())
for mode in modes[1:]:
    print(f"small ragged: {mode} vs {modes[0]}: count mismatches {(res[mode][1] != res[modes[0]][1]).sum()}, pp bit-equal "
          f"{np.array_equal(res[mode][0].view(np.uint32), res[modes[0]][0].view(np.uint32))}", flush=True)

t0 = time.time()
cases = [synth.make_scan_case(500 + i, synth.LYFT, n_traversals=16, frames_per_traversal=F, n_points=NPTS) for i in range(n)]
print(f"{n} scans generated in {time.time() - t0:.1f}s", flush=True)
b = pp_score.pack_batch([c.query_fixed for c in cases], [c.history for c in cases])
ref = None
for mode in modes:
    sc = pp_score.PPScorer(history_pass=mode)
    counts = torch.zeros(b.n_count_total, dtype=torch.int32, device="cuda")
    pp = sc(b, counts=counts)
    torch.cuda.synchronize()
    r = (pp.cpu().numpy(), counts.cpu().numpy())
    if ref is None:
        ref = r
    else:
        print(f"  {mode}: count mismatches {(r[1] != ref[1]).sum()}  pp bit-equal {np.array_equal(r[0].view(np.uint32), ref[0].view(np.uint32))}")
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
    print(f"{mode:6s} history pass {1e3 * hist_ms / n:7.1f} us/scan ({frac * 100:5.2f} % of 6550 GB/s)   "
          f"whole PP stage {1e3 * stage_ms / n:7.1f} us/scan", flush=True)
