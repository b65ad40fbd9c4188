"""Developer probe (GPU): the experimental merged-index + query-centric join path (MODEST_PP_JOIN)."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from modest_b200 import _lib, synth, pp_score
from oracle import modest_oracle as orc
lib = _lib.lib()
for case in (synth.make_scan_case(3, synth.LYFT, n_traversals=3, n_points=6000), synth.make_scan_case(0, n_traversals=4)):
    ref = orc.neighbor_counts(case.query_fixed, case.history)
    ref_pp = orc.persistence_entropy(ref).astype(np.float32)
    os.environ["MODEST_PP_JOIN"] = "1"
    pp, counts = pp_score.count_neighbors_and_score(case.query_fixed, case.history, return_counts=True)
    print("join: count mismatches", int((counts != ref).sum()), "pp bit-equal", np.array_equal(pp, ref_pp), flush=True)
B = 8
pool = [synth.make_scan_case(100 + i, n_traversals=16) for i in range(B)]
batch = pp_score.pack_batch([c.query_fixed for c in pool], [c.history for c in pool])
scorer = pp_score.PPScorer()
buf = (ctypes.c_float * 64)()
for name, mode in (("shipped pp_count_kernel", None), ("history index + join", "1"), ("join only", "2")):
    os.environ.pop("MODEST_PP_JOIN", None)
    if mode:
        os.environ["MODEST_PP_JOIN"] = mode
    for _ in range(2):
        out = scorer(batch)
    torch.cuda.synchronize()
    lib.modest_pp_profile_enable(8)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(6):
        scorer(batch, out=out)
    ev[1].record(); torch.cuda.synchronize()
    n = lib.modest_pp_profile_read(buf, 64)
    kms = float(np.mean([buf[i] for i in range(n)]))
    lib.modest_pp_profile_enable(0)
    print(f"{name:26s} timed part {kms*1e3/B:6.1f} us/scan   whole PP stage {ev[0].elapsed_time(ev[1])/6*1e3/B:6.1f} us/scan", flush=True)
