"""Benchmark of the seed-label hot path (BASELINE.json metric: LiDAR scans/sec through
PP-score + RANSAC + DBSCAN + NMS at 60k points).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one pass of the whole hot path over one batch of synthetic Lyft-shaped scans
(60 000 points, 16 historical traversals of one frame each): PP score -> RANSAC plane ->
masks -> mutual-kNN graph -> DBSCAN -> second plane -> cluster gates -> box fit -> BEV NMS.
`value` times the device work with the inputs resident in HBM (CUDA events, max over ranks);
`e2e` times the public API with HOST inputs: pinned host -> device copies, the same kernels,
device -> host copy of boxes / keep flags, KITTI label text (and, for N > 1, the one
all-gather that collates the label blobs).  `roofline` is the PP neighbour-count kernel:
algorithmic bytes (12 B per query point + 12 B per history point + 4 B per score) over its
CUDA-event duration, against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
`--impl reference` times the reference's own CPU path (oracle port: SciPy cKDTree +
scikit-learn, what the reference's programs call) on the host cores.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POINTS, N_TRAV = 60000, 16
# dram__bytes_read.sum + dram__bytes_write.sum of one pp_count_kernel launch over 24 scans, from the
# `ncu --set full` capture summarised in profiles/r1b_pp_count_bench_launch_ncu_full_summary.csv
# (425.8 MB read + 69.1 MB written); the kernel's traffic is proportional to the scans per launch
PP_COUNT_DRAM_TRAFFIC_PER_SCAN = (425_806_080 + 69_080_320) / 24
METRIC = "LiDAR scans/sec (PP-score+RANSAC+DBSCAN+NMS) @60k pts"
WORKLOAD = "full seed-label pipeline, synthetic Lyft-shape scans (60k pts, 16 traversals x 1 frame)"


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def host_info():
    """CPU model and library versions of the box the CPU numbers were taken on."""
    model = "unknown"
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.lower().startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    import scipy
    import sklearn
    return {"cpu_model": model, "cpu_count": os.cpu_count(), "numpy": np.__version__, "scipy": scipy.__version__,
            "sklearn": sklearn.__version__}


def make_pool(n_scans, seed0):
    from modest_b200 import synth
    return [synth.make_scan_case(seed0 + i, synth.LYFT, n_traversals=N_TRAV, frames_per_traversal=1,
                                 n_points=N_POINTS) for i in range(n_scans)]


# ------------------------------------------------------------------------------------------------
def cpu_one_scan(case):
    """The reference's CPU path for one scan (oracle port), returns the label text."""
    from oracle import modest_oracle as orc
    pp = orc.pp_score(case.query_fixed, case.history)
    cal = orc.Calib(table=case.calib)
    labels, objs = orc.seed_mask_for_scan(case.query, pp, cal, seed=1024 + case.scan_id)
    text, _ = orc.labels_for_scan(objs, cal, lambda b: orc.bev_iou_matrix_f32(b, b))
    return len(text)


def _cpu_worker(scan_id):
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = "1"
    from modest_b200 import synth
    case = synth.make_scan_case(scan_id, synth.LYFT, n_traversals=N_TRAV, frames_per_traversal=1, n_points=N_POINTS)
    t0 = time.perf_counter()
    cpu_one_scan(case)
    return time.perf_counter() - t0


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation on the host cores.  One step =
    one scan per worker process (data-parallel over scans, the reference's total_part sharding),
    generation of the synthetic inputs excluded from the timing."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 32))
    ctx = mp.get_context("spawn")
    per_step = []
    with ctx.Pool(workers) as pool:
        for step in range(args.warmup + args.steps):
            ids = [100000 + step * workers + w for w in range(workers)]
            times = pool.map(_cpu_worker, ids)
            if step >= args.warmup:
                per_step.append(max(times))          # the step ends when its slowest worker ends
    total_t = float(sum(per_step))
    value = workers * args.steps / total_t
    line = {"metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "scans_per_step": workers},
            "cpu_baseline": {"value": value, "unit": "scans/s", "cores": workers, "kind": "port",
                             "sample": f"{workers} scans per step, one per process (cKDTree + sklearn, 1 thread each)",
                             "host": host_info()},
            "e2e": {"value": value, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as td
    from modest_b200 import _lib, dist
    from modest_b200 import pipeline as pl
    from modest_b200 import pp_score as pp_mod

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    numa_cpus = dist.bind_to_gpu_numa(local)      # before any pinned allocation (first touch decides the node)
    if world > 1:
        dist.init("nccl")
    lib = _lib.lib()
    B = args.scans_per_step
    cases = make_pool(B, 1000 + 10007 * rank)

    # ---- host-side (pinned) copies for the e2e path, device-resident copies for `value`
    q_fixed = [torch.from_numpy(c.query_fixed).pin_memory() for c in cases]
    hist = [[torch.from_numpy(h).pin_memory() for h in c.history] for c in cases]
    ptc_host = [torch.from_numpy(c.query).pin_memory() for c in cases]
    calibs = [c.calib for c in cases]
    h2d_bytes = sum(t.numel() * 4 for t in q_fixed) + sum(t.numel() * 4 for h in hist for t in h) + \
        sum(t.numel() * 4 for t in ptc_host)     # == host_batch.h2d_bytes

    scorer = pp_mod.PPScorer()
    pipe = pl.SeedLabelPipeline()
    pp_batch = pp_mod.pack_batch(q_fixed, hist)
    pp_out = torch.empty(pp_batch.n_query_total, dtype=torch.float32, device="cuda")
    scan_batch = pl.make_batch(ptc_host, [torch.zeros(N_POINTS) for _ in cases], calibs)
    scan_batch.pp = pp_out

    # independent pipelines on their own streams: consecutive steps alternate between them so that
    # the many one-CTA-per-scan kernels of one step overlap the wide kernels of the other
    lanes = [(torch.cuda.Stream(), pp_mod.PPScorer(), pl.SeedLabelPipeline(),
              torch.empty(pp_batch.n_query_total, dtype=torch.float32, device="cuda")) for _ in range(args.streams)]
    lane_batches = []
    for (_, _, _, ppbuf) in lanes:
        sbk = pl.make_batch(ptc_host, [torch.zeros(N_POINTS) for _ in cases], calibs)
        sbk.pp = ppbuf
        lane_batches.append(sbk)

    def device_step(seed):
        st, sc, pp_, ppbuf = lanes[seed % len(lanes)]
        with torch.cuda.stream(st):
            sc(pp_batch, out=ppbuf, stream=st)
            return pp_.run(lane_batches[seed % len(lanes)], rng="device", seed=seed, stream=st)

    from modest_b200 import engine as eng
    host_batch = eng.make_host_batch([c.query_fixed for c in cases], [c.history for c in cases],
                                     [c.query for c in cases], calibs, scan_ids=[rank * B + s for s in range(B)])
    engine = eng.SeedLabelEngine()
    d2h_bytes = [0]

    def e2e_run(n_steps):
        """The public streaming API: pinned host batches in, label text out; H2D of batch k+1
        overlaps the kernels of batch k and the text formatting of batch k-1."""
        n_lines = 0
        blobs = {}
        for k, (ids, texts) in enumerate(engine.process(host_batch for _ in range(n_steps))):
            if world > 1:
                blobs.update({k * world * B + i: t.encode() for i, t in zip(ids, texts)})   # scan ids are rank * B + s
            n_lines += sum(t.count("\n") + 1 for t in texts if t)
        if world > 1:       # the path's one collective: label files collated on rank 0 once per run
            dist.gather_blobs(blobs)
        d2h_bytes[0] = engine.d2h_bytes_last
        return n_lines

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    for w in range(max(args.warmup, 3)):
        device_step(w)
    barrier()
    assert lib.modest_pp_profile_enable(min(args.steps, 256)) == 0

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lib.modest_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    main_stream = torch.cuda.current_stream()
    ev0.record(main_stream)
    for st, *_ in lanes:
        st.wait_event(ev0)
    for k in range(args.steps):
        res = device_step(100 + k)
    for st, *_ in lanes:                     # the end event fires when every lane has drained
        e = torch.cuda.Event()
        e.record(st)
        main_stream.wait_event(e)
    ev1.record(main_stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = lib.modest_launch_count() - launches0
    clocks = sampler.stop()
    buf = (ctypes.c_float * 256)()
    n_prof = lib.modest_pp_profile_read(buf, 256)
    pp_ms_overlapped = float(np.mean([buf[i] for i in range(n_prof)])) if n_prof else float("nan")
    # The roofline kernel on its own: in the loop above its launches share the SMs with the other
    # lanes' kernels, which stretches every launch.  The same steps once more on ONE lane give
    # the kernel's own duration and a step it can be compared with (what the ncu launch list of
    # `--streams 1` shows as well); both durations are reported.
    k_single = min(args.steps, 8)
    assert lib.modest_pp_profile_enable(k_single) == 0
    lanes_all, lanes[:] = list(lanes), lanes[:1]
    device_step(0)
    barrier()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record(main_stream)
    lanes[0][0].wait_event(ev2)
    for k in range(k_single):
        device_step(200 + k)
    e = torch.cuda.Event()
    e.record(lanes[0][0])
    main_stream.wait_event(e)
    ev3.record(main_stream)
    barrier()
    ms_single = ev2.elapsed_time(ev3) / k_single
    n_prof = lib.modest_pp_profile_read(buf, 256)
    pp_ms = float(np.mean([buf[i] for i in range(min(n_prof, k_single))])) if n_prof else float("nan")
    lib.modest_pp_profile_enable(0)
    lanes[:] = lanes_all
    n_boxes = int(res.n_boxes.sum().item())

    # ---- e2e (host buffers in, label text out)
    # raw pinned host -> device bandwidth of this box, for context next to the e2e number
    probe = torch.empty(64 << 20, dtype=torch.float32).pin_memory()
    probe_d = torch.empty_like(probe, device="cuda")
    probe_d.copy_(probe, non_blocking=True)
    torch.cuda.synchronize()
    tp = time.perf_counter()
    for _ in range(4):
        probe_d.copy_(probe, non_blocking=True)
    torch.cuda.synchronize()
    h2d_gbs = 4 * probe.numel() * 4 / (time.perf_counter() - tp) / 1e9
    del probe, probe_d
    e2e_run(max(args.warmup, 6))          # every slot of the engine's ring is touched before timing
    barrier()
    engine.host_s.update(upload=0.0, launch=0.0, wait=0.0, text=0.0, batches=0)
    t0 = time.perf_counter()
    e2e_run(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    host_ms = {k: round(1e3 * v / max(engine.host_s["batches"], 1), 2) for k, v in engine.host_s.items() if k != "batches"}

    if world > 1:
        t = torch.tensor([ms, e2e_s * 1e3, pp_ms, ms_single, pp_ms_overlapped], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        ms, e2e_s, pp_ms, ms_single, pp_ms_overlapped = float(t[0]), float(t[1]) / 1e3, float(t[2]), float(t[3]), float(t[4])
    if world > 1:
        td.barrier()
        td.destroy_process_group()
    if rank != 0:
        return
    total_scans = B * world * args.steps
    value = total_scans / (ms * 1e-3)
    peak, peak_src = hbm_peak()
    alg_bytes = pp_batch.algorithmic_bytes
    achieved = alg_bytes / (pp_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "scans_per_step_per_gpu": B, "n_points": N_POINTS, "n_traversals": N_TRAV,
                   "ransac": "device-drawn minimal sets, 100 trials scored, sklearn accept/early-stop replay",
                   "l2": f"inputs larger than L2: {h2d_bytes / 1e6:.0f} MB touched per step",
                   "streams": args.streams,
                   "host_cpus": (f"{len(numa_cpus)} CPUs of the GPU's NUMA node" if numa_cpus else
                                 f"{len(os.sched_getaffinity(0))} (no NUMA binding: topology not exposed or already local)"),
                   "boxes_last_step": n_boxes},
        "clocks": clocks,
        "e2e": {"value": total_scans / e2e_s, "unit": "scans/s", "h2d_bytes_per_step": int(h2d_bytes),
                "d2h_bytes_per_step": int(d2h_bytes[0]), "h2d_link_gbs_measured": round(h2d_gbs, 1),
                "h2d_gbs_used": round(h2d_bytes * args.steps / e2e_s / 1e9, 1),
                "host_ms_per_batch": host_ms},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "pp_count_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": int(PP_COUNT_DRAM_TRAFFIC_PER_SCAN * B),
                     "traffic_source": "ncu --set full of a 24-scan launch (profiles/r1b_pp_count_bench_launch_ncu_full_summary.csv), "
                                       "scaled to this launch's scan count",
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": int(alg_bytes), "kernel_ms": pp_ms,
                     "share_of_step": pp_ms / ms_single,
                     "timed_on": f"{k_single} single-lane steps after the throughput loop ({ms_single:.3f} ms per step); "
                                 "in the overlapped loop the same launch lasts kernel_ms_overlapped",
                     "kernel_ms_overlapped": pp_ms_overlapped},
    }
    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only)
    if world == 1 and not args.no_cpu_baseline:
        t0 = time.perf_counter()
        cpu_one_scan(cases[0])
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": 1.0 / dt, "unit": "scans/s", "cores": 1 if (os.cpu_count() or 1) == 1 else os.cpu_count(),
                                "kind": "port",
                                "sample": "1 scan of the same workload, one process: cKDTree single-threaded, "
                                          "sklearn graph/DBSCAN with n_jobs=-1 as the reference calls them",
                                "host": host_info()}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scans-per-step", type=int, default=48)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=3, help="independent pipeline lanes for the device-resident loop")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
