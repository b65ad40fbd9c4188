"""Benchmark of the seed-label hot path (BASELINE.json metric: LiDAR scans/sec through
PP-score + RANSAC + DBSCAN + NMS at 60k points).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload: a synthetic Lyft-shaped drive -- 16 traversals of one road, a 60 000-point frame every
2 m; every frame is a query scan whose history is the nearest frame of each of the 16 traversals
(BASELINE "60k pts, 16 traversals", one frame per traversal).  K steps x 48 scans = K*48 DISTINCT
scans per GPU, each processed once per timed loop.  A step = one pass of the whole hot path over
one batch of 48 scans: stage B (frames into the scan's fixed frame) -> PP score -> RANSAC plane ->
masks -> mutual-kNN graph -> DBSCAN -> second plane -> cluster gates -> box fit -> BEV NMS.

`value`  device work with the inputs resident in HBM (all raw frames in the device frame cache),
         CUDA events over the engine's lanes, max over ranks.
`e2e`    the public streaming API (SeedLabelEngine.process on frames.JobBatch) from HOST data:
         every raw frame is uploaded from pinned host memory inside the timed region (once: the
         device cache keeps it for the later scans that share it), poses / job tables per batch,
         device -> host copy of boxes / keep flags, KITTI label text; for N > 1 the one
         all-gather that collates the label blobs.
`roofline` the PP history pass (pp_count_kernel): algorithmic bytes (12 B per query point + 12 B
         per history point + 4 B per score) over its CUDA-event duration, against the measured HBM
         copy bandwidth in MEASURED_PEAKS.json.
`nusc`   BASELINE config 5 in small: nuScenes-shaped drive (34k points, 16 traversals x 16 frames
         of history per scan, max_hs=-1.3, centre removal), end to end incl. label files on disk.
`lyft_f36` the Lyft shape at the real history depth (36 frames per traversal, 34.6 M history points
         per scan), end to end.
`--impl reference` times the reference's own CPU path (oracle port: SciPy cKDTree +
scikit-learn, what the reference's programs call) on the host cores, on scans of the same drive.
"""
import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POINTS, N_TRAV = 60000, 16
# dram__bytes_read.sum + dram__bytes_write.sum of one pp_count_kernel launch over 24 scans, from the
# latest `ncu --set full` capture of this bench's own launch (second session of round 2), summarised in
# profiles/r2b_pp_count_knn_beta32_ncu_full_summary.csv (426.0 MB read + 69.6 MB written for 24 scans;
# earlier captures: 425.9 + 69.6, round 1: 425.8 + 69.1); the kernel's traffic is proportional to
# the scans per launch
PP_COUNT_DRAM_TRAFFIC_PER_SCAN = (426_021_632 + 69_637_632) / 24
METRIC = "LiDAR scans/sec (PP-score+RANSAC+DBSCAN+NMS) @60k pts"
WORKLOAD = "full seed-label pipeline, synthetic Lyft-shape drive (60k pts, 16 traversals x 1 frame of history per scan)"


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


def drive(shape_name, n_scans, rank, world, history_frames=1, n_points=None, id_base=0):
    """The synthetic drive of one rank: >= n_scans frames in 16 traversals (frame ids = scan ids)."""
    from modest_b200 import synth
    shape = synth.NUSC if shape_name == "nusc" else synth.LYFT
    per = max(math.ceil(n_scans / N_TRAV), history_frames if history_frames > 1 else 2)
    workers = max(1, min(16, (os.cpu_count() or 1) // max(world, 1)))
    return synth.make_track_dataset(shape, n_traversals=N_TRAV, frames_per_traversal=per, history_frames=history_frames,
                                    n_points=n_points, seed=1024 + 10007 * rank + (7 if shape.nusc else 0),
                                    first_frame_id=rank * 10_000_000 + id_base, workers=workers)


# ------------------------------------------------------------------------------------------------
def cpu_one_scan(case):
    """The reference's CPU path for one scan (oracle port), returns the label text length."""
    from oracle import modest_oracle as orc
    pp = orc.pp_score(case.query_fixed, case.history)
    cal = orc.Calib(table=case.calib)
    labels, objs = orc.seed_mask_for_scan(case.query, pp, cal, seed=1024 + case.scan_id)
    text, _ = orc.labels_for_scan(objs, cal, lambda b: orc.bev_iou_matrix_f32(b, b))
    return len(text)


_CPU_DRIVE = {}


def _cpu_worker(k):
    """One scan of a small drive (generated once per worker process, outside the timing): stage B on
    the host (transform_points arithmetic), then the CPU path."""
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = "1"
    from modest_b200 import synth
    if "ds" not in _CPU_DRIVE:
        _CPU_DRIVE["ds"] = synth.make_track_dataset(synth.LYFT, n_traversals=N_TRAV, frames_per_traversal=4, history_frames=1,
                                                    n_points=N_POINTS, seed=500000)
    ds = _CPU_DRIVE["ds"]
    t0 = time.perf_counter()
    case = synth.scan_case_from_dataset(ds, ds.scan_ids[k % len(ds.scan_ids)])
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
            times = pool.map(_cpu_worker, [step * workers + w for w in range(workers)], chunksize=1)
            if step >= args.warmup:
                per_step.append(max(times))          # the step ends when its slowest worker ends
    total_t = float(sum(per_step))
    value = workers * args.steps / total_t
    line = {"metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "scans_per_step_per_gpu": args.scans_per_step, "n_points": N_POINTS,
                       "n_traversals": N_TRAV, "sample_scans_per_step": workers,
                       "sample": "each step times a bounded sample of the workload: one scan per worker process"},
            "cpu_baseline": {"value": value, "unit": "scans/s", "cores": workers, "kind": "port",
                             "sample": f"{workers} scans per step, one per process (transform_points + cKDTree + sklearn, 1 thread each)",
                             "host": host_info()},
            "e2e": {"value": value, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as td
    from modest_b200 import _lib, dist
    from modest_b200 import engine as eng
    from modest_b200 import frames as fr
    from modest_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    numa_cpus = dist.bind_to_gpu_numa(local)      # before any pinned allocation (first touch decides the node)
    if world > 1:
        dist.init("nccl")
    lib = _lib.lib()
    B = args.scans_per_step
    t_gen = time.perf_counter()
    ds = drive("lyft", args.steps * B, rank, world, n_points=N_POINTS)
    warm = drive("lyft", B, rank, world, n_points=N_POINTS, id_base=5_000_000)
    gen_s = time.perf_counter() - t_gen
    scan_ids = ds.scan_ids[:args.steps * B]
    source = fr.pinned_frame_source({**ds.frames, **warm.frames})
    for f in list(ds.frames) + list(warm.frames):          # pin everything before any timing
        source(f)
    frame_bytes = sum(ds.frames[f].nbytes for f in ds.frames)
    jobs = fr.jobs_from_dataset(ds, scan_ids, B)
    warm_jobs = fr.jobs_from_dataset(warm, warm.scan_ids[:B], B)

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ device-resident `value`
    engine = eng.SeedLabelEngine(frame_source=source, depth=args.streams - 1)
    for f in list(ds.frames) + list(warm.frames):          # inputs resident in HBM: every raw frame cached
        engine.frame_cache.get(f)
    torch.cuda.synchronize()
    slots = engine.slots[:args.streams]

    def device_step(k, jb, slot_list):
        slot = slot_list[k % len(slot_list)]
        engine._upload_jobs(slot, jb)                      # job tables only (all frames are cached)
        engine._compute(slot, k)
        return slot

    def timed_loop(batches, slot_list, per_step_events=False):
        main = torch.cuda.current_stream()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps_ev = []
        ev0.record(main)
        engine.copy_stream.wait_event(ev0)
        for sl in slot_list:
            sl.stream.wait_event(ev0)
        for k, jb in enumerate(batches):
            if per_step_events:
                a = torch.cuda.Event(enable_timing=True)
                sl = slot_list[k % len(slot_list)]
                a.record(sl.stream)
            slot = device_step(k, jb, slot_list)
            if per_step_events:
                b_ = torch.cuda.Event(enable_timing=True)
                b_.record(slot.stream)
                steps_ev.append((a, b_))
        for sl in slot_list:                     # the end event fires when every lane has drained
            main.wait_event(sl.done)
        ev1.record(main)
        return ev0, ev1, steps_ev

    # warm-up: every lane runs the batch shape three times (eager, CUDA-graph capture, one replay)
    n_warm = max(args.warmup, 3 * len(slots))
    for w in range(n_warm):
        device_step(w, warm_jobs[0], slots)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lib.modest_launch_count() + engine.launches_replayed
    replays0 = engine.graph_replays
    barrier()
    ev0, ev1, steps_ev = timed_loop(jobs, slots, per_step_events=True)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = lib.modest_launch_count() + engine.launches_replayed - launches0
    graph_replays = engine.graph_replays - replays0
    clocks = sampler.stop()
    step_ms = sorted(a.elapsed_time(b_) for a, b_ in steps_ev)
    n_boxes = int(slots[(len(jobs) - 1) % len(slots)].result.n_boxes.sum().item())
    # The roofline kernel is timed by event pairs the library records around it, which a replayed
    # graph does not contain: the two probes below launch eagerly (same kernels, same order).
    engine.use_graphs = False
    buf = (ctypes.c_float * 256)()
    k_over = min(args.steps, 2 * len(slots))
    assert lib.modest_pp_profile_enable(k_over) == 0
    timed_loop(jobs[:k_over], slots)
    barrier()
    n_prof = lib.modest_pp_profile_read(buf, 256)
    pp_ms_overlapped = float(np.mean([buf[i] for i in range(min(n_prof, k_over))])) if n_prof else float("nan")
    # The roofline kernel on its own: in the loop above its launches share the SMs with the other
    # lanes' kernels, which stretches every launch.  Some of the steps once more on ONE lane give
    # the kernel's own duration and a step it can be compared with (what the ncu launch list of
    # `--streams 1` shows as well); both durations are reported.
    k_single = min(args.steps, 8)
    assert lib.modest_pp_profile_enable(k_single) == 0
    device_step(0, warm_jobs[0], slots[:1])
    barrier()
    ev2, ev3, _ = timed_loop(jobs[:k_single], slots[:1])
    barrier()
    ms_single = ev2.elapsed_time(ev3) / k_single
    n_prof = lib.modest_pp_profile_read(buf, 256)
    pp_ms = float(np.mean([buf[i] for i in range(min(n_prof, k_single))])) if n_prof else float("nan")
    lib.modest_pp_profile_enable(0)
    pp_alg_bytes = slots[0].pp_batch.algorithmic_bytes
    del engine, slots
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ e2e (host frames in, label text out)
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
    engine = eng.SeedLabelEngine(frame_source=source, depth=args.e2e_depth)

    def e2e_run(batches, collate):
        """The public streaming API: job batches (frame ids + poses) in, label text out; frames the
        device cache does not hold yet are uploaded from pinned host memory on the way."""
        n_lines, blobs = 0, {}
        for ids, texts in engine.process(batches):
            if collate:
                blobs.update({int(i): t.encode() for i, t in zip(ids, texts)})
            n_lines += sum(t.count("\n") + 1 for t in texts if t)
        if collate and world > 1:       # the path's one collective: label files collated on rank 0 once per run
            dist.gather_blobs(blobs)
        return n_lines

    # every slot of the ring sees the batch shape three times (eager, graph capture, replay)
    e2e_run((warm_jobs[0] for _ in range(max(args.warmup, 3 * len(engine.slots)))), False)
    barrier()
    engine.host_s.update(upload=0.0, launch=0.0, wait=0.0, text=0.0, batches=0)
    h2d0 = engine.frame_cache.h2d_bytes + engine.h2d_bytes_tables
    t0 = time.perf_counter()
    # the job batches (pose chains, job tables) are built inside the timed region, lazily
    e2e_run((b for i in range(0, len(scan_ids), B) for b in fr.jobs_from_dataset(ds, scan_ids[i:i + B], B)), True)
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d_bytes = engine.frame_cache.h2d_bytes + engine.h2d_bytes_tables - h2d0
    d2h_bytes = engine.d2h_bytes_last
    host_ms = {k: round(1e3 * v / max(engine.host_s["batches"], 1), 2) for k, v in engine.host_s.items() if k != "batches"}
    del engine
    torch.cuda.empty_cache()

    nusc = lyft36 = None
    if not args.no_nusc:
        nusc = run_secondary(args, rank, world, barrier, "nusc", 16, 8,
                             "nuScenes-shape drive: 34k-pt scans, 16 traversals x 16 history frames per scan, max_hs=-1.3, "
                             "centre removal on history frames, label files written")
        lyft36 = run_secondary(args, rank, world, barrier, "lyft", 36, 4,
                               "Lyft-shape drive at the real history depth: 60k-pt scans, 16 traversals x 36 history frames "
                               "per scan (34.6 M history points, 415 MB of PP input per scan)")

    if world > 1:
        t = torch.tensor([ms, e2e_s * 1e3, pp_ms, ms_single, pp_ms_overlapped], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        ms, e2e_s, pp_ms, ms_single, pp_ms_overlapped = float(t[0]), float(t[1]) / 1e3, float(t[2]), float(t[3]), float(t[4])
    if world > 1:
        td.barrier()
        td.destroy_process_group()
    if rank != 0:
        return
    total_scans = len(scan_ids) * world
    value = total_scans / (ms * 1e-3)
    peak, peak_src = hbm_peak()
    achieved = pp_alg_bytes / (pp_ms * 1e-3) / 1e9
    pct = lambda p: step_ms[min(len(step_ms) - 1, int(p * len(step_ms)))]
    line = {
        "metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": world, "steps": args.steps,
        "warmup": n_warm, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "scans_per_step_per_gpu": B, "n_points": N_POINTS, "n_traversals": N_TRAV,
                   "distinct_scans_per_gpu": len(scan_ids), "frames_per_gpu": len(ds.frames),
                   "stages": "B (frame transform) + C..O",
                   "ransac": "device-drawn minimal sets keyed by scan id, 100 trials scored, sklearn accept/early-stop replay",
                   "l2": f"inputs larger than L2: {frame_bytes / 1e6:.0f} MB of raw frames, every step reads other frames and "
                         f"writes {13.2 * B:.0f} MB of transformed points",
                   "streams": args.streams, "step_ms_on_its_lane": {"p50": pct(0.5), "p95": pct(0.95), "max": step_ms[-1]},
                   "host_cpus": (f"{len(numa_cpus)} CPUs of the GPU's NUMA node" if numa_cpus else
                                 f"{len(os.sched_getaffinity(0))} (no NUMA binding: topology not exposed or already local)"),
                   "boxes_last_step": n_boxes, "dataset_generation_s": round(gen_s, 1)},
        "clocks": clocks,
        "e2e": {"value": total_scans / e2e_s, "unit": "scans/s", "h2d_bytes_per_step": int(h2d_bytes / args.steps),
                "d2h_bytes_per_step": int(d2h_bytes), "h2d_link_gbs_measured": round(h2d_gbs, 1),
                "h2d_gbs_used": round(h2d_bytes / e2e_s / 1e9, 2),
                "what": "SeedLabelEngine.process(JobBatch): every raw frame uploaded once from pinned host memory inside the "
                        "timed region and kept in the device frame cache, poses + job tables per batch, label text out",
                "host_ms_per_batch": host_ms},
        "gpu_launches": int(launches),
        "cuda_graphs": {"replays_in_timed_region": int(graph_replays),
                        "what": "a lane's launch sequence is captured the second time it sees a batch shape and replayed "
                                "afterwards; gpu_launches counts the kernels inside the replays"},
        "roofline": {"bound": "hbm", "kernel": "pp_count_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": int(PP_COUNT_DRAM_TRAFFIC_PER_SCAN * B),
                     "traffic_source": "ncu --set full of a 24-scan launch (profiles/r2b_pp_count_knn_beta32_ncu_full_summary.csv), "
                                       "scaled to this launch's scan count",
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": int(pp_alg_bytes), "kernel_ms": pp_ms,
                     "share_of_step": pp_ms / ms_single,
                     "timed_on": f"{k_single} single-lane steps after the throughput loop ({ms_single:.3f} ms per step); "
                                 "in the overlapped loop the same launch lasts kernel_ms_overlapped",
                     "kernel_ms_overlapped": pp_ms_overlapped},
    }
    if nusc is not None:
        line["nusc"] = nusc
        line["lyft_f36"] = lyft36
    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only)
    if world == 1 and not args.no_cpu_baseline:
        n_cpu = 3
        t0 = time.perf_counter()
        for k in range(n_cpu):
            cpu_one_scan(synth.scan_case_from_dataset(ds, scan_ids[k * 7]))
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n_cpu / dt, "unit": "scans/s", "cores": os.cpu_count() or 1,
                                "kind": "port",
                                "sample": f"{n_cpu} scans of the same drive, one process: transform_points, cKDTree "
                                          "single-threaded, sklearn graph/DBSCAN with n_jobs=-1 as the reference calls them",
                                "host": host_info()}
    print(json.dumps(line))


def run_secondary(args, rank, world, barrier, shape_name, F, Bn, what):
    """A second drive end to end through the engine: nuScenes shape (BASELINE config 5 in small, label files
    written) or Lyft shape with 36 history frames per traversal (the real Lyft history depth)."""
    import torch
    import torch.distributed as td
    from modest_b200 import dist
    from modest_b200 import engine as eng
    from modest_b200 import frames as fr
    from modest_b200 import synth
    shape = synth.NUSC if shape_name == "nusc" else synth.LYFT
    steps = max(2, min(args.steps, 4))
    ds = drive(shape_name, Bn * steps, rank, world, history_frames=F, id_base=2_000_000 if shape.nusc else 3_000_000)
    ids = ds.scan_ids[:Bn * steps]
    source = fr.pinned_frame_source(ds.frames)
    for f in ds.frames:
        source(f)
    cfg = dict(plane_estimate=dict(range=[[-70, 70], [-20, 20]], max_hs=shape.max_hs, offset=0.05),
               image_shape=list(shape.image_shape))
    engine = eng.SeedLabelEngine(cfg, frame_source=source)
    warm_ids = ds.scan_ids[-Bn:]
    list(engine.process(fr.jobs_from_dataset(ds, warm_ids, Bn) * (3 * len(engine.slots))))     # ring touched three times (eager, graph capture, replay)
    engine.frame_cache.clear()
    barrier()
    h2d0 = engine.frame_cache.h2d_bytes + engine.h2d_bytes_tables
    texts = {}
    t0 = time.perf_counter()
    for b_ids, tx in engine.process(b for i in range(0, len(ids), Bn) for b in fr.jobs_from_dataset(ds, ids[i:i + Bn], Bn)):
        texts.update({int(i): t.encode() for i, t in zip(b_ids, tx)})
    if world > 1:
        texts = dist.gather_blobs(texts)
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = engine.frame_cache.h2d_bytes + engine.h2d_bytes_tables - h2d0
    t1 = time.perf_counter()
    n_files = 0
    if rank == 0:
        with tempfile.TemporaryDirectory() as tmp:
            out = os.path.join(tmp, "label_2")
            os.makedirs(out)
            for sid, blob in texts.items():
                with open(os.path.join(out, f"{sid % 1000000:06d}.txt"), "wb") as fh:
                    fh.write(blob)
                n_files += 1
    write_s = time.perf_counter() - t1
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        e2e_s = float(t[0])
    hist_pts = sum(ds.frames[f].shape[0] for g in ds.history_frames(ids[0]) for f in g)
    n_q = ds.frames[ids[0]].shape[0]
    del engine
    torch.cuda.empty_cache()
    return {"workload": what,
            "e2e": {"value": len(ids) * world / e2e_s, "unit": "scans/s", "scans": len(ids) * world,
                    "h2d_bytes_per_scan": int(h2d / len(ids)), "history_points_per_scan": int(hist_pts),
                    "pp_algorithmic_bytes_per_scan": int(12 * n_q + 12 * hist_pts + 4 * n_q)},
            "label_files": {"written": n_files, "seconds": round(write_s, 4)},
            "non_empty_labels": int(sum(1 for v in texts.values() if v))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scans-per-step", type=int, default=48)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-nusc", action="store_true", help="skip the secondary measurements (nuScenes shape, Lyft F=36)")
    ap.add_argument("--streams", type=int, default=3, help="pipeline lanes of the device-resident loop")
    ap.add_argument("--e2e-depth", type=int, default=2, help="batches the engine keeps computing while the host reads back an older one")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
