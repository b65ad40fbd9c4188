"""Wall time of the three drop-in programs on a synthetic KITTI-layout data_root (what a MODEST user runs):
pre_compute_pp_score.py, generate_mask.py (parity mode: numpy RNG stream, one scan per launch; throughput mode:
rng=device batch_size=12) and gen_label_files.py, next to the reference's CPU path (oracle port) on a few of the
same scans.  python scripts/time_cli.py [frames_per_traversal]"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cfg(prog, root, work, **extra):
    from modest_b200 import hydra_compat
    cfg_dir = os.path.join(ROOT, "modest_b200", "generate_cluster_mask", "configs")
    meta = os.path.join(work, "meta")
    ov = [f"data_root={root}", f"data_paths.track_path={meta}/track_list.pkl", f"data_paths.idx_info={meta}/valid_idx_info.pkl",
          f"data_paths.idx_list={meta}/train_idx.txt", f"data_paths.pp_score_path={work}/pp",
          f"data_paths.seg_save_dst={work}/seg", f"data_paths.bbox_info_save_dst={work}/bbox",
          f"data_paths.label_file_save_dst={work}/labels"] + [f"{k}={v}" for k, v in extra.items()]
    return hydra_compat.compose(cfg_dir, prog, ov, cwd=work, run_dir=work)


def main():
    import shutil
    import torch
    from modest_b200 import synth
    from modest_b200.generate_cluster_mask import gen_label_files, generate_mask, pre_compute_pp_score
    fpt = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    with tempfile.TemporaryDirectory() as work:
        root = os.path.join(work, "data")
        info = synth.write_dataset(root, os.path.join(work, "meta"), synth.LYFT, n_traversals=4, frames_per_traversal=fpt,
                                   history_frames=1)
        n = len(info["idx"])
        devnull = open(os.devnull, "w")
        stderr, sys.stderr = sys.stderr, devnull          # the programs print their resolved config

        def timed(fn, *a):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn(*a)
            torch.cuda.synchronize()
            return time.perf_counter() - t0
        rows = []
        timed(pre_compute_pp_score.main, cfg("pp_score.yaml", root, work))            # warm-up (JIT-free, but first CUDA use)
        rows.append(("pre_compute_pp_score.py", timed(pre_compute_pp_score.main, cfg("pp_score.yaml", root, work))))
        np.random.seed(1024)
        timed(generate_mask.main, cfg("generate_mask.yaml", root, work))
        for d in ("seg", "bbox"):
            shutil.rmtree(os.path.join(work, d))
        np.random.seed(1024)
        rows.append(("generate_mask.py (parity mode, rng=numpy, 1 scan per launch)", timed(generate_mask.main, cfg("generate_mask.yaml", root, work))))
        for d in ("seg", "bbox"):
            shutil.rmtree(os.path.join(work, d))
        rows.append(("generate_mask.py rng=device batch_size=12", timed(generate_mask.main, cfg("generate_mask.yaml", root, work, rng="device", batch_size=12))))
        rows.append(("gen_label_files.py", timed(gen_label_files.main, cfg("generate_label_files.yaml", root, work))))
        sys.stderr = stderr
        # the reference's CPU path on 3 of the scans (oracle port: the same library calls)
        from oracle import modest_oracle as orc
        import pickle
        track = pickle.load(open(os.path.join(work, "meta", "track_list.pkl"), "rb"))
        t0 = time.perf_counter()
        k = 0
        for sid in info["idx"][:3]:
            ptc = np.fromfile(os.path.join(root, "velodyne", f"{sid:06d}.bin"), dtype=np.float32).reshape(-1, 4)
            pp = np.load(os.path.join(work, "pp", f"{sid:06d}.npy"))
            cal = orc.Calib(path=os.path.join(root, "calib", f"{sid:06d}.txt"))
            labels, objs = orc.seed_mask_for_scan(ptc, pp, cal, seed=1024 + sid)
            orc.labels_for_scan(objs, cal, lambda b: orc.bev_iou_matrix_f32(b, b))
            k += 1
        cpu = (time.perf_counter() - t0) / k
    print(f"# The drop-in programs on a synthetic data_root: {n} scans of 60 000 points, 4 traversals x 1 history frame\n")
    print(f"device: {torch.cuda.get_device_name(0)}; wall time of `main()` incl. file reads / writes, second run of each program\n")
    print("| program | s total | ms per scan |")
    print("|---|---|---|")
    for name, t in rows:
        print(f"| {name} | {t:.2f} | {1e3 * t / n:.1f} |")
    print(f"\nReference CPU path for generate_mask + gen_label_files (oracle port, same library calls, this host): {cpu:.2f} s per scan.")


if __name__ == "__main__":
    main()
