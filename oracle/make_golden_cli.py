"""TEST INFRASTRUCTURE ONLY -- BASELINE.json configs[0]: run the reference's own CLI `main()`s
(pre_compute_pp_score.py, generate_mask.py) on a synthetic KITTI-layout data_root and store
their output files as tests/golden/cli_lyft.npz.  Build-container only.

gen_label_files.py's main needs a GPU (objs_nms calls .cuda()), so its loop body
(gen_label_files.py:41-52) is replayed with the reference's functions and the reference's CPU
IoU op standing in for its CUDA op; the GPU test compares against these label files byte for
byte (and against the reference CUDA op in test_iou_bit_exact_vs_reference_cuda_kernel).

    python oracle/make_golden_cli.py
"""
import io
import os
import pickle
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from modest_b200 import synth  # noqa: E402
from oracle import reference_harness as rh  # noqa: E402

N_SCANS = 2          # scans processed (the first N of the idx list), one global RNG stream
DATASET = dict(n_traversals=3, frames_per_traversal=2, history_frames=1)


def make_args(ref, root, work):
    meta = os.path.join(work, "meta")
    dp = dict(track_path=os.path.join(meta, "track_list.pkl"), idx_info=os.path.join(meta, "valid_idx_info.pkl"),
              load_precomputed_lidars=None, load_save_precomputed_trans_mat=None,
              idx_list=os.path.join(meta, "train_idx.txt"), pp_score_path=os.path.join(work, "pp"),
              seg_save_dst=os.path.join(work, "seg"), bbox_info_save_dst=os.path.join(work, "bbox"),
              label_file_save_dst=os.path.join(work, "labels"))
    common = dict(work_dir=work, save_dir=work, total_part=1, part=0, data_root=root,
                  calib_path=os.path.join(root, "calib"), ptc_path=os.path.join(root, "velodyne"), data_paths=dp)
    pp = ref.AttrDict(dict(common, seed=1024, max_neighbor_dist=0.3, remove_ground_plane=False, limit_traversals=-1,
                           nusc=False, add_random_noise=0, skip_ephe=False, ephe_type="entropy"))
    gm = ref.AttrDict(dict(common, plane_estimate=dict(range=[[-70, 70], [-20, 20]], max_hs=-1.5, offset=0.05),
                           limit_range=[[-70, 70], [-40, 40]],
                           graph=dict(neighbor_type="radius_mutual_knn", affinity_type="l1", n_neighbors=70, radius=2.),
                           clustering=dict(method="DBSCAN", DBSCAN=dict(eps=0.1, min_samples=10)),
                           filtering=dict(min_points=10, max_volume=120, min_volume=0.5, min_max_height=0.5,
                                          max_min_height=1., percentile=20, min_percentile_pp_score=0.7),
                           bbox_gen=dict(fit_method="closeness_to_edge")))
    return pp, gm, dp


def main():
    ref = rh.load()
    ref_gm = ref.gm
    work = tempfile.mkdtemp(prefix="modest_cli_golden_")
    root = os.path.join(work, "data")
    info = synth.write_dataset(root, os.path.join(work, "meta"), synth.LYFT, **DATASET)
    ids = info["idx"][:N_SCANS]
    with open(os.path.join(work, "meta", "train_idx.txt"), "w") as f:
        f.write("\n".join(f"{x:06d}" for x in ids))
    pp_args, gm_args, dp = make_args(ref, root, work)
    pp_args["data_paths"] = dict(dp, idx_list=dp["idx_list"])
    ref.pp.main(pp_args)                                   # the reference's PP program, unmodified
    np.random.seed(synth.SEED_BASE)                        # one seed for the whole generate_mask run
    ref_gm.main(gm_args)                                   # the reference's seed-mask program, unmodified
    import torch
    from oracle import build_ref
    ext = build_ref.load_cpu()
    out = {}
    for idx in ids:
        objs = pickle.load(open(os.path.join(dp["bbox_info_save_dst"], f"{idx:06d}.pkl"), "rb"))
        calib = ref.ku.Calibration(os.path.join(root, "calib", f"{idx:06d}.txt"))
        if len(objs):
            boxes = torch.from_numpy(np.array([[o.t[0], o.t[2], 0, o.l, o.w, o.h, -o.ry] for o in objs])).float()
            iou = torch.zeros((len(objs), len(objs)))
            ext.boxes_iou_bev_cpu(boxes.contiguous(), boxes.contiguous(), iou)
            iou = iou.numpy()
            keep = np.ones(len(objs), dtype=bool)
            for i in np.diag(iou).argsort()[::-1]:
                if keep[i]:
                    keep[iou[i] > 0.1] = False
                    keep[i] = True
            objs_k = [o for o, k in zip(objs, keep) if k]
        else:
            objs_k = objs
        objs_k = [o for o in objs_k if ref.pc.is_within_fov(o, calib, [1024, 1224])]
        out[f"label_{idx}"] = ref.pc.objs2label(objs_k, calib)
        out[f"pp_{idx}"] = np.load(os.path.join(dp["pp_score_path"], f"{idx:06d}.npy"))
        out[f"seg_{idx}"] = np.load(os.path.join(dp["seg_save_dst"], f"{idx:06d}.npy")).astype(np.int32)
        out[f"boxes_{idx}"] = np.array([[*o.t, o.l, o.w, o.h, o.ry, o.volume] for o in objs]).reshape(-1, 8)
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cli_lyft.npz"), ids=np.array(ids), **out)
    for idx in ids:
        print(idx, out[f"pp_{idx}"].shape, out[f"seg_{idx}"].max(), len(out[f"boxes_{idx}"]),
              out[f"label_{idx}"].count("\n") + 1)
    shutil.rmtree(work)


if __name__ == "__main__":
    main()
