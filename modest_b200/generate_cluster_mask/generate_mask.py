"""Seed-mask program: drop-in for the reference's generate_cluster_mask/generate_mask.py.

Same command line, config keys (configs/generate_mask.yaml), inputs and outputs:
`<seg_save_dst>/%06d.npy` (int64 (N,) labels), `<bbox_info_save_dst>/%06d.pkl` (pickled list of
SimpleNamespace(t, l, w, h, ry, volume)) and a `configs.yaml` beside each.  Every numeric step
of the reference's loop body (generate_mask.py:52-103) runs in libmodest_b200.

By default each scan draws its RANSAC minimal sets from numpy's global RandomState in the
reference's order, so a run seeded like a reference run yields the same files.  `rng=device`
(extra key) switches to on-device draws and lets `batch_size` scans share every launch.
"""
import os
import os.path as osp
import pickle
import sys

import numpy as np

_HERE = osp.dirname(osp.abspath(__file__))
sys.path.insert(0, osp.dirname(osp.dirname(_HERE)))

from modest_b200 import dist, hydra_compat  # noqa: E402
from modest_b200 import pipeline as pl  # noqa: E402
from modest_b200.generate_cluster_mask.utils import kitti_util  # noqa: E402
from modest_b200.generate_cluster_mask.utils.pointcloud_utils import box_namespace, load_velo_scan  # noqa: E402

hydra_main, DictConfig, OmegaConf = hydra_compat.get_hydra()


def eprint(*args, **kwargs):
    print(*args, file=sys.stderr, **kwargs)


def display_args(args):
    eprint("========== clustering info ==========")
    eprint("host: {}".format(os.getenv('HOSTNAME')))
    eprint(OmegaConf.to_yaml(args))
    eprint("=====================================")


def _save_config_once(args, dst):
    os.makedirs(dst, exist_ok=True)
    if not osp.exists(osp.join(dst, "configs.yaml")):
        OmegaConf.save(config=args, f=osp.join(dst, "configs.yaml"))


@hydra_main(config_path="configs/", config_name="generate_mask.yaml")
def main(args: DictConfig):
    display_args(args)
    dist.init()          # no-op unless launched by torchrun
    idx_list = np.array([int(x) for x in open(args.data_paths.idx_list).readlines()])
    total_part, part = dist.resolve_parts(args.total_part, args.part)
    if total_part > 1:
        idx_list = np.array_split(idx_list, total_part)[part]                  # generate_mask.py:35-37
    _save_config_once(args, args.data_paths.seg_save_dst)
    bbox_dst = args.data_paths.get("bbox_info_save_dst", "None")
    if bbox_dst is not None:
        _save_config_once(args, bbox_dst)
    rng_mode = args.get("rng", "numpy")
    batch_size = int(args.get("batch_size", 1)) if rng_mode == "device" else 1
    pipe = pl.SeedLabelPipeline(args)
    todo = []
    for idx in idx_list:
        idx = int(idx)
        # generate_mask.py:48-51 (its args.get("bbox_info_save_dst") looks at the root config and
        # is therefore always the string "None": the pkl has to exist for a skip)
        if osp.exists(osp.join(args.data_paths.seg_save_dst, f"{idx:06d}.npy")) and \
                osp.exists(osp.join(args.data_paths.bbox_info_save_dst, f"{idx:06d}.pkl")):
            continue
        todo.append(idx)
    for s0 in range(0, len(todo), batch_size):
        chunk = todo[s0:s0 + batch_size]
        ptcs = [load_velo_scan(osp.join(args.ptc_path, f"{i:06d}.bin")) for i in chunk]
        pps = [np.load(osp.join(args.data_paths.pp_score_path, f"{i:06d}.npy")) for i in chunk]
        calibs = [kitti_util.Calibration(osp.join(args.calib_path, f"{i:06d}.txt")) for i in chunk]
        batch = pl.make_batch(ptcs, pps, calibs, scan_ids=chunk)
        res = pipe.run(batch, rng=rng_mode, seed=int(args.get("seed", 0)))     # device draws are keyed by scan id: batch-invariant
        pipe.check_flags(res)
        labels = res.labels.cpu().numpy().astype(np.int64)
        boxes, n_boxes = res.boxes.cpu().numpy(), res.n_boxes.cpu().numpy()
        for s, idx in enumerate(chunk):
            objs = [box_namespace(boxes[s, k]) for k in range(int(n_boxes[s]))]
            if bbox_dst is not None:
                pickle.dump(objs, open(osp.join(bbox_dst, f"{idx:06d}.pkl"), "wb"))
            np.save(osp.join(args.data_paths.seg_save_dst, f"{idx:06d}.npy"),
                    labels[batch.h_off[s]:batch.h_off[s + 1]])


if __name__ == "__main__":
    main()
