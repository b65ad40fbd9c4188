"""Label-file program: drop-in for the reference's generate_cluster_mask/gen_label_files.py.

Same command line, config keys (configs/generate_label_files.yaml), input (`bbox_info_save_dst`
pkl files) and output (`<label_file_save_dst>/%06d.txt`, KITTI format).  The K x K BEV IoU runs in
libmodest_b200; FOV gate and text come from the library's host routine.

Under torchrun (one process per GPU) every rank labels its `np.array_split` shard and the label
blobs are collated with one all-gather (NCCL over NVLink when the ranks own GPUs) so that rank 0
writes the complete `label_2` directory; `gather=False` (extra key) keeps the reference's
behaviour of each shard writing only its own files.
"""
import os
import os.path as osp
import pickle
import sys

import numpy as np

_HERE = osp.dirname(osp.abspath(__file__))
sys.path.insert(0, osp.dirname(osp.dirname(_HERE)))

from modest_b200 import dist, hydra_compat  # noqa: E402
from modest_b200.generate_cluster_mask.utils import kitti_util  # noqa: E402
from modest_b200.generate_cluster_mask.utils.pointcloud_utils import is_within_fov, objs2label, objs_nms  # noqa: E402

hydra_main, DictConfig, OmegaConf = hydra_compat.get_hydra()


def eprint(*args, **kwargs):
    print(*args, file=sys.stderr, **kwargs)


def display_args(args):
    eprint("========== kitti_label gen info ==========")
    eprint("host: {}".format(os.getenv('HOSTNAME')))
    eprint(OmegaConf.to_yaml(args))
    eprint("==========================================")


@hydra_main(config_path="configs/", config_name="generate_label_files.yaml")
def main(args: DictConfig):
    display_args(args)
    dist.init()          # no-op unless launched by torchrun
    idx_list = np.array([int(x) for x in open(args.data_paths.idx_list).readlines()])
    total_part, part = dist.resolve_parts(args.total_part, args.part)
    if total_part > 1:
        idx_list = np.array_split(idx_list, total_part)[part]                  # gen_label_files.py:36-38
    os.makedirs(args.data_paths.label_file_save_dst, exist_ok=True)
    blobs = {}
    for idx in idx_list:
        idx = int(idx)
        objs = pickle.load(open(osp.join(args.data_paths.bbox_info_save_dst, f"{idx:06d}.pkl"), "rb"))
        if args.nms.enable and len(objs) > 0:
            objs = objs_nms(objs, nms_threshold=args.nms.threshold)             # :44-45
        calib = kitti_util.Calibration(osp.join(args.calib_path, f"{idx:06d}.txt"))
        if args.fov_only:
            objs = [obj for obj in objs if is_within_fov(obj, calib, args.image_shape)]
        blobs[idx] = objs2label(objs, calib).encode()
    if args.get("gather", True) and dist.world_size() > 1:
        blobs = dist.gather_blobs(blobs)                                       # the one collective
        if dist.rank() != 0:
            return
    for idx, text in blobs.items():
        with open(osp.join(args.data_paths.label_file_save_dst, f"{idx:06d}.txt"), "w") as f:
            f.write(text.decode())


if __name__ == "__main__":
    main()
