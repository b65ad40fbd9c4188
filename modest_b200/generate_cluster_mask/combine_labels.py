"""Self-training label merge: drop-in for the reference's generate_cluster_mask/combine_labels.py
(SURVEY.md 8(f-1)).  Same command line and config keys (configs/combine_labels.yaml).

Per frame of the detector's `result.pkl`: keep detections whose in-box PP-score percentile is at
most `det_filtering.pp_score_threshold` and whose score exceeds `score_filtering`
(combine_labels.py:42-60,101-107), add the seed boxes with an area score (:37-39,108), rank by
score and suppress by BEV IoU (:110-113), FOV gate, KITTI text (:117-121).  The point-in-box
percentile and the IoU run in libmodest_b200.
"""
import os
import os.path as osp
import pickle
import sys
from types import SimpleNamespace

import numpy as np
import torch

_HERE = osp.dirname(osp.abspath(__file__))
sys.path.insert(0, osp.dirname(osp.dirname(_HERE)))

from modest_b200 import _lib, dist, hydra_compat  # noqa: E402
from modest_b200 import pipeline as pl  # noqa: E402
from modest_b200.generate_cluster_mask.utils import kitti_util  # noqa: E402
from modest_b200.generate_cluster_mask.utils.pointcloud_utils import (is_within_fov, load_velo_scan, objs2label,  # noqa: E402
                                                                      objs_nms)

hydra_main, DictConfig, OmegaConf = hydra_compat.get_hydra()


def predicts2objs(preds):
    """combine_labels.py:23-34 -- OpenPCDet KITTI-style prediction dict -> box namespaces
    (dimensions are stored l, h, w)."""
    objs = []
    for i in range(preds['location'].shape[0]):
        obj = SimpleNamespace()
        obj.t = preds['location'][i]
        obj.l, obj.h, obj.w = preds['dimensions'][i][0], preds['dimensions'][i][1], preds['dimensions'][i][2]
        obj.ry = preds['rotation_y'][i]
        obj.score = preds['score'][i]
        objs.append(obj)
    return objs


def add_area_score(objs):
    """combine_labels.py:37-39 -- seed boxes rank below every detection, larger footprints first."""
    for obj in objs:
        obj.score = -999 + obj.w * obj.l


def in_box_pp_percentile(ptc_rect, pp_score, objs, percentile=50):
    """(percentile value f32, point count) per box, on the GPU."""
    n, k = int(ptc_rect.shape[0]), len(objs)
    if k == 0:
        return np.zeros(0, np.float32), np.zeros(0, np.int32)
    lib = _lib.lib()
    rect = torch.from_numpy(np.ascontiguousarray(ptc_rect[:, :3], dtype=np.float64)).cuda()
    pp = torch.from_numpy(np.ascontiguousarray(pp_score, dtype=np.float32)).cuda()
    rows = np.zeros((1, k, 8), dtype=np.float64)
    trig = np.zeros((1, k, 2), dtype=np.float64)
    for i, o in enumerate(objs):
        rows[0, i, :7] = (o.t[0], o.t[1], o.t[2], o.l, o.w, o.h, o.ry)
        trig[0, i] = (np.cos(o.ry), np.sin(o.ry))
    boxes, trig_d = torch.from_numpy(rows).cuda(), torch.from_numpy(trig).cuda()
    off = torch.tensor([0, n], dtype=torch.int64, device="cuda")
    nb = torch.tensor([k], dtype=torch.int32, device="cuda")
    dummy = torch.zeros((max(n, 1), 3), dtype=torch.float32, device="cuda")
    pct = torch.zeros((1, k), dtype=torch.float32, device="cuda")
    cnt = torch.zeros((1, k), dtype=torch.int32, device="cuda")
    ws = torch.empty(int(lib.modest_box_pp_workspace_bytes(1, n, n, k)), dtype=torch.uint8, device="cuda")
    q32 = float(np.float32(percentile) / np.float32(100))
    _lib.check(lib.modest_box_pp_percentile_batch(
        _lib.ptr(dummy), 3, _lib.ptr(off), _lib.ptr(pp), None, _lib.ptr(rect), _lib.ptr(boxes), _lib.ptr(trig_d),
        _lib.ptr(nb), 1, n, n, k, q32, _lib.ptr(pct), _lib.ptr(cnt), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
        "modest_box_pp_percentile_batch")
    return pct.cpu().numpy()[0], cnt.cpu().numpy()[0]


def filter_by_ppscore(ptc_rect, pp_score, obj, percentile=50, threshold=0.5):
    """combine_labels.py:42-60 -- False when the box is empty or its PP percentile exceeds threshold."""
    pct, cnt = in_box_pp_percentile(ptc_rect, pp_score, [obj], percentile)
    return bool(cnt[0] > 0 and not (pct[0] > np.float32(threshold)))


def eprint(*args, **kwargs):
    print(*args, file=sys.stderr, **kwargs)


def display_args(args):
    eprint("========== combine_labels info ==========")
    eprint("host: {}".format(os.getenv('HOSTNAME')))
    eprint(OmegaConf.to_yaml(args))
    eprint("=========================================")


@hydra_main(config_path="configs/", config_name="combine_labels.yaml")
def main(args: DictConfig):
    display_args(args)
    dist.init()
    det_bboxes = pickle.load(open(args.det_result_path, "rb"))
    idx_list = np.array([int(det_bbox['frame_id']) for det_bbox in det_bboxes])
    total_part, part = dist.resolve_parts(args.total_part, args.part)
    if total_part > 1:
        # the reference zips the sharded idx_list with the UNsharded detections (combine_labels.py:83-92),
        # which trips its own assert for part > 0; shard both consistently instead
        pick = np.array_split(np.arange(len(idx_list)), total_part)[part]
        idx_list, det_bboxes = idx_list[pick], [det_bboxes[i] for i in pick]
    os.makedirs(args.save_path, exist_ok=True)
    if args.data_paths.bbox_info_save_dst is None:
        eprint("Warning: not adding generated bboxes")
    for idx, det_bbox in zip(idx_list, det_bboxes):
        idx = int(idx)
        if args.data_paths.bbox_info_save_dst is not None:
            gen_obj = pickle.load(open(osp.join(args.data_paths.bbox_info_save_dst, f'{idx:06d}.pkl'), "rb"))
        else:
            gen_obj = []
        assert idx == int(det_bbox['frame_id'])
        calib = kitti_util.Calibration(osp.join(args.calib_path, f"{idx:06d}.txt"))
        ptc = load_velo_scan(osp.join(args.ptc_path, f"{idx:06d}.bin"))
        ptc_in_rect = calib.project_velo_to_rect(ptc[:, :3])
        pp_score = np.load(osp.join(args.data_paths.pp_score_path, f"{idx:06d}.npy"))
        cand = predicts2objs(det_bbox)
        pct, cnt = in_box_pp_percentile(ptc_in_rect, pp_score, cand, args.det_filtering.pp_score_percentile)
        thr = np.float32(args.det_filtering.pp_score_threshold)
        det_obj = [o for o, p, c in zip(cand, pct, cnt)
                   if (c > 0 and not (p > thr)) & (o.score > args.det_filtering.score_filtering)]
        add_area_score(gen_obj)
        objs = det_obj + gen_obj
        if len(objs) > 0:
            objs = objs_nms(objs, nms_threshold=args.nms.threshold, use_score_rank=True)
        if args.fov_only:
            objs = [obj for obj in objs if is_within_fov(obj, calib, args.image_shape)]
        with open(osp.join(args.save_path, f"{idx:06d}.txt"), "w") as f:
            f.write(objs2label(objs, calib, with_score=args.with_score))


if __name__ == "__main__":
    main()
