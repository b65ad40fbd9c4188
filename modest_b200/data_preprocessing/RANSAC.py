"""Road-plane files: drop-in for the reference's data_preprocessing/RANSAC.py.

Same command line (--calib_dir --lidar_dir --planes_dir --min_h --max_h --split_file), same
output (`<planes_dir>/<idx>.txt`: '# Plane', 'Width 4', 'Height 1', four '{:e}' numbers, no
trailing newline), and -- like the reference (RANSAC.py:83-85) -- nothing is done when
planes_dir already exists.  Rect projection, candidate gate, MAD threshold, hypothesis scoring
and the refit run in libmodest_b200 (float64 throughout, as sklearn sees float64 rect
coordinates); the minimal sets come from numpy's global RandomState exactly as sklearn would
draw them.
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(_HERE)))

from modest_b200 import _lib, ransac_host  # noqa: E402
from modest_b200 import pipeline as pl  # noqa: E402
from modest_b200.generate_cluster_mask.utils import kitti_util as utils  # noqa: E402

MAX_TRIALS = 100


def road_plane(pc_velo, calib, min_h=1.5, max_h=2, rng="numpy", seed=0):
    """(w (3,), h) of RANSAC.py:38-52 for one scan (pc_velo (N,>=3) f32)."""
    lib = _lib.lib()
    n = int(pc_velo.shape[0])
    p4 = np.zeros((n, 4), dtype=np.float32)
    p4[:, :3] = pc_velo[:, :3]
    ptc = torch.from_numpy(p4).cuda()
    off = torch.tensor([0, n], dtype=torch.int64, device="cuda")
    crow = torch.from_numpy(pl.calib_row(calib)[None].copy()).cuda()
    cand = torch.empty((max(n, 1), 3), dtype=torch.float64, device="cuda")
    n_cand = torch.zeros(1, dtype=torch.int32, device="cuda")
    thr = torch.zeros(1, dtype=torch.float64, device="cuda")
    sp = _lib.stream_ptr()
    _lib.check(lib.modest_road_candidates_batch(_lib.ptr(ptc), 4, _lib.ptr(off), _lib.ptr(crow), 1, float(min_h),
                                                float(max_h), _lib.ptr(cand), _lib.ptr(n_cand), _lib.ptr(thr), sp),
               "modest_road_candidates_batch")
    nc = int(n_cand.cpu()[0])
    triples = None
    if rng == "numpy" and nc >= 5:
        triples = torch.from_numpy(ransac_host.peek_triples(nc, MAX_TRIALS)[None]).cuda()
    plane = torch.empty((1, 4), dtype=torch.float64, device="cuda")
    info = torch.zeros((1, 4), dtype=torch.int32, device="cuda")
    ws = torch.empty(int(lib.modest_ransac_workspace_bytes(1, MAX_TRIALS)), dtype=torch.uint8, device="cuda")
    _lib.check(lib.modest_road_plane_fit_batch(_lib.ptr(cand), _lib.ptr(off), _lib.ptr(n_cand), _lib.ptr(thr), 1, n,
                                               _lib.ptr(triples), C.c_uint64(int(seed)), MAX_TRIALS, _lib.ptr(plane),
                                               _lib.ptr(info), _lib.ptr(ws), ws.numel(), sp),
               "modest_road_plane_fit_batch")
    if rng == "numpy" and nc >= 5:
        ransac_host.consume_trials(nc, int(info.cpu()[0, 1]))
    out = plane.cpu().numpy()[0]
    return out[:3], out[3]


def extract_ransac(calib_dir, lidar_dir, planes_dir, min_h=1.5, max_h=2, split_file=None):
    if split_file is not None:
        with open(split_file) as f:
            data_idx_list = sorted([x.strip() for x in f.readlines() if len(x) > 1])
    else:
        data_idx_list = sorted([x[:-4] for x in os.listdir(lidar_dir) if x[-4:] == '.bin'])
    if not os.path.isdir(planes_dir):
        os.makedirs(planes_dir, exist_ok=True)
    for data_idx in data_idx_list:
        print('------------- ', data_idx)
        calib = utils.Calibration(calib_dir + '/' + data_idx + '.txt')
        pc_velo = np.fromfile(lidar_dir + '/' + data_idx + '.bin', dtype=np.float32).reshape(-1, 4)
        w, h = road_plane(pc_velo, calib, min_h, max_h)
        print(w)
        print(h)
        lines = ['# Plane', 'Width 4', 'Height 1', "{:e} {:e} {:e} {:e}".format(w[0], w[1], w[2], h)]
        with open(os.path.join(planes_dir, data_idx + '.txt'), 'w') as f:
            f.write('\n'.join(lines))


if __name__ == '__main__':
    parser = argparse.ArgumentParser()
    parser.add_argument('--calib_dir', default='KITTI/object/training/calib')
    parser.add_argument('--lidar_dir', default='KITTI/object/training/velodyne')
    parser.add_argument('--planes_dir', default='KITTI/object/training/velodyne_planes')
    parser.add_argument('--min_h', type=float, default=1.5)
    parser.add_argument('--max_h', type=float, default=1.8)
    parser.add_argument('--split_file', type=str, default=None)
    args = parser.parse_args()
    if not os.path.isdir(args.planes_dir):
        extract_ransac(args.calib_dir, args.lidar_dir, args.planes_dir, args.min_h, args.max_h, args.split_file)
