"""PP-score program: drop-in for the reference's generate_cluster_mask/pre_compute_pp_score.py.

Same command line (`python pre_compute_pp_score.py data_root=... [key=value ...]`), same config
keys (configs/pp_score.yaml), same inputs (KITTI-layout data_root + meta pkl/txt) and the same
output (`<pp_score_path>/%06d.npy`, float32 (N,)).  What changed is where the numbers are made:
frames are transformed and the neighbour counts / entropy computed by CUDA kernels
(csrc/pp_score.cu) instead of T cKDTree builds + ball queries per scan.

Reference lines are cited inline (pre_compute_pp_score.py:NN).
"""
import os
import os.path as osp
import pickle
import sys

import numpy as np
import torch
from scipy.spatial.transform import Rotation as R

_HERE = osp.dirname(osp.abspath(__file__))
sys.path.insert(0, osp.dirname(osp.dirname(_HERE)))

from modest_b200 import _lib, dist, hydra_compat  # noqa: E402
from modest_b200 import pp_score as pp_mod  # noqa: E402

hydra_main, DictConfig, OmegaConf = hydra_compat.get_hydra()


def eprint(*args, **kwargs):
    print(*args, file=sys.stderr, **kwargs)


def _quaternion_z_transform(angle):
    """pyquaternion's Quaternion(axis=(0,0,1), angle).transformation_matrix, restated: the
    product of the q and conj(q-bar) 4x4 matrices, rows/cols 1..3 (pre_compute_pp_score.py:22-24)."""
    w, x, y, z = np.cos(angle / 2.0), 0.0, 0.0, np.sin(angle / 2.0)
    qm = np.array([[w, -x, -y, -z], [x, w, -z, y], [y, z, w, -x], [z, -y, x, w]])
    qb = np.array([[w, -x, -y, -z], [x, w, z, -y], [y, -z, w, x], [z, y, -x, w]])
    t = np.eye(4)
    t[:3, :3] = np.dot(qm, qb.conj().transpose())[1:][:, 1:]
    return t


_KITTI2NU_lyft = _quaternion_z_transform(np.pi)
_KITTI2NU_nusc = _quaternion_z_transform(np.pi / 2)


def get_relative_pose(fixed_l2e, fixed_ego, query_l2e, query_ego, KITTI2NU=_KITTI2NU_lyft):
    """pre_compute_pp_score.py:27-28 (4x4 pose bookkeeping on the host)."""
    m = query_ego @ query_l2e @ KITTI2NU
    for lhs in (fixed_ego, fixed_l2e, KITTI2NU):
        m = np.linalg.solve(lhs, m)
    return m.astype(np.float32)


def display_args(args):
    eprint("========== ephemerality info ==========")
    eprint("host: {}".format(os.getenv('HOSTNAME')))
    eprint(OmegaConf.to_yaml(args))
    eprint("=======================================")


class FrameStore:
    """velodyne/*.bin frames, read once and kept where the next scan needs them: consecutive query
    scans share most of their history frames (36 frames x T traversals per scan on
    Lyft), so raw frames stay resident on the GPU in an LRU cache bounded by `device_bytes`."""

    def __init__(self, root, device_bytes=16 << 30):
        self.root, self.device_bytes = root, int(device_bytes)
        self._d, self._used = {}, 0
        self.hits = self.misses = 0

    def get(self, fid):
        t = self._d.pop(fid, None)
        if t is not None:
            self._d[fid] = t                       # most recently used goes last
            self.hits += 1
            return t
        self.misses += 1
        arr = np.fromfile(osp.join(self.root, "velodyne", f"{fid:06d}.bin"), dtype=np.float32).reshape(-1, 4)
        t = torch.from_numpy(arr).cuda(non_blocking=False)
        nbytes = t.numel() * 4
        while self._d and self._used + nbytes > self.device_bytes:
            oldest = next(iter(self._d))            # dicts keep insertion order: first = least recently used
            self._used -= self._d.pop(oldest).numel() * 4
        self._d[fid] = t
        self._used += nbytes
        return t


def transform_frames(frames, mats, remove_center):
    """Upload raw (n,4) frames and bring them into the fixed frame with one kernel launch."""
    sizes = [int(f.shape[0]) for f in frames]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    raw = torch.cat([f if f.is_cuda else f.cuda(non_blocking=True) for f in frames])
    T = torch.from_numpy(np.stack([np.asarray(m, np.float32).reshape(16) for m in mats])).cuda()
    out = torch.empty((raw.shape[0], 3), dtype=torch.float32, device="cuda")
    box = np.array([-1.15, 1.75, -0.65, 0.65], dtype=np.float32)      # pre_compute_pp_score.py:48
    _lib.check(_lib.lib().modest_transform_frames_batch(
        _lib.ptr(raw), 4, _lib.ptr(torch.from_numpy(off).cuda()), _lib.ptr(T), len(frames), max(sizes),
        1 if remove_center else 0, box.ctypes.data, _lib.ptr(out), _lib.stream_ptr()), "modest_transform_frames_batch")
    return out, off


@hydra_main(config_path="configs/", config_name="pp_score.yaml")
def main(args: DictConfig):
    display_args(args)
    dist.init()          # no-op unless launched by torchrun
    if args.ephe_type != "entropy":
        raise NotImplementedError()                                             # :73-74
    track_list = pickle.load(open(args.data_paths.track_path, "rb"))            # :86
    valid_idx = pickle.load(open(args.data_paths.idx_info, "rb"))               # :87
    os.makedirs(args.data_paths.pp_score_path, exist_ok=True)
    oxts_path, l2e_path = osp.join(args.data_root, "oxts"), osp.join(args.data_root, "l2e")
    poses, l2es = [], []
    for seq in track_list:                                                      # :92-106
        poses.append([])
        l2es.append([])
        for idx in seq:
            with open(osp.join(oxts_path, f"{idx:06d}.txt"), "r") as f:
                info = np.array([float(x) for x in f.readline().split()])
            trans = np.eye(4)
            trans[:3, 3] = info[:3]
            trans[:3, :3] = R.from_euler('xyz', info[3:]).as_matrix()
            poses[-1].append(trans.astype(np.float32))
            l2es[-1].append(np.load(osp.join(l2e_path, f"{idx:06d}.npy")))
    if args.data_paths.idx_list is not None:                                    # :109-116
        idx_list = [int(x) for x in open(args.data_paths.idx_list).readlines()]
    else:
        idx_list = [x for x in valid_idx]
    total_part, part = dist.resolve_parts(args.total_part, args.part)
    idx_arr = np.array(idx_list)
    if total_part > 1:
        idx_arr = np.array_split(idx_arr, total_part)[part]
    for key in ("load_save_precomputed_trans_mat", "load_precomputed_lidars"):
        if args.data_paths.get(key) is not None:
            os.makedirs(args.data_paths[key], exist_ok=True)
    k2n = _KITTI2NU_nusc if args.nusc else _KITTI2NU_lyft
    store = FrameStore(args.data_root)
    scorer = pp_mod.PPScorer(radius=float(args.max_neighbor_dist))
    for origin_idx in idx_arr:
        origin_idx = int(origin_idx)
        # NB the reference tests the path without ".npy" (:123), so it never skips; kept as is
        if osp.exists(osp.join(args.data_paths.pp_score_path, f"{origin_idx:06d}")):
            continue
        traversals = valid_idx[origin_idx][2]
        assert len(traversals) > 1, origin_idx                                  # :126
        seq0, frames0 = traversals[0]
        first_pose, first_l2e = poses[seq0][frames0[0]], l2es[seq0][frames0[0]]  # :127-130
        frames, mats, trav_sizes, trav_ids = [], [], [], []
        for seq_id, indices in traversals:                                       # :133-150
            n = 0
            for frame in indices:
                f = store.get(track_list[seq_id][frame])
                frames.append(f)
                mats.append(get_relative_pose(first_l2e, first_pose, l2es[seq_id][frame], poses[seq_id][frame], k2n))
                n += int(f.shape[0])
            trav_sizes.append(n)
            trav_ids.append(seq_id)
        hist, _ = transform_frames(frames, mats, bool(args.nusc))
        if args.data_paths.load_precomputed_lidars is not None:                 # :152-155
            h = hist.cpu().numpy()
            offs = np.concatenate([[0], np.cumsum(trav_sizes)])
            combined = {sid: h[offs[i]:offs[i + 1]][~np.isnan(h[offs[i]:offs[i + 1], 0])]
                        for i, sid in enumerate(trav_ids)}
            pickle.dump(combined, open(osp.join(args.data_paths.load_precomputed_lidars,
                                                f"{origin_idx:06d}.pkl"), "wb"))
        oseq, oframe = valid_idx[origin_idx][0], valid_idx[origin_idx][1]        # :156-167
        origin = store.get(track_list[oseq][oframe])
        trans_mat = get_relative_pose(first_l2e, first_pose, l2es[oseq][oframe], poses[oseq][oframe], k2n)
        if args.data_paths.load_save_precomputed_trans_mat is not None:
            np.save(osp.join(args.data_paths.load_save_precomputed_trans_mat, f"{origin_idx:06d}.npy"), trans_mat)
        if args.skip_ephe:
            continue
        query, _ = transform_frames([origin], [trans_mat], False)               # query is never centre-removed
        if args.add_random_noise > 0:                                            # :176-180
            noise = np.random.randn(3)
            noise /= np.linalg.norm(noise)
            noise *= (args.add_random_noise * np.random.uniform())
            query += torch.from_numpy(noise.reshape(1, 3).astype(np.float32)).cuda()
        if args.limit_traversals > 1:                                            # :182-187
            trav_sizes = trav_sizes[:args.limit_traversals]
        offs = np.concatenate([[0], np.cumsum(trav_sizes)])
        history = [hist[offs[i]:offs[i + 1]] for i in range(len(trav_sizes))]
        batch = pp_mod.pack_batch([query], [history])
        H = scorer(batch)                                                        # :188-194 on the GPU
        np.save(osp.join(args.data_paths.pp_score_path, f"{origin_idx:06d}"), H.cpu().numpy().astype(np.float32))


if __name__ == "__main__":
    main()
