"""Synthetic Lyft-/nuScenes-shaped LiDAR data for parity tests and benchmarks.

Nothing here is part of the hot path: it manufactures inputs in the on-disk layout the
reference CLIs read:

  <data_root>/velodyne/%06d.bin   (N,4) f32 [x,y,z,intensity], KITTI lidar frame
  <data_root>/oxts/%06d.txt       "tx ty tz ex ey ez"  (ego pose, euler 'xyz')
  <data_root>/l2e/%06d.npy        4x4 f32 lidar->ego
  <data_root>/calib/%06d.txt      P0..P3, R0_rect, Tr_velo_to_cam, Tr_imu_to_velo ("%.12e")
  <meta>/track_list.pkl           list[list[int]]  frame ids per traversal
  <meta>/valid_idx_info.pkl       {idx: (seq, frame_pos, [(seq_id, [frame_pos...]), ...])}
  <meta>/train_idx.txt            "%06d" per line

The formats follow /root/reference/data_preprocessing/lyft/lyft2kitti.py:261-272,373-393 and
split_traintest.py:110-118; the reader side is pre_compute_pp_score.py:86-106.

Scene model: a fixed static world (tilted ground plane, wall segments, poles) observed from
several traversals of a straight road.  Every frame re-samples the static world with
independent noise and adds its own "dynamic" car-sized boxes, so points on those boxes have no
neighbours in other traversals (low PP score) while the static world is persistent.
Points are emitted beam-major / azimuth-minor like a spinning LiDAR's .bin file.
"""
from __future__ import annotations

import math
import os
import pickle
from dataclasses import dataclass, field

import numpy as np
from scipy.spatial.transform import Rotation

SEED_BASE = 1024  # the reference's (unused) pp_score.yaml `seed`


@dataclass
class Shape:
    name: str = "lyft"
    n_points: int = 60000
    n_beams: int = 64
    sensor_height: float = 1.8
    image_shape: tuple = (1024, 1224)
    max_hs: float = -1.5
    nusc: bool = False


LYFT = Shape()
NUSC = Shape(name="nusc", n_points=34000, n_beams=32, sensor_height=1.6,
             image_shape=(900, 1600), max_hs=-1.3, nusc=True)


def _rz(a):
    c, s = np.cos(a), np.sin(a)
    m = np.eye(4)
    m[0, 0], m[0, 1], m[1, 0], m[1, 1] = c, -s, s, c
    return m


def kitti2nu(nusc: bool) -> np.ndarray:
    """z-rotation the reference applies between KITTI and native lidar axes
    (pre_compute_pp_score.py:22-24: pi for Lyft, pi/2 for nuScenes)."""
    return _rz(np.pi / 2 if nusc else np.pi)


@dataclass
class World:
    """Static scene in world coordinates (ground z ~ 0, road along +X)."""
    tilt: np.ndarray            # (2,) ground slope dz/dx, dz/dy
    walls: np.ndarray           # (W,6) x0,y0,x1,y1,zlo,zhi
    poles: np.ndarray           # (P,4) x,y,radius,height
    rng_seed: int = 0

    def ground_z(self, x, y):
        return self.tilt[0] * x + self.tilt[1] * y


def make_world(seed: int, n_walls: int = 60, n_poles: int = 40, x_range=(-75.0, 95.0)) -> World:
    rng = np.random.default_rng(seed)
    tilt = np.tan(np.deg2rad(rng.uniform(-0.8, 0.8, size=2)))
    walls = np.zeros((n_walls, 6))
    for i in range(n_walls):
        cx = rng.uniform(x_range[0], x_range[1])
        side = rng.choice([-1.0, 1.0])
        cy = side * rng.uniform(7.0, 38.0)
        yaw = rng.uniform(0, np.pi)
        half = rng.uniform(2.5, 10.0)
        dx, dy = np.cos(yaw) * half, np.sin(yaw) * half
        y0, y1 = cy - dy, cy + dy
        # keep the road corridor free
        if side > 0:
            y0, y1 = max(y0, 6.0), max(y1, 6.0)
        else:
            y0, y1 = min(y0, -6.0), min(y1, -6.0)
        walls[i] = (cx - dx, y0, cx + dx, y1, 0.0, rng.uniform(2.0, 5.0))
    poles = np.zeros((n_poles, 4))
    poles[:, 0] = rng.uniform(x_range[0], x_range[1], n_poles)
    poles[:, 1] = rng.choice([-1.0, 1.0], n_poles) * rng.uniform(5.5, 30.0, n_poles)
    poles[:, 2] = rng.uniform(0.08, 0.2, n_poles)
    poles[:, 3] = rng.uniform(3.0, 6.0, n_poles)
    return World(tilt=tilt, walls=walls, poles=poles, rng_seed=seed)


@dataclass
class Pose:
    ego: np.ndarray   # 4x4 f64 ego->world
    l2e: np.ndarray   # 4x4 f64 lidar->ego
    oxts: np.ndarray  # (6,) tx ty tz ex ey ez (what the oxts txt holds)

    def lidar_to_world(self, nusc: bool) -> np.ndarray:
        return self.ego @ self.l2e @ kitti2nu(nusc)


def make_pose(x: float, y: float, yaw: float, shape: Shape, world: World) -> Pose:
    """Vehicle at world (x,y) heading `yaw`; lidar mounted so that the KITTI-frame x axis
    points along the heading (l2e carries the inverse of the KITTI->native z-rotation)."""
    tz = world.ground_z(x, y)
    eul = np.array([0.0, 0.0, yaw])
    ego = np.eye(4)
    ego[:3, :3] = Rotation.from_euler("xyz", eul).as_matrix()
    ego[:3, 3] = (x, y, tz)
    l2e = np.linalg.inv(kitti2nu(shape.nusc))
    l2e[:3, 3] = (0.9, 0.0, shape.sensor_height)
    return Pose(ego=ego, l2e=l2e, oxts=np.array([x, y, tz, *eul]))


def _sample_dynamic_boxes(rng, n_boxes):
    """Car-sized boxes in the SENSOR frame: (cx, cy, l, w, h, yaw). Never axis aligned."""
    out = np.zeros((n_boxes, 6))
    centres = []
    k = 0
    tries = 0
    while k < n_boxes and tries < 10000:
        tries += 1
        cx = rng.uniform(6.0, 60.0)
        cy = rng.uniform(-17.0, 17.0)
        if abs(cy) > 0.55 * cx + 2.0:      # stay roughly inside the camera frustum
            continue
        l, w, h = rng.uniform(3.5, 5.0), rng.uniform(1.6, 2.1), rng.uniform(1.4, 1.9)
        yaw = rng.uniform(0.12, np.pi / 2 - 0.12) + rng.integers(0, 2) * np.pi / 2
        # (plain-python distance test: same decisions as np.min(np.hypot(...)) < 7.5 without ~10k
        #  numpy calls per frame -- the rejection loop usually runs to its try limit)
        if any(math.hypot(px - cx, py - cy) < 7.5 for px, py in centres):
            continue
        out[k] = (cx, cy, l, w, h, yaw)
        centres.append((cx, cy))
        k += 1
    return out[:k]


def _box_surface_points(rng, box, n, sensor_xy=(0.0, 0.0)):
    """n points on the two vertical faces of `box` that face the sensor, plus a sparse roof."""
    cx, cy, l, w, h, yaw = box
    c, s = np.cos(yaw), np.sin(yaw)
    # sensor position in box frame
    dx, dy = sensor_xy[0] - cx, sensor_xy[1] - cy
    sx, sy = c * dx + s * dy, -s * dx + c * dy
    face_x = np.sign(sx) if sx != 0 else 1.0
    face_y = np.sign(sy) if sy != 0 else 1.0
    n_roof = n // 8
    n_long = int((n - n_roof) * l / (l + w))
    n_short = n - n_roof - n_long
    u = np.concatenate([
        np.stack([rng.uniform(-l / 2, l / 2, n_long), np.full(n_long, face_y * w / 2)], 1),
        np.stack([np.full(n_short, face_x * l / 2), rng.uniform(-w / 2, w / 2, n_short)], 1),
        np.stack([rng.uniform(-l / 2, l / 2, n_roof), rng.uniform(-w / 2, w / 2, n_roof)], 1)])
    z = np.concatenate([rng.uniform(0.15, h, n_long + n_short), np.full(n_roof, h)])
    u += rng.normal(0, 0.01, u.shape)
    x = cx + c * u[:, 0] - s * u[:, 1]
    y = cy + s * u[:, 0] + c * u[:, 1]
    return np.stack([x, y, z], 1)   # z relative to local ground


def sample_frame(world: World, pose: Pose, shape: Shape, rng, n_dynamic=(10, 30),
                 n_points: int | None = None, return_boxes: bool = False):
    """One LiDAR frame in the KITTI lidar frame of `pose`: (N,4) f32, beam-major order."""
    N = int(n_points or shape.n_points)
    L2W = pose.lidar_to_world(shape.nusc)
    W2L = np.linalg.inv(L2W)
    n_dyn_boxes = int(rng.integers(n_dynamic[0], n_dynamic[1] + 1)) if n_dynamic[1] > 0 else 0
    boxes = _sample_dynamic_boxes(rng, n_dyn_boxes)
    n_dyn = int(0.15 * N) if len(boxes) else 0
    n_static = int(0.30 * N)
    n_ground = N - n_dyn - n_static

    # ---- ground: radial density ~ 1/r, azimuth uniform, in the sensor's xy then lifted to world
    r = rng.uniform(2.0, 70.0, n_ground)
    az = rng.uniform(-np.pi, np.pi, n_ground)
    gl = np.stack([r * np.cos(az), r * np.sin(az), np.full(n_ground, -shape.sensor_height),
                   np.ones(n_ground)], 1)
    gw = gl @ L2W.T
    gw[:, 2] = world.ground_z(gw[:, 0], gw[:, 1]) + rng.normal(0, 0.02, n_ground)

    # ---- static structure, weights ~ visible area / distance
    sx, sy = L2W[0, 3], L2W[1, 3]
    wl = world.walls
    wmid = 0.5 * (wl[:, 0:2] + wl[:, 2:4])
    wlen = np.hypot(wl[:, 2] - wl[:, 0], wl[:, 3] - wl[:, 1])
    wdist = np.hypot(wmid[:, 0] - sx, wmid[:, 1] - sy)
    wweight = np.where(wdist < 72.0, wlen * wl[:, 5] / np.maximum(wdist, 3.0), 0.0)
    pl = world.poles
    pdist = np.hypot(pl[:, 0] - sx, pl[:, 1] - sy)
    pweight = np.where(pdist < 72.0, 1.5 * pl[:, 3] / np.maximum(pdist, 3.0), 0.0)
    weights = np.concatenate([wweight, pweight])
    if weights.sum() <= 0:
        weights[:] = 1.0
    counts = rng.multinomial(n_static, weights / weights.sum())
    chunks = []
    for i, c in enumerate(counts[:len(wl)]):
        if c == 0:
            continue
        t = rng.uniform(0, 1, c)
        x = wl[i, 0] + t * (wl[i, 2] - wl[i, 0])
        y = wl[i, 1] + t * (wl[i, 3] - wl[i, 1])
        nrm = np.array([-(wl[i, 3] - wl[i, 1]), wl[i, 2] - wl[i, 0]]) / max(wlen[i], 1e-9)
        off = rng.normal(0, 0.015, c)
        z = world.ground_z(x, y) + rng.uniform(0.05, wl[i, 5], c)
        chunks.append(np.stack([x + off * nrm[0], y + off * nrm[1], z], 1))
    for i, c in enumerate(counts[len(wl):]):
        if c == 0:
            continue
        a = rng.uniform(-np.pi, np.pi, c)
        x = pl[i, 0] + pl[i, 2] * np.cos(a)
        y = pl[i, 1] + pl[i, 2] * np.sin(a)
        z = world.ground_z(x, y) + rng.uniform(0.05, pl[i, 3], c)
        chunks.append(np.stack([x, y, z], 1))
    sw = np.concatenate(chunks) if chunks else np.zeros((0, 3))

    # ---- dynamic boxes (sensor frame -> world)
    dw = np.zeros((0, 3))
    if n_dyn:
        area = boxes[:, 2] * boxes[:, 4] / np.maximum(np.hypot(boxes[:, 0], boxes[:, 1]), 4.0)
        per = np.maximum(150, (n_dyn * area / area.sum()).astype(int))
        per[-1] = max(150, n_dyn - per[:-1].sum())
        dchunks = []
        for b, c in zip(boxes, per):
            p = _box_surface_points(rng, b, int(c))
            pl4 = np.stack([p[:, 0], p[:, 1], np.full(len(p), -shape.sensor_height),
                            np.ones(len(p))], 1)
            pw = pl4 @ L2W.T
            pw[:, 2] = world.ground_z(pw[:, 0], pw[:, 1]) + p[:, 2]
            dchunks.append(pw[:, :3])
        dw = np.concatenate(dchunks)

    pts_w = np.concatenate([gw[:, :3], sw, dw])
    pts_l = (np.concatenate([pts_w, np.ones((len(pts_w), 1))], 1) @ W2L.T)[:, :3]
    # trim / pad to exactly N (the dynamic share can overshoot by the 150-point floor)
    if len(pts_l) > N:
        if len(pts_l) - N <= n_ground:
            keep = np.ones(len(pts_l), bool)
            drop = rng.choice(n_ground, len(pts_l) - N, replace=False)
            keep[drop] = False
            pts_l = pts_l[keep]
        else:                       # tiny clouds: the 150-point floor per box exceeds the budget
            pts_l = pts_l[rng.permutation(len(pts_l))[:N]]
    elif len(pts_l) < N:
        extra = N - len(pts_l)
        pts_l = np.concatenate([pts_l, pts_l[:extra] + rng.normal(0, 0.05, (extra, 3))])

    # ---- beam-major / azimuth-minor ordering
    rng_xy = np.hypot(pts_l[:, 0], pts_l[:, 1])
    elev = np.arctan2(pts_l[:, 2], rng_xy)
    lo, hi = np.deg2rad(-42.0), np.deg2rad(20.0)
    beam = np.clip(((elev - lo) / (hi - lo) * shape.n_beams).astype(np.int64), 0,
                   shape.n_beams - 1)
    azim = np.arctan2(pts_l[:, 1], pts_l[:, 0])
    order = np.lexsort((azim, beam))
    pts_l = pts_l[order]
    out = np.empty((N, 4), np.float32)
    out[:, :3] = pts_l.astype(np.float32)
    out[:, 3] = rng.uniform(0, 1, N).astype(np.float32)
    if return_boxes:
        return out, boxes
    return out


def default_calib(shape: Shape) -> dict:
    """KITTI calibration for a forward pinhole camera mounted at the lidar origin."""
    h, w = shape.image_shape
    f = 0.72 * w
    P2 = np.array([[f, 0, w / 2.0, 0], [0, f, h / 2.0, 0], [0, 0, 1, 0]], np.float64)
    v2c = np.array([[0, -1, 0, 0.0], [0, 0, -1, -0.3], [1, 0, 0, -0.1]], np.float64)
    return {"P0": np.zeros((3, 4)), "P1": np.zeros((3, 4)), "P2": P2, "P3": np.zeros((3, 4)),
            "R0_rect": np.eye(3), "Tr_velo_to_cam": v2c, "Tr_imu_to_velo": np.eye(4)[:3]}


def write_calib(path: str, calib: dict):
    with open(path, "w") as f:
        for key, val in calib.items():
            f.write("%s: %s\n" % (key, " ".join("%.12e" % v for v in np.asarray(val).ravel())))


# --------------------------------------------------------------------------------------------
# On-disk dataset (config 1 and the CLI tests)
# --------------------------------------------------------------------------------------------
def write_dataset(root: str, meta_dir: str, shape: Shape = LYFT, n_traversals: int = 3,
                  frames_per_traversal: int = 3, history_frames: int = 1, n_points=None,
                  seed: int = SEED_BASE, prefix: str = "") -> dict:
    """Write a KITTI-layout data_root with `n_traversals` passes over one location.

    Scan ids are traversal-major.  Every frame is a valid query whose history is, for each
    traversal (its own first, as split_traintest.py does), the `history_frames` frames closest
    to it.  Returns {"idx": [...], "track_list": ..., "valid_idx": ...}.
    """
    for d in ("velodyne", "oxts", "l2e", "calib"):
        os.makedirs(os.path.join(root, d), exist_ok=True)
    os.makedirs(meta_dir, exist_ok=True)
    world = make_world(seed)
    calib = default_calib(shape)
    track_list, poses = [], []
    fid = 0
    for t in range(n_traversals):
        rng = np.random.default_rng(seed + 7919 * (t + 1))
        lateral = rng.uniform(-0.6, 0.6)
        start = rng.uniform(-1.0, 1.0)
        track, tposes = [], []
        for k in range(frames_per_traversal):
            x = start + 2.0 * k
            yaw = np.deg2rad(rng.uniform(-2.0, 2.0))
            pose = make_pose(x, lateral, yaw, shape, world)
            frng = np.random.default_rng(SEED_BASE + fid)
            pts = sample_frame(world, pose, shape, frng, n_points=n_points)
            pts.tofile(os.path.join(root, "velodyne", f"{fid:06d}.bin"))
            with open(os.path.join(root, "oxts", f"{fid:06d}.txt"), "w") as f:
                f.write(" ".join(str(v) for v in pose.oxts))
            np.save(os.path.join(root, "l2e", f"{fid:06d}.npy"), pose.l2e.astype(np.float32))
            write_calib(os.path.join(root, "calib", f"{fid:06d}.txt"), calib)
            track.append(fid)
            tposes.append(pose)
            fid += 1
        track_list.append(track)
        poses.append(tposes)
    valid_idx = {}
    for t in range(n_traversals):
        for k in range(frames_per_traversal):
            x = poses[t][k].oxts[0]
            hist = []
            for s in [t] + [u for u in range(n_traversals) if u != t]:
                xs = np.array([p.oxts[0] for p in poses[s]])
                near = np.argsort(np.abs(xs - x), kind="stable")[:history_frames]
                if s == t:
                    # the query's own traversal contributes frames other than itself when it can
                    cand = [j for j in np.argsort(np.abs(xs - x), kind="stable") if j != k]
                    near = np.array(cand[:history_frames]) if cand else near
                hist.append((s, sorted(int(j) for j in near)))
            valid_idx[track_list[t][k]] = (t, k, hist)
    with open(os.path.join(meta_dir, f"{prefix}track_list.pkl"), "wb") as f:
        pickle.dump(track_list, f)
    with open(os.path.join(meta_dir, f"{prefix}valid_idx_info.pkl"), "wb") as f:
        pickle.dump(valid_idx, f)
    with open(os.path.join(meta_dir, f"{prefix}train_idx.txt"), "w") as f:
        f.write("\n".join(f"{x:06d}" for x in valid_idx.keys()))
    return {"idx": list(valid_idx.keys()), "track_list": track_list, "valid_idx": valid_idx}


# --------------------------------------------------------------------------------------------
# In-memory track dataset (streaming engine, benchmarks): the same structure write_dataset puts
# on disk -- several traversals of one road, every frame a valid query whose history is the
# `history_frames` nearest frames of every traversal -- so that consecutive query scans share
# their history frames the way real drives do.
# --------------------------------------------------------------------------------------------
@dataclass
class TrackDataset:
    shape: Shape
    frames: dict                     # frame id -> (N,4) f32, the velodyne/%06d.bin content
    poses: dict                      # frame id -> Pose
    track_list: list                 # [[frame ids of traversal 0], ...]
    valid_idx: dict                  # scan id -> (seq, pos, [(seq_id, [pos, ...]), ...])   (split_traintest.py:110-113)
    calib: dict

    @property
    def scan_ids(self):
        return list(self.valid_idx.keys())

    def fixed_frame(self, scan_id):
        """First frame of the first listed traversal (pre_compute_pp_score.py:127-130)."""
        seq_id, positions = self.valid_idx[scan_id][2][0]
        return self.track_list[seq_id][positions[0]]

    def history_frames(self, scan_id):
        """[[frame ids of traversal 0's history], ...] in the order the reference concatenates them."""
        return [[self.track_list[seq_id][p] for p in positions] for seq_id, positions in self.valid_idx[scan_id][2]]

    def relative_pose(self, scan_id, frame_id):
        return relative_pose_f32(self.poses[self.fixed_frame(scan_id)], self.poses[frame_id], self.shape.nusc)

    def relative_poses_batch(self, scan_ids, frame_ids):
        """(n,4,4) f32: pose of frame_ids[i] in the fixed frame of scan_ids[i], the bits of
        relative_pose() (numpy's stacked solve runs LAPACK gesv per matrix, like the single calls)."""
        chain = self.__dict__.setdefault("_chain", {})
        fixed = self.__dict__.setdefault("_fixed", {})
        k = kitti2nu(self.shape.nusc)
        f64 = lambda a: a.astype(np.float32).astype(np.float64)
        for f in frame_ids:
            if f not in chain:
                chain[f] = f64(self.poses[f].ego) @ f64(self.poses[f].l2e) @ k
        for sid in scan_ids:
            if sid not in fixed:
                p = self.poses[self.fixed_frame(sid)]
                fixed[sid] = (f64(p.ego), f64(p.l2e))
        m = np.stack([chain[f] for f in frame_ids])
        m = np.linalg.solve(np.stack([fixed[sid][0] for sid in scan_ids]), m)
        m = np.linalg.solve(np.stack([fixed[sid][1] for sid in scan_ids]), m)
        m = np.linalg.solve(k[None], m)
        return m.astype(np.float32)

    def relative_poses(self, scan_id, frame_ids):
        """(n,4,4) f32, the bits of relative_pose(): the per-frame product ego @ l2e @ KITTI2NU is kept."""
        cache = self.__dict__.setdefault("_chain", {})
        k = kitti2nu(self.shape.nusc)
        f64 = lambda a: a.astype(np.float32).astype(np.float64)
        for f in frame_ids:
            if f not in cache:
                cache[f] = f64(self.poses[f].ego) @ f64(self.poses[f].l2e) @ k
        fixed = self.poses[self.fixed_frame(scan_id)]
        m = np.stack([cache[f] for f in frame_ids])
        for lhs in (f64(fixed.ego), f64(fixed.l2e), k):
            m = np.linalg.solve(lhs[None], m)
        return m.astype(np.float32)


def _track_frame(args):
    world, pose, shape, fid, n_points = args
    return fid, sample_frame(world, pose, shape, np.random.default_rng(SEED_BASE + fid), n_points=n_points)


def make_track_dataset(shape: Shape = LYFT, n_traversals: int = 16, frames_per_traversal: int = 30,
                       history_frames: int = 1, n_points=None, seed: int = SEED_BASE, first_frame_id: int = 0,
                       workers: int = 0) -> TrackDataset:
    """`n_traversals` passes over one straight road, a frame every 2 m.  Frame ids are traversal-major
    starting at `first_frame_id`.  `workers` > 1 samples the frames in a process pool."""
    span = 2.0 * frames_per_traversal
    world = make_world(seed, n_walls=int(60 * (170.0 + span) / 170.0), n_poles=int(40 * (170.0 + span) / 170.0),
                       x_range=(-75.0, 95.0 + span))
    track_list, poses, todo = [], {}, []
    fid = first_frame_id
    tposes = []
    for t in range(n_traversals):
        rng = np.random.default_rng(seed + 7919 * (t + 1))
        lateral, start = rng.uniform(-0.6, 0.6), rng.uniform(-1.0, 1.0)
        track, tp = [], []
        for k in range(frames_per_traversal):
            pose = make_pose(start + 2.0 * k, lateral, np.deg2rad(rng.uniform(-2.0, 2.0)), shape, world)
            poses[fid] = pose
            todo.append((world, pose, shape, fid, n_points))
            track.append(fid)
            tp.append(pose)
            fid += 1
        track_list.append(track)
        tposes.append(tp)
    if workers and workers > 1:
        import multiprocessing as mp
        with mp.get_context("spawn").Pool(workers) as pool:
            frames = dict(pool.map(_track_frame, todo, chunksize=max(1, len(todo) // (4 * workers))))
    else:
        frames = dict(_track_frame(a) for a in todo)
    valid_idx = {}
    for t in range(n_traversals):
        for k in range(frames_per_traversal):
            x = tposes[t][k].oxts[0]
            hist = []
            for u in [t] + [v for v in range(n_traversals) if v != t]:
                xs = np.array([p.oxts[0] for p in tposes[u]])
                order = [j for j in np.argsort(np.abs(xs - x), kind="stable") if not (u == t and j == k)]
                hist.append((u, sorted(int(j) for j in order[:history_frames])))
            valid_idx[track_list[t][k]] = (t, k, hist)
    return TrackDataset(shape=shape, frames=frames, poses=poses, track_list=track_list, valid_idx=valid_idx,
                        calib=default_calib(shape))


def scan_case_from_dataset(ds: TrackDataset, scan_id: int) -> "ScanCase":
    """The arrays pre_compute_pp_score.py holds for one query scan of the dataset at :188: query and
    per-traversal history in the fixed frame (host arithmetic of the reference: transform_points)."""
    hist = []
    for fids in ds.history_frames(scan_id):
        parts = []
        for f in fids:
            pts = ds.frames[f][:, :3]
            if ds.shape.nusc:
                m = (pts[:, 0] < 1.75) & (pts[:, 0] >= -1.15) & (pts[:, 1] < 0.65) & (pts[:, 1] >= -0.65)
                pts = pts[~m]
            parts.append(transform_points_f32(pts, ds.relative_pose(scan_id, f)))
        hist.append(np.ascontiguousarray(np.concatenate(parts).astype(np.float32)))
    q = ds.frames[scan_id]
    qf = transform_points_f32(q[:, :3], ds.relative_pose(scan_id, scan_id))
    return ScanCase(scan_id=scan_id, query=q, query_fixed=np.ascontiguousarray(qf), history=hist, calib=ds.calib)


# --------------------------------------------------------------------------------------------
# In-memory scans (benchmarks, kernel parity tests): query + T history clouds, all already in
# the fixed frame, i.e. what pre_compute_pp_score.py holds at :188 just before the tree build.
# --------------------------------------------------------------------------------------------
@dataclass
class ScanCase:
    scan_id: int
    query: np.ndarray                 # (N,4) f32 in its own KITTI lidar frame (generate_mask input)
    query_fixed: np.ndarray           # (N,3) f32 query xyz in the fixed frame (PP input)
    history: list = field(default_factory=list)   # T arrays (M_t,3) f32 in the fixed frame
    calib: dict = field(default_factory=dict)


def transform_points_f32(pts_xyz: np.ndarray, tr: np.ndarray) -> np.ndarray:
    """[p,1] @ Tr^T in float32, the arithmetic of the reference's transform_points
    (utils/pointcloud_utils.py:11-19)."""
    hom = np.hstack((pts_xyz.astype(np.float32), np.ones((len(pts_xyz), 1), np.float32)))
    return (hom @ tr.astype(np.float32).T)[:, :3]


def relative_pose_f32(fixed: Pose, query: Pose, nusc: bool) -> np.ndarray:
    """fixed-frame <- query-frame 4x4, f64 solves then f32 cast
    (restates pre_compute_pp_score.py:27-28)."""
    k = kitti2nu(nusc)
    m = query.ego.astype(np.float32).astype(np.float64) @ \
        query.l2e.astype(np.float32).astype(np.float64) @ k
    m = np.linalg.solve(fixed.ego.astype(np.float32).astype(np.float64), m)
    m = np.linalg.solve(fixed.l2e.astype(np.float32).astype(np.float64), m)
    return np.linalg.solve(k, m).astype(np.float32)


def relative_poses_f32(fixed: Pose, queries: list, nusc: bool) -> np.ndarray:
    """relative_pose_f32 for many query frames at once: (n,4,4) f32.  numpy's stacked solve runs the
    same LAPACK gesv per matrix, so every result has the bits of the one-at-a-time call."""
    k = kitti2nu(nusc)
    f64 = lambda a: a.astype(np.float32).astype(np.float64)
    m = np.stack([f64(q.ego) @ f64(q.l2e) @ k for q in queries])
    for lhs in (f64(fixed.ego), f64(fixed.l2e), k):
        m = np.linalg.solve(lhs[None], m)
    return m.astype(np.float32)


def make_scan_case(scan_id: int, shape: Shape = LYFT, n_traversals: int = 16,
                   frames_per_traversal: int = 1, n_points=None, world_seed=None) -> ScanCase:
    rng = np.random.default_rng(SEED_BASE + scan_id)
    world = make_world(SEED_BASE + scan_id if world_seed is None else world_seed)
    qpose = make_pose(rng.uniform(-1, 1), rng.uniform(-0.5, 0.5),
                      np.deg2rad(rng.uniform(-2, 2)), shape, world)
    query = sample_frame(world, qpose, shape, rng, n_points=n_points)
    history = []
    fixed = None
    for t in range(n_traversals):
        frames = []
        x0, lat = rng.uniform(-1.5, 1.5), rng.uniform(-0.6, 0.6)
        for k in range(frames_per_traversal):
            pose = make_pose(x0 + 2.0 * (k - frames_per_traversal // 2), lat,
                             np.deg2rad(rng.uniform(-2, 2)), shape, world)
            if fixed is None:
                fixed = pose
            pts = sample_frame(world, pose, shape, rng, n_points=n_points)[:, :3]
            if shape.nusc:   # pre_compute_pp_score.py:48-52,141-142: history frames only
                m = (pts[:, 0] < 1.75) & (pts[:, 0] >= -1.15) & (pts[:, 1] < 0.65) & \
                    (pts[:, 1] >= -0.65)
                pts = pts[~m]
            frames.append(transform_points_f32(pts, relative_pose_f32(fixed, pose, shape.nusc)))
        history.append(np.ascontiguousarray(np.concatenate(frames).astype(np.float32)))
    qf = transform_points_f32(query[:, :3], relative_pose_f32(fixed, qpose, shape.nusc))
    return ScanCase(scan_id=scan_id, query=query, query_fixed=np.ascontiguousarray(qf),
                    history=history, calib=default_calib(shape))
