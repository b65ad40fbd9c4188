"""Drop-in for the reference's utils/pointcloud_utils.py: same function names, argument order,
defaults and return types; the numerics run in libmodest_b200 on the GPU (no CPU fallback).

Reference lines are cited per function (paths relative to generate_cluster_mask/).  Functions
the seed-label path never calls with the shipped configs (`minimum_bounding_rectangle`,
`PCA_rectangle`, `variance_rectangle`) raise NotImplementedError, as does estimate_plane(it>1).
"""
import ctypes as C
import types

import numpy as np
import torch

from modest_b200 import _lib
from modest_b200 import pipeline as _pl
from . import kitti_util
from .iou3d_nms import iou3d_nms_utils

_PIPE = None
_BIG = 3.0e38


def _pipe():
    global _PIPE
    if _PIPE is None:
        _PIPE = _pl.SeedLabelPipeline()
    return _PIPE


def _as_batch(ptc_xyz, pp=None):
    ptc_xyz = np.asarray(ptc_xyz)
    n = ptc_xyz.shape[0]
    p4 = np.zeros((n, 4), dtype=np.float32)
    p4[:, :3] = ptc_xyz[:, :3]
    pp = np.zeros(n, dtype=np.float32) if pp is None else np.asarray(pp, dtype=np.float32)
    ident = dict(Tr_velo_to_cam=np.eye(4)[:3], R0_rect=np.eye(3), P2=np.eye(4)[:3])
    return _pl.make_batch([p4], [pp], [ident])


def cart2hom(pts_3d):
    """pointcloud_utils.py:8-11"""
    return np.hstack((pts_3d, np.ones((pts_3d.shape[0], 1), dtype=np.float32)))


def transform_points(pts_3d_ref, Tr):
    """pointcloud_utils.py:14-19 -- (N,3) f32 points through a 4x4 (or 3x4) transform, f32."""
    pts = torch.from_numpy(np.ascontiguousarray(pts_3d_ref[:, :3], dtype=np.float32)).cuda()
    T = np.eye(4, dtype=np.float32)
    T[:np.asarray(Tr).shape[0], :] = np.asarray(Tr, dtype=np.float32)
    out = torch.empty((pts.shape[0], 3), dtype=torch.float32, device="cuda")
    foff = torch.tensor([0, pts.shape[0]], dtype=torch.int64, device="cuda")
    Td = torch.from_numpy(T.reshape(1, 16)).cuda()
    _lib.check(_lib.lib().modest_transform_frames_batch(_lib.ptr(pts), 3, _lib.ptr(foff), _lib.ptr(Td), 1,
                                                        int(pts.shape[0]), 0, None, _lib.ptr(out),
                                                        _lib.stream_ptr()), "modest_transform_frames_batch")
    return out.cpu().numpy()


def load_velo_scan(velo_filename):
    """pointcloud_utils.py:22-25"""
    return np.fromfile(velo_filename, dtype=np.float32).reshape((-1, 4))


def load_plane(plane_filename):
    """pointcloud_utils.py:28-41 (road-plane txt reader; file parsing only)"""
    with open(plane_filename, 'r') as f:
        lines = f.readlines()
    plane = np.asarray([float(i) for i in lines[3].split()])
    if plane[1] > 0:
        plane = -plane
    return plane / np.linalg.norm(plane[0:3])


def estimate_plane(origin_ptc, max_hs=-1.5, it=1, ptc_range=((-20, 70), (-20, 20))):
    """pointcloud_utils.py:44-65 -- RANSAC ground plane [a,b,c,d] (c > 0), consuming numpy's
    global RandomState exactly like sklearn's RANSACRegressor would."""
    if it != 1:
        raise NotImplementedError("estimate_plane(it != 1) is never used by the reference's programs")
    b = _as_batch(origin_ptc)
    plane, _ = _pipe().fit_planes(b, max_hs, ptc_range, rng="numpy")
    return plane.cpu().numpy()[0]


def above_plane(ptc, plane, offset=0.05, only_range=((-30, 30), (-30, 30))):
    """pointcloud_utils.py:68-74 -- True for points kept (not below plane+offset inside only_range)."""
    b = _as_batch(ptc)
    pl = torch.from_numpy(np.asarray(plane, dtype=np.float64).reshape(1, 4).copy()).cuda()
    lim = ((-_BIG, _BIG), (-_BIG, _BIG))
    _, _, _, mask = _pipe().ground_masks(b, pl, offset, only_range, lim, want_mask=True)
    return mask.cpu().numpy().astype(bool)


def distance_to_plane(ptc, plane, directional=False):
    """pointcloud_utils.py:76-81 -- (p.n + d)/|n| in f64 (torch on the GPU)."""
    p = torch.from_numpy(np.ascontiguousarray(ptc[:, :3])).cuda().double()
    pl = torch.from_numpy(np.asarray(plane, dtype=np.float64)).cuda()
    d = p @ pl[:3] + pl[3]
    if not directional:
        d = d.abs()
    return (d / torch.sqrt((pl[:3] ** 2).sum())).cpu().numpy()


_FIT_METHOD_IDS = {'min_zx_area_fit': 0, 'PCA': 1, 'variance_to_edge': 2}


def _fit_rectangle(points, method):
    """(corners (4,2), angle, area) from modest_fit_rectangle; points (n,2) f64."""
    xz = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 2)
    n = xz.shape[0]
    if n < 1:
        raise ValueError("empty cluster")
    lib = _lib.lib()
    pipe = _pipe()
    trig, ang, n_ang = pipe._angle_tables("cuda")
    xz_d = torch.from_numpy(xz).cuda()
    out = torch.zeros(11, dtype=torch.float64, device="cuda")
    need = int(lib.modest_fit_rectangle_workspace_bytes(n, n_ang))
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    _lib.check(lib.modest_fit_rectangle(_lib.ptr(xz_d), n, _FIT_METHOD_IDS[method], _lib.ptr(trig), _lib.ptr(ang), n_ang,
                                        _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "modest_fit_rectangle")
    o = out.cpu().numpy()
    if o[10] != 0.0:
        raise ValueError("rectangle fit failed (status %g)" % o[10])
    return o[:8].reshape(4, 2).copy(), np.float64(o[8]), np.float64(o[9])


def minimum_bounding_rectangle(points):
    """pointcloud_utils.py:88-146 -- smallest bounding rectangle with a side on a convex-hull edge.
    Every hull edge is tried; the reference skips the edge that closes scipy's vertex list (whose
    start is a qhull internal), so its rectangle can be marginally larger in that one case."""
    return _fit_rectangle(points, 'min_zx_area_fit')


def PCA_rectangle(cluster_ptc):
    """pointcloud_utils.py:148-165 -- bounding rectangle along the principal axes (sklearn PCA's
    sign convention); equal to sklearn's to rounding."""
    return _fit_rectangle(cluster_ptc, 'PCA')


def variance_rectangle(cluster_ptc, delta=0.1):
    """pointcloud_utils.py:219-275 -- heading that minimises the variance of the distances to the
    nearer edge, 901 headings, numpy's np.var arithmetic."""
    if delta != 0.1:
        raise NotImplementedError("variance_rectangle: only delta=0.1")
    return _fit_rectangle(cluster_ptc, 'variance_to_edge')


def _fit_single(cluster_rect, full_rect):
    """One cluster, rect coordinates given: runs the fit kernels with every gate open."""
    cluster_rect = np.asarray(cluster_rect, dtype=np.float64)
    full_rect = np.asarray(full_rect, dtype=np.float64)
    n_full, n_cl = full_rect.shape[0], cluster_rect.shape[0]
    rect = np.ascontiguousarray(np.concatenate([full_rect[:, :3], cluster_rect[:, :3]]))
    labels = np.full(n_full + n_cl, -1, dtype=np.int32)
    labels[n_full:] = 0
    b = _as_batch(np.zeros((n_full + n_cl, 3), dtype=np.float32))
    gates = np.array([0, np.inf, -np.inf, 0.2, np.inf, -np.inf, np.inf, _pl.CLOSENESS_D0])
    plane = torch.tensor([[0.0, 0.0, 1.0, 0.0]], dtype=torch.float64, device="cuda")
    _, _, boxes, n_boxes, _, flags = _pipe().filter_and_fit(
        b, torch.from_numpy(labels).cuda(), torch.tensor([1], dtype=torch.int32, device="cuda"), plane,
        rect=torch.from_numpy(rect).cuda(), gates=gates)
    if int(n_boxes.cpu()[0]) != 1:
        raise ValueError("box fit failed (empty footprint?) flags=%d" % int(flags.cpu()[0]))
    return boxes.cpu().numpy()[0, 0]


def closeness_rectangle(cluster_ptc, delta=0.1, d0=1e-2):
    """pointcloud_utils.py:167-216 -- (corners (4,2), angle, area) of the closeness-to-edge
    rectangle of (n,2) points.  delta/d0 other than the defaults are not supported."""
    if delta != 0.1 or d0 != 1e-2:
        raise NotImplementedError("closeness_rectangle: only delta=0.1, d0=1e-2")
    xz = np.asarray(cluster_ptc, dtype=np.float64)
    rect = np.stack([xz[:, 0], np.zeros(len(xz)), xz[:, 1]], axis=1)
    row = _fit_single(rect, rect)
    ry, l, w = row[6], row[3], row[4]
    angle = -ry
    c, s = np.cos(angle), np.sin(angle)
    centre = np.array([row[0], row[2]])
    half = np.array([[l / 2, -w / 2], [-l / 2, -w / 2], [-l / 2, w / 2], [l / 2, w / 2]])
    corners = half @ np.array([[c, s], [-s, c]]) + centre
    return corners, angle, l * w


def get_lowest_point_rect(ptc, xz_center, l, w, ry):
    """pointcloud_utils.py:278-290 -- largest rect-y among the points strictly inside the footprint."""
    rect = np.ascontiguousarray(np.asarray(ptc, dtype=np.float64)[:, :3])
    rect_d = torch.from_numpy(rect).cuda()
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    _lib.check(_lib.lib().modest_lowest_point_rect(_lib.ptr(rect_d), int(rect.shape[0]), float(xz_center[0]), float(xz_center[1]),
                                                   float(np.cos(ry)), float(np.sin(ry)), float(l), float(w), _lib.ptr(out),
                                                   _lib.stream_ptr()), "modest_lowest_point_rect")
    bottom = float(out.cpu()[0])
    if np.isnan(bottom):
        raise ValueError("zero-size array to reduction operation maximum which has no identity")   # ys.max() of nothing
    return np.float64(bottom)


def get_obj(ptc, full_ptc, fit_method='min_zx_area_fit'):
    """pointcloud_utils.py:292-317 -- box namespace (t, l, w, h, ry, volume) for one cluster given
    in rect coordinates.  'closeness_to_edge' (the configured method) runs the fused seed-label
    kernels; the other three fitters run the operator-level kernels of csrc/fitters.cu and the
    scalar assembly of :303-316."""
    if fit_method == 'closeness_to_edge':
        return box_namespace(_fit_single(ptc, full_ptc))
    if fit_method not in _FIT_METHOD_IDS:
        raise NotImplementedError(fit_method)
    ptc = np.asarray(ptc, dtype=np.float64)
    corners, ry, area = _fit_rectangle(ptc[:, [0, 2]], fit_method)
    ry = ry * -1
    l = np.linalg.norm(corners[0] - corners[1])
    w = np.linalg.norm(corners[0] - corners[-1])
    c = (corners[0] + corners[2]) / 2
    bottom = get_lowest_point_rect(full_ptc, c, l, w, ry)
    h = bottom - ptc[:, 1].min()
    obj = types.SimpleNamespace()
    obj.t = np.array([c[0], bottom, c[1]])
    obj.l, obj.w, obj.h, obj.ry, obj.volume = l, w, h, ry, area * h
    return obj


def box_namespace(row):
    obj = types.SimpleNamespace()
    obj.t = np.array([row[0], row[1], row[2]])
    obj.l, obj.w, obj.h, obj.ry, obj.volume = (np.float64(row[3]), np.float64(row[4]), np.float64(row[5]),
                                               np.float64(row[6]), np.float64(row[7]))
    return obj


def objs_to_rows(objs):
    return np.array([[o.t[0], o.t[1], o.t[2], o.l, o.w, o.h, o.ry, getattr(o, "volume", 0.0)] for o in objs],
                    dtype=np.float64).reshape(-1, 8)


def objs_nms(objs, use_score_rank=False, nms_threshold=0.1):
    """pointcloud_utils.py:320-344 -- BEV IoU on the GPU, then the reference's own ordering and
    greedy sweep on the K x K matrix (K is tens of boxes)."""
    boxes = np.array([[obj.t[0], obj.t[2], 0, obj.l, obj.w, obj.h, -obj.ry] for obj in objs])
    boxes = torch.from_numpy(boxes).float().cuda()
    overlaps_bev = iou3d_nms_utils.boxes_iou_bev(boxes.contiguous(), boxes.contiguous()).cpu().numpy()
    mask = np.ones(overlaps_bev.shape[0], dtype=bool)
    if use_score_rank:
        order = np.argsort([obj.score for obj in objs])[::-1]
    else:
        order = np.diag(overlaps_bev).argsort()[::-1]
    for idx in order:
        if not mask[idx]:
            continue
        mask[overlaps_bev[idx] > nms_threshold] = False
        mask[idx] = True
    return [objs[i] for i in range(len(objs)) if mask[i]]


def _labels_call(objs, calib, fov_only, image_shape, obj_type, with_score):
    rows = objs_to_rows(objs)
    n = rows.shape[0]
    scores = None
    if with_score:
        scores = np.array([getattr(o, "score", -1) for o in objs], dtype=np.float64)
    P = np.ascontiguousarray(calib.P, dtype=np.float64)
    cap = 256 * max(n, 1) + 16
    buf = C.create_string_buffer(cap)
    ln, cnt = C.c_size_t(0), C.c_int(0)
    kept = np.zeros(max(n, 1), dtype=np.uint8)
    _lib.check(_lib.lib().modest_kitti_labels_host(
        rows.ctypes.data_as(C.c_void_p), n, None, P.ctypes.data_as(C.c_void_p), 1 if fov_only else 0,
        int(image_shape[0]), int(image_shape[1]), obj_type.encode(),
        None if scores is None else scores.ctypes.data_as(C.c_void_p), buf, cap, C.byref(ln), C.byref(cnt),
        kept.ctypes.data_as(C.c_void_p)), "modest_kitti_labels_host")
    return buf.raw[:ln.value].decode(), kept[:n].astype(bool)


def objs2label(objs, calib, obj_type="Dynamic", with_score=False):
    """pointcloud_utils.py:347-370 -- KITTI label lines, '\\n'-joined, no trailing newline."""
    return _labels_call(objs, calib, False, (0, 0), obj_type, with_score)[0]


def is_within_fov(obj, calib, image_shape):
    """pointcloud_utils.py:373-379"""
    return bool(_labels_call([obj], calib, True, image_shape, "Dynamic", False)[1][0])
