"""ctypes binding of libmodest_b200.so (the C ABI declared in include/modest_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, an exception is
raised.  Build the library with `modest_b200/csrc/build.sh` (or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmodest_b200.so")

OK, ERR_ARG, ERR_CUDA, ERR_CAPACITY = 0, -1, -2, -3


class ModestError(RuntimeError):
    pass


_lib = None

_vp, _i32, _i64, _f32, _f64, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_size_t

# name -> (restype, argtypes); must list every symbol include/modest_b200.h declares
SIGNATURES = {
    "modest_abi_version": (C.c_int, []),
    "modest_last_error": (C.c_char_p, []),
    "modest_launch_count": (_i64, []),
    "modest_upload_frames": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp]),
    "modest_transform_frames_batch": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, _i64, C.c_int, _vp, _vp, _vp]),
    "modest_transform_gather_batch": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _i64, _vp, _vp, _vp]),
    "modest_pp_bin_records": (_i64, [_vp, _vp, C.c_int, _i64]),
    "modest_pp_workspace_bytes": (_sz, [C.c_int, _i64, _i64, C.c_int, _i64]),
    "modest_pp_score_batch": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _i64, _i64, _i64,
                                        _i64, _f64, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64,
                                        _vp, _sz, _vp]),
    "modest_pp_profile_enable": (C.c_int, [C.c_int]),
    "modest_pp_profile_read": (C.c_int, [_vp, C.c_int]),
    "modest_plane_candidates_batch": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _f32, _f32, _f32, _f32, _f32,
                                                _vp, _vp, _vp, _vp]),
    "modest_ransac_workspace_bytes": (_sz, [C.c_int, C.c_int]),
    "modest_ransac_fit_batch": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _i64, _vp, C.c_uint64, _vp, C.c_int,
                                          _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "modest_road_candidates_batch": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, _f64, _f64, _vp, _vp, _vp, _vp]),
    "modest_road_plane_fit_batch": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _i64, _vp, C.c_uint64, C.c_int, _vp, _vp, _vp,
                                              _sz, _vp]),
    "modest_ground_mask_batch": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_int, _f64, _vp, _vp, _vp, _vp,
                                           _vp, _vp, _vp]),
    "modest_graph_workspace_bytes": (_sz, [C.c_int, _i64, C.c_int, C.c_int]),
    "modest_affinity_graph_batch": (C.c_int, [_vp, _vp, _vp, C.c_int, _i64, _i64, C.c_int, _f64, C.c_int,
                                              _vp, _vp, _vp, _f64, _vp, _vp, _vp, _sz, _vp]),
    "modest_fit_rectangle_workspace_bytes": (_sz, [C.c_int, C.c_int]),
    "modest_fit_rectangle": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, C.c_int, _vp, _vp, _sz, _vp]),
    "modest_lowest_point_rect": (C.c_int, [_vp, C.c_int, _f64, _f64, _f64, _f64, _f64, _f64, _vp, _vp]),
    "modest_knn_bruteforce": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _f64, _vp, _vp, _vp, _vp, _vp]),
    "modest_radius_graph": (C.c_int, [_vp, C.c_int, C.c_int, _f64, _vp, _vp, _vp, _vp]),
    "modest_edge_affinity": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "modest_dbscan_workspace_bytes": (_sz, [_i64]),
    "modest_dbscan_batch": (C.c_int, [_vp, _vp, _vp, C.c_int, _i64, _i64, C.c_int, _vp, _vp, _vp, _vp, _f64,
                                      C.c_int, _vp, _vp, _vp, _vp, _sz, _vp]),
    "modest_filter_workspace_bytes": (_sz, [C.c_int, _i64, C.c_int]),
    "modest_filter_and_fit_batch": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, _i64, _i64,
                                              C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp,
                                              _vp, _vp, _vp, _sz, _vp]),
    "modest_box_pp_workspace_bytes": (_sz, [C.c_int, _i64, _i64, C.c_int]),
    "modest_box_pp_percentile_batch": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, _i64, _i64, C.c_int,
                                                 _f64, _vp, _vp, _vp, _sz, _vp]),
    "modest_boxes_iou_bev": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp, _vp]),
    "modest_boxes_overlap_bev": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp, _vp]),
    "modest_nms_workspace_bytes": (_sz, [C.c_int]),
    "modest_nms_bev": (C.c_int, [_vp, C.c_int, _f32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "modest_nms_normal": (C.c_int, [_vp, C.c_int, _f32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "modest_seed_nms_batch": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _f32, _vp, _vp, _vp, _vp]),
    "modest_kitti_labels_batch_host": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp, C.c_int, C.c_int, C.c_int,
                                                 C.c_char_p, _vp, _sz, _vp]),
    "modest_kitti_labels_host": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_char_p,
                                           _vp, _vp, _sz, _vp, _vp, _vp]),
}


def lib():
    """Load (once) and return the ctypes handle; raises ModestError when the .so is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ModestError(
                f"{LIB_PATH} not found: the CUDA library is not built. Run "
                "modest_b200/csrc/build.sh (there is no CPU fallback).")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(code: int, what: str = ""):
    if code != OK:
        msg = lib().modest_last_error().decode("utf-8", "replace")
        raise ModestError(f"{what or 'libmodest_b200'} failed with code {code}: {msg}")


def ptr(t):
    """Device/host pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    import torch
    s = torch.cuda.current_stream() if stream is None else stream
    return C.c_void_p(s.cuda_stream)
