"""Stand-in for the reference's pybind11 extension module `iou3d_nms_cuda`
(generate_cluster_mask/utils/iou3d_nms/src/iou3d_nms_api.cpp:11-17): the same five names with
the same argument order and return values, implemented on libmodest_b200's C ABI.

    boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap) -> 1
    boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou)         -> 1
    nms_gpu(boxes, keep_cpu_int64, thresh)               -> number kept
    nms_normal_gpu(boxes, keep_cpu_int64, thresh)        -> number kept
    boxes_iou_bev_cpu(boxes_a, boxes_b, ans_iou)         -> 1   (CPU tensors in/out; evaluated on the GPU)

Unlike the reference (iou3d_nms.cpp:14-38) bad arguments raise instead of calling exit().
"""
import ctypes as C

import torch

from modest_b200 import _lib


def _check_boxes(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise ValueError("tensor must be a CUDA tensor")
        if not t.is_contiguous():
            raise ValueError("tensor must be contiguous")
        if t.dtype != torch.float32:
            raise ValueError("tensor must be float32")


def _pairs(fn_name, boxes_a, boxes_b, out):
    _check_boxes(boxes_a, boxes_b, out)
    fn = getattr(_lib.lib(), fn_name)
    _lib.check(fn(_lib.ptr(boxes_a), int(boxes_a.shape[0]), _lib.ptr(boxes_b), int(boxes_b.shape[0]),
                  _lib.ptr(out), _lib.stream_ptr()), fn_name)
    return 1


def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    return _pairs("modest_boxes_overlap_bev", boxes_a, boxes_b, ans_overlap)


def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou):
    return _pairs("modest_boxes_iou_bev", boxes_a, boxes_b, ans_iou)


def _nms(fn_name, boxes, keep, thresh):
    _check_boxes(boxes)
    if keep.is_cuda or keep.dtype != torch.int64 or not keep.is_contiguous():
        raise ValueError("keep must be a contiguous CPU int64 tensor")
    lib = _lib.lib()
    n = int(boxes.shape[0])
    ws = torch.empty(int(lib.modest_nms_workspace_bytes(n)), dtype=torch.uint8, device=boxes.device)
    num = C.c_int(0)
    _lib.check(getattr(lib, fn_name)(_lib.ptr(boxes), n, float(thresh), None, C.c_void_p(keep.data_ptr()),
                                     C.byref(num), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), fn_name)
    return int(num.value)


def nms_gpu(boxes, keep, nms_overlap_thresh):
    return _nms("modest_nms_bev", boxes, keep, nms_overlap_thresh)


def nms_normal_gpu(boxes, keep, nms_overlap_thresh):
    return _nms("modest_nms_normal", boxes, keep, nms_overlap_thresh)


def boxes_iou_bev_cpu(boxes_a, boxes_b, ans_iou):
    if boxes_a.is_cuda or boxes_b.is_cuda or ans_iou.is_cuda:
        raise ValueError("boxes_iou_bev_cpu takes CPU tensors")
    a, b = boxes_a.float().contiguous().cuda(), boxes_b.float().contiguous().cuda()
    out = torch.zeros((a.shape[0], b.shape[0]), dtype=torch.float32, device="cuda")
    boxes_iou_bev_gpu(a, b, out)
    ans_iou.copy_(out.cpu())
    return 1
