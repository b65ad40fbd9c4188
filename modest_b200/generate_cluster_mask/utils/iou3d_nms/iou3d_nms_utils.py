"""Drop-in for the reference's utils/iou3d_nms/iou3d_nms_utils.py (same names, arguments and
return types), calling libmodest_b200 through the `iou3d_nms_cuda` stand-in next to it.
Reference lines are cited per function (paths relative to generate_cluster_mask/)."""
import numpy as np
import torch

from . import iou3d_nms_cuda


def check_numpy_to_torch(x):
    """iou3d_nms_utils.py:12-15"""
    if isinstance(x, np.ndarray):
        return torch.from_numpy(x).float(), True
    return x, False


def boxes_bev_iou_cpu(boxes_a, boxes_b):
    """iou3d_nms_utils.py:18-34 -- (N,7),(M,7) CPU tensors or arrays -> (N,M) IoU.
    The arithmetic runs on the GPU (this package has no CPU numeric path)."""
    boxes_a, is_numpy = check_numpy_to_torch(boxes_a)
    boxes_b, _ = check_numpy_to_torch(boxes_b)
    assert not (boxes_a.is_cuda or boxes_b.is_cuda), 'Only support CPU tensors'
    assert boxes_a.shape[1] == 7 and boxes_b.shape[1] == 7
    ans_iou = boxes_a.new_zeros(torch.Size((boxes_a.shape[0], boxes_b.shape[0])))
    iou3d_nms_cuda.boxes_iou_bev_cpu(boxes_a.contiguous(), boxes_b.contiguous(), ans_iou)
    return ans_iou.numpy() if is_numpy else ans_iou


def boxes_iou_bev(boxes_a, boxes_b):
    """iou3d_nms_utils.py:37-51 -- CUDA (N,7),(M,7) -> CUDA (N,M) rotated BEV IoU."""
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    ans_iou = torch.zeros((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32, device=boxes_a.device)
    iou3d_nms_cuda.boxes_iou_bev_gpu(boxes_a.contiguous(), boxes_b.contiguous(), ans_iou)
    return ans_iou


def boxes_iou3d_gpu(boxes_a, boxes_b):
    """iou3d_nms_utils.py:54-87 -- BEV overlap x height overlap / union volume."""
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    a_top = (boxes_a[:, 2] + boxes_a[:, 5] / 2).view(-1, 1)
    a_bot = (boxes_a[:, 2] - boxes_a[:, 5] / 2).view(-1, 1)
    b_top = (boxes_b[:, 2] + boxes_b[:, 5] / 2).view(1, -1)
    b_bot = (boxes_b[:, 2] - boxes_b[:, 5] / 2).view(1, -1)
    overlaps_bev = torch.zeros((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32, device=boxes_a.device)
    iou3d_nms_cuda.boxes_overlap_bev_gpu(boxes_a.contiguous(), boxes_b.contiguous(), overlaps_bev)
    overlaps_h = torch.clamp(torch.min(a_top, b_top) - torch.max(a_bot, b_bot), min=0)
    overlaps_3d = overlaps_bev * overlaps_h
    vol_a = (boxes_a[:, 3] * boxes_a[:, 4] * boxes_a[:, 5]).view(-1, 1)
    vol_b = (boxes_b[:, 3] * boxes_b[:, 4] * boxes_b[:, 5]).view(1, -1)
    return overlaps_3d / torch.clamp(vol_a + vol_b - overlaps_3d, min=1e-6)


def nms_gpu(boxes, scores, thresh, pre_maxsize=None, **kwargs):
    """iou3d_nms_utils.py:90-106 -- rotated NMS; returns (kept indices into `boxes`, None)."""
    assert boxes.shape[1] == 7
    order = scores.sort(0, descending=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    boxes = boxes[order].contiguous()
    keep = torch.zeros(boxes.size(0), dtype=torch.long)
    num_out = iou3d_nms_cuda.nms_gpu(boxes, keep, thresh)
    return order[keep[:num_out].to(boxes.device)].contiguous(), None


def nms_normal_gpu(boxes, scores, thresh, **kwargs):
    """iou3d_nms_utils.py:109-122 -- axis-aligned NMS."""
    assert boxes.shape[1] == 7
    order = scores.sort(0, descending=True)[1]
    boxes = boxes[order].contiguous()
    keep = torch.zeros(boxes.size(0), dtype=torch.long)
    num_out = iou3d_nms_cuda.nms_normal_gpu(boxes, keep, thresh)
    return order[keep[:num_out].to(boxes.device)].contiguous(), None
