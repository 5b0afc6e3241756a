"""Drop-in for pcdet/ops/iou3d_nms/iou3d_nms_utils.py.  Reference: iou3d_nms_utils.py:12-152.
Signatures, return types and devices follow the reference; see the notes on nms_*."""
import torch

from . import iou3d_nms_cuda
from .roiaware_pool3d_utils import check_numpy_to_torch


def boxes_bev_iou_cpu(boxes_a, boxes_b):
    """
    Args:
        boxes_a: (N, 7) [x, y, z, dx, dy, dz, heading]   (numpy or CPU tensor)
        boxes_b: (M, 7)
    Returns:
        ans_iou: (N, M), same kind as the inputs.  Executed on the GPU (no CPU path here) with the arithmetic of
        the reference's CUDA kernel (fused multiply-adds as in its sm_100a SASS), NOT of iou3d_cpu.cpp:232, whose
        host-compiled arithmetic rounds every product: values agree with the CPU op to a few ulp (<= 2e-6
        absolute in tests/test_ops_gpu.py), not bit for bit.  A caller that thresholds these IoUs can see a pair
        within that distance of its threshold decided differently.
    """
    boxes_a, is_numpy = check_numpy_to_torch(boxes_a)
    boxes_b, is_numpy = check_numpy_to_torch(boxes_b)
    assert not (boxes_a.is_cuda or boxes_b.is_cuda), 'Only support CPU tensors'
    assert boxes_a.shape[1] == 7 and boxes_b.shape[1] == 7
    dev = torch.device("cuda", torch.cuda.current_device())
    ans = boxes_iou_bev(boxes_a.float().to(dev), boxes_b.float().to(dev)).cpu()
    return ans.numpy() if is_numpy else ans


def boxes_iou_bev(boxes_a, boxes_b):
    """(N,7),(M,7) CUDA -> (N,M) CUDA rotated BEV IoU."""
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    ans_iou = boxes_a.new_empty((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32)
    iou3d_nms_cuda.boxes_iou_bev_gpu(boxes_a.contiguous(), boxes_b.contiguous(), ans_iou)
    return ans_iou


def boxes_iou3d_gpu(boxes_a, boxes_b):
    """(N,7),(M,7) CUDA -> (N,M) 3D IoU.  Same values as the reference's composition of boxes_overlap_bev_gpu and
    eager torch arithmetic (iou3d_nms_utils.py:48-81), computed by one kernel (fnp_boxes_iou3d)."""
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    iou3d = boxes_a.new_empty((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32)
    iou3d_nms_cuda.boxes_iou3d_gpu(boxes_a.contiguous(), boxes_b.contiguous(), iou3d)
    return iou3d


def boxes_aligned_iou3d_gpu(boxes_a, boxes_b):
    """(N,7),(N,7) CUDA -> (N,1) aligned 3D IoU (iou3d_nms_utils.py:83-117), one kernel."""
    assert boxes_a.shape[0] == boxes_b.shape[0]
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    iou3d = boxes_a.new_empty((boxes_a.shape[0], 1), dtype=torch.float32)
    iou3d_nms_cuda.boxes_aligned_iou3d_gpu(boxes_a.contiguous(), boxes_b.contiguous(), iou3d)
    return iou3d


def _nms(native, boxes, scores, thresh, pre_maxsize=None):
    assert boxes.shape[1] == 7
    # Stable descending sort: the reference's unstable sort leaves ties implementation
    # defined; "earlier index wins" is this build's documented rule.  `scores` may live on
    # the CPU (as in frustum_proposals_v1.py:994-1030) -- torch >= 1.12 rejects the
    # reference's order[keep.cuda()] in that case, so indices follow `order`'s device.
    order = scores.sort(stable=True, dim=0, descending=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    boxes = boxes[order.to(boxes.device)].contiguous()
    keep = torch.empty(boxes.size(0), dtype=torch.int64, device=boxes.device)
    num_out = native(boxes, keep, thresh)
    return order[keep[:num_out].to(order.device)].contiguous(), None


def nms_gpu(boxes, scores, thresh, pre_maxsize=None, **kwargs):
    """
    :param boxes: (N, 7) [x, y, z, dx, dy, dz, heading]
    :param scores: (N)
    :param thresh:
    :return: (keep indices into the input, None)
    """
    return _nms(iou3d_nms_cuda.nms_gpu, boxes, scores, thresh, pre_maxsize)


def nms_normal_gpu(boxes, scores, thresh, **kwargs):
    """Axis-aligned (heading ignored) BEV NMS; same contract as nms_gpu."""
    return _nms(iou3d_nms_cuda.nms_normal_gpu, boxes, scores, thresh)
