"""Drop-in for pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py (functions on the Box
Seeker path).  Reference: roiaware_pool3d_utils.py:9-41."""
import numpy as np
import torch

from . import roiaware_pool3d_cuda


def check_numpy_to_torch(x):
    """pcdet/utils/common_utils.py:15-18"""
    if isinstance(x, np.ndarray):
        return torch.from_numpy(x).float(), True
    return x, False


def points_in_boxes_gpu(points, boxes):
    """
    :param points: (B, M, 3)
    :param boxes: (B, T, 7), num_valid_boxes <= T
    :return box_idxs_of_pts: (B, M), default background = -1
    """
    assert boxes.shape[0] == points.shape[0]
    assert boxes.shape[2] == 7 and points.shape[2] == 3
    batch_size, num_points, _ = points.shape
    box_idxs_of_pts = points.new_empty((batch_size, num_points), dtype=torch.int)
    roiaware_pool3d_cuda.points_in_boxes_gpu(boxes.contiguous(), points.contiguous(), box_idxs_of_pts)
    return box_idxs_of_pts


def points_in_boxes_cpu(points, boxes, device=None):
    """
    Args:
        points: (num_points, 3)
        boxes: [x, y, z, dx, dy, dz, heading], (x, y, z) is the box center
    Returns:
        point_indices: (N, num_points) int32 0/1, numpy in -> numpy out

    Same contract as the reference (roiaware_pool3d_utils.py:9-25) but executed on the GPU
    (no CPU compute path in this build): one single-box points_in_boxes launch per box
    batch.  NOTE the reference CPU op uses MARGIN 1e-2 (roiaware_pool3d.cpp:131) while the
    GPU predicate uses 1e-5; this function keeps the GPU predicate, so points within 1 cm
    outside a face differ from the reference CPU op.
    """
    assert boxes.shape[1] == 7
    assert points.shape[1] == 3
    points, is_numpy = check_numpy_to_torch(points)
    boxes, is_numpy = check_numpy_to_torch(boxes)
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    p = points.float().to(dev).contiguous()
    b = boxes.float().to(dev).contiguous()
    n, m = b.shape[0], p.shape[0]
    # (N,1,7) boxes against (N,M,3) broadcast points: index 0 where inside, -1 otherwise
    idx = points_in_boxes_gpu(p.unsqueeze(0).expand(n, m, 3).contiguous(), b.view(n, 1, 7)) if n and m else \
        torch.full((n, m), -1, dtype=torch.int, device=dev)
    out = (idx >= 0).to(torch.int).cpu()
    return out.numpy() if is_numpy else out
