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

    The reference's CPU op (roiaware_pool3d_utils.py:9-25, roiaware_pool3d.cpp:121-168) with ITS predicate --
    MARGIN 1e-2 (not the GPU op's 1e-5), products rounded individually, the host libm's cosf/sinf -- executed on
    the GPU (no CPU compute path in this build): the per-box constants are formed on the host
    (fnp_host_prep_boxes_cpu), the N x P tests run in one kernel (fnp_points_in_boxes_matrix) that writes the
    matrix row by row; no (N, P, 3) expansion of the points is made.  Bit-equal to the reference-compiled op,
    including points within 1 cm of a face (tests/test_ops_gpu.py).
    """
    import ctypes as C

    from .. import _lib
    assert boxes.shape[1] == 7
    assert points.shape[1] == 3
    points, is_numpy = check_numpy_to_torch(points)
    boxes, is_numpy = check_numpy_to_torch(boxes)
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    b_host = boxes.detach().float().cpu().contiguous()
    n, m = int(b_host.shape[0]), int(points.shape[0])
    prep = torch.empty((n, 8), dtype=torch.float32)
    _lib.check(_lib.lib.fnp_host_prep_boxes_cpu(C.c_void_p(b_host.data_ptr()), C.c_void_p(prep.data_ptr()), n),
               "fnp_host_prep_boxes_cpu")
    p = points.detach().float().to(dev).contiguous()
    out = torch.empty((n, m), dtype=torch.int32, device=dev)
    if n and m:
        prep_d = prep.to(dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib.fnp_points_in_boxes_matrix(C.c_void_p(prep_d.data_ptr()), C.c_void_p(p.data_ptr()),
                                                           C.c_void_p(out.data_ptr()), n, m, _lib.current_stream(dev)),
                       "fnp_points_in_boxes_matrix")
    out = out.cpu()
    return out.numpy() if is_numpy else out
