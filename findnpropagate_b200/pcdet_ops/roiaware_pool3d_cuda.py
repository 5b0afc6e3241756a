"""Shim with the pybind11 surface of the reference's `roiaware_pool3d_cuda` extension for
the functions on the Box Seeker path (reference: pcdet/ops/roiaware_pool3d/src/
roiaware_pool3d.cpp:172-177).  RoI-aware pooling forward/backward are out of scope."""
import ctypes as C

import torch

from .. import _lib


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _check_f32_cuda(t, name, last):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    _lib.require_cuda(t)
    if t.dtype != torch.float32:
        raise ValueError("%s must be float32, got %s" % (name, t.dtype))
    if t.shape[-1] != last:
        raise ValueError("%s must have last dimension %d, got %s" % (name, last, tuple(t.shape)))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)


def points_in_boxes_gpu(boxes_tensor, pts_tensor, box_idx_of_points_tensor):
    """boxes (B,T,7), pts (B,M,3), out (B,M) int32 (every entry written: first box or -1).
    Returns 1 like the reference (roiaware_pool3d.cpp:98-118)."""
    _check_f32_cuda(boxes_tensor, "boxes", 7)
    _check_f32_cuda(pts_tensor, "pts", 3)
    out = box_idx_of_points_tensor
    _lib.require_cuda(out)
    if out.dtype != torch.int32 or not out.is_contiguous():
        raise ValueError("box_idx_of_points must be a contiguous int32 tensor")
    B, T = boxes_tensor.shape[0], boxes_tensor.shape[1]
    M = pts_tensor.shape[1]
    if pts_tensor.shape[0] != B or tuple(out.shape) != (B, M):
        raise ValueError("shape mismatch: boxes %s pts %s out %s" % (tuple(boxes_tensor.shape), tuple(pts_tensor.shape), tuple(out.shape)))
    with torch.cuda.device(pts_tensor.device):
        rc = _lib.lib.fnp_points_in_boxes(_ptr(boxes_tensor), _ptr(pts_tensor), _ptr(out), B, T, M,
                                          _lib.current_stream(pts_tensor.device))
    _lib.check(rc, "fnp_points_in_boxes")
    return 1


def points_in_boxes_cpu(boxes_tensor, pts_tensor, pts_indices_tensor):
    raise RuntimeError(
        "points_in_boxes_cpu: this build has no CPU compute path; use "
        "pcdet_ops.roiaware_pool3d_utils.points_in_boxes_cpu (device-executed, same (N,P) 0/1 contract)")
