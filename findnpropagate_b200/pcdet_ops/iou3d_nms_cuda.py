"""Shim with the pybind11 surface of the reference's `iou3d_nms_cuda` extension
(reference: pcdet/ops/iou3d_nms/src/iou3d_nms_api.cpp:11-19).

Differences (deliberate): the NMS entry points accept `keep` on either
device -- a CUDA int64 `keep` avoids the device->host copy the reference performs inside
the op; a CPU `keep` is filled with one copy of num_keep indices, as the reference does."""
import ctypes as C

import torch

from .. import _lib
from .roiaware_pool3d_cuda import _check_f32_cuda, _ptr

_ws_cache = {}


def _workspace(nbytes, device):
    """NMS scratch (bitmask + prepared boxes), one buffer per (device, stream): calls enqueued on different
    streams of one device must not share it.  A buffer that is outgrown is handed back to torch's caching
    allocator, which is stream-ordered, so kernels still reading it are not overtaken."""
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def _pair(fn, name, boxes_a, boxes_b, out):
    _check_f32_cuda(boxes_a, "boxes_a", 7)
    _check_f32_cuda(boxes_b, "boxes_b", 7)
    _check_f32_cuda(out, "out", out.shape[-1])
    N, M = boxes_a.shape[0], boxes_b.shape[0]
    if out.numel() != N * M:
        raise ValueError("%s: output must have %d elements" % (name, N * M))
    with torch.cuda.device(boxes_a.device):
        rc = fn(_ptr(boxes_a), _ptr(boxes_b), _ptr(out), N, M, _lib.current_stream(boxes_a.device))
    _lib.check(rc, name)
    return 1


def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    return _pair(_lib.lib.fnp_boxes_overlap_bev, "fnp_boxes_overlap_bev", boxes_a, boxes_b, ans_overlap)


def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou):
    return _pair(_lib.lib.fnp_boxes_iou_bev, "fnp_boxes_iou_bev", boxes_a, boxes_b, ans_iou)


def boxes_aligned_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    _check_f32_cuda(boxes_a, "boxes_a", 7)
    _check_f32_cuda(boxes_b, "boxes_b", 7)
    N = boxes_a.shape[0]
    if boxes_b.shape[0] != N or ans_overlap.numel() != N:
        raise ValueError("aligned overlap: shape mismatch")
    with torch.cuda.device(boxes_a.device):
        rc = _lib.lib.fnp_boxes_aligned_overlap_bev(_ptr(boxes_a), _ptr(boxes_b), _ptr(ans_overlap), N,
                                                    _lib.current_stream(boxes_a.device))
    _lib.check(rc, "fnp_boxes_aligned_overlap_bev")
    return 1


def boxes_iou3d_gpu(boxes_a, boxes_b, ans_iou):
    """(N,7),(M,7) -> (N,M) 3D IoU in one kernel (no counterpart in the reference's pybind module: the reference
    composes it from boxes_overlap_bev_gpu and a dozen eager torch ops, iou3d_nms_utils.py:48-81)."""
    return _pair(_lib.lib.fnp_boxes_iou3d, "fnp_boxes_iou3d", boxes_a, boxes_b, ans_iou)


def boxes_aligned_iou3d_gpu(boxes_a, boxes_b, ans_iou):
    _check_f32_cuda(boxes_a, "boxes_a", 7)
    _check_f32_cuda(boxes_b, "boxes_b", 7)
    N = boxes_a.shape[0]
    if boxes_b.shape[0] != N or ans_iou.numel() != N:
        raise ValueError("aligned iou3d: shape mismatch")
    with torch.cuda.device(boxes_a.device):
        rc = _lib.lib.fnp_boxes_aligned_iou3d(_ptr(boxes_a), _ptr(boxes_b), _ptr(ans_iou), N,
                                              _lib.current_stream(boxes_a.device))
    _lib.check(rc, "fnp_boxes_aligned_iou3d")
    return 1


def _nms(fn, name, boxes, keep, thresh):
    _check_f32_cuda(boxes, "boxes", 7)
    if keep.dtype != torch.int64 or not keep.is_contiguous() or keep.numel() < boxes.shape[0]:
        raise ValueError("keep must be a contiguous int64 tensor with >= N elements")
    N = boxes.shape[0]
    dev = boxes.device
    nbytes = _lib.lib.fnp_nms_workspace_bytes(N)
    ws = _workspace(nbytes + 8, dev)
    keep_dev = keep if keep.is_cuda else torch.empty(N, dtype=torch.int64, device=dev)
    num = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = fn(_ptr(boxes), N, C.c_float(float(thresh)), _ptr(keep_dev), _ptr(num), _ptr(ws), ws.numel(),
                _lib.current_stream(dev))
    _lib.check(rc, name)
    n = int(num.item())      # the API returns num_to_keep as a Python int, like the reference
    if not keep.is_cuda:
        keep[:n] = keep_dev[:n].cpu()
    return n


def nms_gpu(boxes, keep, nms_overlap_thresh):
    """boxes (N,7) sorted by descending score; keep (N) int64; returns num_to_keep."""
    return _nms(_lib.lib.fnp_nms_rotated, "fnp_nms_rotated", boxes, keep, nms_overlap_thresh)


def nms_normal_gpu(boxes, keep, nms_overlap_thresh):
    return _nms(_lib.lib.fnp_nms_normal, "fnp_nms_normal", boxes, keep, nms_overlap_thresh)


def boxes_iou_bev_cpu(boxes_a, boxes_b, ans_iou):
    raise RuntimeError("boxes_iou_bev_cpu: no CPU compute path in this build; use "
                       "pcdet_ops.iou3d_nms_utils.boxes_bev_iou_cpu (device-executed)")
