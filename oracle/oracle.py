"""TEST INFRASTRUCTURE ONLY -- ctypes/numpy front end of the CPU oracle (fnp_oracle.c).

Imported only by tests/, tools/gen_golden.py, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(findnpropagate_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(HERE, "libfnp_oracle.so")
_SRC = os.path.join(HERE, "fnp_oracle.c")


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
             "-fvisibility=hidden", "-o", _SO, _SRC, "-lm"])
    return _SO


_lib = None
_f = C.c_float
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
_bp = C.POINTER(C.c_uint8)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        for name in ("fnp_o_sinf", "fnp_o_cosf", "fnp_o_exp"):
            getattr(L, name).restype = _f
            getattr(L, name).argtypes = [_f]
        L.fnp_o_atan2f.restype = _f
        L.fnp_o_atan2f.argtypes = [_f, _f]
        L.fnp_o_pt_in_box.restype = C.c_int
        L.fnp_o_pt_in_box.argtypes = [_fp, _fp, _fp, _fp]
        L.fnp_o_points_in_boxes_gpu.restype = None
        L.fnp_o_points_in_boxes_gpu.argtypes = [C.c_int, C.c_int, C.c_int, _fp, _fp, C.c_int, _ip]
        L.fnp_o_count_in_boxes.restype = None
        L.fnp_o_count_in_boxes.argtypes = [C.c_int, _fp, C.c_int, C.c_int, _fp, _ip]
        L.fnp_o_points_in_boxes_cpu.restype = None
        L.fnp_o_points_in_boxes_cpu.argtypes = [C.c_int, C.c_int, _fp, _fp, _ip]
        L.fnp_o_iou_normal.restype = _f
        L.fnp_o_iou_normal.argtypes = [_fp, _fp]
        for name in ("fnp_o_nms_normal", "fnp_o_nms_rotated"):
            getattr(L, name).restype = C.c_int
            getattr(L, name).argtypes = [_fp, C.c_int, _f, _lp]
        for name in ("fnp_o_box_overlap", "fnp_o_iou_bev"):
            getattr(L, name).restype = _f
            getattr(L, name).argtypes = [_fp, _fp]
        for name in ("fnp_o_boxes_overlap_bev", "fnp_o_boxes_iou_bev"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [C.c_int, _fp, C.c_int, _fp, _fp]
        L.fnp_o_project.restype = C.c_int
        L.fnp_o_project.argtypes = [_fp, _f, _f, _f, _f, _f, _fp]
        L.fnp_o_unproject.restype = None
        L.fnp_o_unproject.argtypes = [_fp, _fp, _f, _f, _f, _fp]
        L.fnp_o_frustum_cull.restype = C.c_int
        L.fnp_o_frustum_cull.argtypes = [_fp, C.c_int, C.c_int, _fp, _fp, _fp, _fp, _f, _f, _ip, _fp, _fp]
        L.fnp_o_quantile.restype = _f
        L.fnp_o_quantile.argtypes = [_fp, C.c_int, _f]
        L.fnp_o_centre_line.restype = None
        L.fnp_o_centre_line.argtypes = [_fp, _f, _f, _fp, _fp, _fp, _fp, C.c_int, _fp, C.c_int, _fp, _fp]
        L.fnp_o_hypotheses.restype = None
        L.fnp_o_hypotheses.argtypes = [_fp, _fp, C.c_int, _fp, C.c_int, _fp, _fp, _f, _f, _f, _f, _fp, _fp, _bp]
        L.fnp_o_select.restype = C.c_int
        L.fnp_o_select.argtypes = [_ip, _fp, _bp, C.c_int, _f, _f, _fp]
        L.fnp_o_centre_line_ex.restype = None
        L.fnp_o_centre_line_ex.argtypes = [_fp, _f, _f, _fp, _fp, _fp, _fp, C.c_int, _fp, C.c_int, _f, _fp, _fp]
        L.fnp_o_hypotheses_ex.restype = None
        L.fnp_o_hypotheses_ex.argtypes = [_fp, _fp, C.c_int, _fp, C.c_int, _fp, _fp, _f, _f, _f, _f,
                                          C.c_int, _fp, _fp, _fp, _fp, _fp, _bp, _bp, _fp]
        L.fnp_o_project_kitti.restype = None
        L.fnp_o_project_kitti.argtypes = [_fp, _f, _f, _f, _fp]
        L.fnp_o_unproject_kitti.restype = None
        L.fnp_o_unproject_kitti.argtypes = [_fp, _f, _f, _f, C.c_int, _fp]
        L.fnp_o_frustum_cull_kitti.restype = C.c_int
        L.fnp_o_frustum_cull_kitti.argtypes = [_fp, C.c_int, C.c_int, _fp, _fp, _ip, _fp, _fp]
        L.fnp_o_centre_line_kitti.restype = None
        L.fnp_o_centre_line_kitti.argtypes = [_fp, _f, _f, _fp, _fp, _fp, C.c_int, _fp, C.c_int, _f, _fp, _fp]
        L.fnp_o_hypotheses_kitti.restype = None
        L.fnp_o_hypotheses_kitti.argtypes = [_fp, _fp, C.c_int, _fp, C.c_int, _fp, _fp, _f, _f, _f, _f, _fp,
                                             _fp, _fp, _bp, _bp, _fp]
        L.fnp_o_occl_fail.restype = None
        L.fnp_o_occl_fail.argtypes = [C.c_int, _fp, C.c_int, C.c_int, _fp, _fp, _ip]
        L.fnp_o_select_ex.restype = C.c_int
        L.fnp_o_select_ex.argtypes = [_ip, _fp, _bp, C.c_int, _fp, _bp, _fp, _fp, _f, _f, _f, _f, _f, C.c_int,
                                      _fp, _fp]
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t=_fp):
    return a.ctypes.data_as(t)


# ---------------------------------------------------------------- scalar math
def sinf(x):
    return np.float32(lib().fnp_o_sinf(np.float32(x)))


def cosf(x):
    return np.float32(lib().fnp_o_cosf(np.float32(x)))


def atan2f(y, x):
    return np.float32(lib().fnp_o_atan2f(np.float32(y), np.float32(x)))


def exp(x):
    return np.float32(lib().fnp_o_exp(np.float32(x)))


# ---------------------------------------------------------------- roiaware ops
def points_in_boxes_gpu(points, boxes):
    """(B,M,3),(B,T,7) -> (B,M) int32 first-match index or -1 (GPU-kernel predicate)."""
    points, boxes = _f32(points), _f32(boxes)
    B, M, _ = points.shape
    T = boxes.shape[1]
    out = np.empty((B, M), np.int32)
    lib().fnp_o_points_in_boxes_gpu(B, T, M, _p(boxes), _p(points), 3, _p(out, _ip))
    return out


def count_in_boxes(points, boxes):
    """points (P,>=3) [xyz first], boxes (H,7) -> (H,) int32 counts."""
    points, boxes = _f32(points), _f32(boxes)
    out = np.empty((boxes.shape[0],), np.int32)
    lib().fnp_o_count_in_boxes(points.shape[0], _p(points), points.shape[1], boxes.shape[0],
                               _p(boxes), _p(out, _ip))
    return out


def points_in_boxes_cpu(points, boxes):
    """(P,3),(N,7) -> (N,P) int32 0/1 (CPU-op predicate, MARGIN 1e-2)."""
    points, boxes = _f32(points), _f32(boxes)
    out = np.empty((boxes.shape[0], points.shape[0]), np.int32)
    lib().fnp_o_points_in_boxes_cpu(boxes.shape[0], points.shape[0], _p(boxes), _p(points), _p(out, _ip))
    return out


# ---------------------------------------------------------------- iou / nms ops
def iou_normal(a, b):
    a, b = _f32(a), _f32(b)
    return np.float32(lib().fnp_o_iou_normal(_p(a), _p(b)))


def _nms(fn, boxes, scores, thresh, pre_maxsize=None):
    boxes = _f32(boxes)
    scores = np.asarray(scores, np.float32)
    order = np.argsort(-scores, kind="stable")
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    sb = np.ascontiguousarray(boxes[order])
    keep = np.empty((sb.shape[0],), np.int64)
    n = fn(_p(sb), sb.shape[0], np.float32(thresh), _p(keep, _lp))
    return order[keep[:n]].astype(np.int64)


def nms_normal(boxes, scores, thresh):
    """iou3d_nms_utils.nms_normal_gpu restated (stable descending sort)."""
    return _nms(lib().fnp_o_nms_normal, boxes, scores, thresh)


def nms_rotated(boxes, scores, thresh, pre_maxsize=None):
    """iou3d_nms_utils.nms_gpu restated (stable descending sort)."""
    return _nms(lib().fnp_o_nms_rotated, boxes, scores, thresh, pre_maxsize)


def boxes_overlap_bev(a, b):
    a, b = _f32(a), _f32(b)
    out = np.empty((a.shape[0], b.shape[0]), np.float32)
    lib().fnp_o_boxes_overlap_bev(a.shape[0], _p(a), b.shape[0], _p(b), _p(out))
    return out


def boxes_iou_bev(a, b):
    a, b = _f32(a), _f32(b)
    out = np.empty((a.shape[0], b.shape[0]), np.float32)
    lib().fnp_o_boxes_iou_bev(a.shape[0], _p(a), b.shape[0], _p(b), _p(out))
    return out


def boxes_iou3d(a, b):
    """iou3d_nms_utils.boxes_iou3d_gpu restated (iou3d_nms_utils.py:48-81): overlap-area
    kernel times height overlap; elementwise fp32 torch ops (no contraction)."""
    a, b = _f32(a), _f32(b)
    f = np.float32
    a_max = (a[:, 2] + a[:, 5] / f(2)).reshape(-1, 1)
    a_min = (a[:, 2] - a[:, 5] / f(2)).reshape(-1, 1)
    b_max = (b[:, 2] + b[:, 5] / f(2)).reshape(1, -1)
    b_min = (b[:, 2] - b[:, 5] / f(2)).reshape(1, -1)
    ov_bev = boxes_overlap_bev(a, b)
    ov_h = np.maximum(np.minimum(a_max, b_max) - np.maximum(a_min, b_min), f(0))
    ov3 = ov_bev * ov_h
    va = (a[:, 3] * a[:, 4] * a[:, 5]).reshape(-1, 1)
    vb = (b[:, 3] * b[:, 4] * b[:, 5]).reshape(1, -1)
    return (ov3 / np.maximum(va + vb - ov3, f(1e-6))).astype(np.float32)


# ---------------------------------------------------------------- seeker stages
def project(L, xyz, img_w=1600.0, img_h=900.0):
    L = _f32(L).reshape(-1)
    xyz = _f32(xyz)
    out = np.empty((xyz.shape[0], 3), np.float32)
    on = np.empty((xyz.shape[0],), bool)
    fn = lib().fnp_o_project
    for i in range(xyz.shape[0]):
        on[i] = fn(_p(L), xyz[i, 0], xyz[i, 1], xyz[i, 2], img_w, img_h, _p(out[i]))
    return out, on


def frustum_cull(points, L, combine, trans, box2d, img_w=1600.0, img_h=900.0):
    """points (N,>=3) -> idx (P,), uvd (P,3), xyz (P,3) of the points inside the 2D box."""
    points = _f32(points)
    L, combine, trans, box2d = (_f32(x).reshape(-1) for x in (L, combine, trans, box2d))
    N = points.shape[0]
    idx = np.empty((N,), np.int32)
    uvd = np.empty((N, 3), np.float32)
    xyz = np.empty((N, 3), np.float32)
    n = lib().fnp_o_frustum_cull(_p(points), N, points.shape[1], _p(L), _p(combine), _p(trans),
                                 _p(box2d), img_w, img_h, _p(idx, _ip), _p(uvd), _p(xyz))
    return idx[:n].copy(), uvd[:n].copy(), xyz[:n].copy()


def quantile(x, q):
    x = _f32(x)
    return np.float32(lib().fnp_o_quantile(_p(x), x.shape[0], np.float32(q)))


def centre_line(box2d, dmin, dmax, combine, trans, pmin, pmax, clamp_bottom, mags, search_depth=None):
    box2d, combine, trans, pmin, pmax, mags = (_f32(x).reshape(-1) for x in (box2d, combine, trans, pmin, pmax, mags))
    M = mags.shape[0]
    centres = np.empty((M, 3), np.float32)
    corners = np.empty((8, 3), np.float32)
    lib().fnp_o_centre_line_ex(_p(box2d), np.float32(dmin), np.float32(dmax), _p(combine), _p(trans),
                               _p(pmin), _p(pmax), int(clamp_bottom), _p(mags), M,
                               np.float32(search_depth or 0.0), _p(centres), _p(corners))
    return centres, corners


def unproject(combine, trans, u, v, d):
    combine, trans = _f32(combine).reshape(-1), _f32(trans).reshape(-1)
    out = np.empty(3, np.float32)
    lib().fnp_o_unproject(_p(combine), _p(trans), np.float32(u), np.float32(v), np.float32(d), _p(out))
    return out


def hypotheses(base_boxes, base_corners, centres, L, box2d, max_dist, min_iou,
               img_w=1600.0, img_h=900.0):
    base_boxes, base_corners, centres = _f32(base_boxes), _f32(base_corners), _f32(centres)
    L, box2d = _f32(L).reshape(-1), _f32(box2d).reshape(-1)
    J, M = base_boxes.shape[0], centres.shape[0]
    H = J * M
    boxes = np.empty((H, 7), np.float32)
    iou = np.empty((H,), np.float32)
    valid = np.empty((H,), np.uint8)
    lib().fnp_o_hypotheses(_p(base_boxes), _p(base_corners), J, _p(centres), M, _p(L), _p(box2d),
                           img_w, img_h, np.float32(max_dist), np.float32(min_iou),
                           _p(boxes), _p(iou), _p(valid, _bp))
    return boxes, iou, valid.astype(bool)


def hypotheses_ex(base_boxes, base_corners, centres, L, box2d, max_dist, min_iou, views=None, wc=None,
                  img_w=1600.0, img_h=900.0):
    """hypotheses() with the optional parts: views = (Ls (n,4,4), boxes2d (n,4)) for MULTICAM_IOU,
    wc (3) the weighted centre -> also returns near (H) and dist (H)."""
    base_boxes, base_corners, centres = _f32(base_boxes), _f32(base_corners), _f32(centres)
    L, box2d = _f32(L).reshape(-1), _f32(box2d).reshape(-1)
    J, M = base_boxes.shape[0], centres.shape[0]
    H = J * M
    boxes = np.empty((H, 7), np.float32)
    iou = np.empty((H,), np.float32)
    valid = np.empty((H,), np.uint8)
    near = np.empty((H,), np.uint8)
    dist = np.zeros((H,), np.float32)
    n_views, vL, vb = 0, None, None
    if views is not None:
        vL, vb = _f32(views[0]).reshape(-1, 16), _f32(views[1]).reshape(-1, 4)
        n_views = vL.shape[0]
    wcv = _f32(wc).reshape(-1) if wc is not None else None
    lib().fnp_o_hypotheses_ex(_p(base_boxes), _p(base_corners), J, _p(centres), M, _p(L), _p(box2d),
                              img_w, img_h, np.float32(max_dist), np.float32(min_iou), n_views,
                              _p(vL) if n_views else None, _p(vb) if n_views else None,
                              _p(wcv) if wcv is not None else None, _p(boxes), _p(iou), _p(valid, _bp),
                              _p(near, _bp), _p(dist) if wcv is not None else None)
    return boxes, iou, valid.astype(bool), near.astype(bool), dist


def occl_fail(points, boxes):
    """calc_occl_scores restated: points (P,>=3), boxes (H,7) -> fail (H,) f32 (= n_far * n_out), n_far (H,) int32."""
    points, boxes = _f32(points), _f32(boxes)
    out = np.empty((boxes.shape[0],), np.float32)
    nfar = np.empty((boxes.shape[0],), np.int32)
    lib().fnp_o_occl_fail(points.shape[0], _p(points), points.shape[1], boxes.shape[0], _p(boxes), _p(out),
                          _p(nfar, _ip))
    return out, nfar


def select_ex(counts, iou, valid, dist=None, near=None, fail=None, boxes=None, dns_w=1.0, iou_w=1.0, dst_w=0.0,
              occl_w=0.0, ego_w=0.0, mult=False, occl_mult=False):
    """Score with the optional terms + argmax; returns (h, best score, scores (H) [valid entries])."""
    counts = np.ascontiguousarray(counts, np.int32)
    iou = _f32(iou)
    valid = np.ascontiguousarray(valid, np.uint8)
    H = counts.shape[0]
    dist = _f32(dist) if dist is not None else None
    near = np.ascontiguousarray(near, np.uint8) if near is not None else None
    fail = _f32(fail) if fail is not None else None
    boxes = _f32(boxes) if boxes is not None else None
    scores = np.zeros((H,), np.float32)
    s = C.c_float(0)
    h = lib().fnp_o_select_ex(_p(counts, _ip), _p(iou), _p(valid, _bp), H,
                              _p(dist) if dist is not None else None, _p(near, _bp) if near is not None else None,
                              _p(fail) if fail is not None else None, _p(boxes) if boxes is not None else None,
                              np.float32(dns_w), np.float32(iou_w), np.float32(dst_w), np.float32(occl_w),
                              np.float32(ego_w), int(bool(mult)) | (int(bool(occl_mult)) << 1), _p(scores), C.byref(s))
    return h, np.float32(s.value), scores


def select(counts, iou, valid, dns_w=1.0, iou_w=1.0):
    counts = np.ascontiguousarray(counts, np.int32)
    iou = _f32(iou)
    valid = np.ascontiguousarray(valid, np.uint8)
    s = C.c_float(0)
    h = lib().fnp_o_select(_p(counts, _ip), _p(iou), _p(valid, _bp), counts.shape[0],
                           np.float32(dns_w), np.float32(iou_w), C.byref(s))
    return h, np.float32(s.value)


# ---------------------------------------------------------------- KITTI head (FrustumProposerOGKITTI)
def frustum_cull_kitti(points, K, box2d):
    """points (N,>=3), K (48,) calibration block -> idx (P,), uvd (P,3), xyz (P,3) of the points inside the 2D box
    (no on-image test in that head)."""
    points = _f32(points)
    K, box2d = _f32(K).reshape(-1), _f32(box2d).reshape(-1)
    N = points.shape[0]
    idx = np.empty((N,), np.int32)
    uvd = np.empty((N, 3), np.float32)
    xyz = np.empty((N, 3), np.float32)
    n = lib().fnp_o_frustum_cull_kitti(_p(points), N, points.shape[1], _p(K), _p(box2d), _p(idx, _ip), _p(uvd), _p(xyz))
    return idx[:n].copy(), uvd[:n].copy(), xyz[:n].copy()


def unproject_kitti(K, u, v, d, chain=False):
    K = _f32(K).reshape(-1)
    out = np.empty(3, np.float32)
    lib().fnp_o_unproject_kitti(_p(K), np.float32(u), np.float32(v), np.float32(d), int(chain), _p(out))
    return out


def centre_line_kitti(box2d, dmin, dmax, K, pmin, pmax, clamp_bottom, mags, search_depth=None):
    box2d, K, pmin, pmax, mags = (_f32(x).reshape(-1) for x in (box2d, K, pmin, pmax, mags))
    M = mags.shape[0]
    centres = np.empty((M, 3), np.float32)
    corners = np.empty((8, 3), np.float32)
    lib().fnp_o_centre_line_kitti(_p(box2d), np.float32(dmin), np.float32(dmax), _p(K), _p(pmin), _p(pmax),
                                  int(clamp_bottom), _p(mags), M, np.float32(search_depth or 0.0), _p(centres), _p(corners))
    return centres, corners


def hypotheses_kitti(base_boxes, base_corners, centres, K, box2d, max_dist, min_iou, wc, img_w=1600.0, img_h=900.0):
    base_boxes, base_corners, centres = _f32(base_boxes), _f32(base_corners), _f32(centres)
    K, box2d, wcv = _f32(K).reshape(-1), _f32(box2d).reshape(-1), _f32(wc).reshape(-1)
    J, M = base_boxes.shape[0], centres.shape[0]
    H = J * M
    boxes = np.empty((H, 7), np.float32)
    iou = np.empty((H,), np.float32)
    valid = np.empty((H,), np.uint8)
    near = np.empty((H,), np.uint8)
    dist = np.zeros((H,), np.float32)
    lib().fnp_o_hypotheses_kitti(_p(base_boxes), _p(base_corners), J, _p(centres), M, _p(K), _p(box2d), img_w, img_h,
                                 np.float32(max_dist), np.float32(min_iou), _p(wcv), _p(boxes), _p(iou), _p(valid, _bp),
                                 _p(near, _bp), _p(dist))
    return boxes, iou, valid.astype(bool), near.astype(bool), dist
