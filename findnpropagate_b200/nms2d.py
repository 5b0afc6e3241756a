"""Host-side 2D NMS of the GLIP boxes, batched over every (frame, camera) group at once.

Reference call site: frustum_proposals_v1.py:583-588 -- one
``torchvision.ops.batched_nms(cam_boxes, cam_scores, cam_labels, nms_2d)`` per camera per
frame, each costing tens of microseconds of Python/dispatcher overhead.  The selection runs
in the C-ABI library (``fnp_host_select_candidates``, csrc/fnp_host.cpp: torchvision's CPU
arithmetic restated -- coordinate-trick offsets per group, stable score-descending order,
greedy suppression with ``inter / (area_i + area_j - inter)`` evaluated in fp32 and compared
in fp64) over all groups of a batch in one call, so a 256-frame batch needs one call instead
of 1536.  tests/test_host_cpu.py checks it against torchvision itself, element for element.
"""
import ctypes as C

import numpy as np

from . import _lib

IMAGE_ORDER = (2, 0, 1, 5, 3, 4)       # frustum_proposals_v1.py:201
CAM_RANK = np.zeros(6, dtype=np.int64)  # position of camera c in IMAGE_ORDER
for _r, _c in enumerate(IMAGE_ORDER):
    CAM_RANK[_c] = _r


def _p(a):
    return C.c_void_p(a.ctypes.data)


def frustum_candidates(det_boxes, det_labels, det_scores, det_frame, det_cam, nms_2d, score_thr, n_frames=None,
                       return_starts=False):
    """Candidate frustums of a batch in reference order: frame, then cameras [2,0,1,5,3,4],
    then 2D-NMS order (score descending), keeping score >= score_thr
    (frustum_proposals_v1.py:582-595; the comparison is fp32 like the reference's).
    Returns indices into the detection arrays (and frame_cand_start (n_frames+1) on request)."""
    boxes = np.ascontiguousarray(det_boxes, dtype=np.float32).reshape(-1, 4)
    labels = np.ascontiguousarray(det_labels, dtype=np.int64)
    scores = np.ascontiguousarray(det_scores, dtype=np.float32)
    frame = np.ascontiguousarray(det_frame, dtype=np.int64)
    cam = np.ascontiguousarray(det_cam, dtype=np.int64)
    D = boxes.shape[0]
    if n_frames is None:
        n_frames = int(frame.max()) + 1 if D else 0
    sel = np.empty(max(D, 1), np.int32)
    starts = np.zeros(n_frames + 1, np.int32)
    n = _lib.lib.fnp_host_select_candidates(_p(boxes), _p(labels), _p(scores), _p(frame), _p(cam), D, n_frames,
                                            float(nms_2d), float(score_thr), _p(sel), _p(starts))
    if n < 0:
        _lib.check(n, "fnp_host_select_candidates")
    sel = sel[:n].astype(np.int64)
    return (sel, starts) if return_starts else sel
