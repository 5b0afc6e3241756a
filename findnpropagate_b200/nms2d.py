"""Host-side 2D NMS of the GLIP boxes, batched over every (frame, camera) group at once.

Reference call site: frustum_proposals_v1.py:583-588 -- one
``torchvision.ops.batched_nms(cam_boxes, cam_scores, cam_labels, nms_2d)`` per camera per
frame, each costing tens of microseconds of Python/dispatcher overhead.  This module
restates torchvision's CPU algorithm (coordinate-trick offsets per group, stable
score-descending order, greedy suppression with ``inter / (area_i + area_j - inter) > thr``
evaluated in fp32 and compared in fp64) as numpy array ops over all groups of a batch, so a
256-frame batch needs one pass instead of 1536 calls.  tests/test_nms2d.py checks it
against torchvision itself, element for element.
"""
import numpy as np

IMAGE_ORDER = (2, 0, 1, 5, 3, 4)       # frustum_proposals_v1.py:201
_CAM_RANK = np.zeros(6, dtype=np.int64)
for _r, _c in enumerate(IMAGE_ORDER):
    _CAM_RANK[_c] = _r


def batched_nms_groups(boxes, scores, labels, group, iou_thr):
    """Greedy NMS inside each group, boxes of different labels never interact.

    boxes (D,4) f32 xyxy, scores (D) f32, labels (D) int, group (D) int (arbitrary ids).
    Returns (order, keep): `order` sorts the detections by (group ascending, score
    descending, original index ascending); keep[i] is True when detection order[i] survives.
    """
    boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
    scores = np.asarray(scores, dtype=np.float32)
    labels = np.asarray(labels)
    group = np.asarray(group, dtype=np.int64)
    D = boxes.shape[0]
    if D == 0:
        return np.zeros((0,), np.int64), np.zeros((0,), bool)
    # stable: group, then descending score, then original index
    order = np.lexsort((np.arange(D), -scores.astype(np.float64), group))
    g = group[order]
    b = boxes[order]
    lab = labels[order].astype(np.float32)
    # group boundaries
    start = np.flatnonzero(np.r_[True, g[1:] != g[:-1]])
    count = np.diff(np.r_[start, D])
    G, W = start.shape[0], int(count.max())
    gi = np.repeat(np.arange(G), count)
    pos = np.arange(D) - np.repeat(start, count)
    # torchvision coordinate trick, per group: offset = label * (max_coordinate + 1)
    gmax = np.full(G, -np.inf, np.float32)
    np.maximum.at(gmax, gi, b.max(axis=1))
    off = lab * (gmax[gi] + np.float32(1.0))
    bo = b + off[:, None]
    P = np.zeros((G, W, 4), np.float32)
    P[gi, pos] = bo
    valid = np.zeros((G, W), bool)
    valid[gi, pos] = True
    x1, y1, x2, y2 = P[..., 0], P[..., 1], P[..., 2], P[..., 3]
    area = (x2 - x1) * (y2 - y1)
    alive = valid.copy()
    keep_p = np.zeros((G, W), bool)
    cols = np.arange(W)
    for i in range(W):
        cur = alive[:, i]
        if not cur.any():
            continue
        keep_p[:, i] = cur
        xx1 = np.maximum(x1[:, i:i + 1], x1)
        yy1 = np.maximum(y1[:, i:i + 1], y1)
        xx2 = np.minimum(x2[:, i:i + 1], x2)
        yy2 = np.minimum(y2[:, i:i + 1], y2)
        w = np.maximum(np.float32(0), xx2 - xx1)
        h = np.maximum(np.float32(0), yy2 - yy1)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (area[:, i:i + 1] + area - inter)
        sup = (ovr.astype(np.float64) > float(iou_thr)) & (cols[None, :] > i) & cur[:, None]
        alive &= ~sup
    return order, keep_p[gi, pos]


def frustum_candidates(det_boxes, det_labels, det_scores, det_frame, det_cam, nms_2d, score_thr):
    """Candidate frustums of a batch in reference order: frame, then cameras [2,0,1,5,3,4],
    then 2D-NMS order (score descending), keeping score >= score_thr
    (frustum_proposals_v1.py:582-595; the comparison is fp32 like the reference's).
    Returns indices into the detection arrays."""
    det_frame = np.asarray(det_frame, dtype=np.int64)
    det_cam = np.asarray(det_cam, dtype=np.int64)
    group = det_frame * 6 + _CAM_RANK[det_cam]
    order, keep = batched_nms_groups(det_boxes, det_scores, det_labels, group, nms_2d)
    sc = np.asarray(det_scores, dtype=np.float32)[order]
    ok = keep & ~(sc < np.float32(score_thr))
    return order[ok]
