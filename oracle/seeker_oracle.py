"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the Greedy Box Seeker frame loop.

Restates FrustumProposerOG.get_proposals (reference:
pcdet/models/dense_heads/frustum_proposals_v1.py:523-1067) on top of the C oracle
(fnp_oracle.c) with the shipped option set of
tools/cfgs/nuscenes_box_seeker_proposals.yaml:83 (no img/lidar augmentation, nms_3d = 0),
plus the optional terms of SURVEY.md 8 row f3: dst_w, ego_w, occl_w, search_depth and the flags
MULT, OCCL_MULT, MULTICAM_IOU (passed as keys of `params`), topk > 1 with nms_normal.
seek_frame_kitti restates FrustumProposerOGKITTI.get_proposals (frustum_proposals_v1_kitti.py:292-690) the same way.

Never imported by the product package.  Used by tests/ (as the checker of the CUDA
pipeline), tools/gen_golden.py and the CPU-baseline legs of bench.py.
"""
import numpy as np
import torch

import oracle as O

IMAGE_ORDER = [2, 0, 1, 5, 3, 4]          # frustum_proposals_v1.py:201
IMG_H, IMG_W = 900.0, 1600.0              # frustum_proposals_v1.py:203
FRUSTUM_MIN = np.float32(2.0)             # frustum_proposals_v1.py:240

ANCHORS = [[4.63, 1.97, 1.74], [6.93, 2.51, 2.84], [6.37, 2.85, 3.19], [10.5, 2.94, 3.47],
           [12.29, 2.90, 3.87], [0.50, 2.53, 0.98], [2.11, 0.77, 1.47], [1.70, 0.60, 1.28],
           [0.73, 0.67, 1.77], [0.41, 0.41, 1.07]]   # frustum_proposals_v1.py:270-281

DEFAULTS = dict(lq=0.336, uq=0.356, iou_w=0.95, dst_w=0.226, dns_w=0.05, min_cam_iou=0.3,
                size_min=0.957, size_max=1.2, ry_min=0.0, ry_max=float(torch.pi), cq=0.46,
                num_mags=6, max_dist=50, num_sizes=4, num_rotations=10, topk=1, nms_2d=0.7,
                nms_3d=1.0, score_thr=0.1, nms_normal=0.7, clamp_bottom=0)  # :146-148


def build_tables(params):
    """base_boxes (A, R*S, 7) and base_corners (A, R*S, 8, 3): the constructor tables,
    frustum_proposals_v1.py:282-298 + box_utils.boxes_to_corners_3d (box_utils.py:28-52),
    evaluated with the same torch calls on the CPU."""
    p = dict(DEFAULTS); p.update(params)
    anchors = torch.tensor(ANCHORS, dtype=torch.float32)
    A, R, S = anchors.shape[0], p["num_rotations"], p["num_sizes"]
    size_var = torch.linspace(p["size_min"], p["size_max"], steps=S)
    rots = torch.linspace(p["ry_min"], p["ry_max"], steps=R)
    bb = torch.zeros((A, R, S, 7))
    for i in range(A):
        bb[i, :, :, [3, 4, 5]] = anchors[i]
    for i in range(R):
        bb[:, i, :, -1] = rots[i]
    for i, m in enumerate(size_var):
        bb[:, :, i, [3, 4, 5]] = bb[:, :, i, [3, 4, 5]] * m
    flat = bb.reshape(-1, 7)
    tpl = flat.new_tensor(([1, 1, -1], [1, -1, -1], [-1, -1, -1], [-1, 1, -1],
                           [1, 1, 1], [1, -1, 1], [-1, -1, 1], [-1, 1, 1])) / 2
    c = flat[:, None, 3:6].repeat(1, 8, 1) * tpl[None]
    cosa, sina = torch.cos(flat[:, 6]), torch.sin(flat[:, 6])
    z, o = torch.zeros_like(cosa), torch.ones_like(cosa)
    rot = torch.stack((cosa, sina, z, -sina, cosa, z, z, z, o), dim=1).view(-1, 3, 3)
    c = torch.matmul(c, rot) + flat[:, None, 0:3]
    return bb.reshape(A, -1, 7).numpy(), c.reshape(A, -1, 8, 3).numpy()


def camera_matrices(camera2lidar, camera_intrinsics):
    """combine = cam2lidar_R @ inverse(K) and the translation, per camera
    (frustum_proposals_v1.py:1512-1535), with torch on the host."""
    c2l = torch.as_tensor(camera2lidar).float()
    K = torch.as_tensor(camera_intrinsics).float()[..., :3, :3]
    combine = c2l[..., :3, :3].matmul(torch.inverse(K))
    return combine.numpy(), c2l[..., :3, 3].numpy().copy()


def nms2d_candidates(det_boxes, det_labels, det_scores, det_cam_idx, nms_2d, score_thr, box_format="xyxy"):
    """Frustum candidates of one frame in reference order: cameras [2,0,1,5,3,4], per-camera
    torchvision batched_nms (score-descending), then the score threshold
    (frustum_proposals_v1.py:582-595)."""
    from torchvision.ops import batched_nms
    boxes = torch.as_tensor(det_boxes).float().reshape(-1, 4)
    labels = torch.as_tensor(det_labels).long()
    scores = torch.as_tensor(det_scores).float()
    cams = torch.as_tensor(det_cam_idx).long()
    out = []
    for c in IMAGE_ORDER:
        m = cams == c
        cb, cl, cs = boxes[m], labels[m], scores[m]
        if cb.shape[0] > 0:
            sel = batched_nms(cb, cs, cl, nms_2d)
            cb, cl, cs = cb[sel], cl[sel], cs[sel]
        for b, l, s in zip(cb, cl, cs):
            if s < score_thr:
                continue
            if box_format != "xyxy":       # x, y, w, h: the 2D NMS above saw the raw numbers, the corner
                b = b.clone()              # is formed afterwards in fp32 (frustum_proposals_v1.py:597-601)
                b[2:] += b[0:2]
            out.append((c, b.numpy().copy(), int(l), np.float32(s.item())))
    return out


def seek_frame(points, lidar2image, camera2lidar, camera_intrinsics, dets, params, tables=None,
               keep_intermediates=False, box_format="xyxy"):
    """One frame.  points (N,>=3) xyz first; dets = (boxes, labels, scores, cam_idx).
    Returns dict(pred_boxes (K,7) f32, pred_labels (K) int32, pred_scores (K) f32,
    frustums=[per-frustum intermediates])."""
    p = dict(DEFAULTS); p.update(params)
    assert p["nms_3d"] == 0
    topk = int(p["topk"])
    # optional terms: dst_w, ego_w, occl_w, search_depth, MULT, OCCL_MULT, MULTICAM_IOU;
    # aln_w (randomised pca_lowrank) and rand_center (randn) are not deterministic in the reference
    assert not p.get("aln_w") and not p.get("rand_center")
    dst_w, ego_w, occl_w = float(p["dst_w"]), float(p.get("ego_w") or 0), float(p.get("occl_w") or 0)
    mult, occl_mult, multicam = bool(p.get("MULT")), bool(p.get("OCCL_MULT")), bool(p.get("MULTICAM_IOU"))
    sdepth = p.get("search_depth")
    use_dist, use_fail = dst_w != 0 or mult, occl_w > 0 or occl_mult
    extras = use_dist or use_fail or ego_w > 0 or multicam
    base_boxes, base_corners = tables if tables is not None else build_tables(p)
    combine, trans = camera_matrices(camera2lidar, camera_intrinsics)
    mags = torch.linspace(0.0, 1.0, p["num_mags"]).numpy() if p["num_mags"] > 0 else np.zeros(1, np.float32)
    max_dist = np.float32(p["max_dist"])
    pts = np.ascontiguousarray(points, np.float32)
    cands = nms2d_candidates(*dets, p["nms_2d"], p["score_thr"], box_format)
    culls = [O.frustum_cull(pts, np.ascontiguousarray(lidar2image[c], np.float32), combine[c], trans[c], box2d,
                            IMG_W, IMG_H) for (c, box2d, _, _) in cands]
    boxes_out, labels_out, scores_out, inter = [], [], [], []
    for k, (c, box2d, label, score) in enumerate(cands):
        L = np.ascontiguousarray(lidar2image[c], np.float32)
        idx, uvd, xyz = culls[k]
        rec = dict(cam=c, box2d=box2d, label=label, score=score, n_points=int(idx.shape[0]))
        if idx.shape[0] == 0:
            if keep_intermediates:
                inter.append(rec)
            continue
        d = uvd[:, 2]
        qmin = O.quantile(d, p["lq"])
        # search_depth (:619-623): the far end is the near quantile + depth
        qmax = O.quantile(d, p["uq"]) if sdepth is None else np.float32(qmin + np.float32(sdepth))
        dmin = np.maximum(qmin, FRUSTUM_MIN)
        dmax = np.minimum(qmax, max_dist)
        centres, corners = O.centre_line(box2d, dmin, dmax, combine[c], trans[c], xyz.min(0), xyz.max(0),
                                         p["clamp_bottom"], mags, search_depth=sdepth)
        if not extras:
            hb, iou, valid = O.hypotheses(base_boxes[label - 1], base_corners[label - 1], centres, L, box2d,
                                          max_dist, p["min_cam_iou"], IMG_W, IMG_H)
            near = dist = None
        else:
            # weighted_centre_xyz (:631-636): box centre at the cq depth quantile, unprojected
            wc = O.unproject(combine[c], trans[c], (box2d[0] + box2d[2]) / np.float32(2), (box2d[1] + box2d[3]) / np.float32(2),
                             O.quantile(d, p["cq"]))
            views = None
            if multicam:   # every candidate of the frame with points and the same label, itself included (:879-880)
                same = [i for i, cd in enumerate(cands) if cd[2] == label and culls[i][0].shape[0] > 0]
                views = (np.stack([lidar2image[cands[i][0]] for i in same]), np.stack([cands[i][1] for i in same]))
            hb, iou, valid, near, dist = O.hypotheses_ex(base_boxes[label - 1], base_corners[label - 1], centres, L,
                                                         box2d, max_dist, p["min_cam_iou"], views=views, wc=wc)
            rec.update(wc=wc, near=near, dist=dist)
        counts = np.zeros(hb.shape[0], np.int32)
        fail = np.zeros(hb.shape[0], np.float32) if use_fail else None
        nfar = np.zeros(hb.shape[0], np.int32)
        if valid.any():
            counts[valid] = O.count_in_boxes(xyz, hb[valid])
            if use_fail:
                fail[valid], nfar[valid] = O.occl_fail(xyz, hb[valid])
        if not extras and topk == 1:
            h, s = O.select(counts, iou, valid, p["dns_w"], p["iou_w"])
        else:
            h, s, sc = O.select_ex(counts, iou, valid, dist=dist, near=near, fail=fail, boxes=hb, dns_w=p["dns_w"],
                                   iou_w=p["iou_w"], dst_w=dst_w, occl_w=occl_w, ego_w=ego_w, mult=mult,
                                   occl_mult=occl_mult)
            rec.update(scores=sc, fail=fail, nfar=nfar)
        rec.update(idx=idx, uvd=uvd, xyz=xyz, dmin=dmin, dmax=dmax, centres=centres, corners=corners,
                   hyp_boxes=hb, iou=iou, valid=valid, counts=counts, best=h, best_score=s)
        if keep_intermediates:
            inter.append(rec)
        if h < 0:
            continue
        if topk == 1:
            boxes_out.append(hb[h]); labels_out.append(label); scores_out.append(score)
        else:
            # nms_normal_gpu over the valid hypotheses in descending second-stage score (stable), the first
            # topk survivors in that order (:1030-1046); the frustum's 2D score and label repeat (:1048-1052)
            vi = np.flatnonzero(valid)
            keep = O.nms_normal(hb[vi], rec["scores"][vi], p["nms_normal"])[:topk]
            rec["topk"] = vi[keep]
            for hh in vi[keep]:
                boxes_out.append(hb[hh]); labels_out.append(label); scores_out.append(score)
    return dict(
        pred_boxes=np.asarray(boxes_out, np.float32).reshape(-1, 7),
        pred_labels=np.asarray(labels_out, np.int32),
        pred_scores=np.asarray(scores_out, np.float32),
        frustums=inter)


ANCHORS_KITTI = [[3.9, 1.6, 1.56], [6.37, 2.85, 3.19], [6.93, 2.51, 2.84], [6.93, 2.51, 2.84], [0.8, 0.6, 1.73],
                 [1.76, 0.6, 1.73], [0.8, 0.6, 1.73]]      # frustum_proposals_v1_kitti.py:157-166


def build_tables_kitti(params):
    """The KITTI head's constructor tables (frustum_proposals_v1_kitti.py:167-183): the same torch calls as
    build_tables over its seven anchors."""
    global ANCHORS
    saved, ANCHORS = ANCHORS, ANCHORS_KITTI
    try:
        return build_tables(dict(dict(max_dist=70), **params))
    finally:
        ANCHORS = saved


def kitti_block(P2, R0, V2C):
    """The 48-float calibration block of a KITTI frame (include/fnp.h, FNP_VARIANT_KITTI), formed with the calls
    CalibrationTorch makes (calibration_kitti.py:128-169) -- on the host here; the GPU tests pass the block the engine
    formed on the device, so that both sides read the same numbers."""
    P2, R0, V2C = (torch.as_tensor(np.asarray(x, np.float32)) for x in (P2, R0, V2C))
    m1 = V2C.T @ R0.T
    cu, cv, fu, fv = P2[0, 2], P2[1, 2], P2[0, 0], P2[1, 1]
    tx, ty = P2[0, 3] / (-fu), P2[1, 3] / (-fv)
    r0e = torch.cat((torch.cat((R0, R0.new_zeros((3, 1))), dim=1), R0.new_zeros((1, 4))), dim=0)
    r0e[3, 3] = 1
    v2ce = torch.cat((V2C, V2C.new_zeros((1, 4))), dim=0)
    v2ce[3, 3] = 1
    minv = torch.inverse(torch.matmul(r0e, v2ce).T)
    out = np.zeros(48, np.float32)
    out[0:12] = m1.reshape(-1).numpy()
    out[12:24] = P2.T.reshape(-1).numpy()
    out[24:30] = torch.stack((cu, cv, fu, fv, tx, ty)).numpy()
    out[32:48] = minv.reshape(-1).numpy()
    return out


def seek_frame_kitti(points, K, dets, params, tables=None, keep_intermediates=False):
    """One frame of FrustumProposerOGKITTI.get_proposals (frustum_proposals_v1_kitti.py:292-690) restated on the C
    oracle.  points (N,>=3) xyz first; K (48,) calibration block (kitti_block, or the engine's); dets = (boxes x-y-w-h,
    labels 1..7, scores).  Differences from seek_frame: CalibrationTorch's projection without an on-image test, ONE
    batched first-match points_in_boxes_gpu per frustum (:644-648), densities over their sum, score = dns_w + density +
    iou_w iou + dst_w dists_ranked (:650-654), nms_normal + topk (:657-671)."""
    p = dict(DEFAULTS); p.update(max_dist=70); p.update(params)
    assert p["nms_3d"] == 0 and not p.get("rand_center")
    topk = int(p["topk"])
    dns_w, iou_w, dst_w = np.float32(p["dns_w"]), np.float32(p["iou_w"]), np.float32(p["dst_w"])
    sdepth = p.get("search_depth")
    base_boxes, base_corners = tables if tables is not None else build_tables_kitti(p)
    mags = torch.linspace(0.0, 1.0, p["num_mags"]).numpy() if p["num_mags"] > 0 else np.zeros(1, np.float32)
    max_dist = np.float32(p["max_dist"])
    pts = np.ascontiguousarray(points, np.float32)
    boxes2d, labels, scores = dets
    cands = nms2d_candidates(boxes2d, labels, scores, np.zeros(len(scores), np.int64), p["nms_2d"], p["score_thr"], "xywh")
    boxes_out, labels_out, scores_out, inter = [], [], [], []
    for (c, box2d, label, score) in cands:
        idx, uvd, xyz = O.frustum_cull_kitti(pts, K, box2d)
        rec = dict(cam=c, box2d=box2d, label=label, score=score, n_points=int(idx.shape[0]))
        if idx.shape[0] == 0:                       # :399-401
            if keep_intermediates:
                inter.append(rec)
            continue
        d = uvd[:, 2]
        qmin = O.quantile(d, p["lq"])
        qmax = O.quantile(d, p["uq"]) if sdepth is None else np.float32(qmin + np.float32(sdepth))
        dmin = np.maximum(qmin, FRUSTUM_MIN)
        dmax = np.minimum(qmax, max_dist)
        # weighted_centre_xyz (:392-395): one row through rect_to_lidar -- the small-matrix rounding
        wc = O.unproject_kitti(K, (box2d[0] + box2d[2]) / np.float32(2), (box2d[1] + box2d[3]) / np.float32(2),
                               O.quantile(d, p["cq"]), chain=False)
        centres, corners = O.centre_line_kitti(box2d, dmin, dmax, K, xyz.min(0), xyz.max(0), p["clamp_bottom"], mags,
                                               search_depth=sdepth)
        hb, iou, valid, near, dist = O.hypotheses_kitti(base_boxes[label - 1], base_corners[label - 1], centres, K, box2d,
                                                        max_dist, p["min_cam_iou"], wc)
        counts = np.zeros(hb.shape[0], np.int32)
        sc = np.zeros(hb.shape[0], np.float32)
        vi = np.flatnonzero(valid)
        order = np.zeros(0, np.int64)
        if vi.size:
            first = O.points_in_boxes_gpu(xyz[None], hb[vi][None])[0]
            counts[vi] = np.bincount(first[first >= 0], minlength=vi.size)[:vi.size]
            cnt = counts[vi].astype(np.float32)
            dens = cnt / np.float32(cnt.sum(dtype=np.float32) + np.float32(1e-8))
            dn = dist[near]                          # dists_ranked is normalised over the hypotheses within max_dist (:620-622)
            dr = np.float32(1) - (dist[vi] - dn.min()) / np.float32((dn.max() - dn.min()) + np.float32(1e-8))
            sc[vi] = ((dns_w + dens) + iou_w * iou[vi]) + dr * dst_w
            order = vi[O.nms_normal(hb[vi], sc[vi], p["nms_normal"])[:topk]]
        rec.update(idx=idx, uvd=uvd, xyz=xyz, dmin=dmin, dmax=dmax, wc=wc, centres=centres, corners=corners, hyp_boxes=hb,
                   iou=iou, valid=valid, near=near, dist=dist, counts=counts, scores=sc, topk=order,
                   best=int(order[0]) if order.size else -1)
        if keep_intermediates:
            inter.append(rec)
        for hh in order:
            boxes_out.append(hb[hh]); labels_out.append(label); scores_out.append(score)
    return dict(pred_boxes=np.asarray(boxes_out, np.float32).reshape(-1, 7), pred_labels=np.asarray(labels_out, np.int32),
                pred_scores=np.asarray(scores_out, np.float32), frustums=inter)


def recall_record(pred_boxes, gt_boxes, thresh_list=(0.3, 0.5, 0.7)):
    """Detector3DTemplate.generate_recall_record restated
    (pcdet/models/detectors/detector3d_template.py:315-399) for one frame; returns the
    counters as a dict of ints."""
    known3 = {1, 8, 9}            # car, bicycle, pedestrian          (:17,21)
    known6 = {1, 3, 5, 6, 8, 9}   # + construction_vehicle, trailer, barrier (:18-22)
    rd = {"gt": 0, "num_3known": 0, "num_6known": 0, "num_4unknown": 0, "num_7unknown": 0}
    for t in thresh_list:
        for k in ("rcnn", "rcnn_3known", "rcnn_6known", "rcnn_4unknown", "rcnn_7unknown"):
            rd["%s_%s" % (k, t)] = 0
    gt = np.asarray(gt_boxes, np.float32)
    k = gt.shape[0] - 1
    while k >= 0 and gt[k].sum() == 0:
        k -= 1
    gt = gt[:k + 1]
    if gt.shape[0] == 0:
        return rd
    labels = gt[:, -1].astype(np.int64)
    k3 = np.array([l in known3 for l in labels])
    k6 = np.array([l in known6 for l in labels])
    rd["num_3known"] += int(k3.sum()); rd["num_6known"] += int(k6.sum())
    rd["num_7unknown"] += int((~k3).sum()); rd["num_4unknown"] += int((~k6).sum())
    if pred_boxes.shape[0] > 0:
        iou = O.boxes_iou3d(pred_boxes[:, :7], gt[:, :7])
        mx = iou.max(0)
        for t in thresh_list:
            hit = mx > np.float32(t)
            rd["rcnn_%s" % t] += int(hit.sum())
            rd["rcnn_3known_%s" % t] += int((hit & k3).sum())
            rd["rcnn_6known_%s" % t] += int((hit & k6).sum())
            rd["rcnn_7unknown_%s" % t] += int((hit & ~k3).sum())
            rd["rcnn_4unknown_%s" % t] += int((hit & ~k6).sum())
    rd["gt"] += int(gt.shape[0])
    return rd
