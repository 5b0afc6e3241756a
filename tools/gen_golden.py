"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the reference itself.

Run in the build container (needs /root/reference):   python tools/gen_golden.py

For each fixture frame it stores the synthetic inputs (so the fixture does not depend on
the generator staying bit-stable across hosts), the outputs of the reference's own
FrustumProposerOG.get_proposals (tools/ref_seeker.py documents the CPU run and its
deviations) and the per-frustum intermediates captured at the two native-op call sites
(frustum points, valid hypothesis boxes, per-hypothesis counts, second-stage scores,
keep order).  It also stores op-level vectors produced by the *compiled reference CPU
ops* (points_in_boxes_cpu, boxes_iou_bev_cpu from oracle/_ref).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)

import build_ref  # noqa: E402
import ref_seeker  # noqa: E402
from findnpropagate_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


# option sets of SURVEY.md 8 row f3 (PARAMS keys and the MULT / OCCL_MULT / MULTICAM_IOU switches of
# the head's model_cfg), each pinned by a reference run on a tiny frame
OPTION_SETS = {
    "dst": dict(dst_w=0.226, iou_w=0.95, dns_w=0.05),          # the constructor's own defaults (:146)
    "ego": dict(ego_w=0.3),
    "mult": dict(MULT=True, dst_w=0.7, iou_w=0.9),
    "occl": dict(occl_w=0.5),
    "occlmult": dict(OCCL_MULT=True),
    "multicam": dict(MULTICAM_IOU=True),
    "sdepth": dict(search_depth=4.0),
    "all": dict(dst_w=0.2, ego_w=0.1, occl_w=0.3, MULTICAM_IOU=True, search_depth=6.0),
    "topk": dict(topk=3, nms_normal=0.5),                      # :1030-1046: NMS over the hypotheses, first 3 survivors
    "topkdst": dict(topk=2, nms_normal=0.7, dst_w=0.226, iou_w=0.95, dns_w=0.05),
}


def frame_fixture(name, index, box_format="xyxy", opts=None):
    import json
    cfg = synth.CONFIGS[name]
    params = synth.seeker_params(cfg)
    if opts is not None:
        params.update(OPTION_SETS[opts])
    fr = synth.make_frame(index, cfg)
    if box_format != "xyxy":     # BOX_FORMAT 'xywh' (frustum_proposals_v1.py:597-601): the feeder hands out x, y, w, h
        fr.det_boxes = fr.det_boxes.copy()
        fr.det_boxes[:, 2:] = fr.det_boxes[:, 2:] - fr.det_boxes[:, :2]
    boxes, labels, scores, bidx, cap, head = ref_seeker.run([fr], params, box_format=box_format)
    d = dict(
        cfg=name, index=index, box_format=box_format,
        points=fr.points, lidar2image=fr.lidar2image, camera_intrinsics=fr.camera_intrinsics,
        camera2lidar=fr.camera2lidar, lidar_aug_matrix=fr.lidar_aug_matrix, gt_boxes=fr.gt_boxes,
        det_boxes=fr.det_boxes, det_labels=fr.det_labels, det_scores=fr.det_scores,
        det_cam_idx=fr.det_cam_idx,
        ref_boxes=boxes, ref_labels=labels.astype(np.int32), ref_scores=scores,
        base_boxes=head.base_boxes.numpy(), base_corners=head.base_corners.numpy(),
        n_frustums=len(cap["frustums"]),
        opts=json.dumps(OPTION_SETS[opts] if opts is not None else {}),
    )
    for k, f in enumerate(cap["frustums"]):
        d["f%d_points" % k] = f["points"]
        d["f%d_boxes" % k] = f["boxes"]
        d["f%d_counts" % k] = f["counts"]
        d["f%d_scores" % k] = f["scores"]
        d["f%d_keep" % k] = f["keep"]
    # reference recall record (detector3d_template.py:315) evaluated with the oracle iou3d
    tag = ("" if box_format == "xyxy" else "_" + box_format) + ("" if opts is None else "_opt-" + opts)
    path = os.path.join(OUT, "seeker_%s%s_%d.npz" % (name, tag, index))
    np.savez_compressed(path, **d)
    print("wrote", path, "K =", boxes.shape[0], "frustums =", len(cap["frustums"]))


def op_fixture():
    """Known-answer vectors from the reference's compiled CPU ops."""
    rp = build_ref.load("roiaware_pool3d_cuda")
    iou = build_ref.load("iou3d_nms_cuda")
    rng = np.random.default_rng(7)
    n_box, n_pts = 48, 4000
    boxes = np.zeros((n_box, 7), np.float32)
    boxes[:, 0:3] = rng.uniform(-6, 6, (n_box, 3)) * [1, 1, 0.2]
    boxes[:, 3:6] = synth.PRIORS[rng.integers(0, 10, n_box)] * rng.uniform(0.8, 1.2, (n_box, 3))
    boxes[:, 6] = rng.uniform(-4, 4, n_box)
    boxes[::7, 6] = 0.0
    pts = (rng.uniform(-9, 9, (n_pts, 3)) * [1, 1, 0.25]).astype(np.float32)
    out = torch.zeros((n_box, n_pts), dtype=torch.int32)
    rp.points_in_boxes_cpu(torch.from_numpy(boxes), torch.from_numpy(pts), out)
    a, b = boxes[:24].copy(), boxes[24:].copy()
    b[:6] = a[:6]                      # identical boxes
    b[6:10, :2] = a[6:10, :2] + 0.05   # heavy overlap
    ans = torch.zeros((24, 24))
    iou.boxes_iou_bev_cpu(torch.from_numpy(a), torch.from_numpy(b), ans)
    path = os.path.join(OUT, "ops_cpu_reference.npz")
    np.savez_compressed(path, boxes=boxes, pts=pts, pib_cpu=out.numpy(), iou_a=a, iou_b=b,
                        iou_bev_cpu=ans.numpy())
    print("wrote", path, "inside:", int(out.sum()))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    op_fixture()
    for name, idxs in (("tiny", (0, 1, 2)), ("cfg1", (0,))):
        for i in idxs:
            frame_fixture(name, i)
    frame_fixture("tiny", 3, box_format="xywh")
    for k, o in enumerate(OPTION_SETS):
        frame_fixture("tiny", k % 3, opts=o)
