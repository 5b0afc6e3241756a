"""TEST INFRASTRUCTURE ONLY -- tests/golden/kitti_*.npz from the reference's own KITTI head.

Run in the build container (needs /root/reference):   python tools/gen_golden_kitti.py

Imports FrustumProposerOGKITTI (pcdet/models/dense_heads/frustum_proposals_v1_kitti.py) through the namespace stubs of
tools/ref_seeker.py in its CPU mode (device strings patched, the two native ops emulated by the oracle: batched
first-match points_in_boxes_gpu, nms_normal with a stable sort) and runs get_proposals on KITTI-shaped synthetic
frames (findnpropagate_b200.synth.make_kitti_frame).  Stored per frame: the inputs, the head's outputs, its prior
tables, and per frustum what passed through the two native call sites -- the unprojected frustum points, the valid
hypothesis boxes, the first-match index of every point, the second-stage scores and the keep order.
CPU torch rounds the calibration matmuls (sgemm of MKL) and cdist differently from the GPU in the last ulp; the GPU
tests compare at 1e-5 (boxes) and report what is bit-equal.
"""
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)

import ref_seeker  # noqa: E402
from findnpropagate_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

PARAM_SETS = {
    # the option set of tests/test_reference_gpu.py (deep enough grid, two proposals per frustum)
    "a": dict(lq=0.0, uq=0.25, cq=1.0, iou_w=1.0, nms_normal=1.0, dst_w=0.2, dns_w=1.0, min_cam_iou=0.1, score_thr=0.45,
              nms_2d=0.4, nms_3d=0.0, clamp_bottom=1, num_sizes=1, num_mags=8, num_rotations=6, topk=2),
    # the constructor's own defaults (frustum_proposals_v1_kitti.py:41-44) with the 2D thresholds of the frames
    "b": dict(nms_3d=0.0, score_thr=0.45, nms_2d=0.4),
}


def main():
    mod = ref_seeker.load("cpu", head_file="frustum_proposals_v1_kitti.py")
    Calibration = sys.modules["pcdet.utils.calibration_kitti"].Calibration
    state = {}

    class Feeder:
        def __call__(self, bd):
            pts, calib, boxes, labels, scores = state["frame"]
            z = torch.zeros(len(boxes), dtype=torch.long)
            return torch.from_numpy(boxes.copy()), torch.from_numpy(labels), torch.from_numpy(scores), z, z.clone()
    mod.PreprocessedDetector = lambda paths, class_names=None: Feeder()
    rp, iu = mod.roiaware_pool3d_utils, mod.iou3d_nms_utils
    real_pib, real_nms = rp.points_in_boxes_gpu, iu.nms_normal_gpu
    cap = []

    def pib(points, boxes):
        out = real_pib(points, boxes)
        cap.append(dict(points=points.detach().cpu().numpy().reshape(-1, 3).copy(),
                        boxes=boxes.detach().cpu().numpy().reshape(-1, 7).copy(), first=out.detach().cpu().numpy().reshape(-1).copy()))
        return out

    def nms(boxes, scores, thresh, **kw):
        keep, aux = real_nms(boxes, scores, thresh, **kw)
        cap[-1].update(scores=scores.detach().cpu().numpy().copy(), keep=keep.detach().cpu().numpy().copy())
        return keep, aux
    rp.points_in_boxes_gpu, iu.nms_normal_gpu = pib, nms
    os.makedirs(OUT, exist_ok=True)
    for tag, params in PARAM_SETS.items():
        with contextlib.redirect_stdout(io.StringIO()):
            head = mod.FrustumProposerOGKITTI(model_cfg=ref_seeker.AttrDict(PARAMS=params, PREDS_PATH="unused.json"), class_names=None)
        head.eval()
        for index in (0, 1, 2):
            fr = synth.make_kitti_frame(index)
            state["frame"] = fr
            del cap[:]
            bd = dict(batch_size=1, calib=[Calibration(fr[1])],
                      points=torch.from_numpy(np.c_[np.zeros(len(fr[0]), np.float32), fr[0]]))
            with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
                boxes, labels, scores, bidx = head.get_proposals(bd)
            d = dict(index=index, params=json.dumps(params), points=fr[0], P2=fr[1]["P2"], R0=fr[1]["R0"], V2C=fr[1]["Tr_velo2cam"],
                     det_boxes=fr[2], det_labels=fr[3], det_scores=fr[4],
                     ref_boxes=boxes.cpu().numpy().astype(np.float32), ref_labels=labels.cpu().numpy().astype(np.int32),
                     ref_scores=scores.cpu().numpy().astype(np.float32),
                     base_boxes=head.base_boxes.cpu().numpy(), base_corners=head.base_corners.cpu().numpy(), n_frustums=len(cap))
            for k, c in enumerate(cap):
                for key in ("points", "boxes", "first", "scores", "keep"):
                    d["f%d_%s" % (k, key)] = c[key]
            path = os.path.join(OUT, "kitti_%s_%d.npz" % (tag, index))
            np.savez_compressed(path, **d)
            print("wrote", path, "K =", boxes.shape[0], "frustums scored =", len(cap), "bytes", os.path.getsize(path))


if __name__ == "__main__":
    main()
