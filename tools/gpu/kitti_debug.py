"""Debug aid for tests/test_reference_gpu.py::test_kitti_head_runs_on_the_drop_in_ops: runs the reference's KITTI
head on the reference ops twice (is the head itself deterministic?) and once on the drop-in ops, comparing every
native call (same inputs through both libraries) and the final outputs.  Run on the GPU box from the repo root."""
import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ref_seeker as ref  # noqa: E402
from test_reference_gpu import _kitti_frame  # noqa: E402
from findnpropagate_b200.pcdet_ops import iou3d_nms_utils as our_iou, roiaware_pool3d_utils as our_rp  # noqa: E402

mod = ref.load("cuda", head_file="frustum_proposals_v1_kitti.py")
Calibration = sys.modules["pcdet.utils.calibration_kitti"].Calibration
state = {}


class Feeder:
    def __call__(self, bd):
        pts, calib, boxes, labels, scores = state["frame"]
        z = torch.zeros(len(boxes), dtype=torch.long)
        return torch.from_numpy(boxes.copy()), torch.from_numpy(labels), torch.from_numpy(scores), z, z.clone()


mod.PreprocessedDetector = lambda paths, class_names=None: Feeder()
params = dict(lq=0.0, uq=0.25, cq=1.0, iou_w=1.0, nms_normal=1.0, dst_w=0.2, dns_w=1.0, min_cam_iou=0.1, score_thr=0.45,
              nms_2d=0.4, nms_3d=0.0, clamp_bottom=1, num_sizes=1, num_mags=8, num_rotations=6, topk=2)
with contextlib.redirect_stdout(io.StringIO()):
    head = mod.FrustumProposerOGKITTI(model_cfg=ref.AttrDict(PARAMS=params, PREDS_PATH="unused.json"), class_names=None)
head.eval()
ref_rp, ref_iou = mod.roiaware_pool3d_utils, mod.iou3d_nms_utils
log = []


class BothRP:
    @staticmethod
    def points_in_boxes_gpu(points, boxes):
        a = ref_rp.points_in_boxes_gpu(points, boxes)
        b = our_rp.points_in_boxes_gpu(points, boxes)
        if not torch.equal(a, b):
            d = (a != b).nonzero()
            print("  points_in_boxes_gpu differs: %d of %d points; first" % (len(d), a.numel()), d[:5].tolist(),
                  a[a != b][:5].tolist(), b[a != b][:5].tolist(), "boxes", tuple(boxes.shape), boxes.dtype, boxes.is_contiguous(),
                  "points", tuple(points.shape), points.dtype, points.is_contiguous())
        log.append(("pib", a.clone(), b.clone()))
        return a


class BothIoU:
    @staticmethod
    def nms_normal_gpu(boxes, scores, thresh, **kw):
        a = ref_iou.nms_normal_gpu(boxes, scores, thresh, **kw)
        b = our_iou.nms_normal_gpu(boxes, scores, thresh, **kw)
        same = all(torch.equal(x.cpu(), y.cpu()) if isinstance(x, torch.Tensor) else x == y for x, y in zip(a, b))
        if not same:
            print("  nms_normal_gpu differs:", [x for x in a], [y for y in b], "boxes", tuple(boxes.shape), boxes.device,
                  "scores", scores.device, scores.dtype, scores.tolist())
        return a


def run(fr, ops):
    mod.roiaware_pool3d_utils, mod.iou3d_nms_utils = ops
    bd = dict(batch_size=1, calib=[Calibration(fr[1])],
              points=torch.from_numpy(np.c_[np.zeros(len(fr[0]), np.float32), fr[0]]).cuda())
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        return [o.cpu() for o in head.get_proposals(bd)]


trace = []


def tracing(rp, iou):
    class RP:
        @staticmethod
        def points_in_boxes_gpu(points, boxes):
            out = rp.points_in_boxes_gpu(points, boxes)
            trace.append(("pib", points.detach().cpu().clone(), boxes.detach().cpu().clone(), out.detach().cpu().clone()))
            return out

    class IoU:
        @staticmethod
        def nms_normal_gpu(boxes, scores, thresh, **kw):
            out = iou.nms_normal_gpu(boxes, scores, thresh, **kw)
            trace.append(("nms", boxes.detach().cpu().clone(), scores.detach().cpu().clone(), out[0].detach().cpu().clone()))
            return out
    return RP, IoU


OursRP, OursIoU = tracing(our_rp, our_iou)
RefRP, RefIoU = tracing(ref_rp, ref_iou)


for i in range(3):
    fr = _kitti_frame(i)
    state["frame"] = fr
    print("frame", i, "points", len(fr[0]), "boxes", len(fr[2]))
    r1 = run(fr, (ref_rp, ref_iou))
    r2 = run(fr, (ref_rp, ref_iou))
    rb = run(fr, (BothRP, BothIoU))
    del trace[:]
    rt = run(fr, (RefRP, RefIoU))
    t_ref = list(trace)
    del trace[:]
    ro = run(fr, (OursRP, OursIoU))
    t_our = list(trace)
    print("  traced calls: ref %d, ours %d" % (len(t_ref), len(t_our)))
    for k, (ca, cb) in enumerate(zip(t_ref, t_our)):
        names = ("kind", "in0", "in1", "out")
        bad = [names[q] for q in range(1, 4) if ca[q].shape != cb[q].shape or not torch.equal(ca[q], cb[q])]
        if ca[0] != cb[0] or bad:
            print("  first differing call: #%d %s/%s differs in %s" % (k, ca[0], cb[0], bad))
            for q in range(1, 4):
                if names[q] in bad:
                    x, y = ca[q], cb[q]
                    print("   ", names[q], tuple(x.shape), tuple(y.shape), x.dtype, y.dtype)
                    if x.shape == y.shape:
                        d = (x != y).nonzero()
                        print("    where", d[:8].tolist(), "ref", x[x != y][:8].tolist(), "ours", y[x != y][:8].tolist())
                    if ca[0] == "nms":
                        sc = ca[2]
                        print("    scores (ref trajectory) top:", sc.sort(descending=True)[0][:6].tolist(), "order ref", x[:6].tolist(), "ours", y[:6].tolist())
            break
    for name, x in (("ref again", r2), ("both (ref results returned)", rb), ("ours", ro)):
        for k, (a, b) in enumerate(zip(r1, x)):
            if a.shape != b.shape:
                print("  %s: output %d shape %s vs %s" % (name, k, tuple(a.shape), tuple(b.shape)))
            elif not torch.equal(a, b):
                af, bf = a.double(), b.double()
                bad = (af != bf) & ~(torch.isnan(af) & torch.isnan(bf))
                print("  %s: output %d differs in %d of %d entries (nan %d / %d), max abs diff %.3g, where %s" % (
                    name, k, int(bad.sum()), a.numel(), int(torch.isnan(af).sum()), int(torch.isnan(bf).sum()),
                    float((af - bf)[bad].abs().max()) if bad.any() else 0.0, bad.nonzero()[:6].tolist()))
            else:
                print("  %s: output %d equal" % (name, k))
