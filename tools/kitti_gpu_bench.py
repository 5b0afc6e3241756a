"""Timing of the KITTI single-camera head on the GPU box: the reference's own FrustumProposerOGKITTI.get_proposals
(source unmodified, its kernels compiled for sm_100a -- oracle/_ref) against proposer.FrustumProposerOGKITTI (fused stages,
FNP_VARIANT_KITTI) at the reference's operating point (batch size 1) and the batched engine.  KITTI-shaped synthetic
frames (synth.make_kitti_frame: ~11k points in front of the sensor, ~8 2D boxes), the head's constructor defaults.
Writes gpurun_out/r02y_kitti_gpu.json.  Run from the repo root on a GPU box."""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import ref_seeker as ref  # noqa: E402
from findnpropagate_b200 import proposer, synth  # noqa: E402
from findnpropagate_b200.seeker import KittiFrameInput, SeekerEngine  # noqa: E402

mod = ref.load("cuda", head_file="frustum_proposals_v1_kitti.py")
Calibration = sys.modules["pcdet.utils.calibration_kitti"].Calibration
N = 16
raw = [synth.make_kitti_frame(i) for i in range(N)]
state = {}


class Feeder:
    def __call__(self, bd):
        pts, calib, boxes, labels, scores = state["frame"]
        z = torch.zeros(len(boxes), dtype=torch.long)
        return torch.from_numpy(boxes.copy()), torch.from_numpy(labels), torch.from_numpy(scores), z, z.clone()


mod.PreprocessedDetector = lambda paths, class_names=None: Feeder()
params = dict(nms_3d=0.0, score_thr=0.45, nms_2d=0.4)          # the constructor defaults otherwise (H = 6 x 10 x 4 = 240)
with contextlib.redirect_stdout(io.StringIO()):
    head = mod.FrustumProposerOGKITTI(model_cfg=ref.AttrDict(PARAMS=params, PREDS_PATH="unused.json"), class_names=None)
head.eval()
ours = proposer.FrustumProposerOGKITTI(model_cfg=dict(PARAMS=params), image_detector=Feeder(), device="cuda:0")
bds = []
for fr in raw:
    pts = torch.from_numpy(np.c_[np.zeros(len(fr[0]), np.float32), fr[0]]).cuda()
    bds.append(dict(batch_size=1, calib=[Calibration(fr[1])], points=pts))


def timed(fn, reps):
    for i in range(2):
        state["frame"] = raw[i]
        fn(bds[i])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    k = 0
    for r in range(reps):
        for i in range(N):
            state["frame"] = raw[i]
            k += int(fn(bds[i])[0].shape[0])
            torch.cuda.synchronize()
    return (time.perf_counter() - t0) / (reps * N), k // reps


with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
    t_ref, k_ref = timed(head.get_proposals, 1)
t_ours, k_ours = timed(ours.get_proposals, 5)
eng = SeekerEngine(params, device="cuda:0", box_format="xywh", variant="kitti")
fis = [KittiFrameInput(points=f[0], P2=f[1]["P2"], R0=f[1]["R0"], V2C=f[1]["Tr_velo2cam"], det_boxes=f[2], det_labels=f[3],
                       det_scores=f[4], device="cuda:0") for f in raw] * 16            # 256 frames per batch
for f in fis[:N]:
    f.prepare()
plan = eng.plan(fis)
pts = eng.upload_points(fis)
for _ in range(3):
    eng.finish(eng.execute(plan, pts))
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    res = eng.finish(eng.execute(plan, pts))
torch.cuda.synchronize()
t_batch = (time.perf_counter() - t0) / (10 * len(fis))
out = dict(frames=N, points_per_frame=int(np.mean([len(f[0]) for f in raw])), boxes_2d_per_frame=float(np.mean([len(f[2]) for f in raw])),
           hypotheses_per_frustum=240, proposals={"reference": k_ref, "ours": k_ours},
           reference_gpu_ms_per_frame=1e3 * t_ref, drop_in_head_bs1_ms_per_frame=1e3 * t_ours,
           batched_engine_256_frames_ms_per_frame=1e3 * t_batch,
           frames_per_s=dict(reference_gpu=1 / t_ref, drop_in_head_bs1=1 / t_ours, batched_engine=1 / t_batch))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02y_kitti_gpu.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
