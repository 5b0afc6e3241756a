"""One small batch through the pipeline with the depth-sweep scoring kernel forced on, compared
with the direct kernel (used under compute-sanitizer racecheck / synccheck)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from findnpropagate_b200 import synth  # noqa: E402
from findnpropagate_b200.seeker import FrameInput, SeekerEngine  # noqa: E402

# argv[1] = "cfg2": one full-size frame (frustums of thousands of points: the warp queues fill and drain in place
# several times per item); default: two cfg1 frames with a deeper grid, split into 256-point items
big = len(sys.argv) > 1 and sys.argv[1] == "cfg2"
cfg = synth.CONFIGS["cfg2" if big else "cfg1"]
params = synth.seeker_params(cfg) if big else dict(synth.seeker_params(cfg), num_mags=24)
frames = []
for i in range(1 if big else 2):
    f = synth.make_frame(i, cfg)
    frames.append(FrameInput(points=f.points, lidar2image=f.lidar2image, camera2lidar=f.camera2lidar,
                             camera_intrinsics=f.camera_intrinsics, det_boxes=f.det_boxes, det_labels=f.det_labels,
                             det_scores=f.det_scores, det_cam_idx=f.det_cam_idx, gt_boxes=f.gt_boxes))
out = {}
for mode in ("direct", "sweep"):
    eng = SeekerEngine(params, device="cuda:0", score_mode=mode, split_points=None if big else 256)
    out[mode] = eng.run(frames, nms_thresh=0.1, with_recall=True)
    assert eng.last_score_mode == mode
v = out["direct"]["cand_valid"]
assert np.array_equal(out["direct"]["cand_count"][v], out["sweep"]["cand_count"][v])
assert np.array_equal(out["direct"]["cand_best"], out["sweep"]["cand_best"])
print("sweep smoke ok:", int(v.sum()), "proposals, counts equal to the direct kernel")
