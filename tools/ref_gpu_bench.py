"""B-REF-GPU / B-REF-OPS: the reference on the B200, timed next to this build (GPU box only).

    python tools/ref_gpu_bench.py --out profiles/r02_reference_gpu.json

Table 1 (B-REF-GPU): sync-bracketed wall time per frame of the reference's own
``FrustumProposerOG.get_proposals`` (frustum_proposals_v1.py:523-1067; source unmodified, its own
kernels compiled for sm_100a -- tools/ref_seeker.py) at its operating point, batch size 1
(tools/extract_pseudo_labels.py:36), against the drop-in head ``proposer.FrustumProposerOG`` at batch
size 1 and the batched engine, on the same synthetic frames (cfg1 = shipped-YAML grid, cfg2).

Table 2 (B-REF-OPS): the reference's compiled ops (oracle/_ref) against fnp_* through the same-name
Python wrappers, CUDA events, N in {60, 200, 768, 3072}: points_in_boxes_gpu as the seeker calls it
(one launch + sum + D2H per hypothesis, frustum_proposals_v1.py:930-932) and as one batched call,
nms_gpu, nms_normal_gpu, boxes_iou_bev, boxes_iou3d_gpu.

Reads nothing under /root/reference; TEST/BENCH INFRASTRUCTURE, not product.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

from findnpropagate_b200 import synth  # noqa: E402


def _sync_wall(fn, iters, warmup=1):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        torch.cuda.synchronize()
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t)
    return 1e3 * float(np.median(ts)), 1e3 * float(np.min(ts))


def _events(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(iters):
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def head_table(ref_seeker, cfg_name, n_frames, ref_frames):
    from findnpropagate_b200 import proposer
    from findnpropagate_b200.seeker import FrameInput, SeekerEngine
    cfg = synth.CONFIGS[cfg_name]
    params = synth.seeker_params(cfg)
    frames = [synth.make_frame(i, cfg) for i in range(n_frames)]
    row = dict(config=cfg_name, frames=n_frames, hypotheses_per_frustum=params["num_mags"] * params["num_rotations"] * params["num_sizes"],
               points_per_frame=int(np.mean([f.points.shape[0] for f in frames])))
    # --- reference head, batch size 1
    head = None
    ts, K = [], 0
    for fr in frames[:ref_frames]:
        bd = ref_seeker.batch_dict([fr], "cuda")
        if head is None:
            head = ref_seeker.build_head(params, [fr], device="cuda")
        head.image_detector = ref_seeker.SyntheticFeeder([fr])
        torch.cuda.synchronize()
        t = time.perf_counter()
        with torch.no_grad():
            out = head.get_proposals(bd)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t)
        K += int(out[0].shape[0])
    row["reference_gpu_ms_per_frame"] = 1e3 * float(np.median(ts))
    row["reference_gpu_frames_per_s"] = 1.0 / float(np.median(ts))
    row["reference_gpu_frames_timed"] = len(ts)
    row["reference_gpu_proposals"] = K
    # --- drop-in head, batch size 1 (the reference's operating point), device-resident batch_dict
    ours = proposer.FrustumProposerOG(model_cfg=dict(PARAMS=params), image_detector=proposer.SyntheticGLIP(frames[:1]),
                                      device="cuda:0")
    ours.eval()
    bds = []
    for fr in frames:
        bd = synth.collate([fr])
        for k, v in list(bd.items()):
            if isinstance(v, np.ndarray) and v.dtype.kind == "f":
                bd[k] = torch.from_numpy(v).float().cuda()
        bds.append(bd)
    state = {"i": 0}

    def one():
        i = state["i"] % n_frames
        state["i"] += 1
        ours.image_detector = proposer.SyntheticGLIP([frames[i]])
        ours.forward(bds[i])
    med, mn = _sync_wall(one, iters=max(3 * n_frames, 12), warmup=n_frames)
    row["dropin_head_bs1_ms_per_frame"] = med
    row["dropin_head_bs1_ms_per_frame_min"] = mn
    row["dropin_head_bs1_frames_per_s"] = 1e3 / med
    # --- host-resident batch_dict (collate output before load_data_to_gpu)
    bds_h = [synth.collate([fr]) for fr in frames]

    def one_h():
        i = state["i"] % n_frames
        state["i"] += 1
        ours.image_detector = proposer.SyntheticGLIP([frames[i]])
        ours.forward(bds_h[i])
    med, mn = _sync_wall(one_h, iters=max(3 * n_frames, 12), warmup=n_frames)
    row["dropin_head_bs1_host_input_ms_per_frame"] = med
    # --- batched engine, all frames in one call, points resident
    eng = SeekerEngine(params, device="cuda:0")
    fis = [FrameInput(points=f.points, lidar2image=f.lidar2image, camera2lidar=f.camera2lidar,
                      camera_intrinsics=f.camera_intrinsics, det_boxes=f.det_boxes, det_labels=f.det_labels,
                      det_scores=f.det_scores, det_cam_idx=f.det_cam_idx, gt_boxes=f.gt_boxes) for f in frames]
    pts = eng.upload_points(fis)

    def batched():
        eng.finish(eng.execute(eng.plan(fis), pts))
    med, mn = _sync_wall(batched, iters=10, warmup=2)
    row["engine_batched_ms_per_frame"] = med / n_frames
    row["speedup_dropin_bs1_over_reference_gpu"] = row["reference_gpu_ms_per_frame"] / row["dropin_head_bs1_ms_per_frame"]
    row["speedup_engine_batched_over_reference_gpu"] = row["reference_gpu_ms_per_frame"] / row["engine_batched_ms_per_frame"]
    return row


def ops_table(ref_rp, ref_iou, ref_rp_cuda, ref_iou_cuda):
    from findnpropagate_b200.pcdet_ops import iou3d_nms_utils as our_iou
    from findnpropagate_b200.pcdet_ops import roiaware_pool3d_utils as our_rp
    rng = np.random.RandomState(0)
    rows = []
    P = 5000                                               # frustum points (cfg2 mean ~5k)
    pts = torch.from_numpy((rng.rand(1, P, 3) * [8, 8, 3] + [20, -4, -1.5]).astype(np.float32)).cuda()
    for N in (60, 200, 768, 3072):
        b = np.zeros((N, 7), np.float32)
        b[:, 0:3] = rng.rand(N, 3) * [8, 8, 2] + [20, -4, -1]
        b[:, 3:6] = rng.rand(N, 3) * [4, 2, 1.5] + [0.5, 0.5, 0.5]
        b[:, 6] = rng.rand(N) * np.pi
        boxes = torch.from_numpy(b).cuda()
        scores = torch.from_numpy(rng.rand(N).astype(np.float32)).cuda()
        r = dict(N=N, frustum_points=P)

        def loop(mod):
            def f():
                num = torch.zeros(N)
                for i in range(N):         # frustum_proposals_v1.py:930-932
                    idx = mod.points_in_boxes_gpu(pts, boxes[[i]].reshape(1, -1, 7))
                    num[i] = (idx >= 0).sum()
                return num
            return f
        t_ref, _ = _sync_wall(loop(ref_rp), iters=3, warmup=1)
        t_our, _ = _sync_wall(loop(our_rp), iters=3, warmup=1)
        assert torch.equal(loop(ref_rp)(), loop(our_rp)())
        r["points_in_boxes_gpu_per_hypothesis_loop_ms"] = dict(reference=t_ref, ours=t_our)
        bb = boxes.reshape(1, N, 7)
        r["points_in_boxes_gpu_batched_ms"] = dict(reference=_events(lambda: ref_rp.points_in_boxes_gpu(pts, bb)),
                                                   ours=_events(lambda: our_rp.points_in_boxes_gpu(pts, bb)))
        assert torch.equal(ref_rp.points_in_boxes_gpu(pts, bb), our_rp.points_in_boxes_gpu(pts, bb))
        # the fused count op the seeker stage replaces the loop with (one launch for all N boxes)
        from findnpropagate_b200 import _lib
        p4 = torch.zeros((P, 4), dtype=torch.float32, device="cuda")
        p4[:, :3] = pts[0]
        ps = torch.tensor([0, P], dtype=torch.int32, device="cuda")
        bs = torch.tensor([0, N], dtype=torch.int32, device="cuda")
        cnt = torch.zeros(N, dtype=torch.int32, device="cuda")
        r["fnp_count_in_boxes_ms"] = _events(lambda: _lib.lib.fnp_count_in_boxes(
            p4.data_ptr(), ps.data_ptr(), boxes.data_ptr(), bs.data_ptr(), 1, cnt.data_ptr(), _lib.current_stream()))
        for name in ("nms_gpu", "nms_normal_gpu"):
            fr_, fo_ = getattr(ref_iou, name), getattr(our_iou, name)
            kr, ko = fr_(boxes, scores, 0.3)[0], fo_(boxes, scores, 0.3)[0]
            assert torch.equal(kr.cpu(), ko.cpu()), name
            tr, _ = _sync_wall(lambda: fr_(boxes, scores, 0.3), iters=10, warmup=2)
            to, _ = _sync_wall(lambda: fo_(boxes, scores, 0.3), iters=10, warmup=2)
            r[name + "_ms"] = dict(reference=tr, ours=to, kept=int(kr.numel()))
        r["boxes_iou_bev_ms"] = dict(reference=_events(lambda: ref_iou.boxes_iou_bev(boxes, boxes)),
                                     ours=_events(lambda: our_iou.boxes_iou_bev(boxes, boxes)))
        assert torch.equal(ref_iou.boxes_iou_bev(boxes, boxes), our_iou.boxes_iou_bev(boxes, boxes))
        r["boxes_iou3d_gpu_ms"] = dict(reference=_events(lambda: ref_iou.boxes_iou3d_gpu(boxes, boxes)),
                                       ours=_events(lambda: our_iou.boxes_iou3d_gpu(boxes, boxes)))
        d = (ref_iou.boxes_iou3d_gpu(boxes, boxes) - our_iou.boxes_iou3d_gpu(boxes, boxes)).abs().max().item()
        r["boxes_iou3d_gpu_max_abs_diff"] = d
        rows.append(r)
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_reference_gpu.json"))
    ap.add_argument("--cfg1-frames", type=int, default=4)
    ap.add_argument("--cfg2-frames", type=int, default=4)
    ap.add_argument("--cfg2-ref-frames", type=int, default=2)
    a = ap.parse_args()
    import ref_seeker
    ref_seeker.load("cuda")
    ref_rp = sys.modules["pcdet.ops.roiaware_pool3d.roiaware_pool3d_utils"]
    ref_iou = sys.modules["pcdet.ops.iou3d_nms.iou3d_nms_utils"]
    out = dict(device=torch.cuda.get_device_name(0), torch=torch.__version__,
               note="reference = djamahl99/findnpropagate FrustumProposerOG + pcdet.ops compiled for sm_100a "
                    "(oracle/_ref), run unmodified on this GPU; ours = libfnp_sm100.so behind the same-name API")
    out["heads"] = [head_table(ref_seeker, "cfg1", a.cfg1_frames, a.cfg1_frames),
                    head_table(ref_seeker, "cfg2", a.cfg2_frames, a.cfg2_ref_frames)]
    out["ops"] = ops_table(ref_rp, ref_iou, None, None)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
