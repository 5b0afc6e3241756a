"""CPU tests: the oracle against the golden vectors produced by the reference itself
(tools/gen_golden.py): compiled reference CPU ops + the reference's own get_proposals."""
import glob
import json
import os

import numpy as np
import pytest

import oracle as O
import seeker_oracle as SO
from findnpropagate_b200 import synth


def test_points_in_boxes_cpu_port_matches_reference_cpu_op(golden_dir):
    g = np.load(os.path.join(golden_dir, "ops_cpu_reference.npz"))
    out = O.points_in_boxes_cpu(g["pts"], g["boxes"])
    assert out.shape == g["pib_cpu"].shape
    assert np.array_equal(out, g["pib_cpu"])          # bit-exact (same libm, same expressions)


def test_gpu_predicate_is_subset_of_cpu_margin(golden_dir):
    # MARGIN 1e-5 (GPU kernel) vs 1e-2 (CPU op): every GPU-inside point is CPU-inside
    g = np.load(os.path.join(golden_dir, "ops_cpu_reference.npz"))
    for k in range(0, g["boxes"].shape[0], 5):
        idx = O.points_in_boxes_gpu(g["pts"][None], g["boxes"][k][None, None])[0]
        assert np.all(g["pib_cpu"][k][idx >= 0] == 1)
        assert (idx >= 0).sum() >= g["pib_cpu"][k].sum() - 40


def test_iou_bev_against_reference_cpu_op(golden_dir):
    # the CPU op is compiled without fma and with libm trig; the oracle restates the GPU
    # arithmetic, so agreement is to rounding, not bit-exact
    g = np.load(os.path.join(golden_dir, "ops_cpu_reference.npz"))
    iou = O.boxes_iou_bev(g["iou_a"], g["iou_b"])
    assert np.allclose(iou, g["iou_bev_cpu"], rtol=2e-4, atol=2e-5)
    assert np.allclose(np.diag(iou)[:6], 1.0, atol=1e-5)


def test_first_match_semantics():
    rng = np.random.default_rng(3)
    pts = rng.uniform(-3, 3, (1, 500, 3)).astype(np.float32)
    boxes = np.array([[[0, 0, 0, 2, 2, 2, 0.3], [0, 0, 0, 4, 4, 4, 0.0], [9, 9, 9, 1, 1, 1, 0]]], np.float32)
    idx = O.points_in_boxes_gpu(pts, boxes)[0]
    c = O.count_in_boxes(pts[0], boxes[0])
    assert (idx == 0).sum() == c[0] and (idx >= 0).sum() == c[1] and c[2] == 0
    assert set(np.unique(idx)) <= {-1, 0, 1}


def test_nms_normal_and_rotated_basic():
    b = np.array([[0, 0, 0, 2, 2, 1, 0], [0.1, 0, 0, 2, 2, 1, 0], [5, 5, 0, 2, 2, 1, 0.5], [0, 0, 0, 2, 2, 1, np.pi]], np.float32)
    s = np.array([0.9, 0.8, 0.7, 0.9], np.float32)
    assert list(O.nms_normal(b, s, 0.5)) == [0, 2]
    assert list(O.nms_rotated(b, s, 0.5)) == [0, 2]
    assert list(O.nms_normal(b, s, 1.0)) in ([0, 1, 2], [0, 3, 1, 2], [0, 1, 2, 3])  # thresh 1.0 ~ sort only
    assert O.nms_rotated(np.zeros((0, 7), np.float32), np.zeros(0, np.float32), 0.5).shape == (0,)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "seeker_*.npz"))))
def test_seeker_oracle_against_reference_run(path):
    """Whole-frame parity of the restatement with the reference's get_proposals: same K,
    labels, scores; boxes within 1e-5 relative except yaw 0/pi twins;
    per-frustum intermediates: identical frustum points, valid-hypothesis sets and counts."""
    g = np.load(path)
    cfg = synth.CONFIGS[str(g["cfg"])]
    params = synth.seeker_params(cfg)
    opts = json.loads(str(g["opts"])) if "opts" in g else {}      # row f3 option sets (tools/gen_golden.py)
    params.update(opts)
    out = SO.seek_frame(g["points"], g["lidar2image"], g["camera2lidar"], g["camera_intrinsics"],
                        (g["det_boxes"], g["det_labels"], g["det_scores"], g["det_cam_idx"]), params,
                        tables=(g["base_boxes"], g["base_corners"]), keep_intermediates=True,
                        box_format=str(g["box_format"]) if "box_format" in g else "xyxy")
    assert out["pred_boxes"].shape == g["ref_boxes"].shape
    assert np.array_equal(out["pred_labels"], g["ref_labels"])
    assert np.array_equal(out["pred_scores"], g["ref_scores"])
    rel = np.abs(out["pred_boxes"] - g["ref_boxes"]) / np.maximum(np.abs(g["ref_boxes"]), 1e-3)
    twins = 0
    tol = 1e-5 if not opts else 3e-5     # the option terms add torch ops whose CPU rounding differs (cdist, norm)
    for k in range(rel.shape[0]):
        if rel[k].max() > tol:
            assert rel[k, :6].max() <= tol
            assert abs(abs(out["pred_boxes"][k, 6] - g["ref_boxes"][k, 6]) - np.pi) < 1e-5
            twins += 1
    assert twins <= max(1, rel.shape[0] // 5)
    fr = [f for f in out["frustums"] if f["n_points"] > 0 and f.get("best", -1) >= 0]
    assert len(fr) == int(g["n_frustums"])
    for k, f in enumerate(fr):
        # frustum membership: same number of points, in the same order; the unprojected
        # coordinates agree to rounding (torch's CPU matmul accumulates in another order)
        assert f["xyz"].shape == g["f%d_points" % k].shape
        assert np.allclose(f["xyz"], g["f%d_points" % k], rtol=1e-5, atol=2e-5)
        vb = f["hyp_boxes"][f["valid"]]
        assert vb.shape == g["f%d_boxes" % k].shape                           # same valid set size
        assert np.allclose(vb, g["f%d_boxes" % k], rtol=1e-5, atol=1e-5)
        assert np.array_equal(f["counts"][f["valid"]], g["f%d_counts" % k])   # per-hypothesis counts
        assert np.allclose(f["best_score"], g["f%d_scores" % k].max(), rtol=1e-5 if not opts else 1e-4)
        if "scores" in f:      # second-stage score of every valid hypothesis, optional terms included
            assert np.allclose(f["scores"][f["valid"]], g["f%d_scores" % k], rtol=1e-4, atol=1e-4)


def test_oracle_trig_close_to_libm():
    xs = np.linspace(-7, 7, 2001).astype(np.float32)
    for x in xs[::13]:
        assert abs(float(O.sinf(x)) - np.sin(float(x))) < 3e-7
        assert abs(float(O.cosf(x)) - np.cos(float(x))) < 3e-7
    assert float(O.atan2f(1.0, 1.0)) == pytest.approx(np.pi / 4, abs=2e-7)
    assert float(O.exp(0.0)) == 1.0


def test_quantile_matches_torch():
    import torch
    rng = np.random.default_rng(0)
    for n in (1, 2, 7, 100, 1001):
        x = (rng.random(n) * 50).astype(np.float32)
        for q in (0.0, 0.25, 1.0, 0.336):
            assert float(O.quantile(x, q)) == torch.quantile(torch.from_numpy(x), q).item()


def test_recall_record_counts():
    gt = np.array([[0, 0, 0, 4, 2, 1.5, 0.2, 0, 0, 1], [10, 0, 0, 4, 2, 1.5, 0, 0, 0, 2], [0, 0, 0, 0, 0, 0, 0, 0, 0, 0]], np.float32)
    pred = np.array([[0.1, 0, 0, 4, 2, 1.5, 0.2]], np.float32)
    rd = SO.recall_record(pred, gt)
    assert rd["gt"] == 2 and rd["rcnn_0.5"] == 1 and rd["num_3known"] == 1 and rd["rcnn_7unknown_0.3"] == 0


def test_occlusion_score_is_the_reference_broadcast_product():
    """calc_occl_scores (frustum_proposals_v1.py:408-477) forms ((cur_mags > m1) & (~real_mask)).sum() with a
    (P,1) tensor against a (P,) mask; the oracle restates the resulting (P,P) count as n_far * n_out.  Checked
    against that very torch expression (box_utils.boxes_to_corners_3d restated with torch ops)."""
    import torch
    rng = np.random.default_rng(11)
    pts = (rng.normal(0, 1, (300, 3)) * [3, 3, 0.6] + [12, 4, -0.5]).astype(np.float32)
    boxes = np.zeros((9, 7), np.float32)
    boxes[:, :3] = rng.normal(0, 1.5, (9, 3)) * [1, 1, 0.2] + [12, 4, -0.5]
    boxes[:, 3:6] = synth.PRIORS[rng.integers(0, 10, 9)]
    boxes[:, 6] = rng.uniform(0, np.pi, 9)
    fail, nfar = O.occl_fail(pts, boxes)
    counts = O.count_in_boxes(pts, boxes)
    tp = torch.from_numpy(pts)
    mags = tp.norm(dim=-1, keepdim=True)                       # (P,1), as pts_mags at :1008
    template = torch.tensor([[1, 1, -1], [1, -1, -1], [-1, -1, -1], [-1, 1, -1],
                             [1, 1, 1], [1, -1, 1], [-1, -1, 1], [-1, 1, 1]], dtype=torch.float32) / 2
    for i in range(9):
        b = torch.from_numpy(boxes[i])
        c, s = torch.cos(b[6]), torch.sin(b[6])
        rot = torch.stack([c, s, torch.zeros(()), -s, c, torch.zeros(()), torch.zeros(()), torch.zeros(()), torch.ones(())]).view(3, 3)
        corners = (b[3:6] * template) @ rot + b[:3]
        m1 = corners.norm(dim=-1).min()
        real_mask = torch.from_numpy(O.points_in_boxes_gpu(pts[None], boxes[i][None, None])[0] >= 0)   # (P,)
        ref = int(((mags > m1) & (~real_mask)).sum())          # broadcasts to (P,P)
        assert ref == int((mags[:, 0] > m1).sum()) * int((~real_mask).sum())
        assert abs(int(nfar[i]) - int((mags[:, 0] > m1).sum())) <= 1     # torch's CPU cos/sin/norm may differ in the last ulp
        if int(nfar[i]) == int((mags[:, 0] > m1).sum()):
            assert float(fail[i]) == float(np.float32(ref))
        assert int(nfar[i]) * (pts.shape[0] - int(counts[i])) == int(round(float(fail[i]))) or fail[i] > 2 ** 24
