"""nuScenes batch_dict producer against the reference's own dataset
code: tests/golden/nuscenes_feed.pkl holds a tiny nuScenes-format fixture and what the
reference's NuScenesDataset / DatasetTemplate / DataProcessor functions return for it
(tools/gen_golden_feed.py lifts them out of /root/reference with ast and runs them).  Every key
of every sample and of the collated batch must match exactly, dtype included."""
import os
import pickle
from pathlib import Path

import numpy as np
import pytest

from findnpropagate_b200 import nuscenes_feed as NF

GOLD = os.path.join(os.path.dirname(__file__), "golden", "nuscenes_feed.pkl")
NOT_PINNED = {"camera2ego"}      # pyquaternion is absent where the golden file is made (see gen_golden_feed.py)


@pytest.fixture(scope="module")
def gold():
    with open(GOLD, "rb") as f:
        return pickle.load(f)


@pytest.fixture()
def tree(gold, tmp_path):
    for rel, pts in gold["files"].items():
        p = tmp_path / rel
        p.parent.mkdir(parents=True, exist_ok=True)
        pts.tofile(str(p))
    return tmp_path


def _same(a, b, key):
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        a, b = np.asarray(a), np.asarray(b)
        assert a.dtype == b.dtype, (key, a.dtype, b.dtype)
        assert a.shape == b.shape, (key, a.shape, b.shape)
        if a.dtype.kind == "f":
            assert np.array_equal(a, b, equal_nan=True), key
        else:
            assert np.array_equal(a, b), key
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), key
        for x, y in zip(a, b):
            _same(x, y, key)
    elif isinstance(a, dict):
        assert sorted(a) == sorted(b), key
        for k in a:
            _same(a[k], b[k], key)
    else:
        assert a == b, (key, a, b)


@pytest.mark.parametrize("case", ["seeker_yaml", "ten_sweeps", "no_velocity", "seeker_yaml_extract", "no_velocity_extract"])
def test_samples_and_collated_batch_equal_the_reference(gold, tree, case):
    """*_extract: training=True, the mode tools/extract_pseudo_labels.py builds its loader in -- GT outside the
    range dropped, headings wrapped, the frame without GT re-drawn (SKIP_NO_GT) from numpy's global generator."""
    c = gold["cases"][case]
    feed = NF.NuScenesFeed(tree, gold["infos"], max_sweeps=c["max_sweeps"], pred_velocity=c["pred_velocity"],
                           training=c["training"])
    np.random.seed(gold["seed"])                 # the reference draws the sweeps from the global generator
    samples = [feed[i] for i in range(len(feed))]
    assert len(samples) == len(c["samples"])
    for got, ref in zip(samples, c["samples"]):
        assert sorted(got) == sorted(ref)
        for k in ref:
            if k not in NOT_PINNED:
                _same(got[k], ref[k], k)
        assert got["points"].dtype == np.float32 and got["points"].shape[1] == 5
        assert np.all(np.abs(got["points"][:, :2]) <= 54.0)
    batch = NF.collate_batch(samples)
    assert sorted(batch) == sorted(c["batch"])
    for k in c["batch"]:
        if k not in NOT_PINNED:
            _same(batch[k], c["batch"][k], k)
    assert batch["points"].shape[1] == 6 and batch["gt_boxes"].shape[2] == (10 if c["pred_velocity"] else 8)


def test_quaternion_matrix_is_a_rotation():
    rng = np.random.default_rng(0)
    for _ in range(20):
        R = NF.quaternion_rotation_matrix(rng.normal(size=4))
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and abs(np.linalg.det(R) - 1) < 1e-12
    assert np.allclose(NF.quaternion_rotation_matrix([1, 0, 0, 0]), np.eye(3))


def test_frame_input_and_prefetch(gold, tree):
    """frame_input feeds the engine's FrameInput from the feed + a reference-contract detector;
    prefetch yields the same frames in order, batch by batch."""
    import torch
    feed = NF.NuScenesFeed(tree, gold["infos"], max_sweeps=1)
    calls = []

    def detector(batch_dict):
        assert batch_dict["batch_size"] == 1 and len(batch_dict["image_paths"][0]) == 6
        calls.append(batch_dict["metadata"][0]["token"])
        n = 4
        return (torch.arange(n * 4, dtype=torch.float32).reshape(n, 4), torch.tensor([1, 2, 3, 9]),
                torch.tensor([0.9, 0.8, 0.7, 0.6]), torch.zeros(n, dtype=torch.long), torch.tensor([0, 0, 3, 5]))
    fi, frame_id, meta = feed.frame_input(1, detector)
    assert frame_id == Path(gold["infos"][1]["lidar_path"]).stem and meta["token"] == "token1"
    assert fi.points.dtype == np.float32 and fi.points.flags["C_CONTIGUOUS"] and fi.points.shape[1] == 5
    assert fi.lidar2image.shape == (6, 4, 4) and fi.lidar2image.dtype == np.float32
    assert fi.det_boxes.shape == (4, 4) and fi.det_cam_idx.tolist() == [0, 0, 3, 5]
    assert fi.gt_boxes.shape[1] == 10
    got = list(feed.prefetch(range(len(feed)), detector, batch_frames=2, workers=2))
    assert [len(b[0]) for b in got] == [2, 1]
    assert [i for b in got for i in b[1]] == [Path(i["lidar_path"]).stem for i in gold["infos"]]
    ref = feed.frame_input(2, detector)[0]
    assert np.array_equal(got[1][0][0].points, ref.points)


def test_preprocessed_detector_matches_reference_feeder(golden_dir, tmp_path):
    """proposer.PreprocessedDetector (per-camera COCO result files, the head's other feeder) against what the
    reference's own class returned for the same files (tools/gen_golden_detector.py): values, order, dtypes, a
    single-frame batch, unknown images, and the empty batch's (0,)-shaped tensors."""
    import json
    import torch
    from findnpropagate_b200.proposer import PreprocessedDetector
    cams = ['CAM_BACK', 'CAM_BACK_LEFT', 'CAM_BACK_RIGHT', 'CAM_FRONT', 'CAM_FRONT_LEFT', 'CAM_FRONT_RIGHT']
    cases = json.load(open(os.path.join(golden_dir, "preprocessed_detector.json")))
    assert len(cases) == 3
    for case in cases:
        files = []
        for cam, v in zip(cams, case["views"]):
            files.append(str(tmp_path / ("%s_%s.json" % (case["name"], cam))))
            json.dump(v, open(files[-1], "w"))
        det = PreprocessedDetector(files, class_names=case["class_names"])
        out = det({"image_paths": case["image_paths"], "batch_size": len(case["image_paths"])})
        assert len(out) == 5 and len(out[1]) > 10
        for got, want, dt in zip(out, case["out"], case["out_dtypes"]):
            assert str(got.dtype) == dt
            assert torch.equal(got, torch.tensor(want, dtype=got.dtype))
        one = det({"image_paths": [case["image_paths"][1]], "batch_size": 1})
        for got, want in zip(one, case["out_frame1"]):
            assert got.tolist() == want
        missing = det({"image_paths": [["x/unknown_%d.jpg" % c for c in range(6)]], "batch_size": 1})
        assert [list(t.shape) for t in missing] == case["missing_shapes"]
        assert [str(t.dtype) for t in missing] == case["missing_dtypes"]
    with pytest.raises(TypeError):
        det({"points": None})
