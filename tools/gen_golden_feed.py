"""TEST INFRASTRUCTURE ONLY -- golden vectors for findnpropagate_b200.nuscenes_feed from the
reference's own dataset code.

Run in the build container (needs /root/reference):   python tools/gen_golden_feed.py

The reference's dataset modules cannot be imported here (SharedArray, skimage, pyquaternion,
spconv ... are absent), so the *function definitions themselves* are lifted out of the
reference's source files with ``ast`` at run time -- nothing is copied into this repository --
and executed against a small synthetic nuScenes-format fixture written to a temporary directory:

    pcdet/datasets/nuscenes/nuscenes_dataset.py   NuScenesDataset.get_sweep, get_lidar_with_sweeps,
                                                  fake_crop_image, load_camera_info, __getitem__
    pcdet/datasets/dataset.py                     DatasetTemplate.set_lidar_aug_matrix, prepare_data,
                                                  collate_batch
    pcdet/datasets/processor/data_processor.py    DataProcessor.mask_points_and_boxes_outside_range,
                                                  shuffle_points, forward
    pcdet/datasets/processor/point_feature_encoder.py   (whole module: numpy only)
    pcdet/utils/common_utils.py                   mask_points_by_range, keep_arrays_by_name

One stand-in: ``pyquaternion.Quaternion`` (absent) is replaced by the textbook unit-quaternion
rotation matrix, so ``camera2ego`` -- which the seeker never reads -- is not pinned by this file.
The fixture (raw point files, info dicts) and the reference's outputs go to
tests/golden/nuscenes_feed.npz / .pkl.
"""
import ast
import copy
import os
import pickle
import sys
import tempfile
import types
from collections import defaultdict
from functools import partial
from pathlib import Path

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("FNP_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

CLASS_NAMES = ['car', 'truck', 'construction_vehicle', 'bus', 'trailer', 'barrier', 'motorcycle', 'bicycle',
               'pedestrian', 'traffic_cone']


class AttrDict(dict):
    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


def lift(path, names, glb, cls=None):
    """exec the named top-level functions (or methods of class ``cls``) of a reference file."""
    tree = ast.parse(open(os.path.join(REF, path)).read())
    body = tree.body
    if cls is not None:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    out = {}
    for n in body:
        if isinstance(n, ast.FunctionDef) and n.name in names:
            n.decorator_list = []
            mod = ast.Module(body=[n], type_ignores=[])
            ns = {}
            exec(compile(mod, path, "exec"), glb, ns)
            out[n.name] = ns[n.name]
    missing = set(names) - set(out)
    assert not missing, (path, missing)
    return out


class Quaternion:
    def __init__(self, q):
        q = np.asarray(q, dtype=np.float64)
        self.q = q / np.linalg.norm(q)

    @property
    def rotation_matrix(self):
        w, x, y, z = self.q
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def rand_rot(rng):
    q = rng.normal(size=4)
    return Quaternion(q).rotation_matrix


def make_fixture(root, rng, n_frames=3, n_sweeps=3, n_pts=260):
    """A tiny nuScenes-format tree: LiDAR .bin files (5 floats per point) and info dicts with the
    fields the reference reads (tools' create_data output format)."""
    infos, files = [], {}

    def write(rel, n):
        pts = rng.uniform(-70, 70, (n, 5)).astype(np.float32)
        pts[:, 2] = rng.uniform(-5, 3, n)
        pts[: n // 10, :2] = rng.uniform(-1.2, 1.2, (n // 10, 2))          # some ego points
        pts[n // 10: n // 5, 0] = rng.choice([-54.0, 54.0, 54.000004, -54.000004], n // 5 - n // 10)   # range edge
        p = Path(root) / rel
        p.parent.mkdir(parents=True, exist_ok=True)
        pts.tofile(str(p))
        files[rel] = pts
        return rel

    for i in range(n_frames):
        key = write("samples/LIDAR_TOP/frame%d__LIDAR_TOP__%d.pcd.bin" % (i, 1000 + i), n_pts + 17 * i)
        sweeps = []
        for k in range(n_sweeps):
            T = np.eye(4)
            T[:3, :3] = rand_rot(rng)
            T[:3, 3] = rng.normal(size=3) * 0.5
            sweeps.append(dict(lidar_path=write("sweeps/LIDAR_TOP/frame%d_sweep%d.pcd.bin" % (i, k), n_pts - 11 * k),
                               transform_matrix=T if (i + k) % 4 else None, time_lag=0.05 * (k + 1)))
        cams = {}
        for c, name in enumerate(["CAM_FRONT", "CAM_FRONT_RIGHT", "CAM_FRONT_LEFT", "CAM_BACK", "CAM_BACK_LEFT",
                                  "CAM_BACK_RIGHT"]):
            cams[name] = dict(data_path="samples/%s/frame%d__%s.jpg" % (name, i, name),
                              sensor2lidar_rotation=rand_rot(rng), sensor2lidar_translation=rng.normal(size=3),
                              camera_intrinsics=np.array([[1266.4 + c, 0, 816.3], [0, 1266.4 - c, 491.5], [0, 0, 1.0]]),
                              sensor2ego_rotation=list(rng.normal(size=4)), sensor2ego_translation=list(rng.normal(size=3)))
        G = 5 + i
        names = np.array(rng.choice(CLASS_NAMES + ["animal", "debris"], G))
        gt = rng.uniform(-60, 60, (G, 9)).astype(np.float32)
        gt[:, 2] = rng.uniform(-6, 4, G)                       # most centres inside the z range, some outside
        gt[:, 6] = rng.uniform(-9, 9, G)                       # headings beyond +-pi: wrapped in training mode
        if i == 1:
            gt[:, 0] = rng.choice([-58.0, 57.5], G)            # no GT centre inside the range: SKIP_NO_GT re-draws
        gt[0, 7] = np.nan
        infos.append(dict(lidar_path=key, token="token%d" % i, sweeps=sweeps, cams=cams, gt_boxes=gt, gt_names=names,
                          num_lidar_pts=rng.integers(0, 4, G)))
    return infos, files


def reference_outputs(root, infos, max_sweeps, pred_velocity, training=False):
    g = dict(np=np, Path=Path, copy=copy, defaultdict=defaultdict, partial=partial, torch=torch, Quaternion=Quaternion)
    cu = types.SimpleNamespace(**lift("pcdet/utils/common_utils.py",
                                      ["mask_points_by_range", "keep_arrays_by_name", "limit_period", "check_numpy_to_torch"], g))
    g["common_utils"] = cu
    g["check_numpy_to_torch"] = cu.check_numpy_to_torch
    g["box_utils"] = types.SimpleNamespace(**lift("pcdet/utils/box_utils.py", ["mask_boxes_outside_range_numpy"], g))
    aug = lift("pcdet/datasets/augmentor/data_augmentor.py", ["forward"], g, cls="DataAugmentor")
    pfe_ns = dict(np=np)
    exec(compile(open(os.path.join(REF, "pcdet/datasets/processor/point_feature_encoder.py")).read(), "pfe", "exec"), pfe_ns)
    ds = lift("pcdet/datasets/nuscenes/nuscenes_dataset.py",
              ["get_sweep", "get_lidar_with_sweeps", "fake_crop_image", "load_camera_info", "__getitem__"], g,
              cls="NuScenesDataset")
    dt = lift("pcdet/datasets/dataset.py", ["set_lidar_aug_matrix", "prepare_data", "collate_batch"], g, cls="DatasetTemplate")
    dp = lift("pcdet/datasets/processor/data_processor.py", ["mask_points_and_boxes_outside_range", "shuffle_points", "forward"],
              g, cls="DataProcessor")

    class Proc:
        pass
    proc = Proc()
    proc.point_cloud_range = np.array([-54.0, -54.0, -5.0, 54.0, 54.0, 3.0], dtype=np.float32)
    proc.training, proc.mode = training, ('train' if training else 'test')
    for k, f in dp.items():
        setattr(Proc, k, f)
    proc.data_processor_queue = [
        proc.mask_points_and_boxes_outside_range(config=AttrDict(NAME='mask_points_and_boxes_outside_range',
                                                                 REMOVE_OUTSIDE_BOXES=True)),
        proc.shuffle_points(config=AttrDict(NAME='shuffle_points', SHUFFLE_ENABLED={'train': False, 'test': False}))]

    class DS:
        pass
    for k, f in list(ds.items()) + list(dt.items()):
        setattr(DS, k, staticmethod(f) if k == "collate_batch" else f)
    d = DS()
    d.infos, d.root_path = infos, Path(root)
    d.training, d.class_names = training, CLASS_NAMES
    d.skip_no_gt = True                                  # DatasetTemplate.__init__, dataset.py:23 (SKIP_NO_GT default)

    class Aug:                                           # DataAugmentor with AUG_CONFIG_LIST = [] (extract_pseudo_labels.py:50)
        data_augmentor_queue = []
    Aug.forward = aug["forward"]
    d.data_augmentor = Aug()
    d._merge_all_iters_to_one_epoch = False
    d.use_camera, d.cam_without_image = True, True
    d.camera_image_config = AttrDict(FINAL_DIM=[900, 1600], RESIZE_LIM_TEST=[1.0, 1.0])
    d.dataset_cfg = AttrDict(MAX_SWEEPS=max_sweeps, FILTER_MIN_POINTS_IN_GT=1, SET_NAN_VELOCITY_TO_ZEROS=True,
                             PRED_VELOCITY=pred_velocity)
    d.point_feature_encoder = pfe_ns["PointFeatureEncoder"](
        AttrDict(encoding_type='absolute_coordinates_encoding', used_feature_list=['x', 'y', 'z', 'intensity', 'timestamp'],
                 src_feature_list=['x', 'y', 'z', 'intensity', 'timestamp']), point_cloud_range=proc.point_cloud_range)
    d.data_processor = proc
    DS.__len__ = lambda self: len(infos)
    np.random.seed(1234)                                # the reference draws the sweeps from the global generator
    samples = [d.__getitem__(i) for i in range(len(infos))]
    batch = DS.collate_batch([copy.deepcopy(s) for s in samples])
    return samples, batch


def main():
    rng = np.random.default_rng(20240917)
    os.makedirs(OUT, exist_ok=True)
    with tempfile.TemporaryDirectory() as root:
        infos, files = make_fixture(root, rng)
        cases = {}
        # *_extract: training=True, the mode tools/extract_pseudo_labels.py:47-58 builds its loader in
        for name, (ms, pv, tr) in {"seeker_yaml": (1, True, False), "ten_sweeps": (4, True, False),
                                   "no_velocity": (2, False, False), "seeker_yaml_extract": (1, True, True),
                                   "no_velocity_extract": (3, False, True)}.items():
            samples, batch = reference_outputs(root, copy.deepcopy(infos), ms, pv, tr)
            cases[name] = dict(max_sweeps=ms, pred_velocity=pv, training=tr, samples=samples, batch=batch)
    with open(os.path.join(OUT, "nuscenes_feed.pkl"), "wb") as f:
        pickle.dump(dict(infos=infos, files=files, cases=cases, seed=1234), f, protocol=4)
    for name, c in cases.items():
        print(name, "points", [s["points"].shape for s in c["samples"]], "batch keys", sorted(c["batch"].keys()))


if __name__ == "__main__":
    main()
