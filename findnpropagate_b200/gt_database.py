"""Ground-truth database creation: the other caller of batched first-match ``points_in_boxes_gpu`` on this path
(SURVEY.md section 8 row f4).

Reference: ``NuScenesDataset.create_groundtruth_database`` (pcdet/datasets/nuscenes/nuscenes_dataset.py:346-390):
per frame, every point is assigned to the FIRST GT box that contains it (one ``points_in_boxes_gpu`` call over all
of the frame's boxes), the points of each box are written to ``gt_database_<S>sweeps_withvelo/<frame>_<name>_<i>.bin``
relative to the box centre, and one info dict per box goes into ``nuscenes_dbinfos_<S>sweeps_withvelo.pkl``.

Here the frames of a batch are uploaded together and assigned by ONE ``fnp_points_in_boxes`` launch (B frames x T
boxes x M points, the op's batched form; shorter frames are padded with a far-away point / an empty box), the rest is
the same host-side bookkeeping and the same files, byte for byte.
"""
import pickle
from pathlib import Path
from typing import Optional, Sequence

import numpy as np
import torch

from .nuscenes_feed import lidar_with_sweeps
from .pcdet_ops import roiaware_pool3d_utils

_FAR = 1.0e6        # padding points lie in no box


def assign_points_to_boxes(points_list, boxes_list, device=None):
    """First containing box per point (or -1) for every frame of a batch, one launch.
    points_list[b]: (N_b, >=3) float array, boxes_list[b]: (G_b, >=7).  Returns a list of (N_b,) int64 arrays --
    what ``points_in_boxes_gpu(points[None], boxes[None]).long().squeeze(0).cpu().numpy()`` gives frame by frame."""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    B = len(points_list)
    if B == 0:
        return []
    M = max(max(p.shape[0] for p in points_list), 1)
    T = max(max(g.shape[0] for g in boxes_list), 1)
    pts = np.full((B, M, 3), _FAR, np.float32)
    boxes = np.zeros((B, T, 7), np.float32)
    boxes[:, :, 0:3] = -_FAR                  # padding boxes: zero size, far from every real and padding point
    for b in range(B):
        pts[b, :points_list[b].shape[0]] = np.asarray(points_list[b], np.float32)[:, 0:3]
        boxes[b, :boxes_list[b].shape[0]] = np.asarray(boxes_list[b], np.float32)[:, 0:7]
    idx = roiaware_pool3d_utils.points_in_boxes_gpu(torch.from_numpy(pts).to(dev), torch.from_numpy(boxes).to(dev))
    idx = idx.long().cpu().numpy()
    return [idx[b, :points_list[b].shape[0]] for b in range(B)]


def create_groundtruth_database(root_path, infos: Sequence[dict], used_classes: Optional[Sequence[str]] = None,
                                max_sweeps: int = 10, batch_frames: int = 16, device=None, rng=None):
    """nuscenes_dataset.py:346-390 for a list of info dicts under ``root_path``; returns the db-info dict that is
    also pickled.  ``rng``: private generator for the sweep draw (None: numpy's global one, like the reference)."""
    root = Path(root_path)
    db_dir = root / ('gt_database_%dsweeps_withvelo' % max_sweeps)
    db_info_path = root / ('nuscenes_dbinfos_%dsweeps_withvelo.pkl' % max_sweeps)
    db_dir.mkdir(parents=True, exist_ok=True)
    all_db_infos = {}
    for s in range(0, len(infos), batch_frames):
        chunk = list(range(s, min(s + batch_frames, len(infos))))
        points = [lidar_with_sweeps(root, infos[i], max_sweeps, rng) for i in chunk]       # the reference's order of draws
        assigned = assign_points_to_boxes(points, [infos[i]['gt_boxes'] for i in chunk], device=device)
        for k, idx in enumerate(chunk):
            info, pts, box_of_pt = infos[idx], points[k], assigned[k]
            gt_boxes, gt_names = info['gt_boxes'], info['gt_names']
            for i in range(gt_boxes.shape[0]):
                filepath = db_dir / ('%s_%s_%d.bin' % (idx, gt_names[i], i))
                gt_points = pts[box_of_pt == i]
                gt_points[:, :3] -= gt_boxes[i, :3]
                with open(filepath, 'w') as f:
                    gt_points.tofile(f)
                if (used_classes is None) or gt_names[i] in used_classes:
                    db_info = {'name': gt_names[i], 'path': str(filepath.relative_to(root)), 'image_idx': idx, 'gt_idx': i,
                               'box3d_lidar': gt_boxes[i], 'num_points_in_gt': gt_points.shape[0]}
                    all_db_infos.setdefault(gt_names[i], []).append(db_info)
    with open(db_info_path, 'wb') as f:
        pickle.dump(all_db_infos, f)
    return all_db_infos
