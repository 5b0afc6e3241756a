"""TEST INFRASTRUCTURE ONLY -- golden vectors for findnpropagate_b200.proposer.PreprocessedDetector from the
reference's own class (pcdet/models/preprocessed_detector.py:111-290).

Run in the build container (needs /root/reference):   python tools/gen_golden_detector.py

Writes tests/golden/preprocessed_detector.json: the six synthetic COCO result files (one per camera view),
the batch_dict fields the feeder reads, and what the reference feeder returns for them, for three set-ups:
all class names, a subset of class names in another order, and result files whose annotation category ids
are 1-based over 0-based categories (the reference shifts them, :176-177).
"""
import contextlib
import importlib.util
import io
import json
import os
import tempfile

import numpy as np

REF = os.environ.get("FNP_REFERENCE_ROOT", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CAMS = ['CAM_BACK', 'CAM_BACK_LEFT', 'CAM_BACK_RIGHT', 'CAM_FRONT', 'CAM_FRONT_LEFT', 'CAM_FRONT_RIGHT']
NAMES = ['car', 'truck', 'construction_vehicle', 'bus', 'trailer', 'barrier', 'motorcycle', 'bicycle',
         'pedestrian', 'traffic_cone']


def make_views(rng, n_frames, cat_base, ann_shift, with_ext):
    cats = [{"id": cat_base + i, "name": n} for i, n in enumerate(NAMES)]
    views, paths = [], [[None] * 6 for _ in range(n_frames)]
    for c, cam in enumerate(CAMS):
        images, anns = [], []
        for b in range(n_frames):
            stem = "n%03d__%s__%d" % (b, cam, 1531883530412470 + 37 * b + c)
            fname = "samples/%s/%s.jpg" % (cam, stem)
            images.append({"id": 100 * c + b, "file_name": fname if with_ext else "samples/%s/%s" % (cam, stem)})
            paths[b][c] = "../data/nuscenes/v1.0-trainval/" + fname
            for k in range(int(rng.integers(0, 5))):
                a = {"id": len(anns), "image_id": 100 * c + b,
                     "category_id": int(rng.integers(0, 10)) + cat_base + ann_shift,
                     "bbox": [float(np.float32(v)) for v in rng.uniform(0, 800, 4)]}
                if rng.random() < 0.8:
                    a["score"] = float(np.float32(rng.uniform(0.05, 0.99)))
                anns.append(a)
        views.append({"images": images, "annotations": anns, "categories": cats})
    return views, paths


def main():
    spec = importlib.util.spec_from_file_location("ref_preprocessed_detector",
                                                  os.path.join(REF, "pcdet/models/preprocessed_detector.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(5)
    cases = []
    for name, class_names, cat_base, ann_shift, with_ext in (
            ("all_classes", None, 0, 0, True),
            ("subset_reordered", ['pedestrian', 'car', 'bicycle'], 1, 0, True),
            ("one_based_annotations", NAMES, 0, 1, False)):
        views, paths = make_views(rng, 3, cat_base, ann_shift, with_ext)
        if ann_shift:      # ids that are valid as they are stay: only the out-of-range one (10) is shifted (:176-177)
            pass
        with tempfile.TemporaryDirectory() as td:
            files = []
            for cam, v in zip(CAMS, views):
                files.append(os.path.join(td, "OWL_%s.json" % cam))
                json.dump(v, open(files[-1], "w"))
            with contextlib.redirect_stdout(io.StringIO()):
                det = mod.PreprocessedDetector(files, class_names=class_names)
            bd = {"image_paths": paths, "batch_size": len(paths)}
            out = det(bd)
            one = det({"image_paths": [paths[1]], "batch_size": 1})
            missing = det({"image_paths": [["x/unknown_%d.jpg" % c for c in range(6)]], "batch_size": 1})
        cases.append(dict(name=name, class_names=class_names, views=views, image_paths=paths,
                          out=[t.tolist() for t in out], out_dtypes=[str(t.dtype) for t in out],
                          out_frame1=[t.tolist() for t in one],
                          missing_shapes=[list(t.shape) for t in missing], missing_dtypes=[str(t.dtype) for t in missing]))
        print(name, "D =", len(out[1]), [str(t.dtype) for t in out])
    json.dump(cases, open(os.path.join(ROOT, "tests", "golden", "preprocessed_detector.json"), "w"))


if __name__ == "__main__":
    main()
