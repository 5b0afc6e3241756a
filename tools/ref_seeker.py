"""TEST INFRASTRUCTURE ONLY -- runs the *reference's own* FrustumProposerOG.

Two modes:

* ``load("cuda")`` / ``run(..., device="cuda")`` (GPU box): the head's source UNMODIFIED, its two
  native call sites bound to the reference's own op wrappers and the reference's own kernels compiled
  for sm_100a (oracle/_ref/*.so).  One documented patch: iou3d_nms_utils.py:152/167 index a CPU
  ``order`` with ``keep[:num_out].cuda()`` -- the head passes CPU scores (frustum_proposals_v1.py:919,
  996), and current PyTorch refuses a CUDA index into a CPU tensor, so the index is moved to the device
  of ``order``.  The CPU ``sort`` the wrapper then runs is stable in this PyTorch build.
  The Python files come from /root/reference where it exists, else from the copy oracle/build_ref.py
  installed under oracle/_ref/pysrc (git-ignored, shipped to the GPU box).

* ``load("cpu")`` (build container, no GPU; generates tests/golden): the reference modules are imported
  from where they lie, behind bare namespace stubs so the heavy package __init__ files (spconv, clip,
  kornia ...) never run.  The head hard-codes device='cuda' (frustum_proposals_v1.py:240-303), so its
  source text is loaded with the substitutions  device='cuda' -> device='cpu'  and  .cuda() -> .cpu()
  (nothing else).

Documented deviations of the CPU mode from an unmodified GPU run:
  1. roiaware_pool3d_utils.points_in_boxes_gpu is emulated by the oracle's restatement of
     the GPU kernel predicate (no GPU here);
  2. iou3d_nms_utils.nms_normal_gpu is emulated with a *stable* descending sort and the
     oracle's iou_normal greedy scan (the reference's unstable sort leaves ties undefined);
  3. PreprocessedGLIP is replaced by a feeder returning the synthetic detections;
  4. capture hooks record intermediates.
"""
import importlib
import os
import re
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import build_ref  # noqa: E402
import oracle as O  # noqa: E402


class AttrDict(dict):
    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


_CAPTURE = None


def _pkg(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__package__ = name
    sys.modules[name] = m
    return m


def _emul_points_in_boxes_gpu(points, boxes):
    out = O.points_in_boxes_gpu(points.detach().cpu().numpy(), boxes.detach().cpu().numpy())
    if _CAPTURE is not None:
        _CAPTURE["pib_calls"].append((boxes.detach().cpu().numpy().reshape(-1, 7).copy(), int((out >= 0).sum())))
        _CAPTURE["last_points"] = points.detach().cpu().numpy().reshape(-1, 3)
    return torch.from_numpy(out)


def _emul_nms_normal_gpu(boxes, scores, thresh, **kw):
    b = boxes.detach().cpu().numpy()
    s = scores.detach().cpu().numpy()
    keep = O.nms_normal(b, s, thresh)
    if _CAPTURE is not None:
        calls = _CAPTURE["pib_calls"]
        n = b.shape[0]
        # the density loop makes the first n calls of a frustum; calc_occl_scores (occl_w / OCCL_MULT) makes
        # n more per use, after it
        assert len(calls) % n == 0 and len(calls) >= n
        _CAPTURE["frustums"].append(dict(
            boxes=b.copy(), scores=s.copy(), keep=keep.copy(),
            counts=np.array([c for _, c in calls[:n]], np.int32),
            points=_CAPTURE.get("last_points", np.zeros((0, 3), np.float32)).copy()))
        _CAPTURE["pib_calls"] = []
    return torch.from_numpy(keep), None


def _load_source(modname, path, package, patch=None):
    src = open(path).read()
    if patch is not None:
        src = patch(src)
    mod = types.ModuleType(modname)
    mod.__file__ = path
    mod.__package__ = package
    sys.modules[modname] = mod
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        exec(compile(src, path, "exec"), mod.__dict__)
    return mod


_MODE = None


def load(device="cpu", head_file="frustum_proposals_v1.py"):
    """Import the reference head and return its module.  device: 'cpu' (patched source, emulated native
    ops) or 'cuda' (unmodified head, the reference's own wrappers and compiled kernels)."""
    global _MODE
    key = "fnp_ref_head:" + head_file
    if key in sys.modules:
        assert _MODE == device, "one mode per process (the reference modules are registered globally)"
        return sys.modules[key]
    assert _MODE in (None, device), "one mode per process (the reference modules are registered globally)"
    _MODE = device
    P = os.path.join(build_ref.py_root(), "pcdet")
    for n, p in [("pcdet", P), ("pcdet.models", P + "/models"),
                 ("pcdet.models.dense_heads", P + "/models/dense_heads"),
                 ("pcdet.models.dense_heads.target_assigner", P + "/models/dense_heads/target_assigner"),
                 ("pcdet.models.model_utils", P + "/models/model_utils"), ("pcdet.utils", P + "/utils"),
                 ("pcdet.ops", P + "/ops"), ("pcdet.ops.roiaware_pool3d", P + "/ops/roiaware_pool3d"),
                 ("pcdet.ops.iou3d_nms", P + "/ops/iou3d_nms")]:
        if n not in sys.modules:
            _pkg(n, p)
    sys.modules.setdefault("SharedArray", types.ModuleType("SharedArray"))
    # the compiled reference extension modules (oracle/_ref), importable by name
    for pk, name in (("pcdet.ops.iou3d_nms", "iou3d_nms_cuda"), ("pcdet.ops.roiaware_pool3d", "roiaware_pool3d_cuda")):
        if pk + "." + name in sys.modules:
            continue
        try:
            ext = build_ref.load(name)
        except ImportError:
            if device == "cuda":
                raise
            ext = types.ModuleType(name)
        sys.modules[pk + "." + name] = ext
        setattr(sys.modules[pk], name, ext)
    if device == "cuda":
        # the reference's own wrappers around its own kernels
        if "pcdet.ops.roiaware_pool3d.roiaware_pool3d_utils" not in sys.modules:
            rp = _load_source("pcdet.ops.roiaware_pool3d.roiaware_pool3d_utils",
                              P + "/ops/roiaware_pool3d/roiaware_pool3d_utils.py", "pcdet.ops.roiaware_pool3d")
            sys.modules["pcdet.ops.roiaware_pool3d"].roiaware_pool3d_utils = rp

            def same_device_index(src):
                assert src.count("order[keep[:num_out].cuda()]") == 2
                return src.replace("order[keep[:num_out].cuda()]", "order[keep[:num_out].to(order.device)]")
            iu = _load_source("pcdet.ops.iou3d_nms.iou3d_nms_utils", P + "/ops/iou3d_nms/iou3d_nms_utils.py",
                              "pcdet.ops.iou3d_nms", patch=same_device_index)
            sys.modules["pcdet.ops.iou3d_nms"].iou3d_nms_utils = iu
            _hook_native(rp, iu)
    elif "pcdet.ops.roiaware_pool3d.roiaware_pool3d_utils" not in sys.modules:
        # native op wrappers -> emulations (no GPU in this container)
        rp = types.ModuleType("pcdet.ops.roiaware_pool3d.roiaware_pool3d_utils")
        rp.points_in_boxes_gpu = _emul_points_in_boxes_gpu
        rp.points_in_boxes_cpu = lambda p, b: O.points_in_boxes_cpu(np.asarray(p), np.asarray(b))
        sys.modules[rp.__name__] = rp
        sys.modules["pcdet.ops.roiaware_pool3d"].roiaware_pool3d_utils = rp
        iu = types.ModuleType("pcdet.ops.iou3d_nms.iou3d_nms_utils")
        iu.nms_normal_gpu = _emul_nms_normal_gpu
        iu.boxes_iou3d_gpu = lambda a, b: torch.from_numpy(O.boxes_iou3d(a.cpu().numpy(), b.cpu().numpy()))
        iu.boxes_bev_iou_cpu = lambda a, b: O.boxes_iou_bev(np.asarray(a), np.asarray(b))
        sys.modules[iu.__name__] = iu
        sys.modules["pcdet.ops.iou3d_nms"].iou3d_nms_utils = iu

    path = os.path.join(P, "models/dense_heads", head_file)

    def to_cpu(src):
        src = src.replace("device='cuda'", "device='cpu'").replace(".cuda()", ".cpu()")
        assert "cuda" not in re.sub(r"#.*", "", src).replace("torch.cuda", ""), "unpatched cuda use"
        return src
    mod = _load_source("pcdet.models.dense_heads." + head_file[:-3], path, "pcdet.models.dense_heads",
                       patch=to_cpu if device == "cpu" else None)
    sys.modules[key] = mod
    return mod


def _hook_native(rp, iu):
    """CUDA mode: capture hooks around the reference's two native call sites (pass-through when
    _CAPTURE is None, which is how the timing runs call the head)."""
    real_pib, real_nms = rp.points_in_boxes_gpu, iu.nms_normal_gpu

    def pib(points, boxes):
        out = real_pib(points, boxes)
        if _CAPTURE is not None:
            _CAPTURE["pib_calls"].append((boxes.detach().reshape(-1, 7).cpu().numpy().copy(), int((out >= 0).sum())))
            _CAPTURE["last_points"] = points
        return out

    def nms_normal(boxes, scores, thresh, **kw):
        keep, aux = real_nms(boxes, scores, thresh, **kw)
        if _CAPTURE is not None:
            calls = _CAPTURE["pib_calls"]
            n = boxes.shape[0]
            assert len(calls) % n == 0 and len(calls) >= n
            lp = _CAPTURE.get("last_points")
            _CAPTURE["frustums"].append(dict(
                boxes=boxes.detach().cpu().numpy().copy(), scores=scores.detach().cpu().numpy().copy(),
                keep=keep.detach().cpu().numpy().copy(), counts=np.array([c for _, c in calls[:n]], np.int32),
                points=(lp.detach().reshape(-1, 3).cpu().numpy().copy() if lp is not None else np.zeros((0, 3), np.float32))))
            _CAPTURE["pib_calls"] = []
        return keep, aux
    rp.points_in_boxes_gpu, iu.nms_normal_gpu = pib, nms_normal


class SyntheticFeeder:
    """Stands in for PreprocessedGLIP (preprocessed_detector.py:104): same 5-tensor return."""

    def __init__(self, frames):
        self.frames = frames

    def __call__(self, batch_dict):
        boxes, labels, scores, idx, cam = [], [], [], [], []
        for b, f in enumerate(self.frames):
            boxes.append(torch.from_numpy(f.det_boxes))
            labels.append(torch.from_numpy(f.det_labels))
            scores.append(torch.from_numpy(f.det_scores))
            idx.extend([b] * len(f.det_boxes))
            cam.append(torch.from_numpy(f.det_cam_idx))
        return (torch.cat(boxes), torch.cat(labels), torch.cat(scores), torch.tensor(idx, dtype=torch.long),
                torch.cat(cam))


def build_head(params, frames, box_format="xyxy", device="cpu"):
    mod = load(device)
    mod.PreprocessedGLIP = lambda class_names=None: SyntheticFeeder(frames)
    flags = {k: bool(params.get(k)) for k in ("MULT", "OCCL_MULT", "MULTICAM_IOU")}   # model_cfg-level switches
    cfg = AttrDict(PARAMS={k: v for k, v in params.items() if k not in flags}, PREDS_PATH="PreprocessedGLIP",
                   BOX_FORMAT=box_format, **flags)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        head = mod.FrustumProposerOG(model_cfg=cfg, class_names=None)
    head.eval()
    return head


def batch_dict(frames, device="cpu"):
    """The collated batch the head consumes (pcdet/models/__init__.py:23-36: float arrays -> float
    tensors on the model's device)."""
    from findnpropagate_b200 import synth
    bd = synth.collate(frames)
    for k, v in list(bd.items()):
        if isinstance(v, np.ndarray) and v.dtype.kind == "f":
            bd[k] = torch.from_numpy(v).float().to(device)
    return bd


def run(frames, params, capture=True, box_format="xyxy", device="cpu", head=None):
    """Run reference get_proposals over `frames` (one call, batch_size=len(frames)).
    Returns (boxes (K,7), labels (K), scores (K), batch_idx (K), capture dict, head)."""
    global _CAPTURE
    if head is None:
        head = build_head(params, frames, box_format, device)
    head.image_detector = SyntheticFeeder(frames)
    bd = batch_dict(frames, device)
    _CAPTURE = dict(pib_calls=[], frustums=[]) if capture else None
    try:
        with torch.no_grad():
            boxes, labels, scores, bidx = head.get_proposals(bd)
    finally:
        cap, _CAPTURE = _CAPTURE, None
    return (boxes.cpu().numpy().astype(np.float32), labels.cpu().numpy(), scores.cpu().numpy().astype(np.float32),
            bidx.cpu().numpy(), cap, head)
