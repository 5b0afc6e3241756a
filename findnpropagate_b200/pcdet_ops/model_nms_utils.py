"""Generic NMS dispatcher on top of the device NMS ops, with the call
contract of pcdet/models/model_utils/model_nms_utils.py:6-66: nms_config carries NMS_TYPE
('nms_gpu' | 'nms_normal_gpu'), NMS_THRESH, NMS_PRE_MAXSIZE, NMS_POST_MAXSIZE and is also passed
through as keyword arguments to the op."""
import torch

from . import iou3d_nms_utils


def _cfg(nms_config, key):
    return nms_config[key] if isinstance(nms_config, dict) else getattr(nms_config, key)


def _kwargs(nms_config):
    return dict(nms_config) if isinstance(nms_config, dict) or hasattr(nms_config, "keys") else {}


def _select(scores, boxes, nms_config):
    """Indices (into scores/boxes) surviving top-k pre-selection + NMS + post truncation."""
    if scores.shape[0] == 0:
        return torch.zeros(0, dtype=torch.long, device=scores.device)
    top_scores, top_idx = torch.topk(scores, k=min(int(_cfg(nms_config, 'NMS_PRE_MAXSIZE')), scores.shape[0]))
    op = getattr(iou3d_nms_utils, _cfg(nms_config, 'NMS_TYPE'))
    keep, _ = op(boxes[top_idx][:, 0:7], top_scores, _cfg(nms_config, 'NMS_THRESH'), **_kwargs(nms_config))
    return top_idx[keep[:int(_cfg(nms_config, 'NMS_POST_MAXSIZE'))]]


def class_agnostic_nms(box_scores, box_preds, nms_config, score_thresh=None):
    """-> (selected indices into the inputs, their scores)."""
    if score_thresh is None:
        selected = _select(box_scores, box_preds, nms_config)
    else:
        above = (box_scores >= score_thresh).nonzero().view(-1)
        selected = above[_select(box_scores[above], box_preds[above], nms_config)]
    return selected, box_scores[selected]


def multi_classes_nms(cls_scores, box_preds, nms_config, score_thresh=None):
    """cls_scores (N, num_class), box_preds (N, 7+C) -> (scores, labels, boxes) of the survivors,
    class by class."""
    out_s, out_l, out_b = [], [], []
    for k in range(cls_scores.shape[1]):
        s, b = cls_scores[:, k], box_preds
        if score_thresh is not None:
            m = s >= score_thresh
            s, b = s[m], b[m]
        sel = _select(s, b, nms_config)
        out_s.append(s[sel])
        out_l.append(torch.full((sel.shape[0],), k, dtype=torch.long, device=s.device))
        out_b.append(b[sel])
    return torch.cat(out_s), torch.cat(out_l), torch.cat(out_b)
