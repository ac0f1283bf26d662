"""CPU restatement of the candidate selection of CoDA's stage-2 novel-box discovery.

TEST INFRASTRUCTURE ONLY (see oracle/cpu_step.py).  Follows the reference loop of
models/model_3detr.py:1298-1420 step by step: dummy box (0, 0, 2, 2) + score -1 for a box that was given up
(:1303-1342), torchvision.ops.nms(box2d, scores, 0.25) (:1348), axis-aligned 3-D IoU of every kept box against
every present ground-truth box with `cal_iou` (:868-899, :1374-1386), `box_save` thresholding on
`save_objectness` and the give-up list (:1402-1410); returns the surviving indices in NMS (score) order."""
from __future__ import annotations

import numpy as np
import torch


def _aabb(c):          # (n, 8, 3) -> (n, 6)
    return torch.cat([c.min(dim=1)[0], c.max(dim=1)[0]], dim=1)


def _iou6(p, g):
    lo = torch.maximum(p[:3], g[:3])
    hi = torch.minimum(p[3:], g[3:])
    inter = torch.clamp(hi - lo, min=0).prod()
    return inter / ((p[3:] - p[:3]).prod() + (g[3:] - g[:3]).prod() - inter)


def novel_candidates_ref(boxes2d, valid, objectness, pred_corners, gt_corners, gt_present, nms_iou, gt_iou,
                         min_objectness, cap):
    import torchvision

    b, q, _ = boxes2d.shape
    cand = torch.full((b, cap), -1, dtype=torch.int32)
    count = torch.zeros((b, 2), dtype=torch.int32)
    for i in range(b):
        v = valid[i].bool()
        scores = torch.where(v, objectness[i].float(), torch.full_like(objectness[i].float(), -1.0))
        box = boxes2d[i].float().clone()
        box[~v] = torch.tensor([0.0, 0.0, 2.0, 2.0])
        keep = torchvision.ops.nms(box, scores, iou_threshold=nms_iou)
        gsel = _aabb(gt_corners[i][gt_present[i] > 0].float()) if (gt_present[i] > 0).any() else None
        out = []
        for k in keep.tolist():
            if gsel is not None:
                pk = _aabb(pred_corners[i, k:k + 1].float())[0]
                if any(float(_iou6(pk, gk)) > gt_iou for gk in gsel):
                    continue
            if float(scores[k]) < min_objectness or not bool(v[k]):
                continue
            out.append(k)
        count[i, 0], count[i, 1] = min(len(out), cap), len(out)
        cand[i, : min(len(out), cap)] = torch.tensor(out[:cap], dtype=torch.int32)
    return cand, count
