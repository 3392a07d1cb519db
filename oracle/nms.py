"""CPU restatement of mmdet 2.25 `multiclass_nms` for class-agnostic boxes -- TEST INFRASTRUCTURE.

mmdet / mmcv are third-party packages that are not vendored in the reference (README.md:36-37) and not
installed here, so this follows their published algorithm (mmdet/core/post_processing/bbox_nms.py,
mmcv/ops/nms.py `batched_nms`): every RoI is a candidate for every foreground class with `score > score_thr`;
NMS runs per class (the class-offset trick and the per-class split of `batched_nms` are equivalent to that);
survivors come back sorted by descending score, truncated to `max_num`.  The per-class greedy suppression is
`torchvision.ops.nms` (IoU > threshold suppresses; the same rule as mmcv's `nms`), which pins the kernel to an
independent implementation.  Call sites in the reference: mmdet `BBoxHead.get_bboxes` behind
oadp/dp/roi_heads.py:93-112, and oadp/dp/test_nni.py:88-91."""
from __future__ import annotations

from typing import Tuple

import torch
from torchvision.ops import nms as tv_nms


def multiclass_nms(multi_bboxes: torch.Tensor, multi_scores: torch.Tensor, score_thr: float, iou_threshold: float,
                   max_num: int = -1) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (dets (M,5), labels (M,), candidate index n * K + k (M,))."""
    n, k = multi_scores.shape[0], multi_scores.shape[1] - 1
    boxes, scores = multi_bboxes.float(), multi_scores[:, :k].float()
    dets, labels, flat = [], [], []
    for c in range(k):
        s = scores[:, c]
        valid = (s > score_thr).nonzero().flatten()
        if valid.numel() == 0:
            continue
        keep = valid[tv_nms(boxes[valid], s[valid], iou_threshold)]
        dets.append(torch.cat([boxes[keep], s[keep, None]], 1))
        labels.append(torch.full((keep.numel(), ), c, dtype=torch.long))
        flat.append(keep * k + c)
    if not dets:
        return boxes.new_zeros((0, 5)), torch.zeros(0, dtype=torch.long), torch.zeros(0, dtype=torch.long)
    dets, labels, flat = torch.cat(dets), torch.cat(labels), torch.cat(flat)
    order = torch.argsort(flat)
    dets, labels, flat = dets[order], labels[order], flat[order]
    order = torch.argsort(dets[:, 4], descending=True, stable=True)
    if max_num > 0:
        order = order[:max_num]
    return dets[order], labels[order], flat[order]
