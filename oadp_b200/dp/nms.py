"""Re-softmax + multiclass NMS behind the ViLD ensemble (second half).

Stands where mmdet's `BBoxHead.get_bboxes` calls `F.softmax(cls_score)` and
`mmdet.core.multiclass_nms(bboxes, scores, score_thr, nms, max_per_img)` on the `cls_score` that
`ViLDEnsembleRoIHead._bbox_forward` returns (oadp/dp/roi_heads.py:93-112; configs: score_thr 0.0, iou 0.5,
max_per_img 300), and where oadp/dp/test_nni.py:55-92 calls the same function on its re-weighted scores.
Same signature and return convention as mmdet's function -- `(dets (M,5) = [x1, y1, x2, y2, score], labels (M,))`
sorted by descending score, at most `max_num` -- for class-agnostic boxes (`multi_bboxes` (N,4):
`reg_class_agnostic=True` in the reference's configs).  The overlap relation of the N boxes is computed once and
shared by all classes (oadp_b200/csrc/nms.cu); the final top-`max_num` is a library sort of the few kept scores.
CUDA tensors only."""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Mapping, Optional, Tuple

import torch

from .. import binding

_WS: Dict[int, torch.Tensor] = {}


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def softmax_rows(cls_score: torch.Tensor) -> torch.Tensor:
    """`F.softmax(cls_score, dim=-1)` of mmdet `BBoxHead.get_bboxes` (fp32, one kernel; the padded row pitch of the
    classifier / ensemble outputs is accepted as is)."""
    if not cls_score.is_cuda:
        raise RuntimeError('softmax_rows runs on liboake_b200 (CUDA tensors only)')
    if cls_score.dim() != 2 or cls_score.dtype != torch.float32 or cls_score.stride(1) != 1:
        raise ValueError('cls_score must be a (N, K+1) fp32 matrix with unit column stride')
    n, k1 = cls_score.shape
    out = torch.empty(n, k1, device=cls_score.device, dtype=torch.float32)
    binding.check(binding.load().oake_softmax_rows(cls_score.data_ptr(), n, k1, cls_score.stride(0) if n > 1 else k1,
                                                   out.data_ptr(), k1, _stream(cls_score)))
    return out


def multiclass_nms(multi_bboxes: torch.Tensor, multi_scores: torch.Tensor, score_thr: float, nms_cfg: Mapping[str, Any],
                   max_num: int = -1, score_factors: Optional[torch.Tensor] = None,
                   return_inds: bool = False) -> Tuple[torch.Tensor, ...]:
    """mmdet.core.post_processing.multiclass_nms.  multi_scores (N, K+1): the last column is the background."""
    if not (multi_bboxes.is_cuda and multi_scores.is_cuda):
        raise RuntimeError('multiclass_nms runs on liboake_b200 (CUDA tensors only); there is no CPU fallback')
    if multi_bboxes.dim() != 2 or multi_bboxes.shape[1] != 4:
        raise NotImplementedError('class-specific boxes (N, K*4) are not built: the reference regresses class-agnostic '
                                  'boxes (reg_class_agnostic=True)')
    cfg = dict(nms_cfg)
    if cfg.pop('type', 'nms') != 'nms' or cfg.pop('class_agnostic', False):
        raise NotImplementedError("only nms_cfg = dict(type='nms', iou_threshold=...) is built")
    iou_thr = float(cfg.pop('iou_threshold', cfg.pop('iou_thr', 0.5)))
    n, k = multi_scores.shape[0], multi_scores.shape[1] - 1
    dev = multi_scores.device
    boxes = multi_bboxes.detach().float().contiguous()
    scores = multi_scores.detach().float()
    if score_factors is not None:
        scores = scores * score_factors[:, None]  # mmdet multiplies AFTER the validity test; see below
    if scores.stride(1) != 1:
        scores = scores.contiguous()
    if n == 0 or k <= 0:
        dets, labels = boxes.new_zeros((0, 5)), torch.zeros(0, dtype=torch.long, device=dev)
        return (dets, labels, labels.clone()) if return_inds else (dets, labels)
    if score_factors is not None:
        # validity is decided on the raw scores (`valid_mask = scores > score_thr` precedes the multiplication)
        raw = multi_scores.detach().float()
        scores = torch.where(raw > score_thr, scores, torch.full_like(scores, float('-inf')))
        thr = float('-inf')
    else:
        thr = float(score_thr)
    need = C.c_size_t()
    binding.check(binding.load().oake_nms_workspace_bytes(n, C.byref(need)))
    key = dev.index or 0
    ws = _WS.get(key)
    if ws is None or ws.numel() < need.value:
        ws = _WS[key] = torch.empty(need.value, dtype=torch.uint8, device=dev)
    keep = torch.empty(k, n, dtype=torch.uint8, device=dev)
    binding.check(binding.load().oake_multiclass_nms(boxes.data_ptr(), scores.data_ptr(), n, k,
                                                     scores.stride(0) if n > 1 else scores.shape[1], thr, iou_thr,
                                                     keep.data_ptr(), ws.data_ptr(), ws.numel(), _stream(scores)))
    cls_idx, box_idx = keep.nonzero(as_tuple=True)
    kept = scores[box_idx, cls_idx]
    flat = box_idx * k + cls_idx  # mmdet's candidate index (RoI major, class minor)
    # descending scores; equal scores in candidate order
    order = torch.argsort(flat)
    kept, cls_idx, box_idx, flat = kept[order], cls_idx[order], box_idx[order], flat[order]
    order = torch.argsort(kept, descending=True, stable=True)
    if max_num > 0:
        order = order[:max_num]
    dets = torch.cat([boxes[box_idx[order]], kept[order, None]], dim=1)
    labels = cls_idx[order]
    return (dets, labels, flat[order]) if return_inds else (dets, labels)


def ensemble_detections(cls_score: torch.Tensor, bboxes: torch.Tensor, score_thr: float = 0.0,
                        nms: Optional[Mapping[str, Any]] = None, max_per_img: int = 300) -> Tuple[torch.Tensor, torch.Tensor]:
    """What mmdet does with the `cls_score` of `ViLDEnsembleRoIHead._bbox_forward` and the decoded boxes:
    softmax again, multiclass NMS, keep `max_per_img` (test_cfg.rcnn of vild_ensemble_faster_rcnn_r50_fpn.py:41-44)."""
    return multiclass_nms(bboxes, softmax_rows(cls_score), score_thr, nms or dict(type='nms', iou_threshold=0.5), max_per_img)
