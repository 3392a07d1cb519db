"""RoI heads -- counterpart of oadp/dp/roi_heads.py (registry surface).

    ViLDEnsembleRoIHead     roi_heads.py:20-166   builds `_object_head`; at inference replaces the bbox head's
                                                  class scores by the ViLD ensemble of the two classifiers
    OADPRoIHead             roi_heads.py:169-209  + `_block_head`, `block_forward_train`

Both derive from mmdet's `StandardRoIHead` when mmdet is importable and from the structural stand-in of
`oadp_b200.mmdet_shim` otherwise (RoIAlign, box coding, samplers, NMS stay mmdet's).  What is this package's
own is the arithmetic of `_bbox_forward` at inference (roi_heads.py:93-112): directly downstream of the two
cosine classifier calls (`bbox_head.fc_cls`, `_object_head.fc_cls`), the six tensor lines
`softmax^lambda * softmax^(1-lambda)`, background fix-up, `log` run as ONE kernel (`vild_ensemble`,
oadp_b200/csrc/ensemble.cu).  Attribute names `_object_head` / `_block_head` are part of the contract: the
distiller hook paths of configs/dp/models/*.py go through them.
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional, Sequence

import torch

from .. import binding
from ..registry import HAVE_MMDET, HEADS
from .bbox_heads import BlockMixin, ObjectMixin
from .categories import Globals

if HAVE_MMDET:  # pragma: no cover - mmdet is not installed in this environment
    from mmdet.core import bbox2roi
    from mmdet.models import StandardRoIHead
else:
    from ..mmdet_shim import StandardRoIHead, bbox2roi

__all__ = ['ViLDEnsembleRoIHead', 'OADPRoIHead', 'vild_ensemble', 'ensemble_lambda']


def ensemble_lambda(num_bases: int, num_all: int, device=None) -> torch.Tensor:
    """`_lambda` buffer of roi_heads.py:55-59: 2/3 for base categories, 1/3 for novel + background."""
    lam = torch.ones(num_all + 1, device=device) / 3
    lam[:num_bases] *= 2
    return lam


def vild_ensemble(bbox_logits: torch.Tensor, object_logits: torch.Tensor, lambda_: torch.Tensor) -> torch.Tensor:
    """log(softmax(bbox)^lambda * softmax(object)^(1-lambda)) with the background column replaced by
    log(1 - sum of the others).  (N, K+1) fp32 CUDA tensors (row pitch may exceed K+1: the padded
    logits of `oake_cosine_logits_fwd` are accepted through `.stride(0)`); one kernel launch."""
    if not (bbox_logits.is_cuda and object_logits.is_cuda and lambda_.is_cuda):
        raise RuntimeError('vild_ensemble runs on liboake_b200 (CUDA tensors only); there is no CPU fallback')
    if bbox_logits.shape != object_logits.shape or bbox_logits.dim() != 2:
        raise ValueError(f'shape mismatch: {tuple(bbox_logits.shape)} vs {tuple(object_logits.shape)}')
    n, k1 = bbox_logits.shape
    if lambda_.numel() != k1:
        raise ValueError(f'lambda has {lambda_.numel()} entries, logits have {k1} columns')
    for t in (bbox_logits, object_logits):
        if t.dtype != torch.float32 or t.stride(1) != 1:
            raise ValueError('logits must be fp32 with unit column stride')
    lam = lambda_.to(torch.float32).contiguous()
    out = torch.empty(n, k1, device=bbox_logits.device, dtype=torch.float32)
    stream = torch.cuda.current_stream(bbox_logits.device).cuda_stream
    binding.check(binding.load().oake_vild_ensemble(bbox_logits.data_ptr(), object_logits.data_ptr(), lam.data_ptr(), n,
                                                    k1, bbox_logits.stride(0) if n > 1 else k1,
                                                    object_logits.stride(0) if n > 1 else k1, out.data_ptr(), k1, stream))
    return out


def _get(cfg: Any, key: str) -> Any:
    return cfg[key] if isinstance(cfg, dict) else getattr(cfg, key)


def _set(cfg: Any, key: str, value: Any) -> None:
    if isinstance(cfg, dict):
        cfg[key] = value
    else:
        setattr(cfg, key, value)


@HEADS.register_module()
class ViLDEnsembleRoIHead(StandardRoIHead):

    def __init__(self, *args: Any, bbox_head: Dict[str, Any], object_head: Dict[str, Any],
                 mask_head: Optional[Dict[str, Any]] = None, **kwargs: Any) -> None:
        # automatically detect `num_classes` (roi_heads.py:32-37)
        assert _get(bbox_head, 'num_classes') is None
        _set(bbox_head, 'num_classes', Globals.categories.num_all)
        if mask_head is not None:
            assert _get(mask_head, 'num_classes') is None
            _set(mask_head, 'num_classes', Globals.categories.num_all)
        super().__init__(*args, bbox_head=bbox_head, mask_head=mask_head, **kwargs)
        assert not self.with_shared_head  # `shared_head` is not supported for simplification
        self._object_head: ObjectMixin = HEADS.build(object_head, default_args=bbox_head)
        # lambda for base and novel categories are 2/3 and 1/3, respectively (roi_heads.py:55-59)
        self.register_buffer('_lambda', ensemble_lambda(Globals.categories.num_bases, Globals.categories.num_all),
                             persistent=False)

    @property
    def lambda_(self) -> torch.Tensor:
        return self._lambda

    def _bbox_forward(self, x: Sequence[torch.Tensor], rois: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Training: as `StandardRoIHead`.  Inference: the bbox head's class scores are replaced by the
        calibrated ensemble with the object head (roi_heads.py:64-112)."""
        bbox_results: Dict[str, torch.Tensor] = super()._bbox_forward(x, rois)
        if Globals.training:
            return bbox_results
        bbox_logits = bbox_results['cls_score']
        object_logits, _ = self._object_head(bbox_results['bbox_feats'])
        if os.environ.get('DUMP'):  # `Store.DUMP` (globals_.py:14-16): the NNI search reads these
            self._bbox_logits = bbox_logits
            self._object_logits = object_logits
        bbox_results['cls_score'] = vild_ensemble(bbox_logits.float(), object_logits.float(), self.lambda_)
        return bbox_results

    def _object_forward(self, x: Sequence[torch.Tensor], rois: torch.Tensor) -> None:
        bre = self.bbox_roi_extractor
        object_feats = bre(x[:bre.num_inputs], rois)
        self._object_head(object_feats)  # the output is dropped: the distiller hook on `fc_cls._linear` has it

    def object_forward_train(self, x: Sequence[torch.Tensor], bboxes: List[torch.Tensor]) -> None:
        self._object_forward(x, bbox2roi(bboxes))


@HEADS.register_module()
class OADPRoIHead(ViLDEnsembleRoIHead):

    def __init__(self, *args: Any, bbox_head: Dict[str, Any], block_head: Optional[Dict[str, Any]] = None,
                 **kwargs: Any) -> None:
        super().__init__(*args, bbox_head=bbox_head, **kwargs)
        if block_head is not None:
            self._block_head: BlockMixin = HEADS.build(block_head, default_args=bbox_head)

    @property
    def with_block(self) -> bool:
        return hasattr(self, '_block_head')

    def _block_forward(self, x: Sequence[torch.Tensor], rois: torch.Tensor) -> torch.Tensor:
        bre = self.bbox_roi_extractor
        block_feats = bre(x[:bre.num_inputs], rois)
        logits, _ = self._block_head(block_feats)
        return logits

    def block_forward_train(self, x: Sequence[torch.Tensor], bboxes: List[torch.Tensor],
                            targets: List[torch.Tensor]) -> Dict[str, torch.Tensor]:
        logits = self._block_forward(x, bbox2roi(bboxes))
        return self._block_head.loss(logits[:, :-1], torch.cat(targets))  # the background column is dropped
