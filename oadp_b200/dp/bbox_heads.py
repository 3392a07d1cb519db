"""Block / object bbox heads -- counterpart of oadp/dp/bbox_heads.py (registry surface).

    NotWithRegMixin                         bbox_heads.py:20-24   `with_reg` forced to False
    BlockMixin                              bbox_heads.py:27-42   AsymmetricLoss on sigmoid(logits) + top-k recall
    ObjectMixin                             bbox_heads.py:45-60   frozen background row, last logit := -inf
    Shared2FCBlockBBoxHead                  bbox_heads.py:63-65   HEADS
    Shared4Conv1FCObjectBBoxHead            bbox_heads.py:68-70   HEADS

The bases are mmdet's `BBoxHead` / `Shared2FCBBoxHead` / `Shared4Conv1FCBBoxHead` when mmdet is importable and
the structural stand-ins of `oadp_b200.mmdet_shim` otherwise; `fc_cls` is whatever `cls_predictor_cfg` names --
with the reference's configs the cosine classifiers of `oadp_b200.dp.classifiers` (liboake_b200 kernels), whose
`_linear` output the todd distiller hooks read (`.roi_head._object_head.fc_cls._linear`,
`.roi_head._block_head.fc_cls._linear`; configs/dp/models/*.py).
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import torch

from ..registry import HAVE_MMDET, HEADS, LossRegistry
from . import losses as _losses  # noqa: F401  (registers AsymmetricLoss / RKDLoss / L1Loss / MSELoss)
from .classifiers import BaseClassifier
from .utils import MultilabelTopKRecall

if HAVE_MMDET:  # pragma: no cover - mmdet is not installed in this environment
    from mmdet.models import BBoxHead, Shared2FCBBoxHead, Shared4Conv1FCBBoxHead
else:
    from ..mmdet_shim import BBoxHead, Shared2FCBBoxHead, Shared4Conv1FCBBoxHead

__all__ = ['BlockMixin', 'ObjectMixin', 'Shared2FCBlockBBoxHead', 'Shared4Conv1FCObjectBBoxHead']


class NotWithRegMixin(BBoxHead):
    """Override the `with_reg` argument to False (bbox_heads.py:20-24)."""

    def __init__(self, *args: Any, with_reg: bool = False, **kwargs: Any) -> None:
        super().__init__(*args, with_reg=False, **kwargs)


class BlockMixin(NotWithRegMixin):

    def __init__(self, *args: Any, topk: int, loss: Dict[str, Any], **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self._multilabel_topk_recall = MultilabelTopKRecall(k=topk)
        self._loss = LossRegistry.build(loss)

    def loss(self, logits: torch.Tensor, targets: torch.Tensor) -> Dict[str, torch.Tensor]:
        """bbox_heads.py:34-42.  The caller drops the background column first (`logits[:, :-1]`,
        roi_heads.py:208); the loss takes probabilities, the recall takes the logits."""
        return dict(loss_block=self._loss(logits.sigmoid(), targets),
                    recall_block=self._multilabel_topk_recall(logits, targets))


class ObjectMixin(NotWithRegMixin):

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        # `_bg_embedding` does not get trained, and will not be used during inference (bbox_heads.py:50-55)
        classifier = self.fc_cls
        bg_embedding: Optional[torch.Tensor] = classifier._bg_embedding
        assert bg_embedding is not None
        bg_embedding.requires_grad_(False)
        # a liboake classifier writes the -inf of `forward` below itself, inside its logits kernel
        self._fused_bg = isinstance(classifier, BaseClassifier)
        if self._fused_bg:
            classifier.disable_bg_column = True

    def forward(self, *args: Any, **kwargs: Any) -> Tuple[torch.Tensor, None]:
        logits, _ = super().forward(*args, **kwargs)
        if not self._fused_bg:
            logits[:, -1] = float('-inf')  # disable `_bg_embedding` (bbox_heads.py:57-60)
        return logits, None


@HEADS.register_module()
class Shared2FCBlockBBoxHead(BlockMixin, Shared2FCBBoxHead):
    pass


@HEADS.register_module()
class Shared4Conv1FCObjectBBoxHead(ObjectMixin, Shared4Conv1FCBBoxHead):
    pass
