"""`MultilabelTopKRecall` of oadp/dp/utils.py:13-44 without the host round trip.

The reference computes the training-time `recall_block` / `recall_global` metric by copying targets
and top-k predictions to the CPU and calling `sklearn.metrics.recall_score(..., average='macro',
labels=<labels that occur>, zero_division=0)` every iteration (bbox_heads.py:41, detectors.py:56): a
device synchronisation per head per step.  The same number follows from three reductions on the
device: per label, recall = TP / (TP + FN) = |pred & target| / |target|, averaged over the labels that
occur at least once in `targets`.
"""
from __future__ import annotations

import torch
import torch.nn as nn

__all__ = ['MultilabelTopKRecall', 'NormalizedLinear']


class MultilabelTopKRecall(nn.Module):

    def __init__(self, *args, k: int, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self._k = k

    def forward(self, logits: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        """logits (bs, K) float, targets (bs, K) bool -> one-element tensor, recall in percent."""
        _, indices = logits.topk(self._k)
        preds = torch.zeros_like(targets).scatter(1, indices, 1)
        targets = targets.bool()
        support = targets.sum(0)  # occurrences per label
        hits = (preds.bool() & targets).sum(0)
        present = support > 0
        n_present = present.sum()
        per_label = hits.to(torch.float64) / support.clamp(min=1).to(torch.float64)
        # no label at all: the reference's macro average over an empty label set is NaN -- kept
        recall = torch.where(present, per_label, torch.zeros_like(per_label)).sum() / n_present
        return (recall * 100).to(logits.dtype)


def __getattr__(name: str):
    # `NormalizedLinear` is listed in the reference's oadp/dp/utils.py:47-51; here it lives next to the
    # autograd functions it calls (classifiers.py) -- resolved lazily, the two modules import each other
    if name == 'NormalizedLinear':
        from .classifiers import NormalizedLinear
        return NormalizedLinear
    raise AttributeError(name)
