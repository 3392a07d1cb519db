"""fp32 CPU restatement of the distillation-side losses -- TEST INFRASTRUCTURE.

  * asymmetric_loss: oadp/base/losses.py:10-65 (multi-label ASL on probabilities, called with
    `logits.sigmoid()` by the block head, bbox_heads.py:34-42, and the global head, detectors.py:48-57)
  * rkd_loss: oadp/base/losses.py:68-108 (MSE between the Gram matrices of student and teacher rows;
    configs/dp/models/block.py:31-38)
  * l1_loss / mse_loss: todd L1Loss / MSELoss on the hooked `_linear` rows vs the cached CLIP rows
    (configs/dp/models/{vild_ensemble_faster_rcnn_r50_fpn,block,global_}.py)
Pinned against the reference's own classes in tests/test_ref_golden.py (value and gradient).
"""
from __future__ import annotations

import torch


def _reduce(loss: torch.Tensor, reduction: str, weight: float) -> torch.Tensor:
    if reduction == 'mean':
        loss = loss.mean()
    elif reduction == 'sum':
        loss = loss.sum()
    return weight * loss


def asymmetric_loss(x: torch.Tensor, y: torch.Tensor, gamma_neg: float = 4, gamma_pos: float = 1, clip: float = 0.05,
                    eps: float = 1e-8, reduction: str = 'mean', weight: float = 1.0) -> torch.Tensor:
    x = x.float()
    comp_x = 1 - x
    if clip > 0:
        comp_x = (comp_x + clip).clamp(max=1)
    loss = y * torch.log(x.clamp(min=eps)) + ~y * torch.log(comp_x.clamp(min=eps))
    if gamma_neg > 0 or gamma_pos > 0:
        with torch.no_grad():
            pt = x * y + comp_x * ~y
            w = torch.pow(1 - pt, gamma_pos * y + gamma_neg * ~y)
        loss = loss * w
    return _reduce(-loss, reduction, weight)


def rkd_loss(preds: torch.Tensor, targets: torch.Tensor, reduction: str = 'mean', weight: float = 1.0) -> torch.Tensor:
    p = preds.reshape(-1, preds.shape[-1])
    t = targets.reshape(-1, targets.shape[-1])
    return _reduce((p @ p.T - t @ t.T)**2, reduction, weight)


def l1_loss(pred: torch.Tensor, target: torch.Tensor, reduction: str = 'mean', weight: float = 1.0) -> torch.Tensor:
    return _reduce((pred - target).abs(), reduction, weight)


def mse_loss(pred: torch.Tensor, target: torch.Tensor, reduction: str = 'mean', weight: float = 1.0) -> torch.Tensor:
    return _reduce((pred - target)**2, reduction, weight)
