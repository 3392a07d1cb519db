"""CPU restatement of the cosine classifier of oadp.dp -- TEST INFRASTRUCTURE.

Parity unpinned by the reference (no tests upstream).  Follows, line by line:
  * `NormalizedLinear.forward`  oadp/dp/utils.py:47-51      F.normalize(x W^T + b)
  * `BaseClassifier.embeddings` oadp/dp/classifiers.py:49-57  text rows AS STORED (not re-normalised),
                                                             plus F.normalize(bg) when out == num_all + 1
  * `BaseClassifier.forward`    classifiers.py:59-68         y = h E^T; training: novel columns = -inf
  * `Classifier.forward`        classifiers.py:82-83         y * scaler - bias
  * `ViLDClassifier.forward`    classifiers.py:105-112       y / scaler[train|val]
  * `ObjectMixin.forward`       oadp/dp/bbox_heads.py:57-60  last logit = -inf
Plain differentiable torch fp32, so autograd provides the reference gradients.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn.functional as F


def normalized_linear(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    return F.normalize(x @ weight.T + bias)


def embeddings(text: torch.Tensor, bg: Optional[torch.Tensor]) -> torch.Tensor:
    if bg is None:
        return text
    return torch.cat([text, F.normalize(bg)])


def base_forward(x, weight, bias, text, bg, training: bool, num_bases: int, num_all: int
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (logits (N,K), hooked tensor h (N,512))."""
    h = normalized_linear(x, weight, bias)
    y = h @ embeddings(text, bg).T
    if training:
        y = y.clone()
        y[:, num_bases:num_all] = float('-inf')
    return y, h


def classifier_forward(x, weight, bias, text, bg, training, num_bases, num_all, scaler: float, shift: float):
    y, h = base_forward(x, weight, bias, text, bg, training, num_bases, num_all)
    return y * scaler - shift, h


def vild_forward(x, weight, bias, text, bg, training, num_bases, num_all, scaler_train=0.007, scaler_val=0.01):
    y, h = base_forward(x, weight, bias, text, bg, training, num_bases, num_all)
    return y / (scaler_train if training else scaler_val), h


def object_head_logits(logits: torch.Tensor) -> torch.Tensor:
    out = logits.clone()
    out[:, -1] = float('-inf')
    return out


def vild_ensemble(bbox_logits: torch.Tensor, object_logits: torch.Tensor, lambda_: torch.Tensor) -> torch.Tensor:
    """oadp/dp/roi_heads.py:93-112 (inference branch of ViLDEnsembleRoIHead._bbox_forward), line by
    line: softmax ** lambda, softmax ** (1 - lambda), product, background = 1 - sum, log."""
    bbox_scores = bbox_logits.softmax(-1)**lambda_
    object_scores = object_logits.softmax(-1)**(1 - lambda_)
    cls_score = bbox_scores * object_scores
    cls_score[:, -1] = 1 - cls_score[:, :-1].sum(-1)
    return cls_score.log()
