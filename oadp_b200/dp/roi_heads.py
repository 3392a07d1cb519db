"""ViLD ensemble scoring -- the arithmetic of `ViLDEnsembleRoIHead._bbox_forward` at inference
(oadp/dp/roi_heads.py:93-112), directly downstream of the two cosine classifier calls
(`bbox_head.fc_cls`, `_object_head.fc_cls`).  The RoI head itself (RoIAlign, box coding, NMS)
belongs to mmdet and is out of scope; a maintainer replaces the six
tensor lines of `_bbox_forward` by one `vild_ensemble(...)` call.
"""
from __future__ import annotations

import torch

from .. import binding


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
