"""Distillation-side losses on liboake_b200 -- counterparts of oadp/base/losses.py
(`AsymmetricLoss`, `RKDLoss`) and of the todd `L1Loss` / `MSELoss` the distiller configs apply to the
hooked `fc_cls._linear` output (configs/dp/models/{vild_ensemble_faster_rcnn_r50_fpn,block,global_}.py).

Same constructor surface as the reference / todd (`reduction`, `weight`; ASL: `gamma_neg`,
`gamma_pos`, `clip`, `eps`).  Each forward is one C-ABI call that produces the scalar loss AND the
gradient with respect to the first argument; backward only scales that gradient.  `weight` may be a
float or a `WarmupScheduler`-style dict `{'type': 'WarmupScheduler', 'gain': g, 'end': e}` -- todd's
scheduler is not visible from the reference; it is taken as a linear ramp `g * min(step / e, 1)`
driven by `loss.step(i)` (default: fully warmed up).  CUDA tensors only.

Registered in todd's `LossRegistry` when todd is importable (its own `L1Loss` / `MSELoss` then keep their names:
only `AsymmetricLoss` / `RKDLoss` are added, as oadp/base/losses.py:10,68 does), else in the registry of the same
name in `oadp_b200.registry`, so that `LossRegistry.build(dict(type='AsymmetricLoss', ...))` of
oadp/dp/bbox_heads.py:31 and the distiller's loss configs resolve.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Union

import torch
import torch.nn as nn

from .. import binding
from ..registry import LossRegistry

_WS: Dict[int, torch.Tensor] = {}


def _workspace(device: torch.device, rkd_rows: int = 0) -> torch.Tensor:
    need = C.c_size_t()
    binding.check(binding.load().oake_loss_workspace_bytes(rkd_rows, C.byref(need)))
    key = device.index or 0
    buf = _WS.get(key)
    if buf is None or buf.numel() < need.value:
        buf = torch.empty(need.value, dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf


def _check(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise binding.OakeError('oadp_b200 losses need CUDA tensors (sm_100a); there is no CPU path')
    return t.detach().float().contiguous()


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


class _LossFn(torch.autograd.Function):
    """kind: 'l1' | 'mse' | 'rkd' | 'asl'.  Returns a 0-dim loss; saves d loss / d first argument."""

    @staticmethod
    def forward(ctx, a: torch.Tensor, b: torch.Tensor, kind: str, scale: float, params: tuple) -> torch.Tensor:
        lib = binding.load()
        a32 = _check(a)
        loss = torch.empty((), device=a.device, dtype=torch.float32)
        grad = torch.empty_like(a32) if a.requires_grad else None
        gp = grad.data_ptr() if grad is not None else None
        if kind in ('l1', 'mse'):
            b32 = _check(b)
            if b32.shape != a32.shape:
                raise ValueError(f'shape mismatch: {tuple(a32.shape)} vs {tuple(b32.shape)}')
            ws = _workspace(a.device)
            binding.check(lib.oake_pair_loss(a32.data_ptr(), b32.data_ptr(), a32.numel(), 0 if kind == 'l1' else 1,
                                             scale, loss.data_ptr(), gp, ws.data_ptr(), ws.numel(), _stream(a)))
        elif kind == 'rkd':
            b32 = _check(b)
            s2, t2 = a32.reshape(-1, a32.shape[-1]), b32.reshape(-1, b32.shape[-1])
            if s2.shape[0] != t2.shape[0]:
                raise ValueError('preds and targets must have the same leading shape (losses.py:101)')
            n = s2.shape[0]
            if s2.shape[1] != t2.shape[1]:  # the kernel walks both with one width: separate Gram passes otherwise
                raise ValueError('RKD on liboake_b200 needs student and teacher rows of the same width')
            ws = _workspace(a.device, n)
            binding.check(lib.oake_rkd_loss(s2.data_ptr(), t2.data_ptr(), n, s2.shape[1], scale, loss.data_ptr(), gp,
                                            ws.data_ptr(), ws.numel(), _stream(a)))
        elif kind == 'asl':
            y8 = b.detach().to(torch.uint8).contiguous()
            if y8.shape != a32.shape:
                raise ValueError(f'shape mismatch: {tuple(a32.shape)} vs {tuple(y8.shape)}')
            ws = _workspace(a.device)
            gn, gpos, clip, eps = params
            binding.check(lib.oake_asymmetric_loss(a32.data_ptr(), y8.data_ptr(), a32.numel(), gn, gpos, clip, eps, scale,
                                                   loss.data_ptr(), gp, ws.data_ptr(), ws.numel(), _stream(a)))
        else:
            raise ValueError(kind)
        ctx.grad = grad
        ctx.in_dtype = a.dtype
        return loss

    @staticmethod
    def backward(ctx, dloss: torch.Tensor):
        g = ctx.grad
        if g is None:
            return None, None, None, None, None
        return (g * dloss).to(ctx.in_dtype), None, None, None, None


def _register_unless_present(cls):
    """todd ships `L1Loss` / `MSELoss` itself: with the real registry those names stay todd's."""
    try:
        LossRegistry.register()(cls)
    except Exception:
        pass
    return cls


class BaseLoss(nn.Module):

    def __init__(self, reduction: str = 'mean', weight: Union[float, Dict[str, Any]] = 1.0, **_: Any) -> None:
        super().__init__()
        if reduction not in ('mean', 'sum'):
            raise ValueError(f"reduction must be 'mean' or 'sum', not {reduction!r}")
        self._reduction = reduction
        self._weight = weight
        self._step = None

    def step(self, i: int) -> None:
        self._step = int(i)

    @property
    def weight(self) -> float:
        w = self._weight
        if isinstance(w, dict):
            gain, end = float(w.get('gain', 1.0)), float(w.get('end', 1))
            return gain if self._step is None else gain * min(self._step / end, 1.0)
        return float(w)

    def _scale(self, numel: int) -> float:
        return self.weight / numel if self._reduction == 'mean' else self.weight


@_register_unless_present
class L1Loss(BaseLoss):

    def forward(self, pred: torch.Tensor, target: torch.Tensor, *_: Any, **__: Any) -> torch.Tensor:
        return _LossFn.apply(pred, target, 'l1', self._scale(pred.numel()), ())


@_register_unless_present
class MSELoss(BaseLoss):

    def forward(self, pred: torch.Tensor, target: torch.Tensor, *_: Any, **__: Any) -> torch.Tensor:
        return _LossFn.apply(pred, target, 'mse', self._scale(pred.numel()), ())


@LossRegistry.register()
class RKDLoss(BaseLoss):
    """oadp/base/losses.py:68-108."""

    def forward(self, preds: torch.Tensor, targets: torch.Tensor, *_: Any, **__: Any) -> torch.Tensor:
        n = preds.numel() // preds.shape[-1]
        return _LossFn.apply(preds, targets, 'rkd', self._scale(n * n), ())


@LossRegistry.register()
class AsymmetricLoss(BaseLoss):
    """oadp/base/losses.py:10-65; `x` are probabilities (callers pass `logits.sigmoid()`), `y` bool."""

    def __init__(self, *args: Any, gamma_neg: float = 4, gamma_pos: float = 1, clip: float = 0.05, eps: float = 1e-8,
                 **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self._params = (float(gamma_neg), float(gamma_pos), float(clip), float(eps))

    def forward(self, x: torch.Tensor, y: torch.Tensor, **__: Any) -> torch.Tensor:
        return _LossFn.apply(x, y, 'asl', self._scale(x.numel()), self._params)
