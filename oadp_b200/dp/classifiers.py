"""Cosine classifier -- counterpart of oadp/dp/classifiers.py + `NormalizedLinear` (utils.py:47-51).

Same names, constructor signature (`prompts`, `in_features`, `out_features`[, `scaler`]), attribute
names the rest of the reference reaches into (`_linear`, `_bg_embedding`, `_embeddings`,
`_scaler`) and call-time reads of `Globals.training` / `Globals.categories`.  `_linear` stays an
`nn.Module` invoked through `__call__` whose output is the L2-normalised (N,512) tensor, so the todd
distiller forward hooks of configs/dp/models/*.py keep working unchanged.

On a CUDA device both halves run on liboake_b200 (tcgen05 GEMMs + fused row kernels, forward and
backward); there is no PyTorch fallback on CUDA.  CPU tensors raise: the reference's CPU debug mode
is out of scope for this library (use the oracle in tests).
Registered in mmdet's `LINEAR_LAYERS` when mmdet is importable, else in the registry of the same
name in `oadp_b200.registry` (same decorator form, same `build_linear_layer`).
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Optional

import torch
import torch.nn as nn

from .. import binding
from ..registry import LINEAR_LAYERS
from .categories import Globals

__all__ = ['BaseClassifier', 'Classifier', 'ViLDClassifier', 'NormalizedLinear']

DIM = 512
_DTYPE_CODES = {torch.float32: binding.DTYPE_F32, torch.float16: binding.DTYPE_F16, torch.bfloat16: binding.DTYPE_BF16}


def _kpad(k: int) -> int:
    return (k + 127) // 128 * 128


class _Workspace:
    _cache: Dict[int, torch.Tensor] = {}

    @classmethod
    def get(cls, device: torch.device, n: int, in_features: int, k_pad: int) -> torch.Tensor:
        need = C.c_size_t()
        binding.check(binding.load().oake_classifier_workspace_bytes(n, in_features, k_pad, C.byref(need)))
        key = device.index or 0
        buf = cls._cache.get(key)
        if buf is None or buf.numel() < need.value:
            buf = torch.empty(need.value, dtype=torch.uint8, device=device)
            cls._cache[key] = buf
        return buf


def _require_cuda(t: torch.Tensor) -> None:
    if not t.is_cuda:
        raise binding.OakeError('oadp_b200 classifier kernels need CUDA tensors (sm_100a); there is no CPU path')


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


class _NormalizedLinearFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
        _require_cuda(x)
        lib = binding.load()
        x32 = x.detach().float().contiguous()
        w32 = weight.detach().float().contiguous()
        b32 = bias.detach().float().contiguous()
        n, in_f = x32.shape
        h = torch.empty(n, DIM, device=x.device, dtype=torch.float32)
        inv = torch.empty(n, device=x.device, dtype=torch.float32)
        ws = _Workspace.get(x.device, n, in_f, 128)
        binding.check(lib.oake_normalized_linear_fwd(x32.data_ptr(), w32.data_ptr(), b32.data_ptr(), n, in_f,
                                                     h.data_ptr(), inv.data_ptr(), ws.data_ptr(), ws.numel(),
                                                     _stream(x)))
        ctx.save_for_backward(x32, w32, h, inv)
        ctx.in_dtype = x.dtype
        return h

    @staticmethod
    def backward(ctx, dh: torch.Tensor):
        x32, w32, h, inv = ctx.saved_tensors
        lib = binding.load()
        n, in_f = x32.shape
        dh = dh.float().contiguous()
        need_x, need_w, need_b = ctx.needs_input_grad
        dx = torch.empty_like(x32) if need_x else None
        dw = torch.empty_like(w32) if need_w else None
        db = torch.empty(DIM, device=x32.device) if need_b else None
        ws = _Workspace.get(x32.device, n, in_f, 128)
        binding.check(lib.oake_normalized_linear_bwd(
            x32.data_ptr(), w32.data_ptr(), h.data_ptr(), inv.data_ptr(), dh.data_ptr(), n, in_f,
            dx.data_ptr() if need_x else None, dw.data_ptr() if need_w else None, db.data_ptr() if need_b else None,
            ws.data_ptr(), ws.numel(), _stream(dh)))
        return (dx.to(ctx.in_dtype) if need_x else None), dw, db


class _CosineLogitsFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, h: torch.Tensor, text: torch.Tensor, bg: Optional[torch.Tensor], alpha: float, shift: float,
                ninf_lo: int, ninf_hi: int) -> torch.Tensor:
        _require_cuda(h)
        lib = binding.load()
        h32 = h.detach().float().contiguous()
        text32 = text.detach().float().contiguous()
        bg32 = bg.detach().float().contiguous().reshape(-1) if bg is not None else None
        n, num_all = h32.shape[0], text32.shape[0]
        k = num_all + (1 if bg is not None else 0)
        k_pad = _kpad(k)
        logits = torch.empty(n, k_pad, device=h.device, dtype=torch.float32)
        ws = _Workspace.get(h.device, n, 1024, k_pad)
        binding.check(lib.oake_cosine_logits_fwd(h32.data_ptr(), text32.data_ptr(),
                                                 bg32.data_ptr() if bg32 is not None else None, n, num_all, k_pad,
                                                 alpha, shift, ninf_lo, ninf_hi, logits.data_ptr(), ws.data_ptr(),
                                                 ws.numel(), _stream(h)))
        ctx.save_for_backward(h32, text32, bg32 if bg32 is not None else torch.empty(0, device=h.device))
        ctx.meta = (bg is not None, alpha, ninf_lo, ninf_hi, k, k_pad, bg.shape if bg is not None else None)
        return logits[:, :k].contiguous()  # callers write -inf into it in place (bbox_heads.py:59)

    @staticmethod
    def backward(ctx, dlogits: torch.Tensor):
        h32, text32, bg32 = ctx.saved_tensors
        has_bg, alpha, ninf_lo, ninf_hi, k, k_pad, bg_shape = ctx.meta
        lib = binding.load()
        n, num_all = h32.shape[0], text32.shape[0]
        dl = torch.zeros(n, k_pad, device=h32.device, dtype=torch.float32)
        dl[:, :k] = dlogits
        dh = torch.empty(n, DIM, device=h32.device, dtype=torch.float32)
        want_bg = has_bg and ctx.needs_input_grad[2]
        bg_dead = has_bg and ninf_lo <= num_all < ninf_hi  # the background column was forced to -inf: no gradient
        if want_bg and bg_dead:
            want_bg = False
        dbg = torch.empty(DIM, device=h32.device) if want_bg else None
        ws = _Workspace.get(h32.device, n, 1024, k_pad)
        binding.check(lib.oake_cosine_logits_bwd(h32.data_ptr(), text32.data_ptr(),
                                                 bg32.data_ptr() if has_bg else None, dl.data_ptr(), n, num_all,
                                                 k_pad, alpha, ninf_lo, ninf_hi, dh.data_ptr(),
                                                 dbg.data_ptr() if want_bg else None, ws.data_ptr(), ws.numel(),
                                                 _stream(dl)))
        if has_bg and ctx.needs_input_grad[2] and bg_dead:
            return dh, None, torch.zeros(bg_shape, device=h32.device), None, None, None, None
        return dh, None, (dbg.reshape(bg_shape) if want_bg else None), None, None, None, None


class NormalizedLinear(nn.Linear):
    """F.normalize(Linear(x)) (utils.py:47-51); the hooked module."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _NormalizedLinearFn.apply(x, self.weight, self.bias)


@LINEAR_LAYERS.register_module()
class BaseClassifier(nn.Module):

    def __init__(self, *args, prompts: str, in_features: int, out_features: int, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        prompts_ = torch.load(prompts, 'cpu')
        names = list(prompts_['names'])
        embeddings: torch.Tensor = prompts_['embeddings']
        indices = [names.index(name) for name in Globals.categories.all_]
        embeddings = embeddings[indices].float()
        num_all = Globals.categories.num_all
        if out_features == num_all + 1:  # with a learnable background row
            bg_embedding = nn.Parameter(torch.zeros(1, embeddings.shape[1]))
            nn.init.xavier_uniform_(bg_embedding)
        elif out_features == num_all:
            bg_embedding = None
        else:
            raise RuntimeError(str(out_features))
        self._prompts = prompts_
        self.register_buffer('_embeddings', embeddings, persistent=False)
        self._bg_embedding = bg_embedding
        self._linear = NormalizedLinear(in_features, embeddings.shape[1])
        # set by `ObjectMixin` (bbox_heads.py:57-60 writes -inf into the last column after every forward):
        # the background column then leaves the logits kernel as -inf -- the same values, one launch fewer
        self.disable_bg_column = False

    @property
    def embeddings(self) -> torch.Tensor:
        """Materialised (K,512) table, for code that reads it; the kernels build it on the fly."""
        if self._bg_embedding is None:
            return self._embeddings
        return torch.cat([self._embeddings, nn.functional.normalize(self._bg_embedding)])

    def _affine(self) -> tuple:
        return 1.0, 0.0

    # ------------------------------------------------------------------------------ inference fast path
    def _prepared(self, device: torch.device, k_pad: int):
        """W and E in the tensor-core type, re-made only when a parameter changed (`_version`) or moved."""
        lin, bg = self._linear, self._bg_embedding
        key = (lin.weight._version, lin.weight.data_ptr(), self._embeddings.data_ptr(),
               None if bg is None else (bg._version, bg.data_ptr()), k_pad, str(device))
        cache = getattr(self, '_prepared_cache', None)
        if cache is None or cache[0] != key:
            act = torch.float16 if binding.act_dtype_name() == 'f16' else torch.bfloat16
            w_act = torch.empty(DIM, lin.in_features, device=device, dtype=act)
            e_act = torch.empty(k_pad, DIM, device=device, dtype=act)
            w32 = lin.weight.detach().float().contiguous()
            t32 = self._embeddings.detach().float().contiguous()
            b32 = bg.detach().float().contiguous().reshape(-1) if bg is not None else None
            binding.check(binding.load().oake_classifier_prepare(
                w32.data_ptr(), lin.in_features, t32.data_ptr(), b32.data_ptr() if b32 is not None else None,
                t32.shape[0], k_pad, w_act.data_ptr(), e_act.data_ptr(), _stream(w_act)))
            cache = (key, w_act, e_act, lin.bias.detach().float().contiguous())
            self._prepared_cache = cache
        return cache[1], cache[2], cache[3]

    def _fast_path_ok(self, x: torch.Tensor) -> bool:
        """No gradient wanted and nobody is listening on `_linear`: the two modules may run as one call."""
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return False
        lin = self._linear
        if lin._forward_hooks or lin._forward_pre_hooks or self._forward_hooks or self._forward_pre_hooks:
            return False
        import torch.nn.modules.module as _m
        if _m._global_forward_hooks or _m._global_forward_pre_hooks:
            return False
        return x.is_cuda and x.dim() == 2 and x.dtype in _DTYPE_CODES

    def _ninf_range(self) -> tuple:
        num_all = Globals.categories.num_all
        lo = hi = 0
        if Globals.training:  # novel categories are invisible while training (classifiers.py:62-67)
            lo, hi = Globals.categories.num_bases, num_all
        if self.disable_bg_column and self._bg_embedding is not None:
            lo, hi = (lo if hi else num_all), num_all + 1  # [novels |] background: one contiguous -inf range
        return lo, hi

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        lo, hi = self._ninf_range()
        alpha, shift = self._affine()
        if self._fast_path_ok(x):
            # one C-ABI call: [cast x] -> GEMM -> row normalise -> GEMM (oake_classifier_fwd), prepared W / E
            _require_cuda(x)
            x = x.contiguous()
            n = x.shape[0]
            k = self._embeddings.shape[0] + (1 if self._bg_embedding is not None else 0)
            k_pad = _kpad(k)
            w_act, e_act, b32 = self._prepared(x.device, k_pad)
            h = torch.empty(n, DIM, device=x.device, dtype=torch.float32)
            logits = torch.empty(n, k_pad, device=x.device, dtype=torch.float32)
            ws = _Workspace.get(x.device, n, self._linear.in_features, k_pad)
            binding.check(binding.load().oake_classifier_fwd(
                x.data_ptr(), _DTYPE_CODES[x.dtype], w_act.data_ptr(), b32.data_ptr(), e_act.data_ptr(), n,
                self._linear.in_features, k_pad, alpha, shift, lo, hi, h.data_ptr(), logits.data_ptr(), ws.data_ptr(),
                ws.numel(), _stream(x)))
            return logits[:, :k]  # (a view: the padded pitch is what `vild_ensemble` takes as is)
        h = self._linear(x)  # through __call__: forward hooks see the normalised tensor
        return _CosineLogitsFn.apply(h, self._embeddings, self._bg_embedding, alpha, shift, lo, hi)


@LINEAR_LAYERS.register_module()
class Classifier(BaseClassifier):
    """logits * scaler - bias, both read once from the prompts file (classifiers.py:71-83)."""

    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self._scaler = self._prompts['scaler'].item()
        self._bias = self._prompts['bias'].item()

    def _affine(self) -> tuple:
        return float(self._scaler), float(self._bias)


@LINEAR_LAYERS.register_module()
class ViLDClassifier(BaseClassifier):
    """logits / scaler, scaler chosen by Globals.training at call time (classifiers.py:91-112)."""

    def __init__(self, *args, scaler: Optional[Dict[str, float]] = None, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        if scaler is None:
            scaler = dict(train=0.007, val=0.01)  # the reference's defaults (inverse of its configs)
        self._scaler = scaler

    @property
    def scaler(self) -> float:
        return self._scaler['train'] if Globals.training else self._scaler['val']

    def _affine(self) -> tuple:
        return 1.0 / float(self.scaler), 0.0
