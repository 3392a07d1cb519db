"""Host-side model object that stands where `clip.model.CLIP` stands in the reference.

The reference's OAKE validators only touch this surface of the CLIP model
(oadp/oake/globals.py:47-57, blocks.py:123-129, objects.py:281-330):

    model.encode_image(x)            (B,3,224,224) -> (B,512)
    model.visual(objects, masks)     objects variant after the surgery of objects.py:285-314
    model.visual.grid                7, or 14 after the surgery
    model.visual.patch_size          32
    model.dtype

`OakeModel` provides exactly that, backed by liboake_b200.so: weights live in one device buffer in
the tensor-core element type, every forward is a stream-ordered sequence of sm_100a kernels, and the
fused `F.normalize(...).half()` result the validators store is available from `embed()`.
"""
from __future__ import annotations

import os

import ctypes as C
import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from . import binding

Params = Dict[str, torch.Tensor]

WIDTH, HEADS, PATCH, OUT_DIM, IMAGE, GRID = 768, 12, 32, 512, 224, 7


def resample_positional_embedding(pos: torch.Tensor, grid_to: int, mode: str = 'bilinear') -> torch.Tensor:
    """(1+g*g, D) -> (1+grid_to^2, D): class row kept, grid rows interpolated.

    Stands in for the fork-only `visual.interpolate_positional_embedding` called at
    oadp/oake/objects.py:293-296, whose mode is not visible from the reference;
    `mode` is an explicit knob, recorded by the feature-store writer."""
    g = int(round(math.sqrt(pos.shape[0] - 1)))
    grid = pos[1:].float().reshape(1, g, g, -1).permute(0, 3, 1, 2)
    kw = {} if mode == 'nearest' else {'align_corners': False}
    grid = F.interpolate(grid, size=(grid_to, grid_to), mode=mode, **kw)
    grid = grid.permute(0, 2, 3, 1).reshape(grid_to * grid_to, -1)
    return torch.cat([pos[:1].float(), grid])


def fold_layernorm(weight: torch.Tensor, bias: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                   act: torch.dtype) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """LN(x; gamma, beta) W^T + b  ==  rstd * (x W'^T - mean * s) + c   (include/oake_b200.h).

    Returns (W' = W diag(gamma) rounded to `act`, s = row sums of the ROUNDED W' in fp32,
    c = W beta + b in fp32).  Computed in fp64 so the fold itself adds no error."""
    w64 = weight.double()
    w_folded = (w64 * gamma.double()[None, :]).to(act)
    s = w_folded.double().sum(dim=1).float()
    c = (w64 @ beta.double() + bias.double()).float()
    return w_folded, s, c


def _act_torch_dtype() -> torch.dtype:
    return {'f16': torch.float16, 'bf16': torch.bfloat16}[binding.act_dtype_name()]


class _Packed:
    """All tower weights in one device allocation + the pointer table handed to oake_create."""

    def __init__(self, params: Params, device: torch.device, pos_mode: str) -> None:
        act = _act_torch_dtype()
        layers = 0
        while f'transformer.resblocks.{layers}.ln_1.weight' in params:
            layers += 1
        if layers == 0:
            raise ValueError('params holds no transformer.resblocks.*')
        self.layers = layers
        entries = []  # (key, tensor on cpu in final dtype)

        def add(key: str, t: torch.Tensor, dtype: torch.dtype) -> None:
            entries.append((key, t.detach().to('cpu').to(dtype).contiguous()))

        conv = params['conv1.weight']
        if tuple(conv.shape) != (WIDTH, 3, PATCH, PATCH):
            raise ValueError(f'conv1.weight has shape {tuple(conv.shape)}, expected ViT-B/32')
        add('conv1_w', conv.reshape(WIDTH, 3 * PATCH * PATCH), act)
        add('class_emb', params['class_embedding'], torch.float32)
        pos = params['positional_embedding']
        if pos.shape[0] != GRID * GRID + 1:
            raise ValueError('positional_embedding must be the un-resampled (50,768) table')
        add('pos_t50', pos, torch.float32)
        add('pos_t197', resample_positional_embedding(pos, 2 * GRID, pos_mode), torch.float32)
        for n in ('ln_pre', 'ln_post'):
            add(f'{n}_w', params[f'{n}.weight'], torch.float32)
            add(f'{n}_b', params[f'{n}.bias'], torch.float32)
        add('proj_w', params['proj'].T, act)
        fields = ('qkv_w', 'qkv_s', 'qkv_c', 'out_w', 'out_b', 'fc1_w', 'fc1_s', 'fc1_c', 'fc2_w', 'fc2_b')
        for i in range(layers):
            pre = f'transformer.resblocks.{i}.'
            qw, qs, qc = fold_layernorm(params[pre + 'attn.in_proj_weight'].cpu(), params[pre + 'attn.in_proj_bias'].cpu(),
                                        params[pre + 'ln_1.weight'].cpu(), params[pre + 'ln_1.bias'].cpu(), act)
            fw, fs, fc = fold_layernorm(params[pre + 'mlp.c_fc.weight'].cpu(), params[pre + 'mlp.c_fc.bias'].cpu(),
                                        params[pre + 'ln_2.weight'].cpu(), params[pre + 'ln_2.bias'].cpu(), act)
            add(f'{i}.qkv_w', qw, act)
            add(f'{i}.qkv_s', qs, torch.float32)
            add(f'{i}.qkv_c', qc, torch.float32)
            add(f'{i}.out_w', params[pre + 'attn.out_proj.weight'], act)
            add(f'{i}.out_b', params[pre + 'attn.out_proj.bias'], torch.float32)
            add(f'{i}.fc1_w', fw, act)
            add(f'{i}.fc1_s', fs, torch.float32)
            add(f'{i}.fc1_c', fc, torch.float32)
            add(f'{i}.fc2_w', params[pre + 'mlp.c_proj.weight'], act)
            add(f'{i}.fc2_b', params[pre + 'mlp.c_proj.bias'], torch.float32)

        offsets, total = {}, 0
        for key, t in entries:
            offsets[key] = total
            total += (t.numel() * t.element_size() + 255) // 256 * 256
        host = torch.empty(total, dtype=torch.uint8)
        for key, t in entries:
            n = t.numel() * t.element_size()
            host[offsets[key]:offsets[key] + n] = t.reshape(-1).view(torch.uint8)
        self.buffer = host.to(device)
        base = self.buffer.data_ptr()
        ptr = {k: base + o for k, o in offsets.items()}

        self.layer_array = (binding.LayerWeights * layers)()
        for i in range(layers):
            for field in fields:
                setattr(self.layer_array[i], field, ptr[f'{i}.{field}'])
        w = binding.Weights()
        w.layers, w.width, w.heads, w.patch, w.out_dim, w.image = layers, WIDTH, HEADS, PATCH, OUT_DIM, IMAGE
        for k in ('conv1_w', 'class_emb', 'pos_t50', 'pos_t197', 'ln_pre_w', 'ln_pre_b', 'ln_post_w',
                  'ln_post_b', 'proj_w'):
            setattr(w, k, ptr[k])
        w.layer = C.cast(self.layer_array, C.POINTER(binding.LayerWeights))
        self.struct = w


class OakeEngine:
    """One liboake_b200 handle bound to (device, current stream) + a growable workspace."""

    # chunk sizes chosen so that ceil(rows / 128) lands on a multiple of the 148 SMs (no ragged last wave)
    # ($OAKE_CHUNK_T197 / $OAKE_CHUNK_T50 override them: A/B measurements)
    MAX_CROPS = {binding.VARIANT_T50: int(os.environ.get('OAKE_CHUNK_T50', 1894)),
                 binding.VARIANT_T197: int(os.environ.get('OAKE_CHUNK_T197', 478))}

    def __init__(self, params: Params, device: torch.device | str = 'cuda', pos_mode: str = 'bilinear') -> None:
        self.lib = binding.load()
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise binding.OakeError('the OAKE engine runs on a CUDA (sm_100a) device only')
        if self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        self.pos_mode = pos_mode
        self._packed = _Packed(params, self.device, pos_mode)
        handle = C.c_void_p()
        binding.check(self.lib.oake_create(C.byref(handle), self.device.index, C.byref(self._packed.struct)))
        self._handle = handle
        self._ws: Optional[torch.Tensor] = None

    def close(self) -> None:
        if getattr(self, '_handle', None):
            self.lib.oake_destroy(self._handle)
            self._handle = None

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------------------- encode
    def _workspace(self, crops: int, variant: int) -> torch.Tensor:
        need = C.c_size_t()
        binding.check(self.lib.oake_workspace_bytes(self._handle, crops, variant, C.byref(need)))
        if self._ws is None or self._ws.numel() < need.value:
            self._ws = None
            self._ws = torch.empty(need.value, dtype=torch.uint8, device=self.device)
        return self._ws

    def encode_pixels(self, pixels: torch.Tensor, masks: Optional[torch.Tensor] = None,
                      variant: int = binding.VARIANT_T50,
                      want_raw: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """(B,3,224,224) fp32 CLIP-normalised crops -> (fp16 normalised (B,512), fp32 raw or None)."""
        if pixels.dim() != 4 or tuple(pixels.shape[1:]) != (3, IMAGE, IMAGE):
            raise ValueError(f'pixels must be (B,3,{IMAGE},{IMAGE}), got {tuple(pixels.shape)}')
        pixels = pixels.to(self.device, torch.float32).contiguous()
        n = pixels.shape[0]
        if variant == binding.VARIANT_T197:
            if masks is None:
                raise ValueError('the objects variant needs masks (B,1,14,14)')
            masks = masks.to(self.device, torch.float32).contiguous()
            if masks.numel() != n * 4 * GRID * GRID:
                raise ValueError(f'masks must be (B,1,14,14), got {tuple(masks.shape)}')
        out = torch.empty(n, OUT_DIM, dtype=torch.float16, device=self.device)
        raw = torch.empty(n, OUT_DIM, dtype=torch.float32, device=self.device) if want_raw else None
        step = self.MAX_CROPS[variant]
        ws = self._workspace(min(n, step), variant)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        for s in range(0, n, step):
            b = min(step, n - s)
            binding.check(self.lib.oake_encode_pixels(
                self._handle, pixels[s:].data_ptr(), b, variant,
                masks[s:].data_ptr() if variant == binding.VARIANT_T197 else None,
                out[s:].data_ptr(), raw[s:].data_ptr() if raw is not None else None,
                ws.data_ptr(), ws.numel(), stream))
        return out, raw

    # ------------------------------------------------------------------- instrumentation
    def launch_count(self) -> int:
        v = C.c_longlong()
        binding.check(self.lib.oake_launch_count(self._handle, C.byref(v)))
        return v.value

    def profile(self, enable: bool) -> None:
        binding.check(self.lib.oake_profile_enable(self._handle, int(enable)))

    def profile_collect(self) -> Dict[str, Dict[str, float]]:
        cap = 32
        names = (C.c_char_p * cap)()
        ms = (C.c_double * cap)()
        fl = (C.c_double * cap)()
        ln = (C.c_longlong * cap)()
        n = C.c_int()
        binding.check(self.lib.oake_profile_collect(self._handle, cap, names, ms, fl, ln, C.byref(n)))
        return {names[i].decode(): {'ms': ms[i], 'flops': fl[i], 'launches': ln[i]} for i in range(n.value)}


class _Visual:
    """`model.visual`: callable like the surgically modified VisionTransformer (objects.py:330)."""

    def __init__(self, owner: 'OakeModel') -> None:
        self._owner = owner
        self.grid = GRID
        self.patch_size = PATCH

    def __call__(self, pixels: torch.Tensor, masks: Optional[torch.Tensor] = None) -> torch.Tensor:
        if masks is None:
            return self._owner.encode_image(pixels)
        if self.grid != 2 * GRID:
            raise binding.OakeError('visual(objects, masks) needs the objects surgery (OakeModel.for_objects)')
        _, raw = self._owner.engine.encode_pixels(pixels, masks, binding.VARIANT_T197, want_raw=True)
        return raw


class OakeModel:
    """Drop-in for the `clip.model.CLIP` object the OAKE validators hold."""

    dtype = torch.float32  # what callers cast inputs to (objects.py:328); tensor-core math is internal

    def __init__(self, params: Params, device: torch.device | str = 'cuda', pos_mode: str = 'bilinear') -> None:
        self.engine = OakeEngine(params, device, pos_mode)
        self.visual = _Visual(self)

    def for_objects(self, upsample: int = 2) -> 'OakeModel':
        """The surgery of objects.py:285-314 (denser grid, stride-16 conv, side stream)."""
        if upsample != 2:
            raise binding.OakeError('only upsample=2 (14x14 grid) is built')
        self.visual.grid = GRID * upsample
        return self

    def eval(self) -> 'OakeModel':
        return self

    def requires_grad_(self, flag: bool = False) -> 'OakeModel':
        return self

    def encode_image(self, pixels: torch.Tensor) -> torch.Tensor:
        _, raw = self.engine.encode_pixels(pixels, None, binding.VARIANT_T50, want_raw=True)
        return raw

    def embed(self, pixels: torch.Tensor, masks: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Fused `F.normalize(encode(...)).half()` -- the value the validators store."""
        variant = binding.VARIANT_T50 if masks is None else binding.VARIANT_T197
        out, _ = self.engine.encode_pixels(pixels, masks, variant)
        return out
