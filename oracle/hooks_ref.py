"""Second, structurally different restatement of the objects tower -- TEST INFRASTRUCTURE.

``oracle/vit.py::encode_objects`` is explicit matmul/softmax code.  This file instead
rebuilds the published openai/CLIP module tree (``nn.Conv2d`` + ``nn.MultiheadAttention``
blocks, sequence-first (T,B,D) activations) and drives the side stream through
``register_forward(_pre)_hook`` callbacks with the control flow of
oadp/oake/objects.py:198-266 and the surgery of objects.py:285-314.  The two agree to
fp32 round-off (tests/test_oracle.py), which is what pins the T=197 path: there is no
other implementation of it in the container.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn

from . import vit


class _Block(nn.Module):

    def __init__(self) -> None:
        super().__init__()
        w = vit.WIDTH
        self.attn = nn.MultiheadAttention(w, vit.HEADS)
        self.ln_1 = nn.LayerNorm(w, eps=vit.LN_EPS)
        self.mlp = nn.Sequential(
            OrderedDict(c_fc=nn.Linear(w, 4 * w), gelu=_QuickGELU(), c_proj=nn.Linear(4 * w, w)))
        self.ln_2 = nn.LayerNorm(w, eps=vit.LN_EPS)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        h = self.ln_1(x)
        x = x + self.attn(h, h, h, need_weights=False)[0]
        return x + self.mlp(self.ln_2(x))


class _QuickGELU(nn.Module):

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x * torch.sigmoid(1.702 * x)


class _Transformer(nn.Module):

    def __init__(self, layers: int) -> None:
        super().__init__()
        self.resblocks = nn.Sequential(*[_Block() for _ in range(layers)])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.resblocks(x)


class _Visual(nn.Module):

    def __init__(self, layers: int, tokens: int) -> None:
        super().__init__()
        w = vit.WIDTH
        self.conv1 = nn.Conv2d(3, w, vit.PATCH, vit.PATCH, bias=False)
        self.class_embedding = nn.Parameter(torch.zeros(w))
        self.positional_embedding = nn.Parameter(torch.zeros(tokens, w))
        self.ln_pre = nn.LayerNorm(w, eps=vit.LN_EPS)
        self.transformer = _Transformer(layers)
        self.ln_post = nn.LayerNorm(w, eps=vit.LN_EPS)
        self.proj = nn.Parameter(torch.zeros(w, vit.OUT_DIM))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.conv1(x)
        x = x.flatten(2).permute(0, 2, 1)
        cls_tok = self.class_embedding.expand(x.shape[0], 1, -1)
        x = torch.cat([cls_tok, x], 1) + self.positional_embedding
        x = self.ln_pre(x).permute(1, 0, 2)  # (T,B,D)
        x = self.transformer(x).permute(1, 0, 2)
        return self.ln_post(x[:, 0]) @ self.proj


class _SideStream:
    """State shared by the four kinds of hooks (cf. objects.py ``Hooks``)."""

    def __init__(self) -> None:
        self.y = None
        self.bias = None

    def on_visual_enter(self, module, inputs):
        pixels, masks = inputs
        self.bias = vit.mask_to_bias(masks)  # (B,197), y column last, -100 on background
        return (pixels, )

    def on_transformer_enter(self, module, inputs):
        self.y = inputs[0][:1]

    def on_block_enter(self, block: _Block, inputs):
        x = inputs[0]
        heads = block.attn.num_heads
        bias = self.bias.repeat_interleave(heads, dim=0)[:, None, :]  # (B*h,1,197)
        z = block.ln_1(torch.cat([x[1:], self.y]))
        y = self.y + block.attn(z[-1:], z, z, need_weights=False, attn_mask=bias)[0]
        self.y = y + block.mlp(block.ln_2(y))

    def on_transformer_exit(self, module, inputs, output):
        y, self.y = self.y, None
        return y

    def on_visual_exit(self, module, inputs, output):
        self.bias = None


class HookedVisual(nn.Module):
    """visual(objects, masks) built the way objects.py:285-314 builds it."""

    def __init__(self, p197: vit.Params) -> None:
        super().__init__()
        layers = vit.num_layers(p197)
        tokens = p197['positional_embedding'].shape[0]
        self.visual = _Visual(layers, tokens)
        sd = {k: v.clone() for k, v in p197.items()}
        self.visual.load_state_dict(sd, strict=True)
        conv1 = self.visual.conv1
        conv1.stride = tuple(s // 2 for s in conv1.stride)
        conv1.padding = ((vit.PATCH - 1) // 2, ) * 2
        side = _SideStream()
        self.visual.register_forward_pre_hook(side.on_visual_enter)
        self.visual.register_forward_hook(side.on_visual_exit)
        tr = self.visual.transformer
        tr.register_forward_pre_hook(side.on_transformer_enter)
        tr.register_forward_hook(side.on_transformer_exit)
        for blk in tr.resblocks:
            blk.register_forward_pre_hook(side.on_block_enter)

    @torch.no_grad()
    def forward(self, pixels: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
        return self.visual(pixels.float(), masks.float())
