"""fp32 CPU restatement of the CLIP ViT-B/32 image tower as OAKE uses it.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Parity unpinned by the reference;
pinned to HF CLIP (T=50) and to ``oracle/hooks_ref.py`` (T=197 + side stream).

Reference call sites this follows:
  * oadp/oake/globals.py:54-59, oadp/oake/blocks.py:125-134 -- ``model.encode_image`` on
    (B,3,224,224), un-modified ViT-B/32, T = 7*7+1 = 50 tokens.
  * oadp/oake/objects.py:285-314 -- model surgery: positional embedding resampled
    7x7 -> 14x14, ``conv1.stride`` 32 -> 16, ``conv1.padding`` = 15, T = 197 tokens.
  * oadp/oake/objects.py:198-266 -- the mask-attended CLS side stream ("y") that replaces
    the transformer output.
The architecture itself is the published openai/CLIP ``model.py`` VisionTransformer
(the ``clip`` dependency is not vendored in the reference; README.md:44).

Everything here is plain functional torch on a ``dict[str, Tensor]`` that uses the
OpenAI state-dict names (``conv1.weight``, ``transformer.resblocks.3.attn.in_proj_weight``
...), so real CLIP weights can be dropped in unchanged.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

WIDTH = 768
HEADS = 12
LAYERS = 12
PATCH = 32
GRID = 7
OUT_DIM = 512
LN_EPS = 1e-5


def init_visual_params(seed: int = 0, layers: int = LAYERS) -> Params:
    """Seeded random ViT-B/32 visual-tower weights (no CLIP checkpoint exists offline).

    Scales follow openai/CLIP ``initialize_parameters`` / ``VisionTransformer.__init__``
    (attn std width^-0.5, proj std width^-0.5 * (2*layers)^-0.5, fc std (2*width)^-0.5,
    class/pos/proj scale width^-0.5), but LayerNorm affine parameters and all biases are
    given non-trivial values so that a kernel which drops one of them fails parity.
    """
    g = torch.Generator().manual_seed(seed)

    def rn(*shape: int, std: float = 1.0) -> torch.Tensor:
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    w = WIDTH
    scale = w**-0.5
    attn_std = w**-0.5
    proj_std = (w**-0.5) * ((2 * layers)**-0.5)
    fc_std = (2 * w)**-0.5
    p: Params = {}
    p['conv1.weight'] = rn(w, 3, PATCH, PATCH, std=(3 * PATCH * PATCH)**-0.5)
    p['class_embedding'] = rn(w, std=scale)
    p['positional_embedding'] = rn(GRID * GRID + 1, w, std=scale)
    for name in ('ln_pre', 'ln_post'):
        p[f'{name}.weight'] = 1.0 + rn(w, std=0.1)
        p[f'{name}.bias'] = rn(w, std=0.1)
    for i in range(layers):
        pre = f'transformer.resblocks.{i}.'
        for name in ('ln_1', 'ln_2'):
            p[pre + f'{name}.weight'] = 1.0 + rn(w, std=0.1)
            p[pre + f'{name}.bias'] = rn(w, std=0.1)
        p[pre + 'attn.in_proj_weight'] = rn(3 * w, w, std=attn_std)
        p[pre + 'attn.in_proj_bias'] = rn(3 * w, std=0.02)
        p[pre + 'attn.out_proj.weight'] = rn(w, w, std=proj_std)
        p[pre + 'attn.out_proj.bias'] = rn(w, std=0.02)
        p[pre + 'mlp.c_fc.weight'] = rn(4 * w, w, std=fc_std)
        p[pre + 'mlp.c_fc.bias'] = rn(4 * w, std=0.02)
        p[pre + 'mlp.c_proj.weight'] = rn(w, 4 * w, std=proj_std)
        p[pre + 'mlp.c_proj.bias'] = rn(w, std=0.02)
    p['proj'] = rn(w, OUT_DIM, std=scale)
    return p


def num_layers(p: Params) -> int:
    n = 0
    while f'transformer.resblocks.{n}.ln_1.weight' in p:
        n += 1
    return n


def interpolate_positional_embedding(
    pos: torch.Tensor,
    grid_to: int,
    mode: str = 'bilinear',
) -> torch.Tensor:
    """Class row kept, (g*g, D) grid rows resampled to (grid_to*grid_to, D).

    Follows the call at oadp/oake/objects.py:293-296.  The fork function's interpolation
    mode is not visible from the reference; it is a one-off weight
    transform, so the CUDA path and this oracle consume the SAME resampled table and the
    choice does not affect parity.  Default: bilinear, align_corners=False.
    """
    cls_row, grid_rows = pos[:1], pos[1:]
    g = int(round(math.sqrt(grid_rows.shape[0])))
    assert g * g == grid_rows.shape[0]
    t = grid_rows.reshape(1, g, g, -1).permute(0, 3, 1, 2)
    kwargs = {} if mode == 'nearest' else dict(align_corners=False)
    t = F.interpolate(t, size=(grid_to, grid_to), mode=mode, **kwargs)
    t = t.permute(0, 2, 3, 1).reshape(grid_to * grid_to, -1)
    return torch.cat([cls_row, t])


def objects_surgery(p: Params, upsample: int = 2, mode: str = 'bilinear') -> Params:
    """oadp/oake/objects.py:285-301: same weights, denser positional table."""
    q = dict(p)
    q['positional_embedding'] = interpolate_positional_embedding(
        p['positional_embedding'], GRID * upsample, mode)
    return q


def _ln(x: torch.Tensor, p: Params, name: str) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1], ), p[name + '.weight'], p[name + '.bias'], LN_EPS)


def _quick_gelu(x: torch.Tensor) -> torch.Tensor:
    return x * torch.sigmoid(1.702 * x)


def _heads(t: torch.Tensor) -> torch.Tensor:
    b, n, _ = t.shape
    return t.reshape(b, n, HEADS, WIDTH // HEADS).transpose(1, 2)  # (B,H,n,dh)


def _mlp(x: torch.Tensor, p: Params, pre: str) -> torch.Tensor:
    u = x @ p[pre + 'mlp.c_fc.weight'].T + p[pre + 'mlp.c_fc.bias']
    return _quick_gelu(u) @ p[pre + 'mlp.c_proj.weight'].T + p[pre + 'mlp.c_proj.bias']


def _attend(q, k, v, bias=None):
    """q (B,H,nq,dh), k/v (B,H,nk,dh), additive bias broadcastable to (B,H,nq,nk)."""
    s = (q / math.sqrt(q.shape[-1])) @ k.transpose(-1, -2)
    if bias is not None:
        s = s + bias
    a = torch.softmax(s, dim=-1)
    o = a @ v  # (B,H,nq,dh)
    b, h, nq, dh = o.shape
    return o.transpose(1, 2).reshape(b, nq, h * dh)


def _block(x: torch.Tensor, p: Params, i: int) -> torch.Tensor:
    pre = f'transformer.resblocks.{i}.'
    h = _ln(x, p, pre + 'ln_1')
    qkv = h @ p[pre + 'attn.in_proj_weight'].T + p[pre + 'attn.in_proj_bias']
    q, k, v = qkv.split(WIDTH, dim=-1)
    o = _attend(_heads(q), _heads(k), _heads(v))
    x = x + o @ p[pre + 'attn.out_proj.weight'].T + p[pre + 'attn.out_proj.bias']
    x = x + _mlp(_ln(x, p, pre + 'ln_2'), p, pre)
    return x


def embed(p: Params, pixels: torch.Tensor, stride: int, padding: int) -> torch.Tensor:
    """conv1 -> tokens -> +class -> +pos -> ln_pre.  Returns (B,T,D)."""
    x = F.conv2d(pixels, p['conv1.weight'], None, stride=stride, padding=padding)
    b, d, gh, gw = x.shape
    x = x.reshape(b, d, gh * gw).transpose(1, 2)  # row-major over (gy,gx)
    cls_tok = p['class_embedding'].reshape(1, 1, d).expand(b, 1, d)
    x = torch.cat([cls_tok, x], dim=1) + p['positional_embedding']
    return _ln(x, p, 'ln_pre')


def head(p: Params, tok: torch.Tensor) -> torch.Tensor:
    """ln_post(token) @ proj -> (B,512) (not yet L2-normalised)."""
    return _ln(tok, p, 'ln_post') @ p['proj']


@torch.no_grad()
def encode_image(p: Params, pixels: torch.Tensor) -> torch.Tensor:
    """Un-modified tower, T=50 (globals.py:57, blocks.py:129)."""
    x = embed(p, pixels.float(), stride=PATCH, padding=0)
    for i in range(num_layers(p)):
        x = _block(x, p, i)
    return head(p, x[:, 0])


def mask_to_bias(masks: torch.Tensor) -> torch.Tensor:
    """objects.py:204-214: (B,1,g,g) with 1=background -> additive (B, g*g+1) row.

    Patches first, the y token last; bias = -100 * mask (finite!), 0 for y itself."""
    b = masks.shape[0]
    m = masks.reshape(b, -1).float()
    m = torch.cat([m, m.new_zeros(b, 1)], dim=1)
    return m * -100.0


@torch.no_grad()
def encode_objects(p197: Params, pixels: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
    """Objects variant (objects.py:198-338): stride-16 patch embed, T=197, side stream.

    ``p197`` must already carry the resampled positional table (``objects_surgery``).
    The main stream x runs the plain blocks; before block l touches x the side token y is
    updated by attending (1 query) over ln_1([x[1:]; y]) with the foreground mask bias,
    then its own MLP.  The tower output is y after the last block.
    """
    x = embed(p197, pixels.float(), stride=PATCH // 2, padding=(PATCH - 1) // 2)
    bias = mask_to_bias(masks)[:, None, None, :]  # (B,1,1,197) same for every head
    y = x[:, :1]  # (B,1,D) -- the CLS token right after ln_pre
    for i in range(num_layers(p197)):
        pre = f'transformer.resblocks.{i}.'
        z = _ln(torch.cat([x[:, 1:], y], dim=1), p197, pre + 'ln_1')
        qkv = z @ p197[pre + 'attn.in_proj_weight'].T + p197[pre + 'attn.in_proj_bias']
        q, k, v = qkv.split(WIDTH, dim=-1)
        o = _attend(_heads(q[:, -1:]), _heads(k), _heads(v), bias)
        y = y + o @ p197[pre + 'attn.out_proj.weight'].T + p197[pre + 'attn.out_proj.bias']
        y = y + _mlp(_ln(y, p197, pre + 'ln_2'), p197, pre)
        x = _block(x, p197, i)
    return head(p197, y[:, 0])


def normalize_half(e: torch.Tensor) -> torch.Tensor:
    """globals.py:58-59 / blocks.py:130-133 / objects.py:331-334."""
    return F.normalize(e.float(), dim=-1, eps=1e-12).half()


# --------------------------------------------------------------------------------------
# Independent cross-check: HuggingFace CLIPVisionModelWithProjection
# --------------------------------------------------------------------------------------

def to_hf_state_dict(p: Params) -> Params:
    w = WIDTH
    sd: Params = {
        'vision_model.embeddings.class_embedding': p['class_embedding'],
        'vision_model.embeddings.patch_embedding.weight': p['conv1.weight'],
        'vision_model.embeddings.position_embedding.weight': p['positional_embedding'],
        'vision_model.pre_layrnorm.weight': p['ln_pre.weight'],
        'vision_model.pre_layrnorm.bias': p['ln_pre.bias'],
        'vision_model.post_layernorm.weight': p['ln_post.weight'],
        'vision_model.post_layernorm.bias': p['ln_post.bias'],
        'visual_projection.weight': p['proj'].T.contiguous(),
    }
    for i in range(num_layers(p)):
        src = f'transformer.resblocks.{i}.'
        dst = f'vision_model.encoder.layers.{i}.'
        for j, n in enumerate('qkv'):
            sd[dst + f'self_attn.{n}_proj.weight'] = p[src + 'attn.in_proj_weight'][j * w:(j + 1) * w]
            sd[dst + f'self_attn.{n}_proj.bias'] = p[src + 'attn.in_proj_bias'][j * w:(j + 1) * w]
        sd[dst + 'self_attn.out_proj.weight'] = p[src + 'attn.out_proj.weight']
        sd[dst + 'self_attn.out_proj.bias'] = p[src + 'attn.out_proj.bias']
        for a, b_ in (('ln_1', 'layer_norm1'), ('ln_2', 'layer_norm2')):
            sd[dst + b_ + '.weight'] = p[src + a + '.weight']
            sd[dst + b_ + '.bias'] = p[src + a + '.bias']
        for a, b_ in (('c_fc', 'fc1'), ('c_proj', 'fc2')):
            sd[dst + f'mlp.{b_}.weight'] = p[src + f'mlp.{a}.weight']
            sd[dst + f'mlp.{b_}.bias'] = p[src + f'mlp.{a}.bias']
    return sd


def build_hf_model(p: Params):
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    cfg = CLIPVisionConfig(num_hidden_layers=num_layers(p))  # defaults == ViT-B/32
    model = CLIPVisionModelWithProjection(cfg).eval()
    missing, unexpected = model.load_state_dict(to_hf_state_dict(p), strict=False)
    missing = [m for m in missing if 'position_ids' not in m]
    assert not missing and not unexpected, (missing, unexpected)
    return model
