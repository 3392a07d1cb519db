"""fp32 CPU restatement of the CLIP ViT-B/32 TEXT tower as the prompt builder uses it.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Reference call site: oadp/prompts/vild.py:56-72 --
per prompt template ``tokens = clip.adaptively_tokenize(texts)``, ``model.encode_text(tokens)``,
``F.normalize``; the templates are averaged into ``data/prompts/vild.pth`` = ``{embeddings, names}``,
the ``E`` of the cosine classifier (oadp/dp/classifiers.py:22-68).  The architecture is the published
openai/CLIP ``model.py`` ``CLIP.encode_text`` (the ``clip`` dependency is not vendored, README.md:44):

    x = token_embedding(text) + positional_embedding        (B, n_ctx, 512)
    x = transformer(x) with a causal additive mask          12 blocks, 8 heads of 64, QuickGELU MLP 2048
    x = ln_final(x)
    out = x[arange(B), text.argmax(-1)] @ text_projection   the EOT token has the largest id

Pinned against HuggingFace ``CLIPTextModelWithProjection`` (tests/test_oracle_text.py).  The fork's
``adaptively_tokenize`` is unseen; the tower is causal and pools at the EOT token, so any context
length that holds the longest prompt gives the same result as the stock 77 (checked in the test).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

WIDTH = 512
HEADS = 8
LAYERS = 12
VOCAB = 49408
CONTEXT = 77
OUT_DIM = 512
LN_EPS = 1e-5
SOT, EOT = 49406, 49407


def init_text_params(seed: int = 0, layers: int = LAYERS, vocab: int = VOCAB) -> Params:
    """Seeded random text-tower weights under the OpenAI state-dict names; scales as in openai/CLIP
    ``initialize_parameters``, LayerNorm affine parameters and biases made non-trivial on purpose."""
    g = torch.Generator().manual_seed(seed)

    def rn(*shape: int, std: float = 1.0) -> torch.Tensor:
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    w = WIDTH
    attn_std, proj_std, fc_std = w**-0.5, (w**-0.5) * ((2 * layers)**-0.5), (2 * w)**-0.5
    p: Params = {}
    p['token_embedding.weight'] = rn(vocab, w, std=0.02)
    p['positional_embedding'] = rn(CONTEXT, w, std=0.01)
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
    p['ln_final.weight'] = 1.0 + rn(w, std=0.1)
    p['ln_final.bias'] = rn(w, std=0.1)
    p['text_projection'] = rn(w, OUT_DIM, std=w**-0.5)
    return p


def num_layers(p: Params) -> int:
    n = 0
    while f'transformer.resblocks.{n}.ln_1.weight' in p:
        n += 1
    return n


def synthetic_tokens(n: int, length: int, seed: int = 0, vocab: int = VOCAB) -> torch.Tensor:
    """(n, length) int64 rows shaped like CLIP's tokenizer output: SOT, 1..length-2 word tokens, EOT, zeros."""
    g = torch.Generator().manual_seed(seed)
    out = torch.zeros(n, length, dtype=torch.int64)
    for i in range(n):
        words = int(torch.randint(1, length - 1, (1, ), generator=g))
        out[i, 0] = SOT
        out[i, 1:1 + words] = torch.randint(1, min(vocab, SOT) - 1, (words, ), generator=g)
        out[i, 1 + words] = EOT if vocab > EOT else vocab - 1
    return out


def _ln(x: torch.Tensor, p: Params, name: str) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1], ), p[name + '.weight'], p[name + '.bias'], LN_EPS)


def _block(x: torch.Tensor, p: Params, i: int, mask: torch.Tensor) -> torch.Tensor:
    pre = f'transformer.resblocks.{i}.'
    b, n, w = x.shape
    qkv = _ln(x, p, pre + 'ln_1') @ p[pre + 'attn.in_proj_weight'].T + p[pre + 'attn.in_proj_bias']
    q, k, v = (t.reshape(b, n, HEADS, w // HEADS).transpose(1, 2) for t in qkv.split(w, dim=-1))
    s = (q / math.sqrt(w // HEADS)) @ k.transpose(-1, -2) + mask
    o = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(b, n, w)
    x = x + o @ p[pre + 'attn.out_proj.weight'].T + p[pre + 'attn.out_proj.bias']
    u = _ln(x, p, pre + 'ln_2') @ p[pre + 'mlp.c_fc.weight'].T + p[pre + 'mlp.c_fc.bias']
    u = u * torch.sigmoid(1.702 * u)
    return x + u @ p[pre + 'mlp.c_proj.weight'].T + p[pre + 'mlp.c_proj.bias']


def encode_text(p: Params, tokens: torch.Tensor) -> torch.Tensor:
    """tokens (B, n_ctx <= 77) int64 -> (B, 512) fp32, un-normalised (openai/CLIP ``encode_text``)."""
    n = tokens.shape[1]
    x = p['token_embedding.weight'][tokens] + p['positional_embedding'][:n]
    mask = torch.full((n, n), float('-inf')).triu_(1)
    for i in range(num_layers(p)):
        x = _block(x, p, i, mask)
    x = _ln(x, p, 'ln_final')
    return x[torch.arange(x.shape[0]), tokens.argmax(dim=-1)] @ p['text_projection']


def prompt_embeddings(p: Params, token_batches) -> torch.Tensor:
    """oadp/prompts/vild.py:60-71: one token batch per template (all the category names formatted with
    it) -> encode -> F.normalize -> mean over the templates.  Rows are NOT re-normalised (the classifier
    uses them as stored)."""
    embeddings = [F.normalize(encode_text(p, tokens)) for tokens in token_batches]
    return sum(embeddings) / len(embeddings)


def to_hf_state_dict(p: Params) -> Params:
    w = WIDTH
    sd: Params = {
        'text_model.embeddings.token_embedding.weight': p['token_embedding.weight'],
        'text_model.embeddings.position_embedding.weight': p['positional_embedding'],
        'text_model.final_layer_norm.weight': p['ln_final.weight'],
        'text_model.final_layer_norm.bias': p['ln_final.bias'],
        'text_projection.weight': p['text_projection'].T.contiguous(),
    }
    for i in range(num_layers(p)):
        src, dst = f'transformer.resblocks.{i}.', f'text_model.encoder.layers.{i}.'
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
    from transformers import CLIPTextConfig, CLIPTextModelWithProjection
    cfg = CLIPTextConfig(num_hidden_layers=num_layers(p), vocab_size=p['token_embedding.weight'].shape[0])
    model = CLIPTextModelWithProjection(cfg).eval()  # defaults == the ViT-B/32 text tower
    missing, unexpected = model.load_state_dict(to_hf_state_dict(p), strict=False)
    missing = [m for m in missing if 'position_ids' not in m]
    assert not missing and not unexpected, (missing, unexpected)
    return model
