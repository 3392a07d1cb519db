"""CLIP text tower on the B200 and the prompt builder on top of it (second half).

Stands where `model.encode_text(tokens)` stands in oadp/prompts/vild.py:56-72.  The tower runs on
liboake_b200.so (`oake_encode_text`: the image tower's tcgen05 GEMMs at width 512 plus three small row
kernels, oadp_b200/csrc/text.cu); `build_prompts` reproduces the loop of prompts/vild.py:60-71 -- per
template: format every category name, tokenize, encode, `F.normalize`; then the mean over the templates
-- and returns the `{embeddings, names}` dict the classifiers load (oadp/dp/classifiers.py:27-41).

Parity: tests/test_gpu_text.py (1, 2 and 12 layers, context 16 / 20 / 77) against oracle/text.py, which is
pinned to HuggingFace CLIP; first B200 run, memcheck and timing in profiles/r2_01_text_tower_*.

The tokenizer is not part of this package (CLIP's BPE vocabulary cannot be fetched offline): pass
`clip.tokenize` / the fork's `clip.adaptively_tokenize`, or any callable texts -> int tensor (B, L <= 77).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, Iterable, List, Optional, Sequence

import torch
import torch.nn.functional as F

from . import binding
from .model import _act_torch_dtype, fold_layernorm

Params = Dict[str, torch.Tensor]

WIDTH, HEADS, CONTEXT, OUT_DIM = 512, 8, 77, 512
MAX_ROWS = 128 * 148 * 5  # token rows per call: whole 128-row tiles over 148 SMs, five waves


def pack_text_weights(params: Params, act: torch.dtype):
    """-> ([(key, cpu tensor in its final dtype)], layers, vocab, context): what goes to the device, with
    ln_1 / ln_2 folded into the QKV / c_fc weights exactly as for the image tower (`fold_layernorm`)."""
    layers = 0
    while f'transformer.resblocks.{layers}.ln_1.weight' in params:
        layers += 1
    if layers == 0:
        raise ValueError('params holds no transformer.resblocks.*')
    emb, pos, proj = params['token_embedding.weight'], params['positional_embedding'], params['text_projection']
    if emb.shape[1] != WIDTH or pos.shape[1] != WIDTH or pos.shape[0] > CONTEXT or tuple(proj.shape) != (WIDTH, OUT_DIM):
        raise ValueError('not a ViT-B/32 text tower (width 512, context <= 77, projection 512x512)')
    entries = []

    def add(key: str, t: torch.Tensor, dtype: torch.dtype) -> None:
        entries.append((key, t.detach().to('cpu').to(dtype).contiguous()))

    add('token_emb', emb, torch.float32)
    add('pos', pos, torch.float32)
    add('ln_final_w', params['ln_final.weight'], torch.float32)
    add('ln_final_b', params['ln_final.bias'], torch.float32)
    add('proj_w', proj.T, act)
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
    return entries, layers, emb.shape[0], pos.shape[0]


class OakeTextModel:
    """`encode_text` of the `clip.model.CLIP` object the prompt builders hold (prompts/vild.py:58,65)."""

    dtype = torch.float32

    def __init__(self, params: Params, device: torch.device | str = 'cuda') -> None:
        self.lib = binding.load()
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise binding.OakeError('the text tower runs on a CUDA (sm_100a) device only')
        if self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        entries, layers, vocab, context = pack_text_weights(params, _act_torch_dtype())
        offsets, total = {}, 0
        for key, t in entries:
            offsets[key] = total
            total += (t.numel() * t.element_size() + 255) // 256 * 256
        host = torch.empty(total, dtype=torch.uint8)
        for key, t in entries:
            n = t.numel() * t.element_size()
            host[offsets[key]:offsets[key] + n] = t.reshape(-1).view(torch.uint8)
        self._buffer = host.to(self.device)
        base = self._buffer.data_ptr()
        self._layer_array = (binding.LayerWeights * layers)()
        for i in range(layers):
            for field, _ in binding.LayerWeights._fields_:
                setattr(self._layer_array[i], field, base + offsets[f'{i}.{field}'])
        w = binding.TextWeights()
        w.layers, w.width, w.heads, w.vocab, w.context, w.out_dim = layers, WIDTH, HEADS, vocab, context, OUT_DIM
        for k in ('token_emb', 'pos', 'ln_final_w', 'ln_final_b', 'proj_w'):
            setattr(w, k, base + offsets[k])
        w.layer = C.cast(self._layer_array, C.POINTER(binding.LayerWeights))
        self.context = context
        self.vocab = vocab
        handle = C.c_void_p()
        binding.check(self.lib.oake_text_create(C.byref(handle), self.device.index, C.byref(w)))
        self._handle = handle
        self._ws: Optional[torch.Tensor] = None

    def close(self) -> None:
        if getattr(self, '_handle', None):
            self.lib.oake_text_destroy(self._handle)
            self._handle = None

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass

    def eval(self) -> 'OakeTextModel':
        return self

    def encode_text(self, tokens: torch.Tensor) -> torch.Tensor:
        """tokens (B, L <= 77) integer -> (B, 512) fp32 on the device, un-normalised."""
        if tokens.dim() != 2 or not 1 <= tokens.shape[1] <= self.context:
            raise ValueError(f'tokens must be (B, L) with 1 <= L <= {self.context}, got {tuple(tokens.shape)}')
        if tokens.numel() and bool(((tokens < 0) | (tokens >= self.vocab)).any()):
            # an id outside the embedding table is a tokenizer / vocabulary mismatch, never data
            raise ValueError(f'token ids must lie in [0, {self.vocab}); got [{int(tokens.min())}, {int(tokens.max())}]')
        tokens = tokens.to(self.device, torch.int32).contiguous()
        n, length = tokens.shape
        out = torch.empty(n, OUT_DIM, dtype=torch.float32, device=self.device)
        step = max(1, MAX_ROWS // length)
        need = C.c_size_t()
        binding.check(self.lib.oake_text_workspace_bytes(self._handle, min(n, step), length, C.byref(need)))
        if self._ws is None or self._ws.numel() < need.value:
            self._ws = None
            self._ws = torch.empty(need.value, dtype=torch.uint8, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        for s in range(0, n, step):
            b = min(step, n - s)
            binding.check(self.lib.oake_encode_text(self._handle, tokens[s:].data_ptr(), b, length, out[s:].data_ptr(),
                                                    self._ws.data_ptr(), self._ws.numel(), stream))
        return out


def build_prompts(encode_text: Callable[[torch.Tensor], torch.Tensor], tokenize: Callable[[List[str]], torch.Tensor],
                  templates: Sequence[str], names: Iterable[str]) -> Dict[str, object]:
    """oadp/prompts/vild.py:56-72 with the model and the tokenizer passed in.

    `names` are sorted and de-duplicated as :57 does (`sorted(set(coco.all_ + lvis.all_))`); per template
    every name is formatted, tokenized and encoded, the rows are L2-normalised (:66), and the templates
    are averaged WITHOUT re-normalising (:69) -- rows of the result have norm < 1, which the classifier
    relies on.  Returns `dict(embeddings=(n, 512) fp32 cpu, names=[...])`, the content of
    `data/prompts/vild.pth`."""
    categories = sorted(set(names))
    if not templates:
        raise ValueError('no prompt templates')
    total = None
    with torch.no_grad():
        for template in templates:
            tokens = tokenize([template.format(c) for c in categories])
            e = F.normalize(encode_text(tokens).float())
            total = e if total is None else total + e
    return dict(embeddings=(total / len(templates)).cpu(), names=categories)
