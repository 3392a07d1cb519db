"""CPU tier for the text-tower row (second half): weight packing / LayerNorm fold against the
oracle's plain weights, the prompt-builder loop against oracle.text.prompt_embeddings (prompts/vild.py:56-72),
and the ABI's argument checks (no GPU involved)."""
import ctypes

import pytest
import torch

from oadp_b200 import binding
from oadp_b200 import text as otower
from oracle import text as otext


@pytest.fixture(scope='module')
def params():
    return otext.init_text_params(5, layers=2, vocab=1000)


def test_packed_weights_reproduce_the_layer_math(params):
    """y = LN(x) W^T + b recomputed from the folded triple (W', s, c) the GEMM epilogue uses."""
    entries, layers, vocab, context = otower.pack_text_weights(params, torch.float32)
    e = dict(entries)
    assert (layers, vocab, context) == (2, 1000, 77)
    assert e['token_emb'].shape == (1000, 512) and e['proj_w'].shape == (512, 512)
    assert torch.equal(e['proj_w'], params['text_projection'].T.contiguous())
    x = torch.randn(9, 512, generator=torch.Generator().manual_seed(1))
    mean, var = x.mean(-1, keepdim=True), x.var(-1, unbiased=False, keepdim=True)
    rstd = (var + 1e-5).rsqrt()
    for name, w_key, ln in (('qkv', 'attn.in_proj', 'ln_1'), ('fc1', 'mlp.c_fc', 'ln_2')):
        pre = 'transformer.resblocks.1.'
        weight = params[pre + w_key + ('_weight' if name == 'qkv' else '.weight')]
        bias = params[pre + w_key + ('_bias' if name == 'qkv' else '.bias')]
        want = torch.nn.functional.layer_norm(x, (512, ), params[pre + ln + '.weight'], params[pre + ln + '.bias'],
                                              1e-5) @ weight.T + bias
        got = rstd * (x @ e[f'1.{name}_w'].T - mean * e[f'1.{name}_s']) + e[f'1.{name}_c']
        assert float((got - want).abs().max()) < 1e-4
    for key in ('out_w', 'out_b', 'fc2_w', 'fc2_b'):
        assert f'0.{key}' in e and f'1.{key}' in e


def test_pack_rejects_other_geometries(params):
    bad = dict(params)
    bad['text_projection'] = torch.zeros(512, 256)
    with pytest.raises(ValueError):
        otower.pack_text_weights(bad, torch.float16)
    with pytest.raises(ValueError):
        otower.pack_text_weights({k: v for k, v in params.items() if 'resblocks' not in k}, torch.float16)


def test_build_prompts_follows_the_reference_loop(params):
    names = ['zebra', 'apple', 'car', 'apple']  # sorted + de-duplicated as prompts/vild.py:57
    templates = ['a photo of a {}', 'There is a {} in the scene', 'This is one large {} in the picture']
    vocab = {}

    def tokenize(texts):  # a stand-in word-level tokenizer: SOT, word ids, EOT (largest id), zero padding
        rows = []
        for t in texts:
            ids = [998] + [vocab.setdefault(w, 1 + len(vocab)) for w in t.split()] + [999]
            rows.append(ids + [0] * (12 - len(ids)))
        return torch.tensor(rows)

    got = otower.build_prompts(lambda tok: otext.encode_text(params, tok), tokenize, templates, names)
    assert got['names'] == ['apple', 'car', 'zebra']
    batches = [tokenize([t.format(c) for c in got['names']]) for t in templates]
    want = otext.prompt_embeddings(params, batches)
    assert torch.allclose(got['embeddings'], want, atol=1e-6)
    assert bool((got['embeddings'].norm(dim=-1) < 1.0).all())  # means of unit rows: what the classifier expects
    with pytest.raises(ValueError):
        otower.build_prompts(lambda tok: tok, tokenize, [], names)


def test_text_abi_rejects_bad_arguments_without_gpu(lib):
    w = binding.TextWeights()
    h = ctypes.c_void_p()
    assert lib.oake_text_create(ctypes.byref(h), 0, ctypes.byref(w)) != 0
    assert b'layers' in lib.oake_last_error()
    w.layers, w.width, w.heads, w.vocab, w.context, w.out_dim = 12, 768, 12, 49408, 77, 512
    assert lib.oake_text_create(ctypes.byref(h), 0, ctypes.byref(w)) != 0
    assert b'text geometry' in lib.oake_last_error()
    need = ctypes.c_size_t()
    assert lib.oake_text_workspace_bytes(None, 1, 16, ctypes.byref(need)) != 0
    assert lib.oake_encode_text(None, None, 1, 16, None, None, 0, None) != 0
