"""GPU tier for the text tower (second half): `oake_encode_text` through the C-ABI against the
fp32 oracle (oracle/text.py, pinned to HuggingFace CLIP).  Tolerance: 1 - cosine < 1e-3 per row (the north
star's bar for the image tower) and max-abs < 2e-2 of the row norm.
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import text as otext

pytestmark = pytest.mark.gpu


def _check(got, want):
    got, want = got.float().cpu(), want.float()
    assert got.shape == want.shape
    assert float((1 - F.cosine_similarity(got, want, dim=-1)).max()) < 1e-3
    assert float(((got - want).abs().max(dim=-1).values / want.norm(dim=-1)).max()) < 2e-2


@pytest.mark.parametrize('layers,length,n', [(1, 16, 5), (2, 77, 9), (12, 20, 300)])
def test_encode_text_matches_oracle(lib, layers, length, n):
    from oadp_b200.text import OakeTextModel
    p = otext.init_text_params(11, layers=layers)
    tokens = otext.synthetic_tokens(n, length, seed=layers)
    model = OakeTextModel(p, 'cuda')
    _check(model.encode_text(tokens), otext.encode_text(p, tokens))
    # rows are independent of the batch they travel in, and of the context length
    _check(model.encode_text(tokens[3:4]), otext.encode_text(p, tokens)[3:4])
    if length < 77:
        padded = torch.zeros(n, 77, dtype=torch.int64)
        padded[:, :length] = tokens
        _check(model.encode_text(padded), otext.encode_text(p, tokens))


def test_build_prompts_on_the_gpu(lib):
    from oadp_b200.text import OakeTextModel, build_prompts
    p = otext.init_text_params(12, layers=2)
    model = OakeTextModel(p, 'cuda')
    templates = ['a photo of a {}', 'There is a {} in the scene']
    names = ['person', 'bicycle', 'traffic light']
    words = {}

    def tokenize(texts):
        rows = []
        for t in texts:
            ids = [otext.SOT] + [words.setdefault(w, 1 + len(words)) for w in t.split()] + [otext.EOT]
            rows.append(ids + [0] * (16 - len(ids)))
        return torch.tensor(rows)

    got = build_prompts(model.encode_text, tokenize, templates, names)
    want = otext.prompt_embeddings(p, [tokenize([t.format(c) for c in sorted(names)]) for t in templates])
    _check(got['embeddings'], want)


def test_out_of_vocabulary_token_is_an_error(lib):
    from oadp_b200.text import OakeTextModel
    p = otext.init_text_params(13, layers=1)
    model = OakeTextModel(p, 'cuda')
    tokens = otext.synthetic_tokens(2, 8, seed=0)
    tokens[1, 3] = p['token_embedding.weight'].shape[0]
    with pytest.raises(ValueError, match='token ids'):
        model.encode_text(tokens)
    tokens[1, 3] = -1
    with pytest.raises(ValueError, match='token ids'):
        model.encode_text(tokens)
