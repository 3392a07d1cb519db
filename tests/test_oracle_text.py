"""CPU tier: the text-tower oracle (oracle/text.py, openai/CLIP ``encode_text`` as oadp/prompts/vild.py:56-72
calls it) pinned against an independent implementation, HuggingFace CLIPTextModelWithProjection."""
import pytest
import torch

from oracle import text as otext


@pytest.fixture(scope='module')
def small():
    return otext.init_text_params(3, layers=3, vocab=49408)


def test_hf_cross_check(small):
    tokens = otext.synthetic_tokens(6, 77, seed=1)
    want = otext.build_hf_model(small)(input_ids=tokens).text_embeds
    got = otext.encode_text(small, tokens)
    assert got.shape == (6, 512)
    assert float((got - want).abs().max()) < 2e-5


def test_full_depth_hf_cross_check():
    p = otext.init_text_params(0)
    tokens = otext.synthetic_tokens(3, 77, seed=2)
    with torch.no_grad():
        want = otext.build_hf_model(p)(input_ids=tokens).text_embeds
    assert float((otext.encode_text(p, tokens) - want).abs().max()) < 5e-5


def test_context_length_is_invisible(small):
    """Causal tower pooled at EOT: truncating the context to the longest prompt (what the fork's
    `adaptively_tokenize` presumably does) gives the result of the stock 77-token context."""
    tokens = otext.synthetic_tokens(5, 20, seed=3)
    padded = torch.zeros(5, 77, dtype=torch.int64)
    padded[:, :20] = tokens
    assert float((otext.encode_text(small, tokens) - otext.encode_text(small, padded)).abs().max()) < 1e-5


def test_rows_are_independent(small):
    tokens = otext.synthetic_tokens(4, 16, seed=4)
    full = otext.encode_text(small, tokens)
    assert float((otext.encode_text(small, tokens[2:3]) - full[2:3]).abs().max()) < 1e-5


def test_prompt_embeddings_are_means_of_unit_rows(small):
    batches = [otext.synthetic_tokens(7, 16, seed=s) for s in range(3)]
    e = otext.prompt_embeddings(small, batches)
    assert e.shape == (7, 512) and bool((e.norm(dim=-1) <= 1.0 + 1e-6).all())
