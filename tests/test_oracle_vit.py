"""CPU tier: pins the encoder oracle (oracle/vit.py) -- against HF CLIP, against the hook-driven
second restatement, against the committed golden outputs, and through the invariants SURVEY 8(c)
lists.  The reference itself ships no tests / vectors (parity unpinned upstream)."""
import pathlib
import sys

import pytest
import torch

from oracle import hooks_ref, vit

sys.path.insert(0, str(pathlib.Path(__file__).parent / 'golden'))
import make_golden  # noqa: E402

GOLDEN = torch.load(pathlib.Path(__file__).parent / 'golden' / 'vit_golden.pt')


@pytest.fixture(scope='module')
def small():
    p = vit.init_visual_params(7, layers=2)
    return p, vit.objects_surgery(p)


def test_hf_cross_check_t50(small):
    p, _ = small
    x = torch.randn(3, 3, 224, 224, generator=torch.Generator().manual_seed(0))
    ours = vit.encode_image(p, x)
    with torch.no_grad():
        theirs = vit.build_hf_model(p)(pixel_values=x).image_embeds
    assert (ours - theirs).abs().max() < 2e-5


def test_hooks_cross_check_t197(small):
    _, p197 = small
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 3, 224, 224, generator=g)
    m = (torch.rand(3, 1, 14, 14, generator=g) > 0.5).float()
    a = vit.encode_objects(p197, x, m)
    b = hooks_ref.HookedVisual(p197)(x, m)
    assert (a - b).abs().max() < 2e-5


def test_golden_full_depth():
    """12-layer oracle reproduces the committed outputs (which agree with HF / hooks to 3e-6)."""
    torch.set_num_threads(8)
    pixels, masks = make_golden.golden_inputs()
    assert abs(pixels.double().sum().item() - GOLDEN['pixels_checksum']) < 1e-6
    p = vit.init_visual_params(GOLDEN['weight_seed'])
    assert (vit.encode_image(p, pixels[:2]) - GOLDEN['t50'][:2]).abs().max() < 1e-4
    assert (GOLDEN['t50'] - GOLDEN['t50_hf']).abs().max() < 1e-5
    assert (GOLDEN['t197'] - GOLDEN['t197_hooks']).abs().max() < 1e-5
    out = vit.encode_objects(vit.objects_surgery(p), pixels[:2], masks[:2])
    assert (out - GOLDEN['t197'][:2]).abs().max() < 1e-4


def test_row_independence_and_permutation(small):
    p, p197 = small
    g = torch.Generator().manual_seed(2)
    x = torch.randn(4, 3, 224, 224, generator=g)
    m = (torch.rand(4, 1, 14, 14, generator=g) > 0.5).float()
    full = vit.encode_objects(p197, x, m)
    one = vit.encode_objects(p197, x[2:3], m[2:3])
    assert (full[2:3] - one).abs().max() < 2e-5
    perm = torch.tensor([3, 1, 0, 2])
    assert (vit.encode_objects(p197, x[perm], m[perm]) - full[perm]).abs().max() < 2e-5
    assert (vit.encode_image(p, x)[1:2] - vit.encode_image(p, x[1:2])).abs().max() < 2e-5


def test_mask_semantics(small):
    _, p197 = small
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 224, 224, generator=g)
    zeros = torch.zeros(2, 1, 14, 14)
    ones = torch.ones(2, 1, 14, 14)
    a = vit.encode_objects(p197, x, zeros)
    b = vit.encode_objects(p197, x, ones)
    assert (a - b).abs().max() > 1e-3  # the mask matters
    # bias is -100, not -inf: with everything masked y still attends (to itself, weight ~1)
    assert torch.isfinite(b).all()
    bias = vit.mask_to_bias(ones)
    assert bias.shape == (2, 197) and (bias[:, :196] == -100).all() and (bias[:, 196] == 0).all()


def test_surgery_geometry(small):
    p, p197 = small
    assert p197['positional_embedding'].shape == (197, 768)
    assert torch.equal(p197['positional_embedding'][0], p['positional_embedding'][0])
    tok = vit.embed(p197, torch.zeros(1, 3, 224, 224), stride=16, padding=15)
    assert tok.shape == (1, 197, 768)
    assert vit.embed(p, torch.zeros(1, 3, 224, 224), stride=32, padding=0).shape == (1, 50, 768)
