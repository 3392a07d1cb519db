"""GPU tier: whole-tower parity of the sm_100a path (through the C ABI) against the CPU oracle and
the committed golden outputs.  Bar (BASELINE.json north_star): cosine >= 1 - 1e-3 per crop against
the fp32 reference; we additionally bound the relative L2 error."""
import pathlib
import sys

import pytest
import torch
import torch.nn.functional as F

from oadp_b200 import binding
from oadp_b200.model import OakeModel
from oracle import vit

sys.path.insert(0, str(pathlib.Path(__file__).parent / 'golden'))
import make_golden  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = torch.load(pathlib.Path(__file__).parent / 'golden' / 'vit_golden.pt')

COS_TOL = 1e-3  # north_star: "within 1e-3 cosine"
REL_L2_TOL = 5e-3  # measured 1.3e-3 (fp16 operands and residual stream, fp32 accumulation)
CENTRED_TOL = 2e-3  # 1 - cosine after removing the direction all crops of a random tower share


def check_close(got: torch.Tensor, want: torch.Tensor, what: str):
    got, want = got.float().cpu(), want.float().cpu()
    assert torch.isfinite(got).all(), what
    cos = F.cosine_similarity(got, want, dim=-1)
    rel = (got - want).norm(dim=-1) / want.norm(dim=-1)
    assert (1 - cos).max() < COS_TOL, (what, (1 - cos).max().item())
    assert rel.max() < REL_L2_TOL, (what, rel.max().item())
    return (1 - cos).max().item(), rel.max().item()


@pytest.fixture(scope='module')
def model(lib):
    return OakeModel(vit.init_visual_params(GOLDEN['weight_seed']), 'cuda')


def test_t50_matches_golden(model):
    pixels, _ = make_golden.golden_inputs()
    raw = model.encode_image(pixels.cuda())
    c, r = check_close(raw, GOLDEN['t50'], 't50 raw')
    print(f't50: 1-cos {c:.2e} rel-l2 {r:.2e}')
    emb = model.embed(pixels.cuda())
    assert emb.dtype == torch.float16 and emb.shape == (4, 512)
    check_close(emb, vit.normalize_half(GOLDEN['t50']), 't50 normalised')
    assert ((emb.float().norm(dim=-1) - 1).abs() < 2e-3).all()


def test_t197_matches_golden(model):
    pixels, masks = make_golden.golden_inputs()
    model.for_objects()
    raw = model.visual(pixels.cuda(), masks.cuda())
    c, r = check_close(raw, GOLDEN['t197'], 't197 raw')
    print(f't197: 1-cos {c:.2e} rel-l2 {r:.2e}')
    emb = model.embed(pixels.cuda(), masks.cuda())
    check_close(emb, vit.normalize_half(GOLDEN['t197']), 't197 normalised')


def centred(got: torch.Tensor, want: torch.Tensor, what: str) -> float:
    """Seeded random towers map every crop near ONE common direction, so the plain cosine is blind (2e-7 while
    the relative L2 error is 1e-3).  Remove the oracle's mean row from both sides and compare what is left:
    the part of the embedding that actually depends on the crop."""
    got, want = got.float().cpu(), want.float().cpu()
    mu = want.mean(0, keepdim=True)
    cos = F.cosine_similarity(got - mu, want - mu, dim=-1)
    spread = ((want - mu).norm(dim=-1) / want.norm(dim=-1)).mean().item()
    worst = (1 - cos).max().item()
    print(f'{what}: centred 1-cos {worst:.2e} (crop-dependent part = {spread:.1%} of the row norm)')
    return worst


def test_centered_agreement(model):
    g = torch.Generator().manual_seed(99)
    pixels = torch.randn(24, 3, 224, 224, generator=g) * 1.2
    p = vit.init_visual_params(GOLDEN['weight_seed'])
    want = vit.encode_image(p, pixels)
    got = model.encode_image(pixels.cuda())
    assert centred(got, want, 't50') < CENTRED_TOL
    check_close(got, want, 't50 random crops')


def test_centered_agreement_t197_side_stream(model):
    """The same for the objects tower: T = 197 tokens and the mask-attended side stream (objects.py:198-266),
    with masks of every kind -- empty, full, and random foreground boxes."""
    g = torch.Generator().manual_seed(98)
    n = 12
    pixels = torch.randn(n, 3, 224, 224, generator=g) * 1.2
    masks = torch.ones(n, 1, 14, 14)
    for i in range(n):
        x0, y0 = torch.randint(0, 10, (2, ), generator=g).tolist()
        w, h = torch.randint(2, 14, (2, ), generator=g).tolist()
        masks[i, 0, y0:y0 + h, x0:x0 + w] = 0  # 0 = foreground
    masks[0] = 0
    masks[1] = 1
    p197 = vit.objects_surgery(vit.init_visual_params(GOLDEN['weight_seed']))
    want = vit.encode_objects(p197, pixels, masks)
    model.for_objects()
    got = model.visual(pixels.cuda(), masks.cuda())
    assert centred(got, want, 't197 + side stream') < CENTRED_TOL
    check_close(got, want, 't197 random crops')
    # the side stream really is in play: the same crops with opposite masks give different rows
    other = model.visual(pixels.cuda(), (1 - masks).cuda()).cpu()
    assert (F.cosine_similarity(other - want.mean(0, keepdim=True), want - want.mean(0, keepdim=True), dim=-1) < 0.999).any()


def test_batch_composition_is_invisible(model):
    """chunking / batch composition must not change a crop's result."""
    g = torch.Generator().manual_seed(5)
    pixels = torch.randn(37, 3, 224, 224, generator=g).cuda()
    masks = (torch.rand(37, 1, 14, 14, generator=g) > 0.5).float().cuda()
    full = model.embed(pixels)
    assert torch.equal(full[11:12], model.embed(pixels[11:12]))
    perm = torch.randperm(37, generator=g).cuda()
    assert torch.equal(model.embed(pixels[perm]), full[perm])
    model.for_objects()
    fullo = model.embed(pixels, masks)
    assert torch.equal(fullo[30:31], model.embed(pixels[30:31], masks[30:31]))
    assert torch.equal(model.embed(pixels[perm], masks[perm]), fullo[perm])


def test_mask_changes_objects_embedding(model):
    g = torch.Generator().manual_seed(6)
    pixels = torch.randn(2, 3, 224, 224, generator=g).cuda()
    a = model.embed(pixels, torch.zeros(2, 1, 14, 14).cuda())
    b = model.embed(pixels, torch.ones(2, 1, 14, 14).cuda())
    assert torch.isfinite(b.float()).all()
    assert (a.float() - b.float()).abs().max() > 1e-3


def test_bad_arguments_raise(model):
    with pytest.raises(ValueError):
        model.encode_image(torch.zeros(1, 3, 100, 100).cuda())
    with pytest.raises(ValueError):
        model.engine.encode_pixels(torch.zeros(1, 3, 224, 224).cuda(), None, binding.VARIANT_T197)
    assert model.engine.launch_count() > 0
