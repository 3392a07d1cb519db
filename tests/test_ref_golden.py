"""CPU tier: the oracle against tests/golden/ref_golden.pt -- outputs of the REFERENCE'S OWN source
files (oadp/oake/{base,globals,blocks,objects}.py, oadp/dp/{classifiers,utils}.py, executed in the
build container by tests/golden/make_ref_golden.py on stand-ins for the un-vendored third-party
packages).  This is what pins the oracle: block grids, bboxes, crop pixels, box expansion, masks,
the model surgery and hook-driven side stream, encode_image, and the classifier heads.
/root/reference is not touched here."""
import pathlib
import sys

import PIL.Image
import pytest
import torch

from oracle import classifier as ocls
from oracle import frontend as ofe
from oracle import vit

sys.path.insert(0, str(pathlib.Path(__file__).parent / 'golden'))
import make_ref_golden as mk  # noqa: E402  (helpers only: seeded inputs, checksums)

REF = torch.load(pathlib.Path(__file__).parent / 'golden' / 'ref_golden.pt', weights_only=False)


@pytest.fixture(scope='module')
def inputs():
    return mk.ref_inputs()


@pytest.fixture(scope='module')
def params():
    torch.set_num_threads(8)
    p = vit.init_visual_params(REF['weight_seed'])
    return p, vit.objects_surgery(p)


def test_fixture_came_from_reference_files():
    assert 'oadp/oake/objects.py' in REF['reference_files'] and 'oadp/dp/classifiers.py' in REF['reference_files']


def test_partition_every_length():
    """blocks.py:40-52 for every side length 200..1400."""
    for n, want in REF['partition'].items():
        assert ofe.partition(n) == want, n


def test_blocks_preprocess_bit_exact(inputs):
    """blocks.py:54-109: crop count, bboxes (incl. the xywh-style first row), pixels bit-exact."""
    images, _ = inputs
    for arr, want in zip(images, REF['blocks']):
        got = ofe.blocks_preprocess(PIL.Image.fromarray(arr))
        assert got.blocks.shape[0] == want['n']
        assert torch.equal(got.bboxes, want['bboxes'])
        assert torch.equal(got.bboxes.half(), want['bboxes_half'])
        assert torch.equal(mk.checksums(got.blocks), want['pixels'])


def test_globals_preprocess_bit_exact(inputs):
    images, _ = inputs
    for arr, want in zip(images, REF['globals']):
        got = ofe.globals_preprocess(PIL.Image.fromarray(arr)).unsqueeze(0)
        assert torch.equal(mk.checksums(got), want['pixels'])


def test_objects_preprocess_bit_exact(inputs):
    """objects.py:76-186: min_wh filter, adaptive expansion, PIL crop, masks -- bit-exact."""
    images, proposals = inputs
    for arr, prop, want in zip(images, proposals, REF['objects']):
        got = ofe.objects_preprocess(PIL.Image.fromarray(arr), torch.from_numpy(prop))
        assert torch.equal(got.bboxes, want['bboxes'])
        assert torch.equal(got.objectness, want['objectness'])
        assert torch.equal(got.expanded, want['expanded'])
        assert torch.equal(got.masks.to(torch.uint8), want['masks'])
        assert torch.equal(mk.checksums(got.objects), want['pixels'])
        assert got.bboxes.shape[0] == prop.shape[0] - 2  # the two < 4 px boxes are dropped, the 4 x 4 one kept


def test_objects_surgery_matches_reference_build_model(params):
    """objects.py:285-301: 14 x 14 grid, stride 16, padding 15, resampled positional table."""
    _, p197 = params
    om = REF['objects_model']
    assert om['grid'] == 14 and om['stride'] == (16, 16) and om['padding'] == (15, 15)
    assert torch.equal(p197['positional_embedding'], om['positional_embedding'])


def test_encode_image_vs_reference(params, inputs):
    """model.encode_image as called by globals.py:57 / blocks.py:129 (12 layers, fp32)."""
    p, _ = params
    images, _ = inputs
    for arr, gw, bw in zip(images[:2], REF['globals'], REF['blocks']):
        pil = PIL.Image.fromarray(arr)
        g = vit.encode_image(p, ofe.globals_preprocess(pil).unsqueeze(0))
        assert (g[0] - gw['raw']).abs().max() < 2e-5
        assert torch.equal(vit.normalize_half(g)[0], gw['embedding']) or \
            (vit.normalize_half(g)[0].float() - gw['embedding'].float()).abs().max() <= 2**-11
        b = vit.encode_image(p, ofe.blocks_preprocess(pil).blocks[:6])
        assert (b - bw['raw6']).abs().max() < 2e-5


def test_encode_objects_vs_reference_hooks(params, inputs):
    """model.visual(o, m) through the reference's Hooks + surgery (objects.py:198-314, :330)."""
    _, p197 = params
    images, proposals = inputs
    arr, prop, want = images[0], proposals[0], REF['objects'][0]
    ob = ofe.objects_preprocess(PIL.Image.fromarray(arr), torch.from_numpy(prop))
    got = vit.encode_objects(p197, ob.objects, ob.masks)
    assert (got - want['raw']).abs().max() < 3e-5
    cos = torch.nn.functional.cosine_similarity(got, want['raw'], dim=-1)
    assert (1 - cos).max() < 1e-6
    half = vit.normalize_half(got)
    assert (half.float() - want['embeddings'].float()).abs().max() <= 2**-10


@pytest.mark.parametrize('name', ['base', 'classifier', 'vild', 'vild_default'])
@pytest.mark.parametrize('with_bg', [False, True])
@pytest.mark.parametrize('training', [False, True])
def test_classifier_vs_reference(name, with_bg, training):
    """classifiers.py:19-112 + utils.py:47-51: logits and the hooked `_linear` output."""
    prompts, bases, novels, x, weight, bias, bg = mk.classifier_inputs()
    order = [prompts['names'].index(n) for n in bases + novels]  # classifiers.py:34-35
    text = prompts['embeddings'][order]
    nb, na = len(bases), len(bases) + len(novels)
    bg_ = bg if with_bg else None
    if name == 'base':
        y, h = ocls.base_forward(x, weight, bias, text, bg_, training, nb, na)
    elif name == 'classifier':
        y, h = ocls.classifier_forward(x, weight, bias, text, bg_, training, nb, na, prompts['scaler'].item(),
                                       prompts['bias'].item())
    elif name == 'vild':
        y, h = ocls.vild_forward(x, weight, bias, text, bg_, training, nb, na, 0.01, 0.007)
    else:
        y, h = ocls.vild_forward(x, weight, bias, text, bg_, training, nb, na)
    want = REF['classifier'][(name, with_bg, training)]
    assert torch.equal(torch.isinf(y), torch.isinf(want['logits']))
    fin = ~torch.isinf(y)
    assert (y[fin] - want['logits'][fin]).abs().max() <= 1e-5 * want['logits'][fin].abs().max()
    assert (h - want['hooked']).abs().max() < 1e-6
    if training:
        assert torch.isinf(y[:, nb:na]).all() and not torch.isinf(y[:, :nb]).any()


def test_classifier_bad_out_features_message():
    assert REF['classifier']['bad_out_features'] == '5'  # RuntimeError(str(out_features)), classifiers.py:43-44


# ----------------------------------------------------------------------------------------------
# The PRODUCT's host-side geometry (oadp_b200/frontend.py: what the CUDA front end is driven by)
# against the same reference-generated fixture -- no GPU needed, nothing from oracle/ involved.
def test_product_partition_and_block_plan_vs_reference(inputs):
    from oadp_b200 import frontend
    for n, want in REF['partition'].items():
        assert frontend.partition(n) == want, n
    images, _ = inputs
    for arr, want in zip(images, REF['blocks']):
        h, w = arr.shape[:2]
        plan = frontend.blocks_plan(w, h)
        assert 1 + len(plan.cells) == want['n']
        assert torch.equal(torch.tensor(plan.bboxes, dtype=torch.float32), want['bboxes'])


def test_product_objects_plan_vs_reference(inputs):
    from oadp_b200 import frontend
    images, proposals = inputs
    for arr, prop, want in zip(images, proposals, REF['objects']):
        h, w = arr.shape[:2]
        plan = frontend.objects_plan(prop, (w, h))
        assert torch.equal(torch.from_numpy(plan.bboxes), want['bboxes'])
        assert torch.equal(torch.from_numpy(plan.objectness), want['objectness'])
        assert torch.equal(torch.from_numpy(plan.expanded), want['expanded'])


# ---------------------------------------------------------------------------------------------- losses
@pytest.mark.parametrize('name,kw', [('asl_configs', dict(gamma_neg=4, gamma_pos=0)), ('asl_defaults', dict()),
                                     ('asl_sum_w16', dict(gamma_neg=4, gamma_pos=0, reduction='sum', weight=16.0))])
def test_asymmetric_loss_vs_reference(name, kw):
    """oadp/base/losses.py:10-65, value and gradient (the focusing weight carries no gradient)."""
    from oracle import losses as ol
    probs, targets, _, _ = mk.loss_inputs()
    x = probs.clone().requires_grad_(True)
    loss = ol.asymmetric_loss(x, targets, **kw)
    loss.backward()
    want = REF['losses'][name]
    assert abs(float(loss) - float(want['loss'])) <= 1e-6 * abs(float(want['loss']))
    assert (x.grad - want['grad']).abs().max() <= 1e-5 * want['grad'].abs().max()


@pytest.mark.parametrize('name,kw', [('rkd', dict()), ('rkd_w8', dict(weight=8.0))])
def test_rkd_loss_vs_reference(name, kw):
    from oracle import losses as ol
    _, _, student, teacher = mk.loss_inputs()
    s = student.clone().requires_grad_(True)
    loss = ol.rkd_loss(s, teacher, **kw)
    loss.backward()
    want = REF['losses'][name]
    assert abs(float(loss) - float(want['loss'])) <= 1e-5 * abs(float(want['loss']))
    assert (s.grad - want['grad']).abs().max() <= 1e-5 * want['grad'].abs().max()


@pytest.mark.parametrize('k', [1, 5, 20, 'no_positive'])
def test_topk_recall_vs_reference(k):
    """oadp/dp/utils.py:13-44 (sklearn macro recall on the host) vs the device-side restatement the
    product ships (pure torch reductions; runs on CPU tensors as well)."""
    from oadp_b200.dp.utils import MultilabelTopKRecall
    logits, targets = mk.recall_inputs()
    if k == 'no_positive':
        got = MultilabelTopKRecall(k=5)(logits, torch.zeros_like(targets))
    else:
        got = MultilabelTopKRecall(k=k)(logits, targets)
    want = REF['recall'][k]
    assert got.shape == want.shape and got.dtype == want.dtype
    if torch.isnan(want):
        assert torch.isnan(got)
    else:
        assert abs(float(got) - float(want)) < 1e-4
