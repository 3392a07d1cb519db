"""GPU tier: the registered heads against the fixture produced by executing the
reference's oadp/dp/bbox_heads.py + roi_heads.py (tests/golden/make_ref_heads_golden.py).  The reference's
state dict is loaded into the product's `OADPRoIHead`; the cosine classifiers, the ViLD ensemble and the block
loss then run on liboake_b200 and must reproduce what the reference computed: inference `cls_score`
(roi_heads.py:93-112), the -inf pattern of `ObjectMixin.forward` / `Globals.training`, the hooked `_linear`
rows the distiller reads, `BlockMixin.loss` value, recall and gradients.
Tolerances: the classifiers multiply fp16-rounded operands on the tensor cores (fp32 accumulate) and the
ViLD head divides by 0.007, so log-scores agree to ~1e-1 absolute at |logit| ~ 10; hooked unit rows to 2e-3."""
import pathlib
import sys

import pytest
import torch
import torch.nn as nn

from oadp_b200.dp import categories

GOLDEN = pathlib.Path(__file__).resolve().parent / 'golden'
sys.path.insert(0, str(GOLDEN))
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent))

import make_ref_heads_golden as mrh  # noqa: E402
from test_ref_heads_golden import build_product_head  # noqa: E402

pytestmark = pytest.mark.gpu


class Passthrough(nn.Module):
    """The fixture's RoI extractor: `feats[0]` already holds one (C, 7, 7) feature per RoI."""
    num_inputs = 1

    def forward(self, feats, rois):
        return feats[0]


@pytest.fixture()
def setup(lib, tmp_path, monkeypatch):
    f = torch.load(GOLDEN / 'ref_heads_golden.pt')
    head = build_product_head(tmp_path, monkeypatch, f['state_dict']).cuda()
    head.bbox_roi_extractor = Passthrough()
    return f, head


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def test_inference_ensemble_matches_the_reference(setup, monkeypatch):
    f, head = setup
    feats, rois, *_ = [t.cuda() if torch.is_tensor(t) else t for t in mrh.head_inputs()]
    monkeypatch.setenv('DUMP', '/tmp/unused')
    categories.Globals.training = False
    with torch.no_grad():
        out = head._bbox_forward([feats], rois)
    got, want = out['cls_score'].cpu(), f['eval_cls_score']
    assert got.shape == want.shape and torch.isfinite(got).all()
    assert (got - want).abs().max() < 0.15 and rel(got, want) < 5e-3
    assert (got.exp().sum(-1) - 1).abs().max() < 1e-4
    assert torch.equal(out['bbox_feats'], feats) and rel(out['bbox_pred'].cpu(), f['eval_bbox_pred']) < 1e-5
    # Store.DUMP: the two logit matrices are kept for the NNI search (roi_heads.py:103-105)
    bl, ol = head._bbox_logits.cpu(), head._object_logits.cpu()
    assert rel(bl, f['eval_bbox_logits']) < 5e-3
    assert torch.equal(torch.isinf(ol), torch.isinf(f['eval_object_logits']))  # last column only
    fin = torch.isfinite(ol)
    assert rel(ol[fin], f['eval_object_logits'][fin]) < 5e-3


def test_training_mode_and_distiller_hooks(setup):
    f, head = setup
    feats, rois, *_ = [t.cuda() if torch.is_tensor(t) else t for t in mrh.head_inputs()]
    hooked = {}
    head._object_head.fc_cls._linear.register_forward_hook(lambda m, i, o: hooked.__setitem__('objects', o.detach()))
    categories.Globals.training = True
    try:
        with torch.no_grad():
            got = head._bbox_forward([feats], rois)['cls_score'].cpu()  # as StandardRoIHead: raw bbox-head logits
            head.object_forward_train([feats], [rois[:20, 1:], rois[20:, 1:]])
            obj = head._object_head(feats)[0].cpu()
    finally:
        categories.Globals.training = False
    want = f['train_cls_score']
    assert torch.equal(torch.isinf(got), torch.isinf(want))  # novel columns (classifiers.py:62-67)
    fin = torch.isfinite(want)
    assert rel(got[fin], want[fin]) < 5e-3
    assert (hooked['objects'].cpu() - f['train_object_hooked']).abs().max() < 2e-3
    assert (hooked['objects'].norm(dim=-1) - 1).abs().max() < 1e-3
    # ObjectMixin.forward + training: novel columns AND the background column are -inf, fused into one range
    assert torch.equal(torch.isinf(obj), torch.isinf(f['train_object_logits']))
    fin = torch.isfinite(obj)
    assert rel(obj[fin], f['train_object_logits'][fin]) < 5e-3


def test_block_loss_value_recall_and_gradients(setup):
    f, head = setup
    _, _, block_feats, block_boxes, block_targets = mrh.head_inputs()
    hooked = {}
    head._block_head.fc_cls._linear.register_forward_hook(lambda m, i, o: hooked.__setitem__('blocks', o.detach()))
    bf = block_feats.cuda().requires_grad_(True)
    categories.Globals.training = True
    try:
        losses = head.block_forward_train([bf], [b.cuda() for b in block_boxes], [t.cuda() for t in block_targets])
    finally:
        categories.Globals.training = False
    assert set(losses) == {'loss_block', 'recall_block'}
    assert abs(float(losses['loss_block']) - float(f['block_loss'])) < 5e-3 * float(f['block_loss'])
    assert abs(float(losses['recall_block']) - float(f['block_recall'])) < 1e-3
    assert (hooked['blocks'].cpu() - f['block_hooked']).abs().max() < 2e-3
    losses['loss_block'].backward()
    assert rel(bf.grad.cpu(), f['block_feats_grad']) < 2e-2
    assert rel(head._block_head.fc_cls._linear.weight.grad.cpu(), f['block_fc_cls_weight_grad']) < 2e-2
    # the frozen background row of the object head never receives a gradient; the block head's does
    assert head._object_head.fc_cls._bg_embedding.grad is None
