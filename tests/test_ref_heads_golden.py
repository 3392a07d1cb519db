"""CPU tier: the fixture produced by EXECUTING the reference's oadp/dp/bbox_heads.py + roi_heads.py
(tests/golden/make_ref_heads_golden.py) pins
  * oracle/classifier.py::vild_ensemble / object_head_logits to `ViLDEnsembleRoIHead._bbox_forward` output;
  * the product's head classes: the reference's state dict loads into them key for key (same module names),
    and everything up to `fc_cls` (mmdet's conv / fc stack, plain torch) reproduces the features the
    reference's classifier saw.  The classifier itself needs the GPU: tests/test_gpu_heads.py."""
import copy
import pathlib
import sys

import pytest
import torch

from oadp_b200.dp import categories
from oadp_b200.registry import HEADS
from oracle import classifier as ocls

GOLDEN = pathlib.Path(__file__).resolve().parent / 'golden'
sys.path.insert(0, str(GOLDEN))

import make_ref_golden as mrg  # noqa: E402
import make_ref_heads_golden as mrh  # noqa: E402  (pure helpers: config, seeded inputs; no reference access)


@pytest.fixture(scope='module')
def fixture():
    return torch.load(GOLDEN / 'ref_heads_golden.pt')


def build_product_head(tmp_path, monkeypatch, state_dict):
    prompts, bases, novels, *_ = mrg.classifier_inputs()
    monkeypatch.setattr(categories.Globals, 'categories', categories.Categories(bases, novels), raising=False)
    monkeypatch.setattr(categories.Globals, 'training', False, raising=False)
    ppath = tmp_path / 'prompts.pth'
    torch.save(prompts, ppath)
    cfg = copy.deepcopy(mrh.head_config(str(ppath)))
    cfg['bbox_roi_extractor'] = dict(type='SingleRoIExtractor', roi_layer=dict(type='RoIAlign', output_size=7, sampling_ratio=0),
                                     out_channels=mrh.CHANNELS, featmap_strides=[4])
    head = HEADS.build(dict(type='OADPRoIHead', **cfg)).eval()
    missing, unexpected = head.load_state_dict(state_dict, strict=True)
    assert not missing and not unexpected
    return head


def test_ensemble_oracle_is_pinned_to_the_reference(fixture):
    f = fixture
    # ObjectMixin.forward: the object head's last logit is -inf, with and without Globals.training
    assert torch.isinf(f['eval_object_logits'][:, -1]).all() and torch.isinf(f['train_object_logits'][:, -1]).all()
    assert torch.isinf(f['train_object_logits'][:, 4:6]).all()  # novel columns while training (classifiers.py:62-67)
    want = f['eval_cls_score']
    got = ocls.vild_ensemble(f['eval_bbox_logits'].clone(), f['eval_object_logits'].clone(), f['lambda'])
    assert torch.equal(got, want)  # same lines, same fp32 arithmetic: bit-exact
    assert torch.allclose(f['lambda'], torch.tensor([2 / 3] * 4 + [1 / 3] * 3))
    # training mode leaves the bbox head's scores untouched (roi_heads.py:91-92)
    assert f['train_cls_score'].shape == want.shape and torch.isinf(f['train_cls_score'][:, 4:6]).all()
    # (value of -inf columns: softmax^(1-lambda) of a -inf logit is 0 -> log 0 = -inf for the novel-free bg column)
    assert torch.isfinite(want[:, :6]).all()


def test_reference_state_dict_loads_into_the_product_heads(fixture, tmp_path, monkeypatch):
    head = build_product_head(tmp_path, monkeypatch, fixture['state_dict'])
    assert torch.equal(head.lambda_, fixture['lambda'])
    assert head._object_head.fc_cls._bg_embedding.requires_grad is False
    feats, rois, block_feats, *_ = mrh.head_inputs()
    # mmdet's part of the path (shared convs / fcs, here plain torch on the CPU) up to the classifier input:
    # F.normalize(linear(x)) of those features must be what the reference's hook captured
    with torch.no_grad():
        x = feats
        for conv in head._object_head.shared_convs:
            x = conv(x)
        x = x.flatten(1)
        for fc in head._object_head.shared_fcs:
            x = torch.relu(fc(x))
        lin = head._object_head.fc_cls._linear
        hooked = ocls.normalized_linear(x, lin.weight, lin.bias)
    assert (hooked - fixture['train_object_hooked']).abs().max() < 1e-5


def test_block_loss_oracle_is_pinned(fixture, tmp_path, monkeypatch):
    """BlockMixin.loss on `logits[:, :-1]` (roi_heads.py:208): ASL of the sigmoid + top-k recall, restated with
    oracle/losses.py and oadp_b200.dp.utils on the CPU from the hooked rows."""
    from oadp_b200.dp.utils import MultilabelTopKRecall
    from oracle import losses as olosses
    head = build_product_head(tmp_path, monkeypatch, fixture['state_dict'])
    *_, block_targets = mrh.head_inputs()
    clf = head._block_head.fc_cls
    h = fixture['block_hooked']
    logits = h @ ocls.embeddings(clf._embeddings, clf._bg_embedding.detach()).T
    logits[:, 4:6] = float('-inf')  # Globals.training was True in block_forward_train
    logits = logits * clf._scaler - clf._bias
    targets = torch.cat(block_targets)
    loss = olosses.asymmetric_loss(logits[:, :-1].sigmoid(), targets, gamma_neg=4, gamma_pos=0, weight=16.0)
    assert abs(float(loss) - float(fixture['block_loss'])) < 1e-4 * abs(float(fixture['block_loss']))
    recall = MultilabelTopKRecall(k=2)(logits[:, :-1], targets)
    assert abs(float(recall) - float(fixture['block_recall'])) < 1e-4


def test_expand_mode_constant_is_pinned(fixture):
    """`ExpandMode.CONSTANT` (objects.py:92-93): the reference's `_expand` / `_preprocess` run with that mode
    (same fixture file) pin the product's host geometry and the oracle front end bit-exactly."""
    import numpy as np
    import PIL.Image
    from oadp_b200 import frontend
    from oracle import frontend as ofe
    images, proposals = mrg.ref_inputs()
    for arr, prop, want in zip(images[:2], proposals[:2], fixture['expand_constant']):
        plan = frontend.objects_plan(prop, (arr.shape[1], arr.shape[0]), expand_mode='CONSTANT')
        assert np.array_equal(plan.expanded, want['expanded'].numpy()) and np.array_equal(plan.bboxes, want['bboxes'].numpy())
        side = plan.expanded[:, 2:] - plan.expanded[:, :2]
        assert np.allclose(side, 224.0, atol=1e-3)
        o = ofe.objects_preprocess(PIL.Image.fromarray(arr), torch.from_numpy(prop), expand_mode='CONSTANT')
        assert torch.equal(o.expanded, want['expanded']) and torch.equal(o.masks.to(torch.uint8), want['masks'])
        assert torch.equal(mrg.checksums(o.objects), want['pixels'])
