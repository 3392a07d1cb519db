"""CPU tier: the detector-side registry surface.  Every name the reference registers on the hot
path's boundary -- HEADS (`Shared2FCBlockBBoxHead`, `Shared4Conv1FCObjectBBoxHead`, `ViLDEnsembleRoIHead`,
`OADPRoIHead`), LINEAR_LAYERS (classifiers), PIPELINES (`LoadCLIPFeatures`), todd `LossRegistry`
(`AsymmetricLoss`, `RKDLoss` + the `L1Loss` / `MSELoss` the distiller configs name) -- is built from the config
dicts the reference ships (configs/dp/**, read from /root/reference when it is mounted, and from an inline
copy of the same keys otherwise).  No kernel runs here: construction, attribute names (the distiller hook
paths), frozen background row, `num_classes` detection, the alias import paths."""
import copy
import os
import pathlib

import pytest
import torch

from oadp_b200.compat import Config
from oadp_b200.dp import categories
from oadp_b200.registry import HEADS, LINEAR_LAYERS, PIPELINES, LossRegistry, Registry, build_linear_layer

REF_CFG = pathlib.Path('/root/reference/configs/dp')

# the keys of configs/dp/models/{faster_rcnn_r50_fpn,vild_ensemble_faster_rcnn_r50_fpn,block,global_}.py and
# configs/dp/oadp_ov_coco.py that reach the registries on the boundary (merged as todd.Config merges them)
INLINE_ROI_HEAD = dict(
    type='OADPRoIHead',
    bbox_roi_extractor=dict(type='SingleRoIExtractor', roi_layer=dict(type='RoIAlign', output_size=7, sampling_ratio=0),
                            out_channels=256, featmap_strides=[4, 8, 16, 32]),
    bbox_head=dict(type='Shared4Conv1FCBBoxHead', in_channels=256, fc_out_channels=1024, roi_feat_size=7,
                   bbox_coder=dict(type='DeltaXYWHBBoxCoder', target_means=[0., 0., 0., 0.], target_stds=[0.1, 0.1, 0.2, 0.2]),
                   reg_class_agnostic=True, norm_cfg=dict(type='SyncBN', requires_grad=True),
                   loss_cls=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0),
                   loss_bbox=dict(type='L1Loss', loss_weight=1.0), num_classes=None,
                   cls_predictor_cfg=dict(type='ViLDClassifier', prompts='data/prompts/vild.pth')),
    object_head=dict(type='Shared4Conv1FCObjectBBoxHead',
                     cls_predictor_cfg=dict(type='Classifier', prompts='data/prompts/ml_coco.pth')),
    block_head=dict(type='Shared2FCBlockBBoxHead', topk=5,
                    loss=dict(type='AsymmetricLoss', weight=dict(type='WarmupScheduler', gain=16, end=1000), gamma_neg=4,
                              gamma_pos=0),
                    cls_predictor_cfg=dict(type='Classifier', prompts='data/prompts/ml_coco.pth')),
)
INLINE_GLOBAL_CLASSIFIER = dict(type='Classifier', prompts='data/prompts/ml_coco.pth', out_features=65, in_features=256)
INLINE_DISTILLER_LOSSES = dict(
    loss_clip_objects=dict(type='L1Loss', weight=dict(type='WarmupScheduler', gain=256, end=200)),
    loss_clip_blocks=dict(type='L1Loss', weight=dict(type='WarmupScheduler', gain=128, end=200)),
    loss_clip_block_relations=dict(type='RKDLoss', weight=dict(type='WarmupScheduler', gain=8, end=200)),
    loss_clip_global=dict(type='MSELoss', weight=dict(type='WarmupScheduler', gain=0.5, end=200), reduction='sum'),
)
HOOK_PATHS = ('.roi_head._object_head.fc_cls._linear', '.roi_head._block_head.fc_cls._linear',
              '._global_head._classifier._linear')


@pytest.fixture()
def workdir(tmp_path, monkeypatch):
    """A working directory holding the prompts files at the RELATIVE paths the reference's configs name."""
    g = torch.Generator().manual_seed(3)
    names = sorted(set(categories.coco.all_) | {'zebra crossing', 'unicorn'})
    (tmp_path / 'data' / 'prompts').mkdir(parents=True)
    torch.save(dict(names=names, embeddings=torch.randn(len(names), 512, generator=g) * 0.05),
               tmp_path / 'data' / 'prompts' / 'vild.pth')
    torch.save(dict(names=names, embeddings=torch.randn(len(names), 512, generator=g) * 0.05, scaler=torch.tensor([4.0]),
                    bias=torch.tensor([0.5])), tmp_path / 'data' / 'prompts' / 'ml_coco.pth')
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(categories.Globals, 'categories', categories.coco, raising=False)
    monkeypatch.setattr(categories.Globals, 'training', False, raising=False)
    monkeypatch.delenv('DRY_RUN', raising=False)
    return tmp_path


def _resolve(root, dotted):
    obj = root
    for part in dotted.strip('.').split('.'):
        obj = getattr(obj, part)
    return obj


def _check_roi_head(head):
    from oadp_b200.dp import (Classifier, NormalizedLinear, OADPRoIHead, Shared2FCBlockBBoxHead,
                              Shared4Conv1FCObjectBBoxHead, ViLDClassifier)
    assert isinstance(head, OADPRoIHead) and head.with_block
    assert head.bbox_head.num_classes == 65  # detected from Globals.categories (roi_heads.py:32-34)
    assert isinstance(head.bbox_head.fc_cls, ViLDClassifier) and head.bbox_head.fc_cls._embeddings.shape == (65, 512)
    assert head.bbox_head.fc_cls._bg_embedding.shape == (1, 512)  # num_classes + 1 outputs: learnable background row
    assert head.bbox_head.fc_reg.out_features == 4  # reg_class_agnostic
    obj, blk = head._object_head, head._block_head
    assert isinstance(obj, Shared4Conv1FCObjectBBoxHead) and isinstance(blk, Shared2FCBlockBBoxHead)
    assert isinstance(obj.fc_cls, Classifier) and isinstance(blk.fc_cls, Classifier)
    assert not obj.with_reg and not blk.with_reg and not hasattr(obj, 'fc_reg')  # NotWithRegMixin
    assert obj.fc_cls._bg_embedding.requires_grad is False  # ObjectMixin freezes it (bbox_heads.py:50-55)
    assert blk.fc_cls._bg_embedding.requires_grad is True
    assert obj.fc_cls.disable_bg_column and not blk.fc_cls.disable_bg_column
    assert len(obj.shared_convs) == 4 and len(obj.shared_fcs) == 1 and len(blk.shared_fcs) == 2
    assert obj.fc_cls._linear.in_features == 1024 and blk.fc_cls._linear.in_features == 1024
    lam = head.lambda_
    assert lam.shape == (66, ) and torch.allclose(lam[:48], torch.tensor(2 / 3)) and torch.allclose(lam[48:], torch.tensor(1 / 3))
    assert '_lambda' not in head.state_dict()  # non-persistent buffer (roi_heads.py:59)
    # the todd distiller hook paths of configs/dp/models/*.py resolve to the normalised-linear modules
    holder = type('Detector', (), {})()
    holder.roi_head = head
    for path in HOOK_PATHS[:2]:
        assert isinstance(_resolve(holder, path), NormalizedLinear)
    assert blk._multilabel_topk_recall._k == 5 and type(blk._loss).__name__ == 'AsymmetricLoss'
    assert blk._loss._params[:2] == (4.0, 0.0)


def test_inline_config_builds_every_head(workdir):
    head = HEADS.build(copy.deepcopy(INLINE_ROI_HEAD), default_args=dict(train_cfg=None, test_cfg=dict(score_thr=0.0)))
    _check_roi_head(head)
    from oadp_b200.dp import Classifier, NormalizedLinear
    clf = LINEAR_LAYERS.build(INLINE_GLOBAL_CLASSIFIER)  # GlobalHead's classifier: no background row (65 outputs)
    assert isinstance(clf, Classifier) and clf._bg_embedding is None and isinstance(clf._linear, NormalizedLinear)
    assert (clf._scaler, clf._bias) == (4.0, 0.5)
    for name, cfg in INLINE_DISTILLER_LOSSES.items():
        loss = LossRegistry.build(cfg)
        assert type(loss).__name__ == cfg['type']
        loss.step(100)
        assert loss.weight == pytest.approx(cfg['weight']['gain'] * 0.5)
    assert LossRegistry.build(INLINE_DISTILLER_LOSSES['loss_clip_global'])._reduction == 'sum'


@pytest.mark.skipif(not REF_CFG.exists(), reason='reference tree not mounted (GPU box)')
@pytest.mark.parametrize('config', ['oadp_ov_coco.py', 'vild_ov_coco.py'])
def test_reference_model_configs_build(workdir, config):
    """The reference's own config files, loaded unchanged with their `_base_` chains."""
    cfg = Config.load(REF_CFG / config)
    model = cfg.model
    assert model.type in ('OADP', 'ViLD')
    roi = copy.deepcopy(model.roi_head)
    head = HEADS.build(roi, default_args=dict(train_cfg=model.train_cfg.rcnn, test_cfg=model.test_cfg.rcnn))
    assert type(head).__name__ == model.roi_head.type
    if config == 'oadp_ov_coco.py':
        _check_roi_head(head)
        g = model.global_head
        clf = LINEAR_LAYERS.build(g.classifier)
        assert clf._linear.in_features == 256 and clf._bg_embedding is None
        assert type(LossRegistry.build(g.loss)).__name__ == 'AsymmetricLoss'
    else:
        from oadp_b200.dp import ViLDClassifier
        assert not hasattr(head, '_block_head')
        for h in (head.bbox_head, head._object_head):
            assert isinstance(h.fc_cls, ViLDClassifier) and h.fc_cls._scaler == dict(train=0.01, val=0.007)
    paths = [hook.action.path for hook in model.distiller.student_hooks.values()]
    assert set(paths) <= set(HOOK_PATHS)
    for loss in model.distiller.losses.values():
        assert LossRegistry.build(loss.action) is not None
    # the dataset pipeline names LoadCLIPFeatures; its stores open lazily per key
    steps = [s for s in cfg.trainer.dataloader.dataset.pipeline if s['type'] == 'LoadCLIPFeatures']
    assert len(steps) == 1
    step = PIPELINES.build(steps[0])
    assert step._globals is not None and step._blocks is not None and step._objects is not None


def test_load_clip_features_through_the_registry(workdir):
    from oadp_b200.store import PthStore
    g = torch.Generator().manual_seed(1)
    root = workdir / 'oake'
    PthStore(str(root / 'globals'), 'train2017')['000000000007'] = torch.randn(512, generator=g).half()
    PthStore(str(root / 'objects'), 'train2017')['000000000007'] = dict(
        embeddings=torch.randn(3, 512, generator=g).half(),
        bboxes=torch.tensor([[0, 0, 10, 10], [5, 5, 8, 30], [1, 2, 40, 50]]).half(), objectness=torch.rand(3, 1).half())
    step = PIPELINES.build(dict(type='LoadCLIPFeatures', default=dict(task_name='train2017', type='PthAccessLayer'),
                                globals_=dict(data_root=str(root / 'globals')), objects=dict(data_root=str(root / 'objects'))))
    out = step(dict(img_info=dict(id=7), bbox_fields=[]))
    assert out['clip_global'].shape == (512, ) and out['clip_objects'].shape == (2, 512)  # the 3-px-wide box is dropped
    assert out['bbox_fields'] == ['object_bboxes']


def test_registry_semantics():
    reg = Registry('t')

    @reg.register_module()
    class A:
        def __init__(self, x, y=2):
            self.x, self.y = x, y

    with pytest.raises(KeyError):
        reg.register_module(name='A', module=int)
    reg.register_module(name='A', module=A)  # the same class again is a no-op
    assert reg.build(dict(type='A', x=1), default_args=dict(x=5, y=7)).__dict__ == dict(x=1, y=7)  # cfg wins
    with pytest.raises(KeyError):
        reg.build(dict(type='B'))
    with pytest.raises(KeyError):
        reg.build(dict(x=1))
    assert isinstance(build_linear_layer(None, 4, 3), torch.nn.Linear)
    with pytest.raises(KeyError):
        build_linear_layer(dict(type='Nope'), 4, 3)


def test_alias_import_paths():
    """`oadp.dp.*` / `oadp.base.*` resolve to the same objects as `oadp_b200.dp.*` (oadp/dp/__init__.py:1-6)."""
    import oadp.base
    import oadp.dp
    import oadp.dp.bbox_heads as bh
    import oadp.dp.classifiers as cl
    import oadp.dp.datasets as ds
    import oadp.dp.roi_heads as rh
    import oadp.dp.utils as ut
    import oadp_b200.dp as impl
    assert cl.ViLDClassifier is impl.ViLDClassifier and cl.BaseClassifier is impl.BaseClassifier
    assert bh.Shared2FCBlockBBoxHead is impl.Shared2FCBlockBBoxHead and bh.ObjectMixin is impl.ObjectMixin
    assert rh.OADPRoIHead is impl.OADPRoIHead and rh.ViLDEnsembleRoIHead is impl.ViLDEnsembleRoIHead
    assert ds.LoadCLIPFeatures is impl.LoadCLIPFeatures
    assert ut.NormalizedLinear is impl.NormalizedLinear and ut.MultilabelTopKRecall is impl.MultilabelTopKRecall
    assert oadp.base.Globals is impl.Globals and oadp.base.AsymmetricLoss is impl.AsymmetricLoss
    assert oadp.dp.OADPRoIHead is impl.OADPRoIHead
    for name in ('Shared2FCBlockBBoxHead', 'Shared4Conv1FCObjectBBoxHead', 'ViLDEnsembleRoIHead', 'OADPRoIHead'):
        assert name in HEADS.module_dict
    assert {'BaseClassifier', 'Classifier', 'ViLDClassifier'} <= set(LINEAR_LAYERS.module_dict)
    assert 'LoadCLIPFeatures' in PIPELINES.module_dict
    assert {'AsymmetricLoss', 'RKDLoss'} <= set(LossRegistry.module_dict)


def test_bad_out_features_still_raises(workdir):
    with pytest.raises(RuntimeError, match='64'):
        LINEAR_LAYERS.build(dict(type='Classifier', prompts='data/prompts/ml_coco.pth', in_features=256, out_features=64))
