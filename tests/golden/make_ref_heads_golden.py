"""Generates tests/golden/ref_heads_golden.pt by EXECUTING the reference's oadp/dp/bbox_heads.py and
oadp/dp/roi_heads.py (on top of its classifiers.py / utils.py / base/losses.py / base/globals_.py).

    python tests/golden/make_ref_heads_golden.py          # needs /root/reference (build container only)

Same method as make_ref_golden.py: the reference's own source files are loaded with importlib over the
stand-ins of ref_stubs.py for the un-vendored third parties (here: mmdet's `ConvFCBBoxHead` family,
`StandardRoIHead`, `bbox2roi`, `HEADS`; todd's `LossRegistry`).  Every line that lives in the reference runs
unmodified: `NotWithRegMixin`, `BlockMixin.loss`, `ObjectMixin.__init__/forward` (bbox_heads.py:20-60),
`ViLDEnsembleRoIHead.__init__/_bbox_forward/object_forward_train` (roi_heads.py:20-129),
`OADPRoIHead.__init__/block_forward_train` (roi_heads.py:169-209), the classifiers they build.

Recorded: the head's state dict (mmdet module names), and for seeded RoI features -- inference `cls_score`
(the ViLD ensemble, roi_heads.py:93-112) with the two logit matrices it was computed from; training-mode
`cls_score`; the hooked `_object_head.fc_cls._linear` output of `object_forward_train`; the block head's
`loss_block` / `recall_block` and hooked `_linear` output.  tests/test_gpu_heads.py loads the state dict into the
product's heads and compares; tests/test_ref_golden.py pins oracle/classifier.py::vild_ensemble to it.
"""
from __future__ import annotations

import pathlib
import sys

import torch

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))

import make_ref_golden as base  # noqa: E402

N_ROIS, CHANNELS, FC = 37, 8, 128  # FC: the classifier kernels' backward wants in_features % 128 == 0


def head_config(prompts_path: str) -> dict:
    """configs/dp/models/{vild_ensemble_faster_rcnn_r50_fpn,block}.py + oadp_ov_coco.py, at toy widths."""
    return dict(
        bbox_roi_extractor=dict(type='FixtureRoIExtractor'),
        bbox_head=dict(type='Shared4Conv1FCBBoxHead', in_channels=CHANNELS, conv_out_channels=CHANNELS,
                       fc_out_channels=FC, roi_feat_size=7, reg_class_agnostic=True,
                       norm_cfg=dict(type='SyncBN', requires_grad=True), num_classes=None,
                       cls_predictor_cfg=dict(type='ViLDClassifier', prompts=prompts_path,
                                              scaler=dict(train=0.01, val=0.007))),
        object_head=dict(type='Shared4Conv1FCObjectBBoxHead',
                         cls_predictor_cfg=dict(type='Classifier', prompts=prompts_path)),
        block_head=dict(type='Shared2FCBlockBBoxHead', topk=2,
                        loss=dict(type='AsymmetricLoss', weight=16.0, gamma_neg=4, gamma_pos=0),
                        cls_predictor_cfg=dict(type='Classifier', prompts=prompts_path)),
    )


def head_inputs():
    g = torch.Generator().manual_seed(2024)
    feats = torch.randn(N_ROIS, CHANNELS, 7, 7, generator=g)
    rois = torch.cat([torch.zeros(N_ROIS, 1), torch.rand(N_ROIS, 4, generator=g) * 100], 1)
    block_feats = torch.randn(11, CHANNELS, 7, 7, generator=g)
    block_boxes = [torch.rand(6, 4, generator=g) * 100, torch.rand(5, 4, generator=g) * 100]
    block_targets = [torch.rand(6, 6, generator=g) > 0.7, torch.rand(5, 6, generator=g) > 0.7]
    return feats, rois, block_feats, block_boxes, block_targets


def randomize(module: torch.nn.Module, seed: int) -> None:
    """Seeded values for every parameter and BatchNorm statistic (defaults would leave the norms trivial)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(module.named_parameters()):
            if name.endswith('_bg_embedding') or '_bg_embedding' in name:
                p.copy_(torch.randn(p.shape, generator=g))
            elif p.dim() > 1:
                p.copy_(torch.randn(p.shape, generator=g) * (2.0 / p[0].numel())**0.5)
            else:
                p.copy_(torch.randn(p.shape, generator=g) * 0.1 + (1.0 if name.endswith('bn.weight') else 0.0))
        for name, b in sorted(module.named_buffers()):
            if name.endswith('running_mean'):
                b.copy_(torch.randn(b.shape, generator=g) * 0.1)
            elif name.endswith('running_var'):
                b.copy_(torch.rand(b.shape, generator=g) + 0.5)


def main() -> None:
    m = base._load_reference_modules(with_heads=True)
    stubs = m['ref_stubs']
    prompts, bases, novels, *_ = base.classifier_inputs()
    G = m['globals_']
    G.Globals.categories = G.Categories(bases=bases, novels=novels)
    ppath = HERE / '_ref_prompts.tmp.pth'
    torch.save(prompts, ppath)
    out = dict(reference_files=['oadp/dp/bbox_heads.py', 'oadp/dp/roi_heads.py', 'oadp/dp/classifiers.py',
                                'oadp/dp/utils.py', 'oadp/base/losses.py', 'oadp/base/globals_.py', 'oadp/oake/objects.py'])
    try:
        cfg = head_config(str(ppath))
        cfg = {k: stubs._AttrDict(v) if isinstance(v, dict) else v for k, v in cfg.items()}
        head = m['roi_heads'].OADPRoIHead(**cfg)
    finally:
        ppath.unlink()
    head.eval()
    randomize(head, 7)
    assert head._object_head.fc_cls._bg_embedding.requires_grad is False
    out['state_dict'] = {k: v.clone() for k, v in head.state_dict().items()}
    out['lambda'] = head.lambda_.clone()
    feats, rois, block_feats, block_boxes, block_targets = head_inputs()

    hooked = {}
    h1 = head._object_head.fc_cls._linear.register_forward_hook(lambda mod, i, o: hooked.__setitem__('objects', o.detach().clone()))
    h2 = head._block_head.fc_cls._linear.register_forward_hook(lambda mod, i, o: hooked.__setitem__('blocks', o.detach().clone()))
    with torch.no_grad():
        G.Globals.training = False
        out['eval_bbox_logits'] = head.bbox_head(feats)[0].clone()
        out['eval_bbox_pred'] = head.bbox_head(feats)[1].clone()
        out['eval_object_logits'] = head._object_head(feats)[0].clone()
        out['eval_cls_score'] = head._bbox_forward([feats], rois)['cls_score'].clone()
        G.Globals.training = True
        out['train_cls_score'] = head._bbox_forward([feats], rois)['cls_score'].clone()
        head.object_forward_train([feats], [rois[:20, 1:], rois[20:, 1:]])
        out['train_object_hooked'] = hooked['objects']
        out['train_object_logits'] = head._object_head(feats)[0].clone()
    # block branch with gradients: loss value and d loss / d block features
    bf = block_feats.clone().requires_grad_(True)
    losses = head.block_forward_train([bf], block_boxes, block_targets)
    losses['loss_block'].backward()
    out['block_loss'] = losses['loss_block'].detach().clone()
    out['block_recall'] = losses['recall_block'].detach().clone()
    out['block_hooked'] = hooked['blocks']
    out['block_feats_grad'] = bf.grad.clone()
    out['block_fc_cls_weight_grad'] = head._block_head.fc_cls._linear.weight.grad.clone()
    h1.remove()
    h2.remove()
    G.Globals.training = False

    # ---- objects.py: ExpandMode.CONSTANT (objects.py:92-93), expansion + preprocessing by the reference itself
    import PIL.Image
    from torchvision.datasets.vision import StandardTransform
    ods = m['objects'].COCODataset.__new__(m['objects'].COCODataset)
    ods._grid = 14
    ods._expand_mode = m['objects'].ExpandMode['CONSTANT']
    ods.transforms = StandardTransform(stubs.clip_transform(224), None)
    images, proposals = base.ref_inputs()
    out['expand_constant'] = []
    for arr, prop in zip(images[:2], proposals[:2]):
        pil = PIL.Image.fromarray(arr)
        ods._proposals = {1: torch.tensor(prop, dtype=torch.float32)}
        ob = ods._preprocess(1, pathlib.Path('x.pth'), pil)
        expanded = ods._expand(stubs._BBoxesXYXY(ob.bboxes), torch.tensor(pil.size)).to_tensor()
        out['expand_constant'].append(dict(bboxes=ob.bboxes, expanded=expanded, masks=ob.masks.to(torch.uint8),
                                           pixels=base.checksums(ob.objects)))
    torch.save(out, HERE / 'ref_heads_golden.pt')
    print(f'wrote ref_heads_golden.pt ({(HERE / "ref_heads_golden.pt").stat().st_size / 1024:.0f} KiB); '
          f'eval cls_score {tuple(out["eval_cls_score"].shape)}, block loss {float(out["block_loss"]):.6f}, '
          f'recall {float(out["block_recall"]):.3f}')


if __name__ == '__main__':
    main()
