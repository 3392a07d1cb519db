"""Stand-ins for the mmdet 2.25 classes the reference's heads derive from -- used ONLY when mmdet is absent.

The reference's registry surface is a set of thin subclasses of mmdet classes:

    Shared2FCBlockBBoxHead(BlockMixin, Shared2FCBBoxHead)                 oadp/dp/bbox_heads.py:63-65
    Shared4Conv1FCObjectBBoxHead(ObjectMixin, Shared4Conv1FCBBoxHead)     oadp/dp/bbox_heads.py:68-70
    ViLDEnsembleRoIHead(StandardRoIHead), OADPRoIHead                     oadp/dp/roi_heads.py:20,169

With mmdet installed, `oadp_b200.dp` derives from the real classes and nothing here is imported.  mmdet is
not installed in this environment and the detector is out of scope, so this
module restates just enough of mmdet's PUBLISHED head structure for the registry names to build from the
reference's config dicts and for the path from RoI features to `fc_cls` -- the call site of the cosine
classifier -- to run: module names (`shared_convs.i.conv`, `shared_convs.i.bn`, `shared_fcs.i`, `fc_cls`,
`fc_reg`), constructor arguments and the `(cls_score, bbox_pred)` / `dict(cls_score, bbox_pred, bbox_feats)`
return conventions follow mmdet 2.25 `ConvFCBBoxHead` / `StandardRoIHead`, so a state dict moves between the
two.  Box coding, losses, samplers, NMS and the mask branch stay mmdet's: they are accepted as config and kept,
not built.  These layers are plain PyTorch modules (the conv / fc stack is mmdet's work, not the hot path).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from .registry import HEADS, ROI_EXTRACTORS, build_linear_layer


def _norm(cfg: Optional[Dict[str, Any]], channels: int):
    """-> (attribute name, module) like mmcv `build_norm_layer`; SyncBN runs as BatchNorm2d in one process
    (the reference does the same on CPU, oadp/__init__.py:11-18)."""
    if cfg is None:
        return None, None
    kind = cfg['type']
    if kind in ('BN', 'BN2d', 'SyncBN'):
        layer: nn.Module = nn.BatchNorm2d(channels)
        name = 'bn'
    elif kind == 'GN':
        layer = nn.GroupNorm(cfg['num_groups'], channels)
        name = 'gn'
    else:
        raise KeyError(f'norm type {kind} is not built by the mmdet stand-in')
    for p in layer.parameters():
        p.requires_grad_(cfg.get('requires_grad', True))
    return name, layer


class ConvModule(nn.Module):
    """mmcv ConvModule(conv 3x3 -> norm -> ReLU) with its attribute names (`conv`, `bn` / `gn`)."""

    def __init__(self, in_channels: int, out_channels: int, norm_cfg: Optional[Dict[str, Any]] = None) -> None:
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, 3, padding=1, bias=norm_cfg is None)
        self.norm_name, norm = _norm(norm_cfg, out_channels)
        if norm is not None:
            self.add_module(self.norm_name, norm)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.conv(x)
        if self.norm_name is not None:
            x = getattr(self, self.norm_name)(x)
        return torch.relu(x)


class BBoxHead(nn.Module):
    """Constructor surface of mmdet `BBoxHead`; `fc_cls` is built through `build_linear_layer`, which is how
    `cls_predictor_cfg=dict(type='ViLDClassifier', prompts=...)` reaches the cosine classifier."""

    def __init__(self, with_avg_pool: bool = False, with_cls: bool = True, with_reg: bool = True, roi_feat_size: int = 7,
                 in_channels: int = 256, num_classes: int = 80, bbox_coder: Optional[Dict[str, Any]] = None,
                 reg_class_agnostic: bool = False, reg_decoded_bbox: bool = False,
                 reg_predictor_cfg: Optional[Dict[str, Any]] = None, cls_predictor_cfg: Optional[Dict[str, Any]] = None,
                 loss_cls: Optional[Dict[str, Any]] = None, loss_bbox: Optional[Dict[str, Any]] = None,
                 init_cfg: Any = None, **unused: Any) -> None:
        super().__init__()
        if unused:
            raise TypeError(f'unexpected arguments {sorted(unused)}')
        assert with_cls or with_reg
        self.with_avg_pool, self.with_cls, self.with_reg = with_avg_pool, with_cls, with_reg
        self.roi_feat_size = (roi_feat_size, roi_feat_size) if isinstance(roi_feat_size, int) else tuple(roi_feat_size)
        self.roi_feat_area = self.roi_feat_size[0] * self.roi_feat_size[1]
        self.in_channels, self.num_classes = in_channels, num_classes
        self.reg_class_agnostic, self.reg_decoded_bbox = reg_class_agnostic, reg_decoded_bbox
        self.reg_predictor_cfg = reg_predictor_cfg or dict(type='Linear')
        self.cls_predictor_cfg = cls_predictor_cfg or dict(type='Linear')
        self.bbox_coder_cfg, self.loss_cls_cfg, self.loss_bbox_cfg = bbox_coder, loss_cls, loss_bbox  # mmdet's, unbuilt

    @property
    def cls_channels(self) -> int:
        sigmoid = bool((self.loss_cls_cfg or {}).get('use_sigmoid', False))
        return self.num_classes if sigmoid else self.num_classes + 1  # mmdet: softmax CE carries a background column


class ConvFCBBoxHead(BBoxHead):
    """shared convs -> flatten -> shared fcs -> (cls convs/fcs -> fc_cls, reg convs/fcs -> fc_reg)."""

    def __init__(self, num_shared_convs: int = 0, num_shared_fcs: int = 0, num_cls_convs: int = 0, num_cls_fcs: int = 0,
                 num_reg_convs: int = 0, num_reg_fcs: int = 0, conv_out_channels: int = 256, fc_out_channels: int = 1024,
                 conv_cfg: Any = None, norm_cfg: Optional[Dict[str, Any]] = None, init_cfg: Any = None, *args: Any,
                 **kwargs: Any) -> None:
        super().__init__(*args, init_cfg=init_cfg, **kwargs)
        if num_cls_convs or num_cls_fcs or num_reg_convs or num_reg_fcs:
            raise NotImplementedError('the mmdet stand-in builds shared convs / fcs only (what the reference configures)')
        self.num_shared_convs, self.num_shared_fcs = num_shared_convs, num_shared_fcs
        self.conv_out_channels, self.fc_out_channels = conv_out_channels, fc_out_channels
        self.shared_convs = nn.ModuleList()
        last = self.in_channels
        for _ in range(num_shared_convs):
            self.shared_convs.append(ConvModule(last, conv_out_channels, norm_cfg))
            last = conv_out_channels
        self.shared_fcs = nn.ModuleList()
        if num_shared_fcs > 0:
            if not self.with_avg_pool:
                last *= self.roi_feat_area
            for _ in range(num_shared_fcs):
                self.shared_fcs.append(nn.Linear(last, fc_out_channels))
                last = fc_out_channels
        elif not self.with_avg_pool:
            last *= self.roi_feat_area
        self.shared_out_channels = self.cls_last_dim = self.reg_last_dim = last
        if self.with_cls:
            self.fc_cls = build_linear_layer(self.cls_predictor_cfg, in_features=last, out_features=self.cls_channels)
        if self.with_reg:
            out_reg = 4 if self.reg_class_agnostic else 4 * self.num_classes
            self.fc_reg = build_linear_layer(self.reg_predictor_cfg, in_features=last, out_features=out_reg)

    def forward(self, x: torch.Tensor):
        for conv in self.shared_convs:
            x = conv(x)
        if self.num_shared_fcs > 0:
            if self.with_avg_pool:
                x = x.mean((2, 3))
            x = x.flatten(1)
            for fc in self.shared_fcs:
                x = torch.relu(fc(x))
        elif x.dim() > 2:
            x = x.mean((2, 3)) if self.with_avg_pool else x.flatten(1)
        cls_score = self.fc_cls(x) if self.with_cls else None
        bbox_pred = self.fc_reg(x) if self.with_reg else None
        return cls_score, bbox_pred


@HEADS.register_module()
class Shared2FCBBoxHead(ConvFCBBoxHead):

    def __init__(self, fc_out_channels: int = 1024, *args: Any, **kwargs: Any) -> None:
        super().__init__(num_shared_convs=0, num_shared_fcs=2, fc_out_channels=fc_out_channels, *args, **kwargs)


@HEADS.register_module()
class Shared4Conv1FCBBoxHead(ConvFCBBoxHead):

    def __init__(self, fc_out_channels: int = 1024, *args: Any, **kwargs: Any) -> None:
        super().__init__(num_shared_convs=4, num_shared_fcs=1, fc_out_channels=fc_out_channels, *args, **kwargs)


class BaseRoIExtractor(nn.Module):
    num_inputs: int


@ROI_EXTRACTORS.register_module()
class SingleRoIExtractor(BaseRoIExtractor):
    """mmdet SingleRoIExtractor: each RoI is pooled from ONE pyramid level chosen by its scale
    (`floor(log2(sqrt(w h) / finest_scale + 1e-6))`), RoIAlign (aligned, adaptive sampling) from torchvision."""

    def __init__(self, roi_layer: Dict[str, Any], out_channels: int, featmap_strides: Sequence[int],
                 finest_scale: int = 56, init_cfg: Any = None) -> None:
        super().__init__()
        if roi_layer.get('type') != 'RoIAlign':
            raise KeyError('the mmdet stand-in pools with RoIAlign only')
        size = roi_layer['output_size']
        self.output_size = (size, size) if isinstance(size, int) else tuple(size)
        self.sampling_ratio = int(roi_layer.get('sampling_ratio', 0))
        self.out_channels, self.featmap_strides, self.finest_scale = out_channels, list(featmap_strides), finest_scale

    @property
    def num_inputs(self) -> int:
        return len(self.featmap_strides)

    def map_roi_levels(self, rois: torch.Tensor, num_levels: int) -> torch.Tensor:
        scale = torch.sqrt((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2]))
        return torch.floor(torch.log2(scale / self.finest_scale + 1e-6)).clamp(0, num_levels - 1).long()

    def forward(self, feats: Sequence[torch.Tensor], rois: torch.Tensor) -> torch.Tensor:
        from torchvision.ops import roi_align
        out = feats[0].new_zeros(rois.shape[0], self.out_channels, *self.output_size)
        if rois.shape[0] == 0:
            return out
        levels = self.map_roi_levels(rois, len(feats)) if len(feats) > 1 else rois.new_zeros(rois.shape[0], dtype=torch.long)
        for i, f in enumerate(feats):
            idx = (levels == i).nonzero().flatten()
            if idx.numel():
                out[idx] = roi_align(f, rois[idx], self.output_size, 1.0 / self.featmap_strides[i], self.sampling_ratio,
                                     aligned=True)
        return out


def bbox2roi(bbox_list: List[torch.Tensor]) -> torch.Tensor:
    """mmdet.core.bbox2roi: per-image (n, 4+) boxes -> (sum n, 5) rows [batch index, x1, y1, x2, y2]."""
    rois = []
    for i, b in enumerate(bbox_list):
        if b.shape[0]:
            rois.append(torch.cat([b.new_full((b.shape[0], 1), i), b[:, :4]], dim=-1))
        else:
            rois.append(b.new_zeros((0, 5)))
    return torch.cat(rois, 0)


class StandardRoIHead(nn.Module):
    """Constructor and `_bbox_forward` of mmdet StandardRoIHead (bbox branch only; training, sampling, the mask
    branch and `simple_test_bboxes` are mmdet's)."""

    def __init__(self, bbox_roi_extractor: Optional[Dict[str, Any]] = None, bbox_head: Optional[Dict[str, Any]] = None,
                 mask_roi_extractor: Optional[Dict[str, Any]] = None, mask_head: Optional[Dict[str, Any]] = None,
                 shared_head: Optional[Dict[str, Any]] = None, train_cfg: Any = None, test_cfg: Any = None,
                 pretrained: Any = None, init_cfg: Any = None) -> None:
        super().__init__()
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        if shared_head is not None:
            raise NotImplementedError('shared_head is not built by the mmdet stand-in')
        if bbox_head is not None:
            self.bbox_roi_extractor = ROI_EXTRACTORS.build(bbox_roi_extractor)
            self.bbox_head = HEADS.build(bbox_head)
        self.mask_head_cfg, self.mask_roi_extractor_cfg = mask_head, mask_roi_extractor  # mmdet's, unbuilt

    with_bbox = property(lambda self: hasattr(self, 'bbox_head'))
    with_shared_head = property(lambda self: False)

    def _bbox_forward(self, x: Sequence[torch.Tensor], rois: torch.Tensor) -> Dict[str, torch.Tensor]:
        bbox_feats = self.bbox_roi_extractor(x[:self.bbox_roi_extractor.num_inputs], rois)
        cls_score, bbox_pred = self.bbox_head(bbox_feats)
        return dict(cls_score=cls_score, bbox_pred=bbox_pred, bbox_feats=bbox_feats)
