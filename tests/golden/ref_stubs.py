"""Stand-ins for the third-party packages the reference imports, so that the reference's OWN source
files (/root/reference/oadp/...) can be executed in this container -- FIXTURE GENERATION ONLY.

The reference cannot be imported as a package: ``oadp/__init__.py`` pulls ``todd``, ``mmcv``,
``mmdet``, ``lvis`` and ``clip``, none of which is installed and none of which is vendored
(SURVEY.md section 0).  ``make_ref_golden.py`` therefore loads individual reference modules with
``importlib`` after placing the stubs below in ``sys.modules``.  What is stubbed is exactly the
un-vendored third-party surface; every line of arithmetic that lives in the reference repository
itself (block grid, box expansion, masks, hooks, model surgery, classifier) runs unmodified.

Stubbed behaviour that the reference does not show is chosen explicitly and
recorded in DESIGN.md section 8:
  * todd ``BBoxes.indices(min_wh)``: inclusive (w >= 4 and h >= 4).
  * ``clip`` fork ``interpolate_positional_embedding``: bilinear, align_corners=False.
The ``clip.model`` classes restate the published openai/CLIP ``model.py`` (module tree, parameter
names, sequence-first activations) -- the reference's hooks are written against that tree.
"""
from __future__ import annotations

import enum
import logging
import os
import sys
import types
from collections import OrderedDict
from typing import Any, Iterator

import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision.transforms as T

# --------------------------------------------------------------------------------------- todd


class _StoreMeta(type):
    """Class attributes annotated ``bool`` read the environment (todd.StoreMeta)."""

    def __getattr__(cls, name: str) -> Any:
        ann = {}
        for klass in cls.__mro__:
            ann.update(getattr(klass, '__annotations__', {}))
        if name in ann:
            raw = os.environ.get(name, '')
            if ann[name] in (bool, 'bool'):
                return raw not in ('', '0', 'False', 'false')
            return raw
        raise AttributeError(name)


class _Store(metaclass=_StoreMeta):
    CUDA: bool
    CPU: bool
    MPS: bool
    DRY_RUN: bool
    TRAIN_WITH_VAL_DATASET: bool


class _NonInstantiableMeta(type):

    def __call__(cls, *args, **kwargs):
        raise RuntimeError(f'{cls.__name__} is not instantiable')


class _RegistryMeta(type):
    """todd.Registry: the class itself is the registry (``@DatasetRegistry.register()``)."""

    def __init__(cls, name, bases, ns):
        super().__init__(name, bases, ns)
        cls._items = {}

    def register(cls, *names):

        def deco(obj):
            for n in names or (obj.__name__, ):
                cls._items[n] = obj
            return obj

        return deco

    def build(cls, config, default_config=None):
        cfg = dict(default_config or {})
        cfg.update(config)
        return cls._items[cfg.pop('type')](**cfg)


class _Registry(metaclass=_RegistryMeta):
    pass


class _BBoxes:
    """Subset of todd 0.3.0 ``BBoxes`` the reference touches (objects.py:76-186)."""

    def __init__(self, t: torch.Tensor) -> None:
        self._t = t

    def __len__(self) -> int:
        return self._t.shape[0]

    def __getitem__(self, idx) -> '_BBoxes':
        return type(self)(self._t[idx])

    def __iter__(self) -> Iterator[tuple]:
        for row in self.to(_BBoxesXYXY)._t.tolist():  # python floats: PIL.Image.crop needs them
            yield tuple(row)

    def to_tensor(self) -> torch.Tensor:
        return self._t

    def to(self, cls) -> '_BBoxes':
        if cls is type(self):
            return self
        return cls(cls._from_ltrb(self.lt, self.rb))

    @property
    def wh(self) -> torch.Tensor:
        return self.rb - self.lt

    @property
    def area(self) -> torch.Tensor:
        wh = self.wh
        return wh[:, 0] * wh[:, 1]

    @property
    def center(self) -> torch.Tensor:
        return (self.lt + self.rb) / 2

    def indices(self, min_wh=None) -> torch.Tensor:
        wh = self.wh
        return (wh[:, 0] >= min_wh[0]) & (wh[:, 1] >= min_wh[1])

    def __and__(self, other: '_BBoxes') -> torch.Tensor:
        """Pairwise intersection areas (n, m) -- todd ``BBoxes.intersections`` (datasets.py:192-195)."""
        lt = torch.maximum(self.lt[:, None], other.lt[None])
        rb = torch.minimum(self.rb[:, None], other.rb[None])
        wh = (rb - lt).clamp_min(0)
        return wh[..., 0] * wh[..., 1]


class _BBoxesXYXY(_BBoxes):

    @property
    def lt(self) -> torch.Tensor:
        return self._t[:, :2]

    @property
    def rb(self) -> torch.Tensor:
        return self._t[:, 2:]

    @staticmethod
    def _from_ltrb(lt, rb):
        return torch.cat([lt, rb], -1)

    def translate(self, offset: torch.Tensor) -> '_BBoxesXYXY':
        return _BBoxesXYXY(self._t + torch.cat([offset, offset], -1))


class _BBoxesCXCYWH(_BBoxes):

    @property
    def center(self) -> torch.Tensor:
        return self._t[:, :2]

    @property
    def wh(self) -> torch.Tensor:
        return self._t[:, 2:]

    @property
    def lt(self) -> torch.Tensor:
        return self._t[:, :2] - self._t[:, 2:] / 2

    @property
    def rb(self) -> torch.Tensor:
        return self._t[:, :2] + self._t[:, 2:] / 2

    @staticmethod
    def _from_ltrb(lt, rb):
        return torch.cat([(lt + rb) / 2, rb - lt], -1)

    def translate(self, offset: torch.Tensor) -> '_BBoxesCXCYWH':
        return _BBoxesCXCYWH(torch.cat([self._t[:, :2] + offset, self._t[:, 2:]], -1))


class _Config(dict):
    __getattr__ = dict.__getitem__

    @staticmethod
    def load(path):  # never reached by the fixture generator
        raise NotImplementedError


class _Validator:
    """todd.utils.Validator is only subclassed, never run, by the fixture generator."""

    def __init__(self, *args, **kwargs) -> None:
        pass

    def _control_run_iter(self, batch, memo):
        return None

    def _run_iter(self, batch, memo):
        return None


class _Control(enum.Enum):
    CONTINUE = enum.auto()
    BREAK = enum.auto()


# ------------------------------------------------------------------------------- todd.losses / mmcv


class _BaseLoss(nn.Module):
    """todd 0.3.0 ``losses.BaseLoss`` as the reference uses it: ``reduce`` applies the reduction
    (default mean) and a scalar weight (the configs' WarmupScheduler is a scalar at a given step)."""

    def __init__(self, reduction: str = 'mean', weight: float = 1.0, **kwargs) -> None:
        super().__init__()
        self._reduction = reduction
        self._weight = float(weight)

    def reduce(self, loss: torch.Tensor, **kwargs) -> torch.Tensor:
        if self._reduction == 'mean':
            loss = loss.mean()
        elif self._reduction == 'sum':
            loss = loss.sum()
        return self._weight * loss


class _MSELoss(_BaseLoss):

    def forward(self, pred: torch.Tensor, target: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        return self.reduce(F.mse_loss(pred, target, reduction='none'), **kwargs)


class _L1Loss(_BaseLoss):

    def forward(self, pred: torch.Tensor, target: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        return self.reduce(F.l1_loss(pred, target, reduction='none'), **kwargs)


def _force_fp32(apply_to=()):
    """mmcv.runner.force_fp32: named tensor arguments are cast to fp32 (fp16_enabled is False here)."""
    import functools
    import inspect

    def deco(fn):
        names = list(inspect.signature(fn).parameters)

        @functools.wraps(fn)
        def wrapper(*args, **kwargs):
            args = list(args)
            for i, a in enumerate(args):
                if i < len(names) and names[i] in apply_to and torch.is_tensor(a) and a.is_floating_point():
                    args[i] = a.float()
            return fn(*args, **kwargs)

        return wrapper

    return deco


def _make_todd() -> types.ModuleType:
    todd = types.ModuleType('todd')
    todd.Module = nn.Module
    todd.StoreMeta = _StoreMeta
    todd.NonInstantiableMeta = _NonInstantiableMeta
    todd.Store = _Store
    todd.Registry = _Registry
    todd.BBox = tuple
    todd.BBoxes = _BBoxes
    todd.BBoxesXYXY = _BBoxesXYXY
    todd.BBoxesCXCYWH = _BBoxesCXCYWH
    todd.Config = _Config
    todd.logger = logging.getLogger('todd-stub')
    todd.get_local_rank = lambda: int(os.environ.get('LOCAL_RANK', 0))
    todd.utils = types.ModuleType('todd.utils')
    todd.utils.Validator = _Validator
    todd.utils.Memo = dict
    todd.utils.Control = _Control
    todd.base = types.ModuleType('todd.base')
    todd.base.DictAction = object
    todd.losses = types.ModuleType('todd.losses')
    todd.losses.BaseLoss = _BaseLoss
    todd.losses.MSELoss = _MSELoss
    todd.losses.L1Loss = _L1Loss

    class LossRegistry(_Registry):
        pass

    todd.losses.LossRegistry = LossRegistry

    todd.datasets = types.ModuleType('todd.datasets')

    class AccessLayerRegistry:
        """``ALR.build(config, default)``: the fixture generator registers in-memory mappings under
        the ``task_name`` the merged config names (PthAccessLayer contract: Mapping[key] -> value)."""
        stores = {}

        @classmethod
        def build(cls, config, default):
            cfg = dict(default)
            cfg.update(config)
            return cls.stores[cfg['task_name']]

    todd.datasets.AccessLayerRegistry = AccessLayerRegistry
    return todd


# ------------------------------------------------------------------------- mmdet LINEAR_LAYERS


class _MMRegistry:

    def __init__(self) -> None:
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):

        def deco(obj):
            self.module_dict[name or obj.__name__] = obj
            return obj

        return deco if module is None else deco(module)



# ----------------------------------------------------- mmdet heads (oadp/dp/bbox_heads.py, roi_heads.py)
# Restated from mmdet 2.25's published `ConvFCBBoxHead` / `StandardRoIHead`: module names and the call
# conventions the reference's mixins rely on (`self.fc_cls`, `forward -> (cls_score, bbox_pred)`,
# `_bbox_forward -> dict(cls_score, bbox_pred, bbox_feats)`, `HEADS.build(cfg, default_args=...)`).


def _mm_build(registry, cfg, default_args=None):
    args = dict(default_args or {})
    args.update(cfg)
    return registry.module_dict[args.pop('type')](**args)


class _MMBuildRegistry(_MMRegistry):

    def build(self, cfg, default_args=None):
        return _mm_build(self, cfg, default_args)


def _build_linear_layer(cfg, *args, **kwargs):
    cfg = dict(cfg or dict(type='Linear'))
    kind = cfg.pop('type')
    if kind == 'Linear':
        return nn.Linear(*args, **kwargs, **cfg)
    return sys.modules['mmdet.models.utils.builder'].LINEAR_LAYERS.module_dict[kind](*args, **kwargs, **cfg)


class _MMConvModule(nn.Module):

    def __init__(self, cin, cout, norm_cfg):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 3, padding=1, bias=norm_cfg is None)
        if norm_cfg is not None:
            self.bn = nn.BatchNorm2d(cout)
        self._has_norm = norm_cfg is not None

    def forward(self, x):
        x = self.conv(x)
        return F.relu(self.bn(x) if self._has_norm else x)


class _MMBBoxHead(nn.Module):

    def __init__(self, with_avg_pool=False, with_cls=True, with_reg=True, roi_feat_size=7, in_channels=256,
                 num_classes=80, bbox_coder=None, reg_class_agnostic=False, reg_decoded_bbox=False,
                 reg_predictor_cfg=None, cls_predictor_cfg=None, loss_cls=None, loss_bbox=None, init_cfg=None):
        super().__init__()
        self.with_avg_pool, self.with_cls, self.with_reg = with_avg_pool, with_cls, with_reg
        self.roi_feat_area = roi_feat_size * roi_feat_size
        self.in_channels, self.num_classes, self.reg_class_agnostic = in_channels, num_classes, reg_class_agnostic
        self.reg_predictor_cfg = reg_predictor_cfg or dict(type='Linear')
        self.cls_predictor_cfg = cls_predictor_cfg or dict(type='Linear')


class _MMConvFCBBoxHead(_MMBBoxHead):

    def __init__(self, num_shared_convs=0, num_shared_fcs=0, conv_out_channels=256, fc_out_channels=1024,
                 norm_cfg=None, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.shared_convs = nn.ModuleList()
        last = self.in_channels
        for _ in range(num_shared_convs):
            self.shared_convs.append(_MMConvModule(last, conv_out_channels, norm_cfg))
            last = conv_out_channels
        last *= self.roi_feat_area
        self.shared_fcs = nn.ModuleList()
        for _ in range(num_shared_fcs):
            self.shared_fcs.append(nn.Linear(last, fc_out_channels))
            last = fc_out_channels
        if self.with_cls:
            self.fc_cls = _build_linear_layer(self.cls_predictor_cfg, in_features=last, out_features=self.num_classes + 1)
        if self.with_reg:
            self.fc_reg = _build_linear_layer(self.reg_predictor_cfg, in_features=last,
                                              out_features=4 if self.reg_class_agnostic else 4 * self.num_classes)

    def forward(self, x):
        for conv in self.shared_convs:
            x = conv(x)
        x = x.flatten(1)
        for fc in self.shared_fcs:
            x = F.relu(fc(x))
        return (self.fc_cls(x) if self.with_cls else None), (self.fc_reg(x) if self.with_reg else None)


class _MMShared2FCBBoxHead(_MMConvFCBBoxHead):

    def __init__(self, fc_out_channels=1024, *args, **kwargs):
        super().__init__(num_shared_convs=0, num_shared_fcs=2, fc_out_channels=fc_out_channels, *args, **kwargs)


class _MMShared4Conv1FCBBoxHead(_MMConvFCBBoxHead):

    def __init__(self, fc_out_channels=1024, *args, **kwargs):
        super().__init__(num_shared_convs=4, num_shared_fcs=1, fc_out_channels=fc_out_channels, *args, **kwargs)


class _MMBaseRoIExtractor(nn.Module):
    pass


class _FixtureRoIExtractor(_MMBaseRoIExtractor):
    """The fixture hands over RoI features directly: `feats[0]` holds one (C, 7, 7) feature per RoI, picked
    by the row order of `rois` (RoIAlign is mmcv's arithmetic, not the reference's)."""
    num_inputs = 1

    def forward(self, feats, rois):
        assert feats[0].shape[0] == rois.shape[0]
        return feats[0]


class _MMStandardRoIHead(nn.Module):

    def __init__(self, bbox_roi_extractor=None, bbox_head=None, mask_roi_extractor=None, mask_head=None,
                 shared_head=None, train_cfg=None, test_cfg=None, pretrained=None, init_cfg=None):
        super().__init__()
        heads = sys.modules['mmdet.models'].HEADS
        self.bbox_roi_extractor = _mm_build(heads, bbox_roi_extractor)
        self.bbox_head = _mm_build(heads, bbox_head)

    with_shared_head = property(lambda self: False)

    def _bbox_forward(self, x, rois):
        bbox_feats = self.bbox_roi_extractor(x[:self.bbox_roi_extractor.num_inputs], rois)
        cls_score, bbox_pred = self.bbox_head(bbox_feats)
        return dict(cls_score=cls_score, bbox_pred=bbox_pred, bbox_feats=bbox_feats)


def _bbox2roi(bbox_list):
    return torch.cat([torch.cat([b.new_full((b.shape[0], 1), i), b[:, :4]], dim=-1) for i, b in enumerate(bbox_list)], 0)


class _AttrDict(dict):
    """todd.Config / mmcv.ConfigDict as roi_heads.py uses them: `bbox_head.num_classes` get and set."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__

# ----------------------------------------------------------------------- clip (openai layout)


class LayerNorm(nn.LayerNorm):
    """fp32 LayerNorm, cast back (openai/CLIP model.py)."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return super().forward(x.float()).type(x.dtype)


class QuickGELU(nn.Module):

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x * torch.sigmoid(1.702 * x)


class ResidualAttentionBlock(nn.Module):

    def __init__(self, d_model: int, n_head: int, attn_mask=None) -> None:
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(
            OrderedDict([('c_fc', nn.Linear(d_model, d_model * 4)), ('gelu', QuickGELU()),
                         ('c_proj', nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = LayerNorm(d_model)
        self.attn_mask = attn_mask

    def attention(self, x: torch.Tensor) -> torch.Tensor:
        return self.attn(x, x, x, need_weights=False, attn_mask=self.attn_mask)[0]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = x + self.attention(self.ln_1(x))
        x = x + self.mlp(self.ln_2(x))
        return x


class Transformer(nn.Module):

    def __init__(self, width: int, layers: int, heads: int) -> None:
        super().__init__()
        self.width = width
        self.layers = layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads) for _ in range(layers)])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.resblocks(x)


class VisionTransformer(nn.Module):

    def __init__(self, input_resolution: int, patch_size: int, width: int, layers: int, heads: int,
                 output_dim: int) -> None:
        super().__init__()
        self.input_resolution = input_resolution
        self.output_dim = output_dim
        self.patch_size = patch_size  # fork attribute (objects.py:301)
        self.grid = input_resolution // patch_size  # fork attribute (objects.py:281,297)
        self.conv1 = nn.Conv2d(3, width, patch_size, patch_size, bias=False)
        scale = width**-0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn(self.grid**2 + 1, width))
        self.ln_pre = LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))

    def interpolate_positional_embedding(self, size) -> torch.Tensor:
        """Fork method (objects.py:293-295); mode unseen -- bilinear, align_corners=False."""
        pe = self.positional_embedding.detach()
        cls_row, grid = pe[:1], pe[1:]
        d = grid.shape[1]
        grid = grid.reshape(1, self.grid, self.grid, d).permute(0, 3, 1, 2)
        grid = F.interpolate(grid, size=tuple(size), mode='bilinear', align_corners=False)
        grid = grid.permute(0, 2, 3, 1).reshape(size[0] * size[1], d)
        return torch.cat([cls_row, grid])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.conv1(x)
        x = x.reshape(x.shape[0], x.shape[1], -1)
        x = x.permute(0, 2, 1)
        cls_tok = self.class_embedding.to(x.dtype) + torch.zeros(
            x.shape[0], 1, x.shape[-1], dtype=x.dtype, device=x.device)
        x = torch.cat([cls_tok, x], dim=1)
        x = x + self.positional_embedding.to(x.dtype)
        x = self.ln_pre(x)
        x = x.permute(1, 0, 2)
        x = self.transformer(x)
        x = x.permute(1, 0, 2)
        x = self.ln_post(x[:, 0, :])
        if self.proj is not None:
            x = x @ self.proj
        return x


class CLIP(nn.Module):
    """Image half only (the text tower is not on the OAKE path)."""

    def __init__(self, layers: int = 12) -> None:
        super().__init__()
        self.visual = VisionTransformer(224, 32, 768, layers, 12, 512)

    @property
    def dtype(self) -> torch.dtype:
        return self.visual.conv1.weight.dtype

    def encode_image(self, image: torch.Tensor) -> torch.Tensor:
        return self.visual(image.type(self.dtype))


def clip_transform(n_px: int = 224) -> T.Compose:
    """openai/CLIP ``clip.py::_transform``."""
    return T.Compose([
        T.Resize(n_px, interpolation=T.InterpolationMode.BICUBIC),
        T.CenterCrop(n_px),
        lambda image: image.convert('RGB'),
        T.ToTensor(),
        T.Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)),
    ])


_STATE = {'params': None, 'layers': 12}


def set_clip_weights(params: dict) -> None:
    """OpenAI-named visual-tower state dict used by the next ``clip.load_default``."""
    _STATE['params'] = params
    _STATE['layers'] = 1 + max(int(k.split('.')[2]) for k in params if k.startswith('transformer.resblocks.'))


def _load_default(jit: bool = False):
    model = CLIP(_STATE['layers'])
    model.visual.load_state_dict({k: v.clone() for k, v in _STATE['params'].items()}, strict=True)
    return model.eval(), clip_transform(224)


def install() -> None:
    """Places the stub packages in ``sys.modules`` (idempotent)."""
    if 'todd' in sys.modules and getattr(sys.modules['todd'], '_oadp_b200_stub', False):
        return
    todd = _make_todd()
    todd._oadp_b200_stub = True
    sys.modules['todd'] = todd
    sys.modules['todd.utils'] = todd.utils
    sys.modules['todd.base'] = todd.base
    sys.modules['todd.losses'] = todd.losses
    sys.modules['todd.datasets'] = todd.datasets
    clip = types.ModuleType('clip')
    clip.model = types.ModuleType('clip.model')
    for k in (CLIP, VisionTransformer, Transformer, ResidualAttentionBlock, LayerNorm, QuickGELU):
        setattr(clip.model, k.__name__, k)
    clip.load_default = _load_default
    sys.modules['clip'] = clip
    sys.modules['clip.model'] = clip.model
    for name in ('mmdet', 'mmdet.models', 'mmdet.models.utils', 'mmdet.models.utils.builder', 'mmcv', 'mmcv.runner'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['mmdet.models.utils.builder'].LINEAR_LAYERS = _MMRegistry()
    mm = sys.modules['mmdet.models']
    mm.HEADS = _MMBuildRegistry()
    mm.BBoxHead, mm.Shared2FCBBoxHead, mm.Shared4Conv1FCBBoxHead = _MMBBoxHead, _MMShared2FCBBoxHead, _MMShared4Conv1FCBBoxHead
    mm.BaseRoIExtractor, mm.StandardRoIHead = _MMBaseRoIExtractor, _MMStandardRoIHead
    for cls in (_MMShared2FCBBoxHead, _MMShared4Conv1FCBBoxHead, _FixtureRoIExtractor):
        mm.HEADS.module_dict[cls.__name__.replace('_MM', '').lstrip('_')] = cls
    core = types.ModuleType('mmdet.core')
    core.bbox2roi = _bbox2roi
    sys.modules['mmdet.core'] = core
    sys.modules['mmcv'].ConfigDict = _AttrDict
    sys.modules['mmcv.runner'].force_fp32 = _force_fp32
    # mmdet.datasets / lvis: only names that oadp/dp/datasets.py imports at module level
    md = types.ModuleType('mmdet.datasets')
    md.DATASETS, md.PIPELINES = _MMRegistry(), _MMRegistry()
    for cname in ('CocoDataset', 'CustomDataset', 'LVISV1Dataset'):
        setattr(md, cname, type(cname, (), {}))
    sys.modules['mmdet.datasets'] = md
    aw = types.ModuleType('mmdet.datasets.api_wrappers')
    aw.COCO, aw.COCOeval = type('COCO', (), {}), type('COCOeval', (), {})
    sys.modules['mmdet.datasets.api_wrappers'] = aw
    lv = types.ModuleType('lvis')
    lv.LVIS = type('LVIS', (), {})
    sys.modules['lvis'] = lv
