"""Registries the reference's detector-side classes are registered in.

The reference decorates its classes with mmdet / todd registries:

    LINEAR_LAYERS  BaseClassifier, Classifier, ViLDClassifier                 oadp/dp/classifiers.py:19,71,91
    HEADS          Shared2FCBlockBBoxHead, Shared4Conv1FCObjectBBoxHead        oadp/dp/bbox_heads.py:63-70
                   ViLDEnsembleRoIHead, OADPRoIHead                            oadp/dp/roi_heads.py:20,169
    PIPELINES      LoadCLIPFeatures                                            oadp/dp/datasets.py:137-138
    LossRegistry   AsymmetricLoss, RKDLoss (todd)                              oadp/base/losses.py:10,68

so that the config dicts under configs/dp/ (`type='OADPRoIHead'`, ...) can build them.  When mmdet / todd are
importable the real registries are used and the classes of this package drop into an existing mmdet
installation; neither is installed in this environment, so the same decorator and `build`
surface is provided here: `register_module(name=None, force=False, module=None)` (mmcv) and `register(*names)`
(todd) on one `Registry` class.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Mapping, Optional


class Registry:
    """mmcv.utils.Registry / todd.Registry subset: name -> class, built from `dict(type=..., **kwargs)`."""

    def __init__(self, name: str) -> None:
        self._name = name
        self._module_dict: Dict[str, type] = {}

    name = property(lambda self: self._name)
    module_dict = property(lambda self: self._module_dict)

    def __len__(self) -> int:
        return len(self._module_dict)

    def __contains__(self, key: str) -> bool:
        return key in self._module_dict

    def __repr__(self) -> str:
        return f'Registry(name={self._name}, items={sorted(self._module_dict)})'

    def get(self, key: str) -> Optional[type]:
        return self._module_dict.get(key)

    def _register(self, cls: type, name: Optional[str], force: bool) -> None:
        key = name or cls.__name__
        if not force and key in self._module_dict and self._module_dict[key] is not cls:
            raise KeyError(f'{key} is already registered in {self._name}')
        self._module_dict[key] = cls

    def register_module(self, name: Optional[str] = None, force: bool = False, module: Optional[type] = None):
        """`@REG.register_module()` (decorator) or `REG.register_module(name=..., module=cls)` (call)."""
        if module is not None:
            self._register(module, name, force)
            return module

        def deco(cls: type) -> type:
            self._register(cls, name, force)
            return cls

        return deco

    def register(self, *names: str) -> Callable[[type], type]:
        """todd form: `@LossRegistry.register()`."""

        def deco(cls: type) -> type:
            for n in names or (cls.__name__, ):
                self._register(cls, n, force=False)
            return cls

        return deco

    def build(self, cfg: Mapping[str, Any], default_args: Optional[Mapping[str, Any]] = None, **kwargs: Any) -> Any:
        """`default_args` fills the keys the config does not set (mmcv `build_from_cfg`, todd `default_config`)."""
        if not isinstance(cfg, Mapping) or 'type' not in cfg:
            raise KeyError(f'{self._name}: a config needs a `type` key, got {cfg!r}')
        args = dict(default_args or {})
        args.update(kwargs)
        args.update(cfg)
        kind = args.pop('type')
        cls = kind if isinstance(kind, type) else self._module_dict.get(kind)
        if cls is None:
            raise KeyError(f'{kind} is not in the {self._name} registry')
        return cls(**args)


def _external(module: str, attr: str):
    try:  # pragma: no cover - neither mmdet nor todd is installed in this environment
        import importlib
        return getattr(importlib.import_module(module), attr)
    except Exception:
        return None


HAVE_MMDET = _external('mmdet.models', 'HEADS') is not None

HEADS = _external('mmdet.models', 'HEADS') or Registry('models/heads')
LINEAR_LAYERS = _external('mmdet.models.utils.builder', 'LINEAR_LAYERS') or Registry('linear layers')
PIPELINES = _external('mmdet.datasets', 'PIPELINES') or Registry('pipeline')
ROI_EXTRACTORS = _external('mmdet.models', 'ROI_EXTRACTORS') or HEADS
LossRegistry = _external('todd.losses', 'LossRegistry') or Registry('losses')


def build_linear_layer(cfg: Optional[Mapping[str, Any]], *args: Any, **kwargs: Any):
    """mmdet.models.utils.build_linear_layer: `dict(type='Linear')` by default, otherwise a registered layer;
    positional / keyword arguments (`in_features`, `out_features`) are appended to the config's."""
    import torch.nn as nn
    cfg = dict(cfg) if cfg is not None else dict(type='Linear')
    kind = cfg.pop('type')
    if kind == 'Linear' and 'Linear' not in LINEAR_LAYERS.module_dict:
        return nn.Linear(*args, **kwargs, **cfg)
    cls = LINEAR_LAYERS.get(kind)
    if cls is None:
        raise KeyError(f'Unrecognized linear type {kind}')
    return cls(*args, **kwargs, **cfg)
