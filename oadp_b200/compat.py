"""Minimal stand-ins for the third-party plumbing the reference's OAKE entry points lean on
(`todd.Config`, `todd.Store`, `todd.base.DictAction`, torchvision `CocoDetection` + pycocotools).
None of them is installed in this environment and none is on the hot path; only
the surface the reference actually touches is provided:

  * Config: python-file configs with `_base_` inheritance (relative paths), recursive dict merge,
    `_delete_=True`, attribute access, `.override({'.a.b.c': v})`  (oadp/oake/base.py:66-72,117-120;
    configs/oake/*.py).
  * Store: boolean environment flags -- `DRY_RUN` (README.md:184-186), `CUDA` / `CPU` by device
    availability (oadp/oake/base.py:82-88,122).
  * CocoImages: the `ids` / `_load_image` part of torchvision.datasets.CocoDetection that
    BaseDataset uses (oadp/oake/base.py:28-54), read straight from the annotation json.
"""
from __future__ import annotations

import argparse
import ast
import json
import os
import pathlib
from typing import Any, Dict, List, Mapping


class Config(dict):
    """dict with attribute access and the loading rules of todd-style python configs."""

    def __getattr__(self, name: str) -> Any:
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name: str, value: Any) -> None:
        self[name] = value

    @staticmethod
    def _wrap(obj: Any) -> Any:
        if isinstance(obj, Mapping):
            return Config({k: Config._wrap(v) for k, v in obj.items()})
        if isinstance(obj, (list, tuple)):
            return type(obj)(Config._wrap(v) for v in obj)
        return obj

    @staticmethod
    def _merge(base: Dict[str, Any], update: Mapping[str, Any]) -> Dict[str, Any]:
        for k, v in update.items():
            if isinstance(v, Mapping) and v.get('_delete_', False):
                base[k] = {kk: vv for kk, vv in v.items() if kk != '_delete_'}
            elif isinstance(v, Mapping) and isinstance(base.get(k), Mapping):
                base[k] = Config._merge(dict(base[k]), v)
            else:
                base[k] = v
        return base

    @classmethod
    def _load_raw(cls, path: pathlib.Path) -> Dict[str, Any]:
        scope: Dict[str, Any] = {}
        exec(compile(path.read_text(), str(path), 'exec'), scope)  # configs are python files
        own = {k: v for k, v in scope.items() if not k.startswith('__') and not callable(v) and
               not isinstance(v, type(os))}
        bases = own.pop('_base_', [])
        if isinstance(bases, str):
            bases = [bases]
        merged: Dict[str, Any] = {}
        for b in bases:
            merged = cls._merge(merged, cls._load_raw((path.parent / b).resolve()))
        return cls._merge(merged, own)

    @classmethod
    def load(cls, path: str | os.PathLike) -> 'Config':
        return cls._wrap(cls._load_raw(pathlib.Path(path).resolve()))

    def override(self, items: Mapping[str, Any]) -> None:
        """{'.train.dataloader.num_workers': 0} -- dotted paths, leading dot optional."""
        for dotted, value in items.items():
            keys = [k for k in dotted.split('.') if k]
            node: Any = self
            for k in keys[:-1]:
                if k not in node:
                    node[k] = Config()
                node = node[k]
            node[keys[-1]] = Config._wrap(value)


class DictAction(argparse.Action):
    """`--override .a.b:1 .c::text`: `key:value` with a python-literal value, `key::value` for a raw
    string (README.md:216,282 of the reference shows both forms)."""

    def __call__(self, parser, namespace, values, option_string=None):
        out: Dict[str, Any] = getattr(namespace, self.dest, None) or {}
        for item in values if isinstance(values, list) else [values]:
            if '::' in item:
                k, v = item.split('::', 1)
                out[k] = v
                continue
            k, v = item.split(':', 1)
            try:
                out[k] = ast.literal_eval(v)
            except (ValueError, SyntaxError):
                out[k] = v
        setattr(namespace, self.dest, out)


def _env_flag(name: str) -> bool:
    v = os.environ.get(name, '')
    return v not in ('', '0', 'False', 'false')


class _Store:

    @property
    def DRY_RUN(self) -> bool:  # noqa: N802
        return _env_flag('DRY_RUN')

    @property
    def CUDA(self) -> bool:  # noqa: N802
        import torch
        return torch.cuda.is_available() and not _env_flag('CPU')

    @property
    def CPU(self) -> bool:  # noqa: N802
        return not self.CUDA


Store = _Store()


class CocoImages:
    """ids + image loading of a COCO-format annotation file, without pycocotools."""

    def __init__(self, root: str, annFile: str, transform=None, **_: Any) -> None:  # noqa: N803
        self.root = pathlib.Path(root)
        with open(annFile) as f:
            ann = json.load(f)
        self.imgs: Dict[int, Dict[str, Any]] = {im['id']: im for im in ann['images']}
        self.ids: List[int] = sorted(self.imgs)  # CocoDetection: list(sorted(self.coco.imgs.keys()))
        self.transform = transform

    def __len__(self) -> int:
        return len(self.ids)

    def image_path(self, id_: int) -> pathlib.Path:
        return self.root / self.imgs[id_]['file_name']

    def _load_image(self, id_: int):
        import PIL.Image
        return PIL.Image.open(self.image_path(id_)).convert('RGB')
