"""Per-proposal object CLIP features -- counterpart of oadp/oake/objects.py.

Per image: proposals (N,5) -> min_wh filter -> ADAPTIVE square expansion -> crop (+zero pad) ->
CLIP transform; 14x14 foreground mask per crop; ViT-B/32 with stride-16 patch embedding, resampled
positional table and the mask-attended CLS side stream (objects.py:43-186,198-338).  Stored as
{'embeddings': f16 (No,512), 'bboxes': f16 (No,4), 'objectness': f16 (No,1)}.
"""
from __future__ import annotations

import enum
import os
import pathlib
import pickle
from typing import Any, Dict, List, Tuple

import numpy as np

from ..compat import Config, Store
from ..model import OakeModel
from .base import BaseDataset, BaseValidator, Item, default_params


class ExpandMode(enum.Enum):
    RECTANGLE = enum.auto()
    LONGEST_EDGE = enum.auto()
    CONSTANT = enum.auto()
    ADAPTIVE = enum.auto()


class DatasetRegistry:
    """`type=` lookup used by configs/oake/objects_*.py (objects.py:39-44,189-190)."""
    _registry: Dict[str, type] = {}

    @classmethod
    def register(cls):
        def deco(klass: type) -> type:
            cls._registry[klass.__name__] = klass
            return klass
        return deco

    @classmethod
    def build(cls, config: Config, default_config: Dict[str, Any] | None = None) -> Any:
        cfg = dict(default_config or {})
        cfg.update(config)
        return cls._registry[cfg.pop('type')](**cfg)


@DatasetRegistry.register()
class COCODataset(BaseDataset):

    def __init__(self, *args, grid: int, expand_mode: str = 'ADAPTIVE', proposal_file: str,
                 proposal_sorted: bool, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self._grid = grid
        self._expand_mode = ExpandMode[expand_mode]
        if self._expand_mode is not ExpandMode.ADAPTIVE:
            # the reference's LONGEST_EDGE / RECTANGLE branches cannot run and no
            # shipped config selects CONSTANT
            raise NotImplementedError(f'expand_mode={expand_mode}: only ADAPTIVE is built')
        with open(proposal_file, 'rb') as f:
            proposals = pickle.load(f)
        ids = self.ids if proposal_sorted else list(self.imgs.keys())
        self._proposals = {id_: np.asarray(p, dtype=np.float32) for id_, p in zip(ids, proposals)}

    def _extra(self, id_: int) -> np.ndarray:
        return self._proposals[id_]

    def cost(self, index: int) -> float:
        return float(len(self._proposals[self.ids[index]]))


@DatasetRegistry.register()
class LVISDataset(COCODataset):

    def image_path(self, id_: int) -> pathlib.Path:
        url = self.imgs[id_]['coco_url']
        return self.root / url.replace('http://images.cocodataset.org/', '')


class Validator(BaseValidator):

    def __init__(self, *args, mini_batch_size: int = 512, **kwargs) -> None:
        # the tower chunks by its own SM-aligned size; results do not depend on the chunking
        self._mini_batch_size = mini_batch_size
        super().__init__(*args, **kwargs)

    @classmethod
    def _build_model(cls, upsample: int = 2) -> Tuple[OakeModel, Any]:
        return OakeModel(default_params(), 'cuda').for_objects(upsample), None

    def _build_dataset(self, config: Config) -> BaseDataset:
        config.pop('transform', None)
        return DatasetRegistry.build(config, default_config=dict(grid=self._model.visual.grid))

    def _submit(self, items: List[Item]):
        return self._pipeline.submit_objects([it.image for it in items], [it.extra for it in items],
                                             dry_run=Store.DRY_RUN)


if __name__ == '__main__':
    Validator.main()
