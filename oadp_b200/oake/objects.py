"""Per-proposal object CLIP features -- counterpart of oadp/oake/objects.py.

Per image: proposals (N,5) -> min_wh filter -> ADAPTIVE square expansion -> crop (+zero pad) ->
CLIP transform; 14x14 foreground mask per crop; ViT-B/32 with stride-16 patch embedding, resampled
positional table and the mask-attended CLS side stream (objects.py:43-186,198-338).  Stored as
{'embeddings': f16 (No,512), 'bboxes': f16 (No,4), 'objectness': f16 (No,1)}.
"""
from __future__ import annotations

import enum
import pathlib
import pickle
from typing import Any, Dict, List, NamedTuple, Optional, Tuple

import numpy as np
import torch

from ..compat import Config, Store
from ..model import OakeModel
from .base import BaseDataset, BaseValidator, DataLoader, Memo, default_params


class Batch(NamedTuple):
    """objects.py:24-29.  Native form: `objects` = the uint8 HWC image (or `jpeg.JpegSource`), `bboxes` /
    `objectness` = the image's UNFILTERED proposals (N,4) / (N,1), `masks=None` -- the min_wh filter, the
    expansion, the crops and the 14x14 masks are the pipeline's (objects.py:157-186 on the GPU).  The
    reference's form -- float (No,3,224,224) crops, filtered boxes, (No,1,14,14) masks -- is accepted by
    `_run_iter` as well."""
    output: pathlib.Path
    objects: Any
    bboxes: Any
    objectness: Any
    masks: Optional[torch.Tensor]


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
class COCODataset(BaseDataset[Batch]):

    def __init__(self, *args, grid: int, expand_mode: str = 'ADAPTIVE', proposal_file: str,
                 proposal_sorted: bool, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self._grid = grid
        self._expand_mode = ExpandMode[expand_mode]
        if self._expand_mode not in (ExpandMode.ADAPTIVE, ExpandMode.CONSTANT):
            # the reference's LONGEST_EDGE / RECTANGLE branches cannot run (objects.py:90-91,100-101:
            # a (values, indices) tuple as length / an always-true assert)
            raise NotImplementedError(f'expand_mode={expand_mode}: only ADAPTIVE and CONSTANT can run upstream')
        with open(proposal_file, 'rb') as f:
            proposals = pickle.load(f)
        ids = self.ids if proposal_sorted else list(self.imgs.keys())
        self._proposals = {id_: np.asarray(p, dtype=np.float32) for id_, p in zip(ids, proposals)}

    expand_mode = property(lambda self: self._expand_mode.name)

    def _preprocess(self, id_: int, output: pathlib.Path, image: Any) -> Batch:
        proposals = self._proposals[id_].reshape(-1, 5)
        return Batch(output, image, proposals[:, :4], proposals[:, 4:], None)

    def cost(self, index: int) -> float:
        return float(len(self._proposals[self.ids[index]]))


@DatasetRegistry.register()
class LVISDataset(COCODataset):

    def image_path(self, id_: int) -> pathlib.Path:
        url = self.imgs[id_]['coco_url']
        return self.root / url.replace('http://images.cocodataset.org/', '')


class Validator(BaseValidator[Batch]):

    def __init__(self, *args, mini_batch_size: int = 512, **kwargs) -> None:
        # the tower chunks by its own SM-aligned size; results do not depend on the chunking
        self._mini_batch_size = mini_batch_size
        super().__init__(*args, **kwargs)

    @classmethod
    def _build_model(cls, upsample: int = 2) -> Tuple[OakeModel, Any]:
        return OakeModel(default_params(), 'cuda').for_objects(upsample), None

    def _build_dataloader(self, config: Config) -> DataLoader[Batch]:
        """objects.py:275-283: the dataset type comes from the config, the grid from the model."""
        dataset = Config(config.dataset)
        dataset.pop('transform', None)
        config.dataset = DatasetRegistry.build(dataset, default_config=dict(grid=self._model.visual.grid))
        return super()._build_dataloader(config)

    @staticmethod
    def _proposals(batch: Batch) -> np.ndarray:
        return np.concatenate([np.asarray(batch.bboxes, dtype=np.float32).reshape(-1, 4),
                               np.asarray(batch.objectness, dtype=np.float32).reshape(-1, 1)], axis=1)

    def _run_iter(self, batch: Batch, memo: Memo) -> torch.Tensor:
        """objects.py:316-338 for one image."""
        if torch.is_tensor(batch.objects):  # the reference's batch: crops, filtered boxes, masks
            emb = self._model.embed(batch.objects, batch.masks)
            memo['result'] = dict(embeddings=emb.cpu(), bboxes=batch.bboxes.half(), objectness=batch.objectness.half())
        else:
            memo['result'] = self._pipeline.encode_objects([batch.objects], [self._proposals(batch)], Store.DRY_RUN,
                                                           self._dataset.expand_mode)[0]
        return super()._run_iter(batch, memo)

    def _submit(self, batches: List[Batch]):
        return self._pipeline.submit_objects([b.objects for b in batches], [self._proposals(b) for b in batches],
                                             dry_run=Store.DRY_RUN, expand_mode=self._dataset.expand_mode)


if __name__ == '__main__':
    Validator.main()
