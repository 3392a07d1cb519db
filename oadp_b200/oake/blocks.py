"""Block-grid CLIP features -- counterpart of oadp/oake/blocks.py.

Per image: the global crop + every 224x224 block of a /1.5 pyramid (stride <= 112), encoded by the
un-modified ViT-B/32, stored as {'embeddings': f16 (Nb,512), 'bboxes': f16 (Nb,4)}
(blocks.py:40-109,125-135; the first bbox row keeps the reference's (x0,y0,side,side) form)."""
from __future__ import annotations

import pathlib
from typing import Any, List, NamedTuple, Optional

import torch

from .. import frontend
from ..compat import Config
from .base import BaseDataset, BaseValidator, DataLoader, Memo


class Batch(NamedTuple):
    """blocks.py:19-22.  `blocks`: the uint8 HWC image (or `jpeg.JpegSource`) with `bboxes=None` -- the grid,
    the pyramid and the crops are the pipeline's; or, the reference's way, float (Nb,3,224,224) crops with
    their (Nb,4) boxes."""
    output: pathlib.Path
    blocks: Any
    bboxes: Optional[torch.Tensor]


class Dataset(BaseDataset[Batch]):

    def __init__(self, *args, block_size: int = 224, max_stride: int = 112, rescale: float = 1.5, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        if (block_size, max_stride, rescale) != (224, 112, 1.5):
            raise NotImplementedError('only the reference defaults block_size=224, max_stride=112, rescale=1.5')

    def _preprocess(self, id_: int, output: pathlib.Path, image: Any) -> Batch:
        return Batch(output, image, None)

    def cost(self, index: int) -> float:
        info = self.imgs[self.ids[index]]
        if 'width' in info and 'height' in info:
            return 1.0 + len(frontend.blocks_plan(info['width'], info['height']).cells)
        return 27.0


class Validator(BaseValidator[Batch]):

    def _build_dataloader(self, config: Config) -> DataLoader[Batch]:
        dataset = Config(config.dataset)
        dataset.pop('transform', None)
        config.dataset = Dataset(**dataset)
        return super()._build_dataloader(config)

    def _run_iter(self, batch: Batch, memo: Memo) -> torch.Tensor:
        """blocks.py:125-135 for one image."""
        if torch.is_tensor(batch.blocks):
            memo['result'] = dict(embeddings=self._model.embed(batch.blocks).cpu(), bboxes=batch.bboxes.half())
        else:
            memo['result'] = self._pipeline.encode_blocks([batch.blocks])[0]
        return super()._run_iter(batch, memo)

    def _submit(self, batches: List[Batch]):
        return self._pipeline.submit_blocks([b.blocks for b in batches])


if __name__ == '__main__':
    Validator.main()
