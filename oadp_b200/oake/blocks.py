"""Block-grid CLIP features -- counterpart of oadp/oake/blocks.py.

Per image: the global crop + every 224x224 block of a /1.5 pyramid (stride <= 112), encoded by the
un-modified ViT-B/32, stored as {'embeddings': f16 (Nb,512), 'bboxes': f16 (Nb,4)}
(blocks.py:40-109,125-135; the first bbox row keeps the reference's (x0,y0,side,side) form)."""
from __future__ import annotations

from typing import Any, List

from .. import frontend
from .base import BaseDataset, BaseValidator, Item


class Dataset(BaseDataset):

    def __init__(self, *args, block_size: int = 224, max_stride: int = 112, rescale: float = 1.5, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        if (block_size, max_stride, rescale) != (224, 112, 1.5):
            raise NotImplementedError('only the reference defaults block_size=224, max_stride=112, rescale=1.5')

    def cost(self, index: int) -> float:
        info = self.imgs[self.ids[index]]
        if 'width' in info and 'height' in info:
            return 1.0 + len(frontend.blocks_plan(info['width'], info['height']).cells)
        return 27.0


class Validator(BaseValidator):
    DATASET = Dataset

    def _submit(self, items: List[Item]):
        return self._pipeline.submit_blocks([it.image for it in items])


if __name__ == '__main__':
    Validator.main()
