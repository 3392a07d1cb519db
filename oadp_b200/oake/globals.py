"""Global (whole-image) CLIP features -- counterpart of oadp/oake/globals.py.

Per image: CLIP transform of the whole image -> un-modified ViT-B/32 -> L2 normalise -> fp16 (512,)
(globals.py:26-33,49-60).  Images are batched across the tower instead of B=1 (results are
row-independent; the reference's B=1 is launch-bound on any GPU)."""
from __future__ import annotations

from typing import Any, List

from .base import BaseDataset, BaseValidator, Item


class Dataset(BaseDataset):
    pass


class Validator(BaseValidator):
    DATASET = Dataset

    def _submit(self, items: List[Item]):
        return self._pipeline.submit_globals([it.image for it in items])


if __name__ == '__main__':
    Validator.main()
