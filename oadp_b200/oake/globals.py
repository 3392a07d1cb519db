"""Global (whole-image) CLIP features -- counterpart of oadp/oake/globals.py.

Per image: CLIP transform of the whole image -> un-modified ViT-B/32 -> L2 normalise -> fp16 (512,)
(globals.py:26-33,49-60).  `run()` batches images across the tower instead of B=1 (results are
row-independent; the reference's B=1 is launch-bound on any GPU)."""
from __future__ import annotations

import pathlib
from typing import Any, List, NamedTuple

import torch

from ..compat import Config
from .base import BaseDataset, BaseValidator, DataLoader, Memo


class Batch(NamedTuple):
    """globals.py:19-21.  `image`: the uint8 HWC image (or `jpeg.JpegSource`) -- the CLIP transform runs on
    the GPU; a float (3,224,224) tensor preprocessed the reference's way is accepted by `_run_iter` too."""
    output: pathlib.Path
    image: Any


class Dataset(BaseDataset[Batch]):

    def _preprocess(self, id_: int, output: pathlib.Path, image: Any) -> Batch:
        return Batch(output, image)


class Validator(BaseValidator[Batch]):

    def __init__(self, *args, batch_images: int = 256, **kwargs) -> None:
        # one crop per image: the tower needs hundreds of images per call to fill 148 SMs (8 crops per call run at
        # 6 % of what the same kernels reach at 512; the reference's B = 1 is launch-bound on any GPU)
        super().__init__(*args, batch_images=batch_images, **kwargs)

    def _build_dataloader(self, config: Config) -> DataLoader[Batch]:
        config.pop('transform', None)
        dataset = Config(config.dataset)
        dataset.pop('transform', None)
        config.dataset = Dataset(**dataset)
        return super()._build_dataloader(config)

    def _run_iter(self, batch: Batch, memo: Memo) -> torch.Tensor:
        """globals.py:49-60 for one image."""
        if torch.is_tensor(batch.image):  # already CLIP-preprocessed pixels, as the reference's Dataset yields
            memo['result'] = self._model.embed(batch.image.unsqueeze(0)).squeeze(0).cpu()
        else:
            memo['result'] = self._pipeline.encode_globals([batch.image])[0]
        return super()._run_iter(batch, memo)

    def _submit(self, batches: List[Batch]):
        return self._pipeline.submit_globals([b.image for b in batches])

    def _collate_embeddings(self) -> bool:
        return True


if __name__ == '__main__':
    Validator.main()
