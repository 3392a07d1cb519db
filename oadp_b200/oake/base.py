"""Shared skeleton of the three OAKE extractors -- the counterpart of oadp/oake/base.py.

Kept from the reference (the drop-in contract):
  * CLI `python|torchrun -m oadp.oake.<task> NAME CONFIG [--override .k.k:v ...]` (base.py:66-72)
  * `BaseDataset`: COCO image ids, output `{output_dir}/{id:012d}.pth`, items already on disk are
    skipped, `auto_fix=True` re-loads them to detect truncated files (base.py:28-54)
  * `BaseValidator.main()`: build the model once, run the `val` split, then `train` (base.py:115-152)
  * `torch.save(result, output)` per image (base.py:106-113)
Changed on purpose (B200-first):
  * the dataset item is the decoded uint8 image (+ proposals); cropping / resizing / normalising
    happen on the GPU (`oadp_b200.pipeline`), not with PIL in DataLoader workers
  * images are grouped into batches of `batch_images` so that the tower always sees SM-filling
    crop counts, and sharded across ranks by crop count (`oadp_b200.dist`) without duplicates
  * files are written by a small thread pool while the GPU works on the next batch
  * `--override .store:packed` writes one packed shard per rank (`oadp_b200.store.PackedStore`,
    SURVEY 8f-1) instead of one pickle per image; resume then skips the keys already in a shard
  * `--override .decode:gpu` hands the compressed JPEG files to the pipeline, which
    decodes them on the GPU bit-identically to Pillow (`oadp_b200.jpeg`); the default `pillow` decodes
    on host threads as the reference's DataLoader workers do
"""
from __future__ import annotations

import argparse
import collections
import concurrent.futures
import itertools
import pathlib
import time
from abc import ABC, abstractmethod
from typing import Any, Dict, Generic, Iterator, List, NamedTuple, Optional, Sequence, Tuple, TypeVar

import numpy as np
import torch

from .. import dist as oake_dist
from .. import jpeg as oake_jpeg
from ..compat import CocoImages, Config, DictAction, Store
from ..model import OakeModel
from ..pipeline import OakePipeline
from ..store import PackedStore, PackedWriter, key_of


class Item(NamedTuple):
    id_: int
    output: pathlib.Path
    image: Any  # uint8 HWC RGB array, or the still-compressed file (`jpeg.JpegSource`) with decode='gpu'
    extra: Any = None


T = TypeVar('T')


class BaseDataset(CocoImages, ABC, Generic[T]):

    def __init__(self, *args, auto_fix: bool = False, output_dir: str, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self._auto_fix = auto_fix
        self.gpu_decode = False  # set by the validator (`decode='gpu'`)
        self._output_dir = pathlib.Path(output_dir)
        self._output_dir.mkdir(parents=True, exist_ok=True)

    def output_path(self, id_: int) -> pathlib.Path:
        return self._output_dir / f'{id_:012d}.pth'

    def use_packed(self) -> None:
        """Resume against the packed shards in the output directory instead of `.pth` files."""
        self._packed_done = set(PackedStore(str(self._output_dir)).keys())

    def is_done(self, id_: int) -> bool:
        if getattr(self, '_packed_done', None) is not None:
            return key_of(id_) in self._packed_done
        output = self.output_path(id_)
        if not output.exists():
            return False
        if not self._auto_fix:
            return True
        try:
            torch.load(output, 'cpu')
            return True
        except Exception:
            print(f'Fixing {output}', flush=True)
            return False

    def __getitem__(self, index: int) -> Optional[Item]:
        id_ = self.ids[index]
        if self.is_done(id_):
            return None
        if self.gpu_decode:
            image = oake_jpeg.load(self.image_path(id_))
        else:
            image = np.asarray(self._load_image(id_), dtype=np.uint8)
        return Item(id_, self.output_path(id_), image, self._extra(id_))

    def _extra(self, id_: int) -> Any:
        return None

    def cost(self, index: int) -> float:
        """Relative amount of GPU work of item `index` (for the balanced shard)."""
        return 1.0


def parse_args(argv: Optional[Sequence[str]] = None) -> argparse.Namespace:
    parser = argparse.ArgumentParser(description='OAKE feature extraction')
    parser.add_argument('name', type=str)
    parser.add_argument('config', type=Config.load)
    parser.add_argument('--override', action=DictAction, nargs='+')
    return parser.parse_args(argv)


def default_params() -> Dict[str, torch.Tensor]:
    """Stand-in for `clip.load_default`: a ViT-B/32 state dict named by $OAKE_CLIP_WEIGHTS
    (OpenAI `visual.*` names, e.g. exported from the official checkpoint).  Without it the call FAILS,
    as `clip.load_default` does when the checkpoint is missing -- features of random weights written
    into a real output directory would be taken for finished work by every later resume.  Seeded
    random weights (no CLIP checkpoint exists offline) are an explicit opt-in for tests, the bench
    and dry runs: `DRY_RUN=True` or `OAKE_ALLOW_RANDOM_WEIGHTS=1`."""
    import os
    path = os.environ.get('OAKE_CLIP_WEIGHTS')
    if path:
        sd = torch.load(path, 'cpu')
        sd = sd.get('state_dict', sd)
        return {k[len('visual.'):] if k.startswith('visual.') else k: v.float() for k, v in sd.items()
                if k.startswith('visual.') or k in ('proj', 'class_embedding', 'positional_embedding') or
                k.startswith(('conv1.', 'ln_pre.', 'ln_post.', 'transformer.'))}
    if not (Store.DRY_RUN or os.environ.get('OAKE_ALLOW_RANDOM_WEIGHTS') == '1'):
        raise RuntimeError('OAKE_CLIP_WEIGHTS is not set: point it at a CLIP ViT-B/32 state dict (visual.* names). '
                           'Set OAKE_ALLOW_RANDOM_WEIGHTS=1 (or DRY_RUN=True) to run on seeded random weights.')
    from .. import synth
    print('OAKE_CLIP_WEIGHTS is not set: using seeded random ViT-B/32 weights (explicit opt-in)', flush=True)
    return synth.visual_params(0)


class BaseValidator(ABC, Generic[T]):
    """One split of one task: iterate images, encode on the GPU, write `.pth` files."""

    DATASET = BaseDataset

    def __init__(self, name: str, model: OakeModel, *, dataloader: Config, log: Optional[Config] = None,
                 batch_images: int = 8, store: str = 'pth', decode: str = 'pillow', **_: Any) -> None:
        self._name = name
        self._model = model
        self._pipeline = OakePipeline(model.engine)
        self._log_interval = int((log or {}).get('interval', 50))
        self._batch_images = 1 if Store.DRY_RUN else int(batch_images)
        self._dataset = self._build_dataset(Config(dataloader.dataset))
        if decode not in ('pillow', 'gpu'):
            raise ValueError(f"decode must be 'pillow' or 'gpu', not {decode!r}")
        self._dataset.gpu_decode = decode == 'gpu'
        if store not in ('pth', 'packed'):
            raise ValueError(f"store must be 'pth' or 'packed', not {store!r}")
        self._packed: Optional[PackedWriter] = None
        workers = int(dataloader.get('num_workers', 2)) or 1
        self._decode_workers = 0 if Store.DRY_RUN else int(dataloader.get('num_workers', 2))
        if store == 'packed':
            self._dataset.use_packed()
            out_dir = self._dataset._output_dir
            rank, _world = oake_dist.rank_world()
            serial = len(list(out_dir.glob(f'shard-{rank:05d}-*.idx.json')))
            self._packed = PackedWriter(str(out_dir), '', f'shard-{rank:05d}-{serial:03d}')
            workers = 1  # appends to one file, in order
        self._writer = concurrent.futures.ThreadPoolExecutor(max_workers=workers)

    def _write(self, result: Any, item: Item) -> None:
        if self._packed is not None:
            self._packed.add(key_of(item.id_), result)
        else:
            torch.save(result, item.output)

    # ------------------------------------------------------------------ to be provided per task
    @classmethod
    def _build_model(cls) -> Tuple[OakeModel, Any]:
        """(model, preprocess) like the reference; preprocess is None: it lives on the GPU."""
        return OakeModel(default_params(), 'cuda'), None

    def _build_dataset(self, config: Config) -> BaseDataset:
        config.pop('transform', None)
        return self.DATASET(**config)

    @abstractmethod
    def _submit(self, items: List[Item]):
        """-> a `Pending` whose result() is one entry per item, in the layout the reference stores."""

    # ---------------------------------------------------------------------------------- the loop
    def _shard(self) -> List[int]:
        rank, world = oake_dist.rank_world()
        todo = list(range(len(self._dataset)))
        if world == 1:
            return todo
        costs = [self._dataset.cost(i) for i in todo]
        return oake_dist.balanced_partition(costs, world)[rank]

    def _items(self, indices: List[int]) -> Iterator[Optional[Item]]:
        """Dataset items in order, decoded ahead of the GPU by a few host threads (PIL releases the
        GIL while it decodes) -- the role of the reference's DataLoader workers (base.py:78-89), minus
        the PIL crop / resize work, which lives on the GPU here.  DRY_RUN: no threads (base.py:82-83)."""
        if self._decode_workers <= 0:
            for i in indices:
                yield self._dataset[i]
            return
        depth = max(2 * self._batch_images, 2 * self._decode_workers)
        with concurrent.futures.ThreadPoolExecutor(max_workers=self._decode_workers) as pool:
            window: 'collections.deque' = collections.deque()
            it = iter(indices)
            for i in itertools.islice(it, depth):
                window.append(pool.submit(self._dataset.__getitem__, i))
            while window:
                item = window.popleft().result()
                nxt = next(it, None)
                if nxt is not None:
                    window.append(pool.submit(self._dataset.__getitem__, nxt))
                yield item

    def _batches(self, indices: List[int]) -> Iterator[List[Item]]:
        batch: List[Item] = []
        for item in self._items(indices):
            if item is None:  # already on disk (base.py:45-47)
                continue
            batch.append(item)
            if len(batch) == self._batch_images:
                yield batch
                batch = []
        if batch:
            yield batch

    def run(self) -> int:
        indices = self._shard()
        if Store.DRY_RUN:
            indices = indices[:3]
        done, t0, pending = 0, time.perf_counter(), []
        in_flight = None  # (batch, Pending): the GPU works on it while the next batch is decoded / staged

        def drain(entry):
            batch_, ticket = entry
            for item, result in zip(batch_, ticket.result()):
                pending.append(self._writer.submit(self._write, result, item))

        try:
            for batch in itertools.chain(self._batches(indices), [None]):
                nxt = (batch, self._submit(batch)) if batch is not None else None
                if in_flight is not None:
                    drain(in_flight)
                in_flight = nxt
                if batch is None:
                    break
                done += len(batch)
                if done % max(self._log_interval, 1) < len(batch):
                    dt = time.perf_counter() - t0
                    print(f'[{self._name}] {done}/{len(indices)} images, {done / dt:.1f} img/s', flush=True)
                still = []
                for f in pending:
                    if f.done():
                        f.result()  # surface write errors
                    else:
                        still.append(f)
                pending = still
            for f in pending:
                f.result()
        finally:
            # whatever was encoded is published even if the run dies: the shard index is written last
            self._writer.shutdown(wait=True)
            if self._packed is not None:
                self._packed.close()
        return done

    @classmethod
    def main(cls, argv: Optional[Sequence[str]] = None) -> None:
        args = parse_args(argv)
        config: Config = args.config
        if args.override:
            config.override(args.override)
        if not Store.CUDA:
            raise RuntimeError('oadp_b200 OAKE needs a CUDA (sm_100a) device; there is no CPU path')
        import os
        local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local_rank % torch.cuda.device_count())
        if int(os.environ.get('WORLD_SIZE', '1')) > 1 and not torch.distributed.is_initialized():
            torch.distributed.init_process_group(backend='nccl',
                                                 device_id=torch.device('cuda', torch.cuda.current_device()))

        model, _ = cls._build_model()
        train = config.pop('train')
        val = config.pop('val')
        for split in (val, train):  # val first, then train (base.py:136-152)
            cls(args.name, model, **split, **config).run()
