"""Shared skeleton of the three OAKE extractors -- the counterpart of oadp/oake/base.py.

Kept from the reference (the drop-in contract):
  * CLI `python|torchrun -m oadp.oake.<task> NAME CONFIG [--override .k.k:v ...]` (base.py:66-72)
  * `BaseDataset`: COCO image ids, output `{output_dir}/{id:012d}.pth`, items already on disk are
    skipped, `auto_fix=True` re-loads them to detect truncated files (base.py:28-54)
  * `BaseValidator.main()`: build the model once, run the `val` split, then `train` (base.py:115-152)
  * `torch.save(result, output)` per image (base.py:106-113)
  * the validator class surface: `_build_model() -> (model, preprocess)`,
    `_build_dataloader(config)`, `_control_run_iter(batch, memo)`, `_run_iter(batch, memo) -> Tensor`,
    per-task `Batch` NamedTuples with the reference's field names, `Dataset._preprocess(id_, output, image)`
    (base.py:56-63,78-113).  `_run_iter` encodes ONE batch synchronously like the reference's; `run()` drives
    the same datasets through the asynchronous, multi-image form of the same call (`_submit`)
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
import enum
import itertools
import json
import pathlib
import time
from abc import ABC, abstractmethod
from typing import Any, Dict, Generic, Iterable, Iterator, List, Optional, Protocol, Sequence, Tuple, TypeVar

import numpy as np
import torch

from .. import dist as oake_dist
from .. import jpeg as oake_jpeg
from ..compat import CocoImages, Config, DictAction, Store
from ..model import OakeModel
from ..pipeline import OakePipeline
from ..store import PackedStore, PackedWriter, key_of


class Control(enum.Enum):
    """todd.utils.Control: what `_control_run_iter` may answer."""
    BREAK = enum.auto()
    CONTINUE = enum.auto()


Memo = dict  # todd.utils.Memo


class Batch(Protocol):
    """What a dataset item must offer the loop (base.py:18-23); the tasks define NamedTuples with the
    reference's field names (globals.py:19-21, blocks.py:19-22, objects.py:24-29)."""

    @property
    def output(self) -> pathlib.Path:
        ...


def key_of_batch(batch: Batch) -> str:
    return batch.output.stem  # f'{image_id:012d}'


T = TypeVar('T', bound=Batch)


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

    def exists(self, id_: int) -> bool:
        """Cheap form of `is_done` (no `auto_fix` re-load): which items the shard leaves out."""
        if getattr(self, '_packed_done', None) is not None:
            return key_of(id_) in self._packed_done
        return self.output_path(id_).exists()

    def __getitem__(self, index: int) -> Optional[T]:
        id_ = self.ids[index]
        if self.is_done(id_):
            return None
        if self.gpu_decode:
            image = oake_jpeg.load(self.image_path(id_))
        else:
            image = np.asarray(self._load_image(id_), dtype=np.uint8)
        return self._preprocess(id_, self.output_path(id_), image)

    @abstractmethod
    def _preprocess(self, id_: int, output: pathlib.Path, image: Any) -> T:
        """base.py:56-63.  `image` is the decoded uint8 HWC array (or the still-compressed
        `jpeg.JpegSource` with decode='gpu'), not a PIL image: cropping / resizing / normalising are the
        pipeline's kernels, so the batch carries the image itself in the field the reference fills with
        preprocessed tensors."""

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


class DataLoader(Generic[T]):
    """What `_build_dataloader` returns -- the role of `torch.utils.data.DataLoader(batch_size=None, sampler=
    DistributedSampler(shuffle=False), num_workers=...)` at base.py:78-89: iterating yields the rank's items
    in order, `None` for those already on disk.  Items are decoded ahead of the GPU by `num_workers` host
    threads (PIL and the JPEG stager release the GIL) instead of worker processes: the PIL crop / resize work
    that made processes necessary lives on the GPU here.  `num_workers=0`: no threads (DRY_RUN, base.py:82-83).
    The rank's share is a crop-count balanced partition of the items still to do (`oadp_b200.dist`), not the
    sampler's round robin with wrap-around duplicates."""

    def __init__(self, dataset: BaseDataset, indices: Sequence[int], num_workers: int = 2, prefetch: int = 16) -> None:
        self.dataset = dataset
        self.indices = list(indices)
        self.num_workers = int(num_workers)
        self.prefetch = int(prefetch)

    def __len__(self) -> int:
        return len(self.indices)

    def __iter__(self) -> Iterator[Optional[T]]:
        if self.num_workers <= 0:
            for i in self.indices:
                yield self.dataset[i]
            return
        depth = max(self.prefetch, 2 * self.num_workers)
        with concurrent.futures.ThreadPoolExecutor(max_workers=self.num_workers) as pool:
            window: 'collections.deque' = collections.deque()
            it = iter(self.indices)
            for i in itertools.islice(it, depth):
                window.append(pool.submit(self.dataset.__getitem__, i))
            while window:
                item = window.popleft().result()
                nxt = next(it, None)
                if nxt is not None:
                    window.append(pool.submit(self.dataset.__getitem__, nxt))
                yield item


class BaseValidator(ABC, Generic[T]):
    """One split of one task: iterate images, encode on the GPU, write `.pth` files (base.py:75-113)."""

    def __init__(self, name: str, model: OakeModel, *, dataloader: Config, log: Optional[Config] = None,
                 batch_images: int = 8, store: str = 'pth', decode: str = 'pillow', collate: bool = False,
                 **_: Any) -> None:
        self._name = name
        self._model = model
        self._pipeline = OakePipeline(model.engine)
        self._log_interval = int((log or {}).get('interval', 50))
        self._batch_images = 1 if Store.DRY_RUN else int(batch_images)
        if decode not in ('pillow', 'gpu'):
            raise ValueError(f"decode must be 'pillow' or 'gpu', not {decode!r}")
        if store not in ('pth', 'packed'):
            raise ValueError(f"store must be 'pth' or 'packed', not {store!r}")
        self._decode, self._store = decode, store
        self._collate = bool(collate)
        self._rows = 0  # crops encoded by this rank in this split
        self._manifest: List[Tuple[int, int, Optional[torch.Tensor]]] = []  # (image id, rows, global embedding)
        self._packed: Optional[PackedWriter] = None
        dataloader = Config(dataloader)
        workers = int(dataloader.get('num_workers', 2)) or 1
        self._dataloader: DataLoader[T] = self._build_dataloader(dataloader)
        self._dataset: BaseDataset = self._dataloader.dataset
        if store == 'packed':
            out_dir = self._dataset._output_dir
            rank, _world = oake_dist.rank_world()
            serial = len(list(out_dir.glob(f'shard-{rank:05d}-*.idx.json')))
            self._packed = PackedWriter(str(out_dir), '', f'shard-{rank:05d}-{serial:03d}')
            workers = 1  # appends to one file, in order
        self._writer = concurrent.futures.ThreadPoolExecutor(max_workers=workers)

    def _write(self, result: Any, batch: T) -> None:
        if self._packed is not None:
            self._packed.add(key_of_batch(batch), result)
        else:
            torch.save(result, batch.output)

    # ------------------------------------------------------------------ the reference's class surface
    @classmethod
    def _build_model(cls) -> Tuple[OakeModel, Any]:
        """(model, preprocess) like the reference (base.py:91-94); preprocess is None: it lives on the GPU."""
        return OakeModel(default_params(), 'cuda'), None

    def _build_dataloader(self, config: Config) -> DataLoader[T]:
        """base.py:78-89.  `config.dataset` is the built dataset (the task's override does that, as the
        reference's do); DRY_RUN forces `num_workers = 0`; the rank's share replaces the DistributedSampler."""
        if Store.DRY_RUN:
            config.num_workers = 0
        dataset: BaseDataset = config.dataset
        dataset.gpu_decode = self._decode == 'gpu'
        if self._store == 'packed':
            dataset.use_packed()
        workers = int(config.get('num_workers', 2))
        return DataLoader(dataset, self._shard(dataset), workers, prefetch=2 * self._batch_images)

    def _control_run_iter(self, batch: Optional[T], memo: Memo) -> Optional[Control]:
        """base.py:96-104: an item that is already on disk comes back as None and is skipped."""
        if batch is None:
            return Control.CONTINUE
        return None

    def _run_iter(self, batch: T, memo: Memo) -> torch.Tensor:
        """base.py:106-113: the task's override has put the record to store in `memo['result']`."""
        self._write(memo['result'], batch)
        return torch.tensor(0.0)

    @abstractmethod
    def _submit(self, batches: List[T]):
        """Asynchronous multi-image form of `_run_iter`: -> a `Pending` whose result() is one record per
        batch, in the layout the reference stores."""

    # ---------------------------------------------------------------------------------- the loop
    def _shard(self, dataset: BaseDataset) -> List[int]:
        """This rank's indices: the items NOT yet on disk, partitioned by crop count.  Rank 0 decides and
        broadcasts, so that all ranks cut the same list even if some of them start after others have
        already written files.  (`auto_fix` keeps existing files in the list: they are re-loaded.)"""
        rank, world = oake_dist.rank_world()
        todo = list(range(len(dataset)))
        if world == 1:
            return todo
        if not dataset._auto_fix:
            todo = [i for i in todo if not dataset.exists(dataset.ids[i])]
        todo = oake_dist.broadcast_object(todo)
        costs = [dataset.cost(i) for i in todo]
        return [todo[j] for j in oake_dist.balanced_partition(costs, world)[rank]]

    def _batches(self) -> Iterator[List[T]]:
        group: List[T] = []
        memo: Memo = Memo()
        for batch in self._dataloader:
            if self._control_run_iter(batch, memo) is Control.CONTINUE:  # already on disk (base.py:45-47)
                continue
            group.append(batch)
            if len(group) == self._batch_images:
                yield group
                group = []
        if group:
            yield group

    def run(self) -> int:
        total = len(self._dataloader)
        if Store.DRY_RUN:
            self._dataloader.indices = self._dataloader.indices[:3]
        done, t0, pending = 0, time.perf_counter(), []
        in_flight = None  # (batches, Pending): the GPU works on it while the next group is decoded / staged

        def drain(entry):
            batches_, ticket = entry
            for batch, result in zip(batches_, ticket.result()):
                pending.append(self._writer.submit(self._write, result, batch))
                self._rows += result['embeddings'].shape[0] if isinstance(result, dict) else 1
                if self._collate:
                    rows = result['embeddings'].shape[0] if isinstance(result, dict) else 1
                    self._manifest.append((int(key_of_batch(batch)), rows, None if isinstance(result, dict) else result))

        try:
            for group in itertools.chain(self._batches(), [None]):
                nxt = (group, self._submit(group)) if group is not None else None
                if in_flight is not None:
                    drain(in_flight)
                in_flight = nxt
                if group is None:
                    break
                done += len(group)
                if done % max(self._log_interval, 1) < len(group):
                    dt = time.perf_counter() - t0
                    print(f'[{self._name}] {done}/{total} images, {done / dt:.1f} img/s', flush=True)
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
        elapsed = time.perf_counter() - t0
        rank, world = oake_dist.rank_world()
        # one machine-readable line per rank and split (tools/bench_cli_scaling.py adds them up)
        print('[oake-timing] ' + json.dumps(dict(name=self._name, rank=rank, world=world, images=done, crops=self._rows,
                                                 seconds=round(elapsed, 4), decode=self._decode, store=self._store,
                                                 output_dir=str(self._dataset._output_dir))), flush=True)
        if self._collate:
            self.collate()
        return done

    def collate(self) -> Optional[Dict[str, torch.Tensor]]:
        """`--override .collate:True`: the ranks' outputs collated with ONE exchange at the end of the split --
        `dist.all_gather_embeddings` (NCCL on GPUs: an all_gather of the counts + all_gather_into_tensor of the
        padded rows) over (image id, rows written[, the 512-d global embedding]).  Rank 0 writes
        `{output_dir}/manifest.pth` = {ids (N,) int64 ascending, rows (N,) int64[, embeddings (N,512) f16]}: for
        the globals task that is the whole feature table in one file, for blocks / objects the index of what the
        per-image files (or packed shards) hold.  The per-image outputs themselves never travel."""
        dev = self._pipeline.device if hasattr(self._pipeline, 'device') else torch.device('cpu')
        ids = torch.tensor([m[0] for m in self._manifest], dtype=torch.int64, device=dev)
        rows = torch.tensor([m[1] for m in self._manifest], dtype=torch.int64, device=dev).reshape(-1, 1)
        with_emb = self._manifest[0][2] is not None if self._manifest else self._collate_embeddings()
        all_rows, all_ids = oake_dist.all_gather_embeddings(rows, ids)
        out: Dict[str, torch.Tensor] = dict()
        if with_emb:
            emb = (torch.stack([m[2] for m in self._manifest]) if self._manifest else torch.zeros(0, 512, dtype=torch.float16)).to(dev)
            all_emb, _ = oake_dist.all_gather_embeddings(emb, ids)
        order = torch.argsort(all_ids)
        out['ids'], out['rows'] = all_ids[order].cpu(), all_rows[order, 0].cpu()
        if with_emb:
            out['embeddings'] = all_emb[order].cpu()
        rank, _ = oake_dist.rank_world()
        if rank == 0:
            torch.save(out, self._dataset._output_dir / 'manifest.pth')
        self._manifest = []
        return out if rank == 0 else None

    def _collate_embeddings(self) -> bool:
        """Whether this task's manifest carries the embeddings themselves (globals: one row per image)."""
        return False

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
