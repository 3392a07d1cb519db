"""Feature store on either side of the OAKE hot path.

The reference keeps one tiny pickle per image and task -- `data/<ds>/oake/<task>/<split>/{id:012d}.pth`
written by `torch.save` (oadp/oake/base.py:112) -- and the detector's `LoadCLIPFeatures` pipeline
step re-opens three of them per sample through todd's `PthAccessLayer`
(oadp/dp/datasets.py:137-214; configs/dp/datasets/ov_coco.py:23-32).  With ~118 k train images x 3
tasks that is ~350 k small files read every epoch by DataLoader workers.

Two access layers with the same `Mapping[key] -> value` contract (`key` = `f'{image_id:012d}'`,
value = exactly what `torch.load` of the reference file returns):

  * `PthStore`      -- the reference layout, file per key (read and write); drop-in for
                       `PthAccessLayer(data_root=..., task_name=...)`.
  * `PackedStore`   -- one shard per writer (rank): `<task>/<name>.bin` holds the raw fp16 payloads
                       back to back (64-byte aligned), `<task>/<name>.idx.json` the per-key layout.
                       Readers memory-map the shards, so a lookup is two slices and no pickle;
                       values come back as torch views of the map (copy-on-write, fork-safe).

`LoadCLIPFeatures` mirrors the reference pipeline step on top of either layer, including its quirks:
block boxes are used as stored (fp16, first row `(x0, y0, side, side)`), the
objects' `min_wh=(4,4)` filter is re-applied on the fp16 boxes, pseudo labels (`>= num_all`) are
dropped before the block multi-labels are built.
"""
from __future__ import annotations

import json
import mmap
import os
import pathlib
from collections.abc import Mapping
from typing import Any, Dict, Iterator, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

Value = Union[torch.Tensor, Dict[str, torch.Tensor]]
_ALIGN = 64
_DTYPES = {'float16': np.float16, 'float32': np.float32}


def key_of(image_id: int) -> str:
    return f'{int(image_id):012d}'


# ------------------------------------------------------------------------------------ file per key
class PthStore(Mapping):
    """`data_root/task_name/{key}.pth` -- todd `PthAccessLayer` layout (datasets.py:146-160)."""

    def __init__(self, data_root: str, task_name: str = '', **_: Any) -> None:
        self._dir = pathlib.Path(data_root) / task_name

    def __getitem__(self, key: str) -> Value:
        path = self._dir / f'{key}.pth'
        if not path.exists():
            raise KeyError(key)
        return torch.load(path, 'cpu')

    def __setitem__(self, key: str, value: Value) -> None:
        self._dir.mkdir(parents=True, exist_ok=True)
        torch.save(value, self._dir / f'{key}.pth')

    def __iter__(self) -> Iterator[str]:
        return (p.stem for p in sorted(self._dir.glob('*.pth')))

    def __len__(self) -> int:
        return sum(1 for _ in self._dir.glob('*.pth'))


# ------------------------------------------------------------------------------------------ packed
class PackedWriter:
    """Appends values to one shard.  One writer per process (rank); `close()` publishes the index."""

    def __init__(self, data_root: str, task_name: str, shard: str = 'shard-00000') -> None:
        self._dir = pathlib.Path(data_root) / task_name
        self._dir.mkdir(parents=True, exist_ok=True)
        self._bin_path = self._dir / f'{shard}.bin'
        self._idx_path = self._dir / f'{shard}.idx.json'
        self._bin = open(self._bin_path, 'wb')
        self._off = 0
        self._index: Dict[str, Any] = {}

    def _put(self, t: torch.Tensor) -> Dict[str, Any]:
        a = t.detach().cpu().contiguous().numpy()
        if a.dtype.name not in _DTYPES:
            raise TypeError(f'unsupported dtype {a.dtype}')
        pad = (-self._off) % _ALIGN
        if pad:
            self._bin.write(b'\0' * pad)
            self._off += pad
        self._bin.write(a.tobytes())
        ent = dict(o=self._off, s=list(a.shape), d=a.dtype.name)
        self._off += a.nbytes
        return ent

    def add(self, key: str, value: Value) -> None:
        if key in self._index:
            raise KeyError(f'duplicate key {key}')
        if isinstance(value, torch.Tensor):
            self._index[key] = self._put(value)
        else:
            self._index[key] = {name: self._put(t) for name, t in value.items()}

    def close(self) -> None:
        self._bin.flush()
        os.fsync(self._bin.fileno())
        self._bin.close()
        tmp = self._idx_path.with_suffix('.tmp')
        tmp.write_text(json.dumps(dict(version=1, bytes=self._off, keys=self._index)))
        os.replace(tmp, self._idx_path)  # the index appears only once the payload is complete

    def __enter__(self) -> 'PackedWriter':
        return self

    def __exit__(self, *exc: Any) -> None:
        self.close()


class PackedStore(Mapping):
    """Read side: every `*.idx.json` under `data_root/task_name` is one shard."""

    def __init__(self, data_root: str, task_name: str = '', **_: Any) -> None:
        self._dir = pathlib.Path(data_root) / task_name
        self._where: Dict[str, Tuple[int, Any]] = {}
        self._bins: List[pathlib.Path] = []
        self._maps: List[Optional[np.ndarray]] = []
        for idx in sorted(self._dir.glob('*.idx.json')):
            meta = json.loads(idx.read_text())
            b = idx.with_name(idx.name[:-len('.idx.json')] + '.bin')
            if b.stat().st_size < meta['bytes']:
                raise IOError(f'{b} is shorter than its index says (truncated shard)')
            n = len(self._bins)
            self._bins.append(b)
            self._maps.append(None)
            for k, ent in meta['keys'].items():
                self._where[k] = (n, ent)

    def _map(self, n: int) -> np.ndarray:
        if self._maps[n] is None:  # opened lazily: DataLoader workers map after fork
            with open(self._bins[n], 'rb') as f:
                mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_COPY)
            self._maps[n] = np.frombuffer(mm, dtype=np.uint8)
        return self._maps[n]

    def _get(self, n: int, ent: Dict[str, Any]) -> torch.Tensor:
        dt = np.dtype(_DTYPES[ent['d']])
        count = int(np.prod(ent['s'])) if ent['s'] else 1
        raw = self._map(n)[ent['o']:ent['o'] + count * dt.itemsize]
        return torch.from_numpy(raw.view(dt).reshape(ent['s']))

    def __getitem__(self, key: str) -> Value:
        n, ent = self._where[key]
        if 'o' in ent:
            return self._get(n, ent)
        return {name: self._get(n, e) for name, e in ent.items()}

    def __iter__(self) -> Iterator[str]:
        return iter(sorted(self._where))

    def __len__(self) -> int:
        return len(self._where)

    def __getstate__(self) -> Dict[str, Any]:  # picklable for spawn-ed workers: maps are re-opened
        st = dict(self.__dict__)
        st['_maps'] = [None] * len(self._bins)
        return st


def pack(src: Mapping, data_root: str, task_name: str, shard: str = 'shard-00000',
         keys: Optional[Sequence[str]] = None) -> int:
    """Converts any store (e.g. a reference `.pth` directory) into one packed shard."""
    n = 0
    with PackedWriter(data_root, task_name, shard) as w:
        for k in (keys if keys is not None else list(src)):
            w.add(k, src[k])
            n += 1
    return n


ACCESS_LAYERS = {'PthAccessLayer': PthStore, 'PthStore': PthStore, 'PackedStore': PackedStore}


def build_access_layer(config: Optional[Mapping], default: Mapping) -> Optional[Mapping]:
    """`ALR.build(config, default)` of the reference (datasets.py:153-160): `config` overrides
    `default`; `type` picks the layer (the reference's configs say `PthAccessLayer`)."""
    if config is None:
        return None
    cfg = dict(default)
    cfg.update(config)
    return ACCESS_LAYERS[cfg.pop('type', 'PthAccessLayer')](**cfg)


# --------------------------------------------------------------------------- LoadCLIPFeatures
def pairwise_intersection(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """todd `BBoxesXYXY.__and__`: (n, m) intersection areas of xyxy boxes."""
    lt = torch.maximum(a[:, None, :2], b[None, :, :2])
    rb = torch.minimum(a[:, None, 2:], b[None, :, 2:])
    wh = (rb - lt).clamp_min(0)
    return wh[..., 0] * wh[..., 1]


class LoadCLIPFeatures:
    """Pipeline step of oadp/dp/datasets.py:137-214 on top of `build_access_layer` stores."""

    def __init__(self, default: Mapping, globals_: Optional[Mapping] = None, blocks: Optional[Mapping] = None,
                 objects: Optional[Mapping] = None, num_all: Optional[int] = None) -> None:
        assert globals_ is not None or blocks is not None or objects is not None
        default = dict(default)
        if os.environ.get('TRAIN_WITH_VAL_DATASET', '') not in ('', '0', 'False', 'false'):
            default['task_name'] = default['task_name'].replace('train', 'val')  # datasets.py:150-152
        self._globals = build_access_layer(globals_, default)
        self._blocks = build_access_layer(blocks, default)
        self._objects = build_access_layer(objects, default)
        self._num_all = num_all
        self._dry_key: Optional[str] = None
        if os.environ.get('DRY_RUN', '') not in ('', '0', 'False', 'false'):  # datasets.py:163-169
            keys = [set(m.keys()) for m in (self._globals, self._blocks, self._objects) if m is not None]
            self._dry_key = sorted(set.intersection(*keys))[0]

    def _categories_num_all(self) -> int:
        if self._num_all is not None:
            return self._num_all
        from .dp.categories import Globals
        return Globals.categories.num_all

    def __call__(self, results: Dict[str, Any]) -> Dict[str, Any]:
        key = self._dry_key if self._dry_key is not None else key_of(results['img_info']['id'])
        bbox_fields: List[str] = results['bbox_fields']

        if self._globals is not None:
            results['clip_global'] = self._globals[key].squeeze(0)

        if self._blocks is not None:
            blocks = self._blocks[key]
            block_bboxes = blocks['bboxes']
            if 'gt_bboxes' in results:
                num_all = self._categories_num_all()
                gt_bboxes = np.asarray(results['gt_bboxes'])
                gt_labels = np.asarray(results['gt_labels'])
                keep = gt_labels < num_all  # pseudo labels out (datasets.py:185-188)
                gt_bboxes, gt_labels = gt_bboxes[keep], gt_labels[keep]
                inter = pairwise_intersection(block_bboxes, torch.as_tensor(gt_bboxes).reshape(-1, 4))
                block_ids, gt_ids = torch.where(inter > 0)
                labels = np.zeros((block_bboxes.shape[0], num_all), dtype=bool)
                labels[block_ids.numpy(), gt_labels[gt_ids.numpy()]] = True
                results['block_labels'] = labels
            results['clip_blocks'] = blocks['embeddings']
            results['block_bboxes'] = block_bboxes.float().numpy()
            bbox_fields.append('block_bboxes')

        if self._objects is not None:
            objects = self._objects[key]
            object_bboxes = objects['bboxes']
            wh = object_bboxes[:, 2:] - object_bboxes[:, :2]  # fp16 arithmetic, as stored
            indices = (wh[:, 0] >= 4) & (wh[:, 1] >= 4)
            results['clip_objects'] = objects['embeddings'][indices]
            results['object_bboxes'] = object_bboxes[indices].float().numpy()
            bbox_fields.append('object_bboxes')

        return results
