"""CPU tier: host-side logic of the reference-facing mirror -- config loading (the reference's own
config files, when /root/reference is mounted), CLI override parsing, the balanced image shard and
the gloo world-size-2 path of the output collation."""
import os
import pathlib
import socket
import subprocess
import sys
import textwrap

import pytest

from oadp_b200 import dist as odist
from oadp_b200.compat import Config, DictAction

REF_CFG = pathlib.Path('/root/reference/configs/oake')
ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.mark.skipif(not REF_CFG.exists(), reason='reference tree not mounted (GPU box)')
def test_reference_configs_load_unchanged():
    c = Config.load(REF_CFG / 'objects_lvis.py')  # _base_ chain: objects_lvis -> objects_coco -> base
    ds = c.train.dataloader.dataset
    assert ds.type == 'LVISDataset' and ds.root == 'data/coco' and ds.proposal_sorted is True
    assert ds.output_dir == 'data/lvis_v1/oake/objects/train2017'
    assert c.mini_batch_size == 512 and c.log.interval == 5 and c.val.dataloader.num_workers == 2
    g = Config.load(REF_CFG / 'globals.py')
    assert g.val.dataloader.dataset.annFile == 'data/coco/annotations/instances_val2017.json'
    assert g.log.interval == 50


def test_config_base_delete_and_override(tmp_path):
    (tmp_path / 'base.py').write_text('a = dict(x=1, y=dict(z=2, w=3))\nb = [1, 2]\n')
    (tmp_path / 'child.py').write_text("_base_ = ['base.py']\na = dict(y=dict(_delete_=True, q=9), k=5)\n")
    c = Config.load(tmp_path / 'child.py')
    assert c.a.x == 1 and c.a.k == 5 and dict(c.a.y) == {'q': 9} and c.b == [1, 2]
    c.override({'.a.y.q': 10, 'n.m': 'v'})
    assert c.a.y.q == 10 and c.n.m == 'v'
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument('--override', action=DictAction, nargs='+')
    ns = ap.parse_args(['--override', '.train.dataloader.num_workers:0', '.name::007', '.f:1.5'])
    assert ns.override == {'.train.dataloader.num_workers': 0, '.name': '007', '.f': 1.5}


def test_balanced_partition_properties():
    costs = [300, 5, 290, 17, 300, 44, 1, 280, 300, 120, 9]
    for world in (1, 2, 3, 8):
        shards = odist.balanced_partition(costs, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(len(costs)))  # every image exactly once: no wrap-around duplicates
        loads = [sum(costs[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(costs)
    assert odist.balanced_partition([], 4) == [[], [], [], []]
    # the reference's round-robin on the same costs is worse or equal
    rr = [sum(costs[i] for i in range(r, len(costs), 2)) for r in range(2)]
    lpt = [sum(costs[i] for i in s) for s in odist.balanced_partition(costs, 2)]
    assert max(lpt) <= max(rr)


WORKER = textwrap.dedent('''
    import os, sys, torch, torch.distributed as dist
    sys.path.insert(0, {root!r})
    from oadp_b200 import dist as odist
    dist.init_process_group('gloo')
    rank, world = odist.rank_world()
    costs = [float(3 + (i * 7) % 11) for i in range(23)]
    mine = odist.balanced_partition(costs, world)[rank]
    emb = torch.stack([torch.full((512,), float(i)) for i in mine]).half() if mine else torch.zeros(0, 512).half()
    ids = torch.tensor(mine, dtype=torch.int64)
    e, i = odist.all_gather_embeddings(emb, ids)
    order = torch.argsort(i)
    assert i[order].tolist() == list(range(23)), i
    assert torch.equal(e[order][:, 0].float(), torch.arange(23).float())
    dist.destroy_process_group()
    print('rank', rank, 'ok', len(mine))
''')


def test_world_size_2_gloo_shard_and_collate(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER.format(root=str(ROOT)))
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE='2', LOCAL_RANK=str(r), MASTER_ADDR='127.0.0.1',
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all('ok' in o for o in outs)


def test_synthetic_dataset_and_skip_logic(tmp_path):
    import torch
    from oadp_b200 import synth
    from oadp_b200.oake import blocks as oblocks
    from oadp_b200.oake import objects as oobjects
    info = synth.write_coco_dataset(tmp_path, 3, seed=1, n_proposals=10)
    cfg = Config.load(info['configs']['objects'])
    ds = oobjects.DatasetRegistry.build(Config(cfg.val.dataloader.dataset), default_config=dict(grid=14))
    assert ds.ids == info['ids'] and len(ds) == 3
    item = ds[1]
    assert isinstance(item, oobjects.Batch) and item._fields == ('output', 'objects', 'bboxes', 'objectness', 'masks')
    assert item.objects.dtype.name == 'uint8' and item.objects.shape == (427, 640, 3)  # the image: crops are the GPU's
    assert item.bboxes.shape == (10, 4) and item.objectness.shape == (10, 1) and item.masks is None
    assert item.output.name == f'{info["ids"][1]:012d}.pth'
    assert ds.cost(0) == 10.0
    torch.save(dict(a=1), item.output)
    assert ds[1] is None  # already on disk -> skipped (base.py:45-47)
    item.output.write_bytes(b'truncated')
    assert ds[1] is None
    fix = oobjects.DatasetRegistry.build(Config(cfg.val.dataloader.dataset, auto_fix=True), default_config=dict(grid=14))
    assert fix[1] is not None  # auto_fix re-loads and finds the file broken (base.py:48-52)
    for mode in ('RECTANGLE', 'LONGEST_EDGE'):  # cannot run upstream either
        with pytest.raises(NotImplementedError):
            oobjects.DatasetRegistry.build(Config(cfg.val.dataloader.dataset, expand_mode=mode),
                                           default_config=dict(grid=14))
    const = oobjects.DatasetRegistry.build(Config(cfg.val.dataloader.dataset, expand_mode='CONSTANT'),
                                           default_config=dict(grid=14))
    assert const.expand_mode == 'CONSTANT' and ds.expand_mode == 'ADAPTIVE'
    b = oblocks.Dataset(**Config.load(info['configs']['blocks']).train.dataloader.dataset)
    assert [b.cost(i) for i in range(3)] == [27.0, 22.0, 27.0]


def test_dataset_hands_out_compressed_files_with_gpu_decode(tmp_path, lib):
    """`decode='gpu'`: the item is the still-compressed file for JPEGs the GPU decoder covers,
    Pillow's pixels for everything else -- either way with the shape the planner needs."""
    import numpy as np
    import PIL.Image
    from oadp_b200 import jpeg as oake_jpeg
    from oadp_b200 import synth
    from oadp_b200.oake import globals as oglobals
    info = synth.write_coco_dataset(tmp_path, 3, seed=2, fmt='jpg')
    cfg = Config.load(info['configs']['globals'])
    ds = oglobals.Dataset(**cfg.val.dataloader.dataset)
    assert isinstance(ds[0].image, np.ndarray)  # default: decoded on the host as the reference does
    ds.gpu_decode = True
    items = [ds[i] for i in range(3)]
    assert all(isinstance(it.image, oake_jpeg.JpegSource) for it in items)
    for it, (w, h) in zip(items, synth.COCO_SIZES):
        assert it.image.shape == (h, w, 3) and it.image.dtype == np.uint8 and it.image.ndim == 3
    # a progressive file in the same directory falls back to the reference's loader
    path = ds.image_path(ds.ids[1])
    PIL.Image.open(path).save(path, 'JPEG', progressive=True)
    got = ds[1].image
    assert isinstance(got, np.ndarray) and np.array_equal(got, np.asarray(PIL.Image.open(path).convert('RGB')))


# ------------------------------------------------------------------ the reference's validator class surface
class _FakePending:

    def __init__(self, value):
        self._value = value

    def result(self):
        return self._value


class _FakePipeline:
    """Stands where `OakePipeline` stands: records what the validator hands over, returns layouts."""

    def __init__(self):
        self.calls = []

    def _record(self, kind, images, extra=None):
        import torch
        self.calls.append((kind, len(images), extra))
        if kind == 'globals':
            return [torch.zeros(512, dtype=torch.float16) for _ in images]
        return [dict(embeddings=torch.zeros(2, 512, dtype=torch.float16), bboxes=torch.zeros(2, 4, dtype=torch.float16))
                for _ in images]

    def encode_globals(self, images):
        return self._record('globals', images)

    def submit_globals(self, images):
        return _FakePending(self._record('globals', images))

    def encode_objects(self, images, proposals, dry_run=False, expand_mode='ADAPTIVE'):
        return self._record('objects', images, (proposals[0].shape, dry_run, expand_mode))

    def submit_objects(self, images, proposals, dry_run=False, expand_mode='ADAPTIVE'):
        return _FakePending(self._record('objects', images, (proposals[0].shape, dry_run, expand_mode)))


class _FakeModel:
    engine = None

    class visual:  # noqa: N801
        grid = 14


def _validator(module, cfg_split, monkeypatch, **kwargs):
    from oadp_b200.oake import base as obase
    fake = _FakePipeline()
    monkeypatch.setattr(obase, 'OakePipeline', lambda engine: fake)
    return module.Validator('t', _FakeModel(), **cfg_split, **kwargs), fake


def test_validator_surface_matches_the_reference(tmp_path, monkeypatch):
    """`_build_dataloader(config)`, `_control_run_iter(batch, memo)`, `_run_iter(batch, memo) ->
    Tensor` writing `memo['result']` to `batch.output` (base.py:78-113), per-task `Batch` field names
    (globals.py:19-21, blocks.py:19-22, objects.py:24-29) -- and `run()` as the batched form of the same."""
    import torch
    from oadp_b200 import synth
    from oadp_b200.oake import base as obase
    from oadp_b200.oake import blocks as oblocks
    from oadp_b200.oake import globals as oglobals
    from oadp_b200.oake import objects as oobjects
    monkeypatch.delenv('DRY_RUN', raising=False)
    assert oglobals.Batch._fields == ('output', 'image')
    assert oblocks.Batch._fields == ('output', 'blocks', 'bboxes')
    assert oobjects.Batch._fields == ('output', 'objects', 'bboxes', 'objectness', 'masks')
    for mod in (oglobals, oblocks, oobjects):
        for name in ('_build_model', '_build_dataloader', '_control_run_iter', '_run_iter', 'main', 'run'):
            assert callable(getattr(mod.Validator, name)), (mod.__name__, name)
    info = synth.write_coco_dataset(tmp_path, 4, seed=5, n_proposals=7)

    cfg = Config.load(info['configs']['globals'])
    v, fake = _validator(oglobals, cfg.val, monkeypatch, batch_images=3)
    loader = v._dataloader
    assert isinstance(loader, obase.DataLoader) and isinstance(loader.dataset, oglobals.Dataset) and len(loader) == 4
    batches = list(loader)
    assert all(isinstance(b, oglobals.Batch) for b in batches)
    memo = obase.Memo()
    assert v._control_run_iter(None, memo) is obase.Control.CONTINUE and v._control_run_iter(batches[0], memo) is None
    out = v._run_iter(batches[0], memo)  # one image, synchronously, like the reference's loop body
    assert torch.is_tensor(out) and float(out) == 0.0 and batches[0].output.exists()
    assert torch.load(batches[0].output).shape == (512, ) and fake.calls == [('globals', 1, None)]
    assert v.run() == 3  # the other three, as one group of `batch_images`; the first one is skipped (None)
    assert fake.calls[1:] == [('globals', 3, None)]
    assert sorted(p.name for p in batches[0].output.parent.glob('*.pth')) == [f'{i:012d}.pth' for i in info['ids']]

    cfg = Config.load(info['configs']['objects'])
    v, fake = _validator(oobjects, cfg.val, monkeypatch, batch_images=8)
    assert isinstance(v._dataset, oobjects.COCODataset) and v._dataset._grid == 14  # grid from the model (objects.py:281)
    b0 = next(iter(v._dataloader))
    v._run_iter(b0, obase.Memo())
    assert fake.calls == [('objects', 1, ((7, 5), False, 'ADAPTIVE'))]
    assert set(torch.load(b0.output)) == {'embeddings', 'bboxes'}  # whatever the pipeline returned is what is stored


def test_dry_run_uses_no_workers_and_three_items(tmp_path, monkeypatch):
    from oadp_b200 import synth
    from oadp_b200.oake import globals as oglobals
    monkeypatch.setenv('DRY_RUN', 'True')
    info = synth.write_coco_dataset(tmp_path, 5, seed=6)
    cfg = Config.load(info['configs']['globals'])
    v, fake = _validator(oglobals, cfg.val, monkeypatch, batch_images=8)
    assert v._dataloader.num_workers == 0  # base.py:82-83
    assert v.run() == 3 and [c[1] for c in fake.calls] == [1, 1, 1]  # one image per step in DRY_RUN


SHARD_WORKER = textwrap.dedent('''
    import os, sys, json, pathlib, torch, torch.distributed as dist
    sys.path.insert(0, {root!r})
    sys.path.insert(0, {tests!r})
    import test_host_mirror as thm
    from oadp_b200.compat import Config
    from oadp_b200.oake import base as obase, globals as oglobals
    dist.init_process_group('gloo')
    rank = dist.get_rank()
    cfg = Config.load({config!r})
    out_dir = pathlib.Path(cfg.val.dataloader.dataset.output_dir)
    ids = {ids!r}
    if rank == 1:  # a rank that starts late sees a file its peer has written meanwhile: its view must not count
        seen = oglobals.Dataset.exists
        oglobals.Dataset.exists = lambda self, id_: seen(self, id_) or id_ == ids[5]
    fake = thm._FakePipeline()
    obase.OakePipeline = lambda engine: fake
    v = oglobals.Validator('t', thm._FakeModel(), **cfg.val, collate=True)
    mine = [v._dataset.ids[i] for i in v._dataloader.indices]
    v.run()
    print(json.dumps(dict(rank=rank, mine=mine)))
    dist.destroy_process_group()
''')


def test_world_size_2_shard_skips_done_items_and_collates(tmp_path):
    """Two gloo ranks: the share is cut from the items NOT yet on disk (rank 0's view, broadcast), no image is
    encoded twice, and `collate` leaves one manifest with every id of the run."""
    import json
    import torch
    from oadp_b200 import synth
    info = synth.write_coco_dataset(tmp_path / 'ds', 9, seed=8)
    cfg = Config.load(info['configs']['globals'])
    out_dir = pathlib.Path(cfg.val.dataloader.dataset.output_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    ids = info['ids']
    for done in (ids[0], ids[3]):
        torch.save(torch.zeros(512).half(), out_dir / f'{done:012d}.pth')
    script = tmp_path / 'shard_worker.py'
    script.write_text(SHARD_WORKER.format(root=str(ROOT), tests=str(ROOT / 'tests'), config=info['configs']['globals'], ids=ids))
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE='2', LOCAL_RANK=str(r), MASTER_ADDR='127.0.0.1',
                   MASTER_PORT=str(port), OAKE_ALLOW_RANDOM_WEIGHTS='1')
        env.pop('DRY_RUN', None)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=180) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    shares = [json.loads(o[0].strip().splitlines()[-1])['mine'] for o in outs]
    todo = [i for i in ids if i not in (ids[0], ids[3])]  # rank 0's view (ids[5] included), broadcast to rank 1
    assert sorted(shares[0] + shares[1]) == todo and not set(shares[0]) & set(shares[1])
    assert abs(len(shares[0]) - len(shares[1])) <= 1
    manifest = torch.load(out_dir / 'manifest.pth')
    assert manifest['ids'].tolist() == todo and manifest['rows'].tolist() == [1] * len(todo)
    assert manifest['embeddings'].shape == (len(todo), 512) and manifest['embeddings'].dtype == torch.float16
    assert sorted(p.stem for p in out_dir.glob('0*.pth')) == [f'{i:012d}' for i in ids]
