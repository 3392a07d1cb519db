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
    assert item.image.dtype.name == 'uint8' and item.image.shape == (427, 640, 3)
    assert item.extra.shape == (10, 5) and item.output.name == f'{info["ids"][1]:012d}.pth'
    assert ds.cost(0) == 10.0
    torch.save(dict(a=1), item.output)
    assert ds[1] is None  # already on disk -> skipped (base.py:45-47)
    item.output.write_bytes(b'truncated')
    assert ds[1] is None
    fix = oobjects.DatasetRegistry.build(Config(cfg.val.dataloader.dataset, auto_fix=True), default_config=dict(grid=14))
    assert fix[1] is not None  # auto_fix re-loads and finds the file broken (base.py:48-52)
    with pytest.raises(NotImplementedError):
        oobjects.DatasetRegistry.build(Config(cfg.val.dataloader.dataset, expand_mode='RECTANGLE'),
                                       default_config=dict(grid=14))
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
