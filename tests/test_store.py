"""CPU tier: feature store -- the reference's file-per-key layout, the packed shards,
and the LoadCLIPFeatures pipeline step (oadp/dp/datasets.py:137-214) on top of both."""
import pickle

import numpy as np
import pytest
import torch

from oadp_b200 import store


def make_values(seed=0, n=5):
    g = torch.Generator().manual_seed(seed)
    vals = {}

    def boxes(m):
        lt = torch.rand(m, 2, generator=g) * 400
        return torch.cat([lt, lt + 5 + torch.rand(m, 2, generator=g) * 200], 1).half()

    for i in range(n):
        nb, no = 3 + i, 7 + 2 * i
        vals[store.key_of(100 + i)] = dict(
            globals=torch.randn(512, generator=g).half(),
            blocks=dict(embeddings=torch.randn(nb, 512, generator=g).half(),
                        bboxes=boxes(nb)),
            objects=dict(embeddings=torch.randn(no, 512, generator=g).half(),
                         bboxes=boxes(no),
                         objectness=torch.rand(no, 1, generator=g).half()))
    return vals


def same(a, b):
    if isinstance(a, dict):
        return set(a) == set(b) and all(same(a[k], b[k]) for k in a)
    return a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b)


def test_pth_store_is_the_reference_layout(tmp_path):
    vals = make_values()
    s = store.PthStore(str(tmp_path), 'coco/oake/objects/val')
    for k, v in vals.items():
        s[k] = v['objects']
    f = tmp_path / 'coco/oake/objects/val' / '000000000100.pth'  # {id:012d}.pth, torch.save (base.py:44,112)
    assert f.exists() and same(torch.load(f, 'cpu'), vals['000000000100']['objects'])
    assert sorted(s) == sorted(vals) and len(s) == len(vals)
    with pytest.raises(KeyError):
        s['000000000999']


@pytest.mark.parametrize('task', ['globals', 'blocks', 'objects'])
def test_packed_round_trip_two_shards(tmp_path, task):
    vals = make_values()
    keys = sorted(vals)
    with store.PackedWriter(str(tmp_path), task, 'shard-00000-000') as w:
        for k in keys[:3]:
            w.add(k, vals[k][task])
    with store.PackedWriter(str(tmp_path), task, 'shard-00001-000') as w:
        for k in keys[3:]:
            w.add(k, vals[k][task])
    s = store.PackedStore(str(tmp_path), task)
    assert list(s) == keys
    for k in keys:
        assert same(s[k], vals[k][task])
    s2 = pickle.loads(pickle.dumps(s))  # spawn-ed DataLoader workers
    assert same(s2[keys[-1]], vals[keys[-1]][task])


def test_packed_rejects_duplicates_and_truncation(tmp_path):
    vals = make_values(n=2)
    w = store.PackedWriter(str(tmp_path), 't')
    k = sorted(vals)[0]
    w.add(k, vals[k]['globals'])
    with pytest.raises(KeyError):
        w.add(k, vals[k]['globals'])
    w.close()
    b = tmp_path / 't' / 'shard-00000.bin'
    b.write_bytes(b.read_bytes()[:100])
    with pytest.raises(IOError):
        store.PackedStore(str(tmp_path), 't')


def test_pack_converts_a_reference_directory(tmp_path):
    vals = make_values()
    src = store.PthStore(str(tmp_path / 'pth'), 'blocks')
    for k, v in vals.items():
        src[k] = v['blocks']
    assert store.pack(src, str(tmp_path / 'packed'), 'blocks') == len(vals)
    dst = store.PackedStore(str(tmp_path / 'packed'), 'blocks')
    assert all(same(dst[k], src[k]) for k in src)


def reference_load_clip_features(results, g, b, o, num_all):
    """Line-by-line restatement of datasets.py:171-214 (torch.where over pairwise intersections)."""
    out = dict(results)
    out['bbox_fields'] = list(results['bbox_fields'])
    out['clip_global'] = g.squeeze(0)
    bb = b['bboxes']
    keep = results['gt_labels'] < num_all
    gtb, gtl = results['gt_bboxes'][keep], results['gt_labels'][keep]
    labels = np.zeros((bb.shape[0], num_all), dtype=bool)
    for i in range(bb.shape[0]):
        for j in range(gtb.shape[0]):
            w = min(float(bb[i, 2]), float(gtb[j, 2])) - max(float(bb[i, 0]), float(gtb[j, 0]))
            h = min(float(bb[i, 3]), float(gtb[j, 3])) - max(float(bb[i, 1]), float(gtb[j, 1]))
            if max(w, 0.0) * max(h, 0.0) > 0:
                labels[i, gtl[j]] = True
    out['block_labels'] = labels
    out['clip_blocks'] = b['embeddings']
    out['block_bboxes'] = bb.float().numpy()
    ob = o['bboxes']
    idx = ((ob[:, 2] - ob[:, 0]) >= 4) & ((ob[:, 3] - ob[:, 1]) >= 4)
    out['clip_objects'] = o['embeddings'][idx]
    out['object_bboxes'] = ob[idx].float().numpy()
    out['bbox_fields'] += ['block_bboxes', 'object_bboxes']
    return out


@pytest.mark.parametrize('layer', ['PthAccessLayer', 'PackedStore'])
def test_load_clip_features(tmp_path, layer):
    vals = make_values()
    # tiny boxes so that the objects min_wh filter has something to drop
    for v in vals.values():
        v['objects']['bboxes'][0] = torch.tensor([10.0, 10.0, 13.0, 40.0]).half()
    for task in ('globals', 'blocks', 'objects'):
        for split in ('train', ):
            name = f'coco/oake/{task}/{split}'
            if layer == 'PackedStore':
                with store.PackedWriter(str(tmp_path), name) as w:
                    for k, v in vals.items():
                        w.add(k, v[task])
            else:
                s = store.PthStore(str(tmp_path), name)
                for k, v in vals.items():
                    s[k] = v[task]
    default = dict(type=layer, data_root=str(tmp_path))  # configs/dp/datasets/ov_coco.py:23-32 shape
    step = store.LoadCLIPFeatures(default, globals_=dict(task_name='coco/oake/globals/train'),
                                  blocks=dict(task_name='coco/oake/blocks/train'),
                                  objects=dict(task_name='coco/oake/objects/train'), num_all=65)
    rng = np.random.default_rng(0)
    for k, v in vals.items():
        xy = rng.uniform(0, 400, size=(6, 2)).astype(np.float32)
        gt = np.concatenate([xy, xy + rng.uniform(5, 200, size=(6, 2)).astype(np.float32)], 1)
        labels = np.array([3, 64, 65, 70, 0, 12])  # 65, 70 are pseudo labels (>= num_all)
        res = dict(img_info=dict(id=int(k)), bbox_fields=['gt_bboxes'], gt_bboxes=gt, gt_labels=labels)
        got = step(dict(res, bbox_fields=list(res['bbox_fields'])))
        want = reference_load_clip_features(res, v['globals'], v['blocks'], v['objects'], 65)
        assert got['bbox_fields'] == want['bbox_fields']
        for name in ('clip_global', 'clip_blocks', 'clip_objects'):
            assert torch.equal(got[name], want[name])
        for name in ('block_labels', 'block_bboxes', 'object_bboxes'):
            assert np.array_equal(got[name], want[name])
        assert got['clip_objects'].shape[0] == v['objects']['embeddings'].shape[0] - 1
        assert not got['block_labels'][:, 65:].any() if got['block_labels'].shape[1] > 65 else True


@pytest.mark.parametrize('layer', ['PthAccessLayer', 'PackedStore'])
def test_load_clip_features_vs_reference_class(tmp_path, layer):
    """oadp_b200.store.LoadCLIPFeatures against the reference's own `LoadCLIPFeatures.__call__`
    (oadp/dp/datasets.py:171-214, executed on stubs by tests/golden/make_ref_golden.py): every result
    key, dtype and value, for both storage layouts."""
    import pathlib
    import sys
    sys.path.insert(0, str(pathlib.Path(__file__).parent / 'golden'))
    import make_ref_golden as mk
    ref = torch.load(pathlib.Path(__file__).parent / 'golden' / 'ref_golden.pt', weights_only=False)['load_clip_features']
    records, results = mk.feature_store_inputs()
    for task in ('globals', 'blocks', 'objects'):
        name = f'coco/oake/{task}/train'
        if layer == 'PackedStore':
            with store.PackedWriter(str(tmp_path), name) as w:
                for k, v in records.items():
                    w.add(k, v[task])
        else:
            s = store.PthStore(str(tmp_path), name)
            for k, v in records.items():
                s[k] = v[task]
    step = store.LoadCLIPFeatures(dict(type=layer, data_root=str(tmp_path)),
                                  globals_=dict(task_name='coco/oake/globals/train'),
                                  blocks=dict(task_name='coco/oake/blocks/train'),
                                  objects=dict(task_name='coco/oake/objects/train'), num_all=6)
    for key, res in results.items():
        got = step(dict(res, bbox_fields=list(res['bbox_fields'])))
        want = ref[key]
        assert got['bbox_fields'] == want['bbox_fields']
        for name in ('clip_global', 'clip_blocks', 'clip_objects'):
            assert got[name].dtype == want[name].dtype and torch.equal(got[name], want[name]), name
        for name in ('block_bboxes', 'block_labels', 'object_bboxes'):
            assert got[name].dtype == want[name].dtype and np.array_equal(got[name], want[name]), name
        assert not got['block_labels'].all() and got['block_labels'].any()
