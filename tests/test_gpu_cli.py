"""GPU tier: the reference-facing entry points end to end -- `oadp.oake.{globals,blocks,objects}`
Validator.main on a synthetic COCO-format dataset: output files, layouts, resume, oracle parity."""
import pathlib

import numpy as np
import PIL.Image
import pytest
import torch
import torch.nn.functional as F

from oadp_b200 import synth
from oracle import frontend as ofe
from oracle import vit

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dataset(tmp_path_factory, lib):
    return synth.write_coco_dataset(tmp_path_factory.mktemp('coco'), 5, seed=2, n_proposals=20)


@pytest.fixture(autouse=True)
def random_weights_opt_in(monkeypatch):
    # no CLIP checkpoint exists offline: the CLI runs on seeded random weights only when told to
    monkeypatch.setenv('OAKE_ALLOW_RANDOM_WEIGHTS', '1')


def cos_ok(got, want):
    return float((1 - F.cosine_similarity(got.float(), want.float(), dim=-1)).max()) < 1e-3


def test_cli_all_tasks(dataset, monkeypatch):
    monkeypatch.delenv('DRY_RUN', raising=False)
    monkeypatch.delenv('OAKE_CLIP_WEIGHTS', raising=False)
    import oadp.oake.blocks as cli_blocks
    import oadp.oake.globals as cli_globals
    import oadp.oake.objects as cli_objects
    root = pathlib.Path(dataset['root'])
    p = vit.init_visual_params(0)  # what default_params() falls back to (seeded random ViT-B/32)
    p197 = vit.objects_surgery(p)
    id0 = dataset['ids'][0]
    pil = PIL.Image.open(root / 'images' / f'{id0:012d}.png').convert('RGB')
    props = torch.from_numpy(synth.proposals(*pil.size, 20, seed=2 * 7919 + 0))

    cli_globals.Validator.main(['t', dataset['configs']['globals'], '--override', '.batch_images:4'])
    for split in ('train', 'val'):
        files = sorted((root / 'oake' / 'globals' / split).glob('*.pth'))
        assert [f.stem for f in files] == [f'{i:012d}' for i in dataset['ids']]
    g = torch.load(root / 'oake' / 'globals' / 'val' / f'{id0:012d}.pth')
    assert g.shape == (512, ) and g.dtype == torch.float16
    # each file holds its own rows only, not the storage of the whole batch behind a view
    assert g.untyped_storage().nbytes() == 1024
    assert (root / 'oake' / 'globals' / 'val' / f'{id0:012d}.pth').stat().st_size < 4096
    assert cos_ok(g[None], vit.normalize_half(vit.encode_image(p, ofe.globals_preprocess(pil)[None])))

    cli_blocks.Validator.main(['t', dataset['configs']['blocks']])
    b = torch.load(root / 'oake' / 'blocks' / 'train' / f'{id0:012d}.pth')
    rb = ofe.blocks_preprocess(pil)
    assert set(b) == {'embeddings', 'bboxes'} and b['embeddings'].dtype == torch.float16
    assert torch.equal(b['bboxes'], rb.bboxes.half())
    assert cos_ok(b['embeddings'], vit.normalize_half(vit.encode_image(p, rb.blocks)))
    assert b['embeddings'].untyped_storage().nbytes() == b['embeddings'].numel() * 2

    cli_objects.Validator.main(['t', dataset['configs']['objects']])
    o = torch.load(root / 'oake' / 'objects' / 'val' / f'{id0:012d}.pth')
    ro = ofe.objects_preprocess(pil, props)
    assert set(o) == {'embeddings', 'bboxes', 'objectness'}
    assert o['objectness'].shape == (ro.objectness.shape[0], 1) and torch.equal(o['bboxes'], ro.bboxes.half())
    assert cos_ok(o['embeddings'], vit.normalize_half(vit.encode_objects(p197, ro.objects, ro.masks)))
    assert o['embeddings'].untyped_storage().nbytes() == o['embeddings'].numel() * 2
    size = (root / 'oake' / 'objects' / 'val' / f'{id0:012d}.pth').stat().st_size
    assert size < o['embeddings'].numel() * 2 + 8192

    # resume: everything is on disk, a second run must not rewrite a single file
    victim = root / 'oake' / 'objects' / 'val' / f'{dataset["ids"][2]:012d}.pth'
    before = {f: f.stat().st_mtime_ns for f in (root / 'oake' / 'objects' / 'val').glob('*.pth')}
    victim.unlink()
    cli_objects.Validator.main(['t', dataset['configs']['objects']])
    after = {f: f.stat().st_mtime_ns for f in (root / 'oake' / 'objects' / 'val').glob('*.pth')}
    assert victim in after and all(after[f] == t for f, t in before.items() if f != victim)


def test_cli_dry_run(dataset, monkeypatch, tmp_path):
    monkeypatch.setenv('DRY_RUN', 'True')
    import oadp.oake.objects as cli_objects
    out = tmp_path / 'dry'
    cli_objects.Validator.main(['t', dataset['configs']['objects'], '--override',
                                f'.val.dataloader.dataset.output_dir::{out}/val',
                                f'.train.dataloader.dataset.output_dir::{out}/train'])
    files = sorted((out / 'val').glob('*.pth'))
    assert 1 <= len(files) <= 3
    assert torch.load(files[0])['embeddings'].shape[0] <= 5  # objects.py:166-167


def test_cli_packed_store_matches_pth(dataset, monkeypatch, tmp_path):
    """`--override .store:packed`: one shard per rank holding exactly what the `.pth`
    files hold; a second run resumes from the shard index and writes nothing new."""
    monkeypatch.delenv('DRY_RUN', raising=False)
    import oadp.oake.blocks as cli_blocks
    from oadp_b200 import store
    root = pathlib.Path(dataset['root'])
    out = tmp_path / 'packed'
    ov = ['--override', '.store:packed', f'.val.dataloader.dataset.output_dir::{out}/val',
          f'.train.dataloader.dataset.output_dir::{out}/train']
    cli_blocks.Validator.main(['t', dataset['configs']['blocks']] + ov)
    packed = store.PackedStore(str(out), 'val')
    assert list(packed) == [f'{i:012d}' for i in dataset['ids']]
    if not (root / 'oake' / 'blocks' / 'val').exists():
        cli_blocks.Validator.main(['t', dataset['configs']['blocks']])
    pth = store.PthStore(str(root / 'oake' / 'blocks'), 'val')
    for k in packed:
        a, b = packed[k], pth[k]
        assert torch.equal(a['embeddings'], b['embeddings']) and torch.equal(a['bboxes'], b['bboxes'])
    cli_blocks.Validator.main(['t', dataset['configs']['blocks']] + ov)  # resume
    again = store.PackedStore(str(out), 'val')
    assert len(again) == len(packed)
    assert sum(1 for _ in (out / 'val').glob('*.idx.json')) == 2  # the second run's (empty) shard


def test_cli_refuses_random_weights_unless_told(dataset, monkeypatch):
    """`clip.load_default` fails without a checkpoint; so does this CLI (no silent random features)."""
    monkeypatch.delenv('DRY_RUN', raising=False)
    monkeypatch.delenv('OAKE_CLIP_WEIGHTS', raising=False)
    monkeypatch.delenv('OAKE_ALLOW_RANDOM_WEIGHTS', raising=False)
    import oadp.oake.globals as cli_globals
    with pytest.raises(RuntimeError, match='OAKE_CLIP_WEIGHTS'):
        cli_globals.Validator.main(['t', dataset['configs']['globals']])


def test_run_iter_is_the_single_image_form_of_run(dataset, monkeypatch, tmp_path):
    """The reference's loop body `_run_iter(batch, memo)` (base.py:106-113) and the batched `run()` write the
    same bytes; a batch preprocessed the reference's way (float crops + masks) goes through the same tower."""
    monkeypatch.delenv('DRY_RUN', raising=False)
    from oadp_b200.compat import Config
    from oadp_b200.oake import base as obase
    from oadp_b200.oake import objects as oobjects
    cfg = Config.load(dataset['configs']['objects'])
    model, preprocess = oobjects.Validator._build_model()
    assert preprocess is None and model.visual.grid == 14
    outs = {}
    for name in ('single', 'batched'):
        split = Config(cfg.val)
        split.dataloader.dataset.output_dir = str(tmp_path / name)
        v = oobjects.Validator('t', model, **split, **{k: v for k, v in cfg.items() if k not in ('train', 'val')})
        if name == 'single':
            for batch in v._dataloader:
                memo = obase.Memo()
                if v._control_run_iter(batch, memo) is obase.Control.CONTINUE:
                    continue
                assert float(v._run_iter(batch, memo)) == 0.0 and set(memo['result']) == {'embeddings', 'bboxes', 'objectness'}
        else:
            assert v.run() == len(dataset['ids'])
        outs[name] = {p.name: torch.load(p) for p in sorted((tmp_path / name).glob('*.pth'))}
    assert outs['single'].keys() == outs['batched'].keys() and len(outs['single']) == len(dataset['ids'])
    for k, a in outs['single'].items():
        for field in ('embeddings', 'bboxes', 'objectness'):
            assert torch.equal(a[field], outs['batched'][k][field])
    # the reference's own batch layout: PIL-preprocessed crops, filtered boxes, 14x14 masks (objects.py:157-186)
    root = pathlib.Path(dataset['root'])
    id0 = dataset['ids'][0]
    pil = PIL.Image.open(root / 'images' / f'{id0:012d}.png').convert('RGB')
    ro = ofe.objects_preprocess(pil, torch.from_numpy(synth.proposals(*pil.size, 20, seed=2 * 7919 + 0)))
    ref_batch = oobjects.Batch(tmp_path / 'ref_style.pth', ro.objects, ro.bboxes, ro.objectness, ro.masks)
    memo = obase.Memo()
    v._run_iter(ref_batch, memo)
    stored = torch.load(tmp_path / 'ref_style.pth')
    native = outs['batched'][f'{id0:012d}.pth']
    assert torch.equal(stored['bboxes'], native['bboxes']) and torch.equal(stored['objectness'], native['objectness'])
    assert cos_ok(stored['embeddings'], native['embeddings'])  # PIL pixels vs GPU pixels: same crops up to fp16 rounding


def test_cli_expand_mode_constant(dataset, monkeypatch, tmp_path):
    """`expand_mode='CONSTANT'` (objects.py:92-93): 224-px squares instead of sqrt(8 w h)."""
    monkeypatch.delenv('DRY_RUN', raising=False)
    import oadp.oake.objects as cli_objects
    root = pathlib.Path(dataset['root'])
    out = tmp_path / 'const'
    cli_objects.Validator.main(['t', dataset['configs']['objects'], '--override',
                                '.val.dataloader.dataset.expand_mode::CONSTANT', '.train.dataloader.dataset.expand_mode::CONSTANT',
                                f'.val.dataloader.dataset.output_dir::{out}/val',
                                f'.train.dataloader.dataset.output_dir::{out}/train'])
    id0 = dataset['ids'][0]
    pil = PIL.Image.open(root / 'images' / f'{id0:012d}.png').convert('RGB')
    props = torch.from_numpy(synth.proposals(*pil.size, 20, seed=2 * 7919 + 0))
    ro = ofe.objects_preprocess(pil, props, expand_mode='CONSTANT')
    p197 = vit.objects_surgery(vit.init_visual_params(0))
    o = torch.load(out / 'val' / f'{id0:012d}.pth')
    assert torch.equal(o['bboxes'], ro.bboxes.half())
    assert cos_ok(o['embeddings'], vit.normalize_half(vit.encode_objects(p197, ro.objects, ro.masks)))
    adaptive = torch.load(root / 'oake' / 'objects' / 'val' / f'{id0:012d}.pth') if (root / 'oake' / 'objects' / 'val').exists() else None
    if adaptive is not None:
        assert not torch.equal(adaptive['embeddings'], o['embeddings'])
