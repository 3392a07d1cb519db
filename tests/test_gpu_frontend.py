"""GPU tier: the sm_100a front end against PIL / torchvision (the reference's own libraries):
uint8 crops bit-exact, masks exact, and whole-pipeline embeddings against the CPU oracle."""
import numpy as np
import PIL.Image
import pytest
import torch
import torch.nn.functional as F
import torchvision.transforms as T

from oadp_b200 import frontend, synth
from oadp_b200.model import OakeEngine
from oadp_b200.pipeline import OakePipeline
from oracle import frontend as ofe
from oracle import vit

pytestmark = pytest.mark.gpu

WEIGHT_SEED = 1234


@pytest.fixture(scope='module')
def pipe(lib):
    return OakePipeline(OakeEngine(synth.visual_params(WEIGHT_SEED), 'cuda'))


def pil_clip_u8(img: PIL.Image.Image, box) -> np.ndarray:
    crop = img.crop(tuple(box))
    crop = T.CenterCrop(224)(T.Resize(224, interpolation=T.InterpolationMode.BICUBIC)(crop))
    return np.asarray(crop.convert('RGB'))


def test_object_crops_bit_exact(pipe):
    w, h = 640, 480
    arr = synth.image(w, h, 21)
    img = PIL.Image.fromarray(arr)
    props = synth.proposals(w, h, 120, seed=3)
    props[1] = [0.0, 0.0, 640.0, 480.0, 0.9]  # crop far larger than the image: zero padded, scale ~7
    props[2] = [300.0, 200.0, 304.5, 204.2, 0.9]  # tiny: 12 px crop upsampled 18x
    props[3] = [0.0, 0.0, 50.0, 30.0, 0.8]  # at the corner
    plan = frontend.objects_plan(props, (w, h))
    jobs = frontend.crop_jobs(0, w, h, plan.boxes_int, 1 << 21)
    got = pipe.debug_crops_u8([arr], jobs)
    assert got.shape[0] == plan.boxes_int.shape[0] > 100
    for i, box in enumerate(plan.expanded.tolist()):
        ref = pil_clip_u8(img, box)
        assert np.array_equal(got[i], ref), (i, box, np.abs(got[i].astype(int) - ref.astype(int)).max())


@pytest.mark.parametrize('wh', [(640, 427), (500, 375), (333, 517)])
def test_pyramid_level_bit_exact(pipe, wh):
    w, h = wh
    arr = synth.image(w, h, 5)
    img = PIL.Image.fromarray(arr)
    lw, lh = int(w / 1.5), int(h / 1.5)
    ref = np.asarray(img.resize((lw, lh)))
    job = frontend.level_job(0, w, h, 1 << 21, lw, lh)
    offs, img_bytes = pipe._place_images([arr])
    pipe._arena.reserve(img_bytes, (1 << 21) + lw * lh * 3)
    pipe._meta.reserve(1024)
    pipe._arena.host.numpy()[:arr.size] = arr.reshape(-1)
    pipe._meta.host.numpy()[:job.nbytes] = job.view(np.uint8).reshape(-1)
    pipe._arena.upload(img_bytes)
    pipe._meta.upload(job.nbytes)
    from oadp_b200 import binding
    p = pipe._arena.dev.data_ptr()
    binding.check(pipe.lib.oake_resize_u8(p, p, pipe._meta.dev.data_ptr(), 1, frontend.max_tiles(job),
                                          pipe._err.data_ptr(), pipe._stream()))
    torch.cuda.synchronize()
    assert int(pipe._err.item()) == 0
    got = pipe._arena.dev[1 << 21:(1 << 21) + lw * lh * 3].cpu().numpy().reshape(lh, lw, 3)
    assert np.array_equal(got, ref)


def test_masks_exact(pipe):
    w, h = 640, 480
    props = synth.proposals(w, h, 200, seed=8)
    plan = frontend.objects_plan(props, (w, h))
    n = plan.bboxes.shape[0]
    fg = torch.from_numpy(plan.foregrounds).cuda()
    box = torch.from_numpy(plan.expanded).cuda()
    masks = torch.empty(n, 1, 14, 14, device='cuda')
    from oadp_b200 import binding
    binding.check(pipe.lib.oake_object_masks(fg.data_ptr(), box.data_ptr(), n, masks.data_ptr(), pipe._stream()))
    torch.cuda.synchronize()
    ref = torch.cat([ofe.object_mask(tuple(f), tuple(b)) for f, b in
                     zip(plan.foregrounds.tolist(), plan.expanded.tolist())])
    assert torch.equal(masks.cpu(), ref)
    assert 0 < ref.mean() < 1


def check_emb(got, want, what):
    got, want = got.float(), want.float()
    cos = F.cosine_similarity(got, want, dim=-1)
    assert (1 - cos).max() < 1e-3, (what, (1 - cos).max().item())
    assert ((got - want).norm(dim=-1) / want.norm(dim=-1)).max() < 3e-2, what


def test_pipeline_globals_config0(pipe):
    """BASELINE.json configs[0]: 4 synthetic 640x480 images, ViT-B/32, feature parity."""
    imgs = [synth.image(640, 480, s) for s in range(4)]
    got = pipe.encode_globals(imgs)
    p = vit.init_visual_params(WEIGHT_SEED)
    px = torch.stack([ofe.globals_preprocess(PIL.Image.fromarray(a)) for a in imgs])
    want = vit.normalize_half(vit.encode_image(p, px))
    assert got[0].shape == (512, ) and got[0].dtype == torch.float16
    check_emb(torch.stack(got), want, 'globals')


def test_pipeline_blocks(pipe):
    imgs = [synth.image(640, 480, 7), synth.image(500, 375, 8), synth.image(300, 200, 9)]
    got = pipe.encode_blocks(imgs)
    p = vit.init_visual_params(WEIGHT_SEED)
    for g, a in zip(got, imgs):
        ref = ofe.blocks_preprocess(PIL.Image.fromarray(a))
        want = vit.normalize_half(vit.encode_image(p, ref.blocks))
        assert g['embeddings'].shape == want.shape
        check_emb(g['embeddings'], want, 'blocks')
        assert torch.equal(g['bboxes'], ref.bboxes.half())
    assert [g['embeddings'].shape[0] for g in got] == [27, 17, 1]


def test_pipeline_objects(pipe):
    w, h = 640, 427
    imgs = [synth.image(w, h, 31), synth.image(480, 640, 32)]
    props = [synth.proposals(w, h, 24, seed=1), synth.proposals(480, 640, 17, seed=2)]
    got = pipe.encode_objects(imgs, props)
    p197 = vit.objects_surgery(vit.init_visual_params(WEIGHT_SEED))
    for g, a, pr in zip(got, imgs, props):
        ref = ofe.objects_preprocess(PIL.Image.fromarray(a), torch.from_numpy(pr))
        want = vit.normalize_half(vit.encode_objects(p197, ref.objects, ref.masks))
        check_emb(g['embeddings'], want, 'objects')
        assert torch.equal(g['bboxes'], ref.bboxes.half())
        assert torch.equal(g['objectness'], ref.objectness.half())
    dry = pipe.encode_objects(imgs, props, dry_run=True)
    assert all(d['embeddings'].shape[0] <= 5 for d in dry)
    assert torch.equal(dry[0]['embeddings'], got[0]['embeddings'][:dry[0]['embeddings'].shape[0]])


def test_fused_front_end_is_the_unfused_one(pipe, monkeypatch):
    """oake_resize_to_patches (pixels straight into the tower's front-end matrix, zero border included) must give the
    embeddings of the two-step path (uint8 crops, then the matrix kernel) bit for bit: same table, same rounding."""
    from oadp_b200 import pipeline as pl
    imgs = synth.images(3, seed=4)
    props = [synth.proposals(im.shape[1], im.shape[0], 150, seed=40 + i) for i, im in enumerate(imgs)]
    props[0][1] = [0.0, 0.0, 640.0, 480.0, 0.9]  # expanded far beyond the image: BIG class, tiles outside the image
    got = {}
    for fused in (True, False):
        monkeypatch.setattr(pl, '_FUSED_FRONTEND', fused)
        objs = pipe.encode_objects(imgs, props)
        glob = pipe.encode_globals(imgs)
        got[fused] = (torch.cat([o['embeddings'] for o in objs]), torch.stack(glob))
    assert got[True][0].shape[0] > 400
    assert torch.equal(got[True][0], got[False][0])
    assert torch.equal(got[True][1], got[False][1])
    # the fused path in passes of a few jobs (crop index = pass offset + job index) and without coefficient tables
    monkeypatch.setattr(pl, '_FUSED_FRONTEND', True)
    for knob in ('OAKE_RESIZE_SLICE', 'OAKE_RESIZE_NO_TABLES'):
        monkeypatch.setenv(knob, '5' if knob.endswith('SLICE') else '1')
        objs = pipe.encode_objects(imgs, props)
        assert torch.equal(torch.cat([o['embeddings'] for o in objs]), got[True][0])
        monkeypatch.delenv(knob)


def test_resize_paths_agree(pipe, monkeypatch):
    """The three ways a crop can travel through oake_resize_u8 give the same bytes: per-job coefficient tables (FAST and
    BIG classes), per-tile generation when the scratch has no room for tables, and passes of a few jobs at a time."""
    w, h = 640, 480
    arr = synth.image(w, h, 33)
    props = synth.proposals(w, h, 90, seed=8)
    props[1] = [0.0, 0.0, 640.0, 480.0, 0.9]   # BIG class (scale ~7), mostly outside the image
    props[2] = [100.0, 50.0, 600.0, 470.0, 0.9]  # BIG class, mostly inside
    plan = frontend.objects_plan(props, (w, h))
    jobs = frontend.crop_jobs(0, w, h, plan.boxes_int, 1 << 21)
    ref = pipe.debug_crops_u8([arr], jobs)
    monkeypatch.setenv('OAKE_RESIZE_NO_TABLES', '1')
    assert np.array_equal(pipe.debug_crops_u8([arr], jobs), ref)
    monkeypatch.delenv('OAKE_RESIZE_NO_TABLES')
    monkeypatch.setenv('OAKE_RESIZE_SLICE', '7')
    assert np.array_equal(pipe.debug_crops_u8([arr], jobs), ref)
