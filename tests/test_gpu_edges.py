"""GPU tier: edge cases of the hot path -- paths the mainstream parity tests do not reach: soft
(non 0/1) side-stream masks, large-magnitude attention scores, empty / single-crop / chunk-boundary
batches, crops that lie outside the image, and the resize kernel's documented limits."""
import numpy as np
import PIL.Image
import pytest
import torch
import torchvision.transforms as T

from oadp_b200 import binding, frontend, synth
from oadp_b200.model import OakeEngine, OakeModel
from oadp_b200.pipeline import OakePipeline
from oracle import vit

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def act_dtype():
    return torch.float16 if binding.act_dtype_name() == 'f16' else torch.bfloat16


def stream():
    return torch.cuda.current_stream().cuda_stream


def ref_side(qkv, mask, B, P):
    """fp32 reference of the side row only (objects.py:224-247) on rounded inputs."""
    W = 768
    q, k, v = qkv.float().split(W, dim=-1)
    ys = slice(B * P + B, B * P + 2 * B)
    qy = q[ys].reshape(B, 1, 12, 64).transpose(1, 2)
    ky = torch.cat([k[:B * P].reshape(B, P, 12, 64), k[ys].reshape(B, 1, 12, 64)], 1).transpose(1, 2)
    vy = torch.cat([v[:B * P].reshape(B, P, 12, 64), v[ys].reshape(B, 1, 12, 64)], 1).transpose(1, 2)
    bias = torch.cat([mask * -100.0, mask.new_zeros(B, 1)], 1)[:, None, None, :]
    oy = torch.softmax(qy @ ky.transpose(-1, -2) / 8 + bias, -1) @ vy
    return oy.transpose(1, 2).reshape(B, W)


@pytest.mark.parametrize('side_only', [0, 1])
def test_attention_soft_masks(lib, side_only):
    """Masks that are not 0/1: the reference's `attn_mask *= -100` (objects.py:212) holds for any
    value, so the persistent kernel must leave its bit-mask fast path."""
    B, P = 40, 196
    g = torch.Generator(device=DEV).manual_seed(21)
    R = B * (P + 2)
    qkv = (torch.randn(R, 2304, device=DEV, generator=g) * 1.5).to(act_dtype())
    mask = torch.rand(B, P, device=DEV, generator=g) * 0.05  # bias in [-5, 0]: every key still matters
    mask[0] = (torch.rand(P, device=DEV, generator=g) > 0.5).float()  # one binary crop among soft ones
    out = torch.zeros(R, 768, device=DEV, dtype=act_dtype())
    binding.check(lib.oake_test_attention_side(qkv.data_ptr(), mask.data_ptr(), out.data_ptr(), B, P, side_only, stream()))
    torch.cuda.synchronize()
    ys = slice(B * P + B, R)
    assert (out[ys].float() - ref_side(qkv, mask, B, P)).abs().max() < 6e-3


def test_attention_large_scores(lib):
    """Scores of +-60 nats: softmax must subtract the row maximum before exp2 / fp16 packing."""
    B, P = 13, 196
    g = torch.Generator(device=DEV).manual_seed(22)
    R = B * (P + 2)
    qkv = (torch.randn(R, 2304, device=DEV, generator=g) * 1.5)
    qkv[:, :1536] *= 4.0  # q and k: scores x 16
    qkv = qkv.to(act_dtype())
    mask = (torch.rand(B, P, device=DEV, generator=g) > 0.5).float()
    out = torch.zeros(R, 768, device=DEV, dtype=act_dtype())
    binding.check(lib.oake_test_attention_side(qkv.data_ptr(), mask.data_ptr(), out.data_ptr(), B, P, 0, stream()))
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    q, k, v = qkv.float().split(768, dim=-1)

    def gather(t):
        return torch.cat([t[:B * P].reshape(B, P, 12, 64), t[B * P:B * P + B].reshape(B, 1, 12, 64)], 1).transpose(1, 2)

    o = torch.softmax(gather(q) @ gather(k).transpose(-1, -2) / 8, -1) @ gather(v)
    o = o.transpose(1, 2).reshape(B, P + 1, 768)
    assert (out[:B * P].float() - o[:, :P].reshape(B * P, 768)).abs().max() < 2e-2
    assert (out[B * P + B:].float() - ref_side(qkv, mask, B, P)).abs().max() < 2e-2


@pytest.fixture(scope='module')
def pipe(lib):
    return OakePipeline(OakeEngine(synth.visual_params(5, layers=2), 'cuda'))


def test_empty_and_degenerate_proposals(pipe):
    """No proposal survives `min_wh` (the reference would raise in torch.stack([])):
    empty, correctly shaped outputs -- alone and next to a non-empty image."""
    img = synth.image(320, 240, 1)
    none = np.array([[10, 10, 12, 50, 0.9], [5, 5, 100, 7, 0.8]], dtype=np.float32)  # w < 4, h < 4
    some = synth.proposals(320, 240, 6, seed=1)
    got = pipe.encode_objects([img, img, img], [none, some, np.zeros((0, 5), np.float32)])
    assert got[0]['embeddings'].shape == (0, 512) and got[0]['bboxes'].shape == (0, 4)
    assert got[0]['objectness'].shape == (0, 1) and got[2]['embeddings'].shape == (0, 512)
    alone = pipe.encode_objects([img], [some])[0]
    assert got[1]['embeddings'].shape[0] > 0 and torch.equal(got[1]['embeddings'], alone['embeddings'])
    assert pipe.encode_objects([], []) == []
    assert pipe.encode_globals([]) == [] and pipe.encode_blocks([]) == []


def test_image_smaller_than_a_block(pipe):
    """Images under 224 px on a side have no block grid: only the global crop (blocks.py:40-43)."""
    got = pipe.encode_blocks([synth.image(200, 150, 2), synth.image(224, 224, 3)])
    assert got[0]['embeddings'].shape == (1, 512) and got[0]['bboxes'].tolist() == [[25.0, 0.0, 150.0, 150.0]]
    assert got[1]['embeddings'].shape == (2, 512)  # global crop + the single 224 x 224 block


def test_chunk_boundary_rows_identical(lib):
    """One crop more than a tower chunk (478 for T197): the tail chunk of 1 gives the same bits."""
    p = vit.init_visual_params(3, layers=1)
    m = OakeModel(p, 'cuda')
    m.for_objects()
    g = torch.Generator().manual_seed(8)
    n = 479
    px = torch.randn(n, 3, 224, 224, generator=g).cuda()
    mk = (torch.rand(n, 1, 14, 14, generator=g) > 0.5).float().cuda()
    full = m.embed(px, mk)
    assert full.shape == (n, 512) and torch.isfinite(full.float()).all()
    assert torch.equal(full[478:], m.embed(px[478:], mk[478:]))
    assert torch.equal(full[:3], m.embed(px[:3], mk[:3]))


def test_crop_outside_the_image_and_scale_limit(pipe):
    """PIL `crop` pads with zeros outside the image; scale factors above the kernel limit are
    reported through the error flag instead of producing wrong pixels."""
    w, h = 200, 160
    arr = synth.image(w, h, 4)
    img = PIL.Image.fromarray(arr)
    boxes = np.array([[-300, -300, -100, -100], [150, 100, 350, 300], [-20, -10, 30, 40]], dtype=np.int64)
    jobs = frontend.crop_jobs(0, w, h, boxes, 1 << 21)
    got = pipe.debug_crops_u8([arr], jobs)
    for i, box in enumerate(boxes.tolist()):
        crop = T.CenterCrop(224)(T.Resize(224, interpolation=T.InterpolationMode.BICUBIC)(img.crop(tuple(box))))
        assert np.array_equal(got[i], np.asarray(crop.convert('RGB'))), i
    assert not got[0].any()
    big = synth.image(3000, 2900, 5)  # 2900 / 224 > 11: beyond the documented limit
    with pytest.raises(ValueError):  # refused on the host ...
        frontend.crop_jobs(0, 3000, 2900, np.array([[0, 0, 3000, 2900]], dtype=np.int64), 3000 * 2900 * 3 + 4096)
    jobs = frontend.crop_jobs(0, 3000, 2900, np.array([[0, 0, 2400, 2400]], dtype=np.int64), 3000 * 2900 * 3 + 4096)
    jobs['box_w'], jobs['box_h'] = 2900, 2900  # ... and, if a caller skips that check, flagged by the kernel
    with pytest.raises(binding.OakeError):
        pipe.debug_crops_u8([big], jobs)


def test_full_size_step_properties(lib):
    """BASELINE-size step (8 COCO-shaped images, 300 proposals each, full 12-layer tower) checked
    through size-independent properties: unit-norm fp16 rows, bit-identical results when the same
    images are encoded in different batch compositions and on a second run, and an object crop
    that equals the whole image giving the same pixels as the global crop path."""
    pipe = OakePipeline(OakeEngine(synth.visual_params(0), 'cuda'))
    sizes = [(640, 480), (640, 427), (480, 640), (427, 640), (500, 375), (640, 640), (612, 612), (640, 426)]
    imgs = [synth.image(w, h, 100 + k) for k, (w, h) in enumerate(sizes)]
    props = [synth.proposals(w, h, 300, seed=200 + k) for k, (w, h) in enumerate(sizes)]
    full = pipe.encode_objects(imgs, props)
    n = sum(o['embeddings'].shape[0] for o in full)
    assert 2300 < n <= 2400  # ~2 % of the proposals are degenerate
    for o in full:
        e = o['embeddings'].float()
        assert torch.isfinite(e).all() and ((e.norm(dim=-1) - 1).abs() < 2e-3).all()
    halves = pipe.encode_objects(imgs[:3], props[:3]) + pipe.encode_objects(imgs[3:], props[3:])
    again = pipe.encode_objects(list(reversed(imgs)), list(reversed(props)))[::-1]
    for a, b, c in zip(full, halves, again):
        assert torch.equal(a['embeddings'], b['embeddings']) and torch.equal(a['embeddings'], c['embeddings'])
        assert torch.equal(a['bboxes'], b['bboxes']) and torch.equal(a['objectness'], c['objectness'])
    blocks = pipe.encode_blocks(imgs)
    assert [b['embeddings'].shape[0] for b in blocks] == [27, 22, 27, 22, 17, 39, 39, 22]
    g = pipe.encode_globals(imgs)
    for gi, bi in zip(g, blocks):
        assert torch.equal(gi, bi['embeddings'][0])  # block 0 is the global crop (blocks.py:95)
