"""CPU tier: the product's host geometry (oadp_b200/frontend.py) against the PIL/torchvision oracle
(oracle/frontend.py) and against known answers probed from the reference algorithm."""
import numpy as np
import PIL.Image
import pytest
import torch
import torchvision.transforms as T

from oadp_b200 import frontend, synth
from oracle import frontend as ofe


@pytest.mark.parametrize('length,expect', [(100, []), (224, [0]), (225, [0, 1]), (336, [0, 112]), (337, [0, 57, 113]),
                                           (480, [0, 86, 171, 256]), (640, None)])
def test_partition_matches_oracle(length, expect):
    got = frontend.partition(length)
    assert got == ofe.partition(length)
    if expect is not None:
        assert got == expect
    if got:
        assert got[0] == 0 and got[-1] == length - 224
        assert all(b - a <= 112 for a, b in zip(got, got[1:]))


@pytest.mark.parametrize('wh,count', [((640, 480), 27), ((640, 427), 22), ((500, 375), 17), ((640, 640), 39),
                                      ((200, 300), 1), ((224, 224), 2)])
def test_blocks_plan_counts_and_bboxes(wh, count):
    w, h = wh
    plan = frontend.blocks_plan(w, h)
    assert 1 + len(plan.cells) == count == ofe.crops_per_image_blocks(w, h)
    img = PIL.Image.fromarray(synth.image(w, h, 3))
    ref = ofe.blocks_preprocess(img)
    assert ref.bboxes.shape[0] == count
    assert torch.equal(torch.tensor(plan.bboxes, dtype=torch.float32), ref.bboxes)
    # quirk: row 0 is (x0, y0, side, side), not xyxy
    side = min(w, h)
    assert plan.bboxes[0, 2] == side and plan.bboxes[0, 3] == side


def test_clip_window_matches_torchvision():
    rng = np.random.default_rng(0)
    for _ in range(300):
        cw, ch = int(rng.integers(5, 2000)), int(rng.integers(5, 2000))
        if max(cw, ch) / min(cw, ch) > 30:
            continue
        ow, oh, wx, wy = [int(v[0]) for v in frontend.clip_resize_window(np.array([cw]), np.array([ch]))]
        im = PIL.Image.new('RGB', (cw, ch))
        r = T.Resize(224, interpolation=T.InterpolationMode.BICUBIC)(im)
        assert r.size == (ow, oh), (cw, ch)
        # CenterCrop offsets: torchvision int(round((n - 224) / 2.0))
        assert wx == int(round((ow - 224) / 2.0)) and wy == int(round((oh - 224) / 2.0))
    # half-to-even cases
    assert [int(v[0]) for v in frontend.clip_resize_window(np.array([224]), np.array([225]))][2:] == [0, 0]
    assert [int(v[0]) for v in frontend.clip_resize_window(np.array([224]), np.array([227]))][2:] == [0, 2]


def test_objects_plan_matches_oracle():
    w, h = 640, 480
    props = synth.proposals(w, h, 300, seed=11)
    props[5] = [630.0, 470.0, 639.5, 479.0, 0.5]  # near the corner: pushed inside
    props[6] = [0.0, 0.0, 640.0, 480.0, 0.4]  # square larger than the image: not moved, zero padded
    props[7] = [10.0, 10.0, 14.0, 13.9, 0.3]  # h < 4: filtered
    props[8] = [10.0, 10.0, 14.0, 14.0, 0.3]  # exactly 4x4: kept (inclusive)
    plan = frontend.objects_plan(props, (w, h))
    img = PIL.Image.fromarray(synth.image(w, h, 4))
    ref = ofe.objects_preprocess(img, torch.from_numpy(props[:40]))
    plan40 = frontend.objects_plan(props[:40], (w, h))
    assert np.array_equal(plan40.bboxes, ref.bboxes.numpy())
    assert np.array_equal(plan40.objectness, ref.objectness.numpy())
    assert np.array_equal(plan40.expanded, ref.expanded.numpy())
    keep = frontend.objects_plan(props[7:9], (w, h))
    assert keep.bboxes.shape[0] == 1
    e = plan.expanded
    side = e[:, 2] - e[:, 0]
    area = (plan.bboxes[:, 2] - plan.bboxes[:, 0]) * (plan.bboxes[:, 3] - plan.bboxes[:, 1])
    assert np.allclose(side, np.sqrt(8 * area), rtol=1e-5)
    fits = (side <= w) & (side <= h)
    assert (e[fits, 0] >= -1e-3).all() and (e[fits, 2] <= w + 1e-3).all()
    # PIL rounds half to even
    assert np.array_equal(plan.boxes_int, np.array([[int(round(float(v))) for v in row] for row in e]))
    # dry run keeps at most the valid ones among the first five proposals
    assert frontend.objects_plan(props, (w, h), dry_run=True).bboxes.shape[0] <= 5


def test_expand_rejects_nothing_and_handles_empty():
    plan = frontend.objects_plan(np.zeros((0, 5), np.float32), (640, 480))
    assert plan.bboxes.shape == (0, 4) and plan.boxes_int.shape == (0, 4)


def test_job_records_layout():
    jobs = frontend.crop_jobs(4096, 640, 480, np.array([[10, 20, 110, 121], [-5, -5, 700, 700]]), 1 << 20)
    assert jobs.dtype.itemsize == 72 and jobs.shape == (2, )
    assert jobs['out_w'][0] == 224 and jobs['out_h'][0] == int(224 * 101 / 100)
    assert jobs['dst_off'][1] - jobs['dst_off'][0] == 224 * 224 * 3
    assert frontend.max_tiles(jobs) == 49
    with pytest.raises(ValueError):
        frontend.crop_jobs(0, 640, 480, np.array([[0, 0, 3000, 3000]]), 0)
    lvl = frontend.level_job(0, 640, 480, 999, 426, 320)
    assert frontend.max_tiles(lvl) == 14 * 10


def test_fused_stage_predicate():
    """Which plans may skip the uint8 crops (oake_resize_to_patches): one stage whose jobs are the crops, whole windows."""
    from oadp_b200 import frontend as fe
    boxes = np.array([[0, 0, 640, 480], [10, 20, 200, 300]], dtype=np.int64)
    jobs = fe.crop_jobs(0, 640, 480, boxes, 1 << 20)
    crops = np.zeros(2, dtype=fe.CROP_SRC)
    crops['off'] = jobs['dst_off']
    crops['pitch_px'] = fe.SIZE
    assert fe.fused_stage([jobs], crops)
    assert not fe.fused_stage([jobs, jobs], crops)  # pyramid stages (blocks)
    assert not fe.fused_stage([jobs[:1]], crops)  # crops that are not resize outputs (windows into a level)
    moved = crops.copy()
    moved['off'][1] += 3
    assert not fe.fused_stage([jobs], moved)
    level = fe.level_job(0, 640, 480, 1 << 20, 426, 320)
    one = np.zeros(1, dtype=fe.CROP_SRC)
    one['off'] = level['dst_off']
    one['pitch_px'] = 426
    assert not fe.fused_stage([level], one)  # a whole-image level is not a 224 x 224 window
    assert not fe.fused_stage([jobs[:0]], crops[:0])
