"""GPU tier: re-softmax + multiclass NMS (second half) against the CPU restatement of mmdet's
`multiclass_nms` built on torchvision.ops.nms (oracle/nms.py).  Index work: the kept candidates, their order and
the labels must be IDENTICAL; boxes / scores are copies, so they are bit-exact too."""
import pytest
import torch

from oadp_b200.dp import nms as onms
from oadp_b200.dp import roi_heads
from oracle import classifier as ocls
from oracle import nms as ref

pytestmark = pytest.mark.gpu


def proposals(n, seed, size=800.0, clustered=True):
    g = torch.Generator().manual_seed(seed)
    if clustered:  # RPN-like: many boxes around a few objects, so that suppression really happens
        centres = torch.rand(12, 2, generator=g) * size
        c = centres[torch.randint(0, 12, (n, ), generator=g)] + torch.randn(n, 2, generator=g) * 12
        wh = (40 + torch.rand(n, 2, generator=g) * 120) * (1 + 0.1 * torch.randn(n, 2, generator=g)).abs()
    else:
        c = torch.rand(n, 2, generator=g) * size
        wh = 10 + torch.rand(n, 2, generator=g) * 200
    return torch.cat([c - wh / 2, c + wh / 2], 1).clamp(0, size)


@pytest.mark.parametrize('n,k,thr,max_num', [(1000, 65, 0.0, 300), (1000, 1203, 0.0, 300), (1000, 65, 0.05, 100),
                                             (37, 5, 0.0, -1), (1, 3, 0.0, 300), (129, 2, 0.3, 10)])
def test_multiclass_nms_matches_the_restatement(lib, n, k, thr, max_num):
    g = torch.Generator().manual_seed(n + k)
    boxes = proposals(n, n * 7 + k)
    scores = torch.softmax(torch.randn(n, k + 1, generator=g) * 3, dim=-1)
    want = ref.multiclass_nms(boxes, scores, thr, 0.5, max_num)
    got = onms.multiclass_nms(boxes.cuda(), scores.cuda(), thr, dict(type='nms', iou_threshold=0.5), max_num, return_inds=True)
    assert torch.equal(got[2].cpu(), want[2]), (got[2][:10], want[2][:10])
    assert torch.equal(got[1].cpu(), want[1]) and torch.equal(got[0].cpu(), want[0])
    if max_num > 0:
        assert got[0].shape[0] <= max_num
    assert (got[0][1:, 4] <= got[0][:-1, 4]).all()  # descending scores


def test_edges(lib):
    cfg = dict(type='nms', iou_threshold=0.5)
    # nothing above the threshold
    d, l = onms.multiclass_nms(proposals(10, 1).cuda(), torch.full((10, 4), 0.01).cuda(), 0.5, cfg, 100)
    assert d.shape == (0, 5) and l.shape == (0, )
    # no RoIs at all
    d, l = onms.multiclass_nms(torch.zeros(0, 4).cuda(), torch.zeros(0, 4).cuda(), 0.0, cfg, 100)
    assert d.shape == (0, 5) and l.dtype == torch.long
    # identical boxes: one survivor per class, the best-scoring one
    boxes = torch.tensor([[10., 10., 50., 50.]]).repeat(6, 1)
    scores = torch.tensor([[.1, .6, .3], [.5, .2, .3], [.2, .7, .1], [.4, .1, .5], [.3, .3, .4], [.05, .05, .9]])
    d, l = onms.multiclass_nms(boxes.cuda(), scores.cuda(), 0.0, cfg, -1)
    assert l.tolist() == [1, 0] and d[:, 4].tolist() == pytest.approx([0.7, 0.5])
    # IoU exactly at the threshold does not suppress (strict >), degenerate (zero-area) boxes never overlap
    boxes = torch.tensor([[0., 0., 10., 10.], [0., 0., 10., 20.], [5., 5., 5., 5.], [5., 5., 5., 5.]])  # IoU(0,1) = 0.5
    scores = torch.tensor([[.9, .1], [.8, .2], [.7, .3], [.6, .4]])
    d, l, idx = onms.multiclass_nms(boxes.cuda(), scores.cuda(), 0.0, cfg, -1, return_inds=True)
    want = ref.multiclass_nms(boxes, scores, 0.0, 0.5, -1)
    assert torch.equal(idx.cpu(), want[2]) and idx.numel() == 4
    # a NaN score is not a candidate; class-specific boxes and CPU tensors are refused
    scores[1, 0] = float('nan')
    d, l = onms.multiclass_nms(boxes.cuda(), scores.cuda(), 0.0, cfg, -1)
    assert d.shape[0] == 3 and torch.isfinite(d).all()
    with pytest.raises(NotImplementedError):
        onms.multiclass_nms(torch.zeros(4, 8).cuda(), scores.cuda(), 0.0, cfg)
    with pytest.raises(RuntimeError):
        onms.multiclass_nms(boxes, scores, 0.0, cfg)


def test_ensemble_tail_end_to_end(lib):
    """logits of the two heads -> vild_ensemble -> softmax again -> NMS -> 300 detections, against the oracle chain
    (roi_heads.py:93-112, then mmdet get_bboxes)."""
    n, num_bases, num_all = 1000, 48, 65
    g = torch.Generator().manual_seed(5)
    bbox = torch.randn(n, num_all + 1, generator=g) * 4
    obj = torch.randn(n, num_all + 1, generator=g) * 4
    obj[:, -1] = float('-inf')
    lam = roi_heads.ensemble_lambda(num_bases, num_all)
    boxes = proposals(n, 11)
    cls_score = roi_heads.vild_ensemble(bbox.cuda(), obj.cuda(), lam.cuda())
    probs = onms.softmax_rows(cls_score)
    want_probs = ocls.vild_ensemble(bbox.double(), obj.double(), lam.double()).softmax(-1)
    assert (probs.cpu().double() - want_probs).abs().max() < 1e-5
    dets, labels = onms.ensemble_detections(cls_score, boxes.cuda(), 0.0, dict(type='nms', iou_threshold=0.5), 300)
    assert dets.shape == (300, 5) and labels.shape == (300, ) and int(labels.max()) < num_all
    # the same scores through the restatement give the same detections (scores taken from the GPU: the sort is
    # sensitive to the last bit, which is the softmax's, not the NMS's)
    want = ref.multiclass_nms(boxes, probs.cpu(), 0.0, 0.5, 300)
    assert torch.equal(labels.cpu(), want[1]) and torch.equal(dets.cpu(), want[0])
