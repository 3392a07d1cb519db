"""GPU tier: ViLD ensemble scoring kernel against the line-by-line restatement of
oadp/dp/roi_heads.py:93-112 in oracle/classifier.py."""
import pytest
import torch

from oadp_b200.dp import roi_heads
from oracle import classifier as ocls

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('n,num_bases,num_all', [(1000, 48, 65), (1000, 866, 1203), (1, 48, 65), (37, 3, 5)])
def test_vild_ensemble(lib, n, num_bases, num_all):
    g = torch.Generator().manual_seed(n + num_all)
    k1 = num_all + 1
    bbox = torch.randn(n, k1, generator=g) * 4
    obj = torch.randn(n, k1, generator=g) * 4
    obj[:, -1] = float('-inf')  # ObjectMixin.forward, bbox_heads.py:57-60
    lam = roi_heads.ensemble_lambda(num_bases, num_all)
    want = ocls.vild_ensemble(bbox.double(), obj.double(), lam.double())
    got = roi_heads.vild_ensemble(bbox.cuda(), obj.cuda(), lam.cuda()).cpu()
    assert got.shape == want.shape
    assert torch.isfinite(got).all()
    assert (got.double() - want).abs().max() < 2e-4  # fp32 exp/pow/log on values in [log 1e-30, 0]
    # the scores of a row sum to one (background = 1 - foreground)
    assert (got.exp().sum(-1) - 1).abs().max() < 1e-4


def test_vild_ensemble_padded_pitch(lib):
    """Logits straight out of the classifier are padded to a multiple of 128 columns."""
    g = torch.Generator().manual_seed(3)
    n, k1 = 50, 66
    bbox = torch.randn(n, 128, generator=g).cuda()
    obj = torch.randn(n, 128, generator=g).cuda()
    lam = roi_heads.ensemble_lambda(48, 65).cuda()
    got = roi_heads.vild_ensemble(bbox[:, :k1], obj[:, :k1], lam)
    want = ocls.vild_ensemble(bbox[:, :k1].cpu().double(), obj[:, :k1].cpu().double(), lam.cpu().double())
    assert (got.cpu().double() - want).abs().max() < 2e-4


def test_vild_ensemble_rejects_cpu(lib):
    with pytest.raises(RuntimeError):
        roi_heads.vild_ensemble(torch.zeros(2, 3), torch.zeros(2, 3), torch.ones(3) / 3)
