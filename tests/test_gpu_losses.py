"""GPU tier: distillation-side loss kernels against the reference-generated fixture
(oadp/base/losses.py executed by tests/golden/make_ref_golden.py) and against autograd through the
CPU oracle at the configs' sizes."""
import pathlib
import sys

import pytest
import torch

from oadp_b200.dp import losses as L
from oracle import losses as ol

sys.path.insert(0, str(pathlib.Path(__file__).parent / 'golden'))
import make_ref_golden as mk  # noqa: E402  (helpers only)

pytestmark = pytest.mark.gpu
REF = torch.load(pathlib.Path(__file__).parent / 'golden' / 'ref_golden.pt', weights_only=False)


def close(got, want, rtol):
    got, want = got.detach().cpu().double(), want.detach().cpu().double()
    return (got - want).abs().max() <= rtol * want.abs().max()


@pytest.mark.parametrize('name,kw', [('asl_configs', dict(gamma_neg=4, gamma_pos=0)), ('asl_defaults', dict()),
                                     ('asl_sum_w16', dict(gamma_neg=4, gamma_pos=0, reduction='sum', weight=16.0))])
def test_asymmetric_loss_vs_reference(lib, name, kw):
    probs, targets, _, _ = mk.loss_inputs()
    x = probs.cuda().requires_grad_(True)
    loss = L.AsymmetricLoss(**kw)(x, targets.cuda())
    loss.backward()
    want = REF['losses'][name]
    assert close(loss, want['loss'], 2e-6)
    assert close(x.grad, want['grad'], 2e-5)


@pytest.mark.parametrize('name,kw', [('rkd', dict()), ('rkd_w8', dict(weight=8.0))])
def test_rkd_loss_vs_reference(lib, name, kw):
    _, _, student, teacher = mk.loss_inputs()
    s = student.cuda().requires_grad_(True)
    loss = L.RKDLoss(**kw)(s, teacher.cuda())
    loss.backward()
    want = REF['losses'][name]
    assert close(loss, want['loss'], 2e-5)
    assert close(s.grad, want['grad'], 2e-5)


@pytest.mark.parametrize('kind,reduction,weight,n', [('l1', 'mean', 256.0, 600), ('l1', 'mean', 128.0, 54),
                                                     ('mse', 'sum', 0.5, 2), ('mse', 'mean', 1.0, 1000)])
def test_l1_mse_vs_oracle(lib, kind, reduction, weight, n):
    """configs/dp/models: L1 x 256 on objects, L1 x 128 on blocks, MSE(sum) x 0.5 on the global rows."""
    g = torch.Generator().manual_seed(n)
    s = torch.nn.functional.normalize(torch.randn(n, 512, generator=g), dim=-1)
    t = torch.nn.functional.normalize(torch.randn(n, 512, generator=g), dim=-1).half().float()
    s[0, :7] = t[0, :7]  # exact ties: sign(0) = 0
    a = s.clone().requires_grad_(True)
    want = (ol.l1_loss if kind == 'l1' else ol.mse_loss)(a, t, reduction, weight)
    want.backward()
    b = s.cuda().requires_grad_(True)
    mod = (L.L1Loss if kind == 'l1' else L.MSELoss)(reduction=reduction, weight=weight)
    got = mod(b, t.cuda())
    (got * 3.0).backward()  # upstream gradient other than 1
    assert close(got, want, 1e-5)
    assert close(b.grad, a.grad * 3.0, 1e-5)


def test_rkd_block_sized_batch_and_warmup(lib):
    g = torch.Generator().manual_seed(5)
    s = torch.nn.functional.normalize(torch.randn(2, 39, 512, generator=g), dim=-1)
    t = torch.nn.functional.normalize(torch.randn(2, 39, 512, generator=g), dim=-1)
    a = s.clone().requires_grad_(True)
    want = ol.rkd_loss(a, t, 'mean', 8.0 * 0.25)
    want.backward()
    mod = L.RKDLoss(weight=dict(type='WarmupScheduler', gain=8, end=200))
    mod.step(50)  # a quarter of the warm-up
    b = s.cuda().requires_grad_(True)
    got = mod(b, t.cuda())
    got.backward()
    assert close(got, want, 2e-5) and close(b.grad, a.grad, 2e-5)
    assert not L.L1Loss()(s.cuda(), t.cuda()).requires_grad  # no gradient requested: loss only


def test_losses_reject_cpu_tensors(lib):
    from oadp_b200 import binding
    with pytest.raises(binding.OakeError):
        L.L1Loss()(torch.zeros(2, 4), torch.zeros(2, 4))
