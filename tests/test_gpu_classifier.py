"""GPU tier: the cosine classifier (forward + backward, through the C ABI via the nn.Modules)
against the CPU oracle and its autograd gradients.  fp16 tensor-core operands, fp32 accumulation:
tolerances are relative to the magnitude of each tensor."""
import pytest
import torch
import torch.nn.functional as F

from oadp_b200.dp import categories
from oadp_b200.dp import classifiers as C
from oracle import classifier as oc

pytestmark = pytest.mark.gpu
BWD_TOL = 5e-3  # gradients: fp16 tensor-core operands rescaled by a power of two of their max-abs, fp32 accumulation


def make_prompts(tmp_path, names, seed=0):
    g = torch.Generator().manual_seed(seed)
    emb = F.normalize(torch.randn(len(names), 512, generator=g)) * 0.8
    path = tmp_path / 'prompts.pth'
    torch.save(dict(names=list(names), embeddings=emb, scaler=torch.tensor([50.0]), bias=torch.tensor([3.0])), path)
    return str(path)


def rel_err(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


@pytest.fixture
def coco_globals(lib):
    categories.Globals.categories = categories.coco
    categories.Globals.training = False
    yield
    categories.Globals.training = False


@pytest.mark.parametrize('cls_name,out_features,in_features,n', [('Classifier', 66, 1024, 1000),
                                                                ('ViLDClassifier', 66, 1024, 515),
                                                                ('BaseClassifier', 65, 256, 3)])
def test_forward_matches_oracle(tmp_path, coco_globals, cls_name, out_features, in_features, n):
    path = make_prompts(tmp_path, sorted(categories.coco.all_))
    clf = getattr(C, cls_name)(prompts=path, in_features=in_features, out_features=out_features).cuda()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(n, in_features, generator=g)
    captured = []
    clf._linear.register_forward_hook(lambda m, i, o: captured.append(o))  # what the todd distiller does
    for training in (False, True):
        categories.Globals.training = training
        captured.clear()
        y = clf(x.cuda())
        w, b = clf._linear.weight.detach().cpu(), clf._linear.bias.detach().cpu()
        bg = clf._bg_embedding.detach().cpu() if clf._bg_embedding is not None else None
        text = clf._embeddings.cpu()
        if cls_name == 'Classifier':
            want, h = oc.classifier_forward(x, w, b, text, bg, training, 48, 65, 50.0, 3.0)
        elif cls_name == 'ViLDClassifier':
            want, h = oc.vild_forward(x, w, b, text, bg, training, 48, 65)
        else:
            want, h = oc.base_forward(x, w, b, text, bg, training, 48, 65)
        assert y.shape == want.shape == (n, out_features)
        inf = torch.isinf(want)
        assert torch.equal(torch.isinf(y.cpu()), inf) and (y.cpu()[inf] < 0).all()
        assert rel_err(y.cpu()[~inf], want[~inf]) < 3e-3
        assert len(captured) == 1 and captured[0].shape == (n, 512)
        assert rel_err(captured[0], h) < 1e-3
        assert torch.allclose(captured[0].norm(dim=-1).cpu(), torch.ones(n), atol=1e-4)
    # ObjectMixin writes -inf into the result in place (bbox_heads.py:59)
    y[:, -1] = float('-inf')


def test_lvis_sized_head(tmp_path, lib):
    names = [f'cat{i:04d}' for i in range(1203)]
    categories.Globals.categories = categories.Categories(names[:866], names[866:])
    categories.Globals.training = True
    try:
        path = make_prompts(tmp_path, names)
        clf = C.ViLDClassifier(prompts=path, in_features=1024, out_features=1204, scaler=dict(train=0.01, val=0.007)).cuda()
        x = torch.randn(700, 1024, generator=torch.Generator().manual_seed(5))
        y = clf(x.cuda()).cpu()
        want, _ = oc.vild_forward(x, clf._linear.weight.detach().cpu(), clf._linear.bias.detach().cpu(),
                                  clf._embeddings.cpu(), clf._bg_embedding.detach().cpu(), True, 866, 1203, 0.01, 0.007)
        assert y.shape == (700, 1204) and torch.isinf(y[:, 866:1203]).all()
        fin = ~torch.isinf(want)
        assert rel_err(y[fin], want[fin]) < 3e-3
    finally:
        categories.Globals.categories = categories.coco
        categories.Globals.training = False


def test_backward_matches_autograd(tmp_path, coco_globals):
    path = make_prompts(tmp_path, sorted(categories.coco.all_))
    n, in_f = 300, 1024
    clf = C.Classifier(prompts=path, in_features=in_f, out_features=66).cuda()
    g = torch.Generator().manual_seed(9)
    x = torch.randn(n, in_f, generator=g)
    labels = torch.randint(0, 48, (n, ), generator=g)
    labels[::7] = 65  # some background
    target = F.normalize(torch.randn(n, 512, generator=g))  # cached CLIP features (distillation target)
    categories.Globals.training = True

    xs = x.cuda().requires_grad_(True)
    hooked = []
    clf._linear.register_forward_hook(lambda m, i, o: hooked.append(o))
    y = clf(xs)
    loss = F.cross_entropy(y, labels.cuda()) + 256 * F.l1_loss(hooked[0], target.cuda())
    loss.backward()

    w = clf._linear.weight.detach().cpu().requires_grad_(True)
    b = clf._linear.bias.detach().cpu().requires_grad_(True)
    bg = clf._bg_embedding.detach().cpu().requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    yr, hr = oc.classifier_forward(xr, w, b, clf._embeddings.cpu(), bg, True, 48, 65, 50.0, 3.0)
    loss_r = F.cross_entropy(yr, labels) + 256 * F.l1_loss(hr, target)
    loss_r.backward()

    errs = dict(loss=abs(float(loss) - float(loss_r)) / abs(float(loss_r)), dx=rel_err(xs.grad, xr.grad),
                dw=rel_err(clf._linear.weight.grad, w.grad), db=rel_err(clf._linear.bias.grad, b.grad),
                dbg=rel_err(clf._bg_embedding.grad, bg.grad))
    print('backward relative errors', {k: f'{v:.2e}' for k, v in errs.items()})
    assert max(errs.values()) < BWD_TOL, errs


@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float16])
def test_backward_with_half_precision_inputs(tmp_path, lib, dtype):
    """BASELINE config 5 (LVIS-1203 head, bf16 training step) and the mmcv fp16 hook of the COCO configs: x
    arrives in bf16 / fp16, the gradient goes back in the same type; parameters and their gradients are fp32."""
    names = [f'cat{i:04d}' for i in range(1203)]
    categories.Globals.categories = categories.Categories(names[:866], names[866:])
    categories.Globals.training = True
    try:
        path = make_prompts(tmp_path, names)
        clf = C.ViLDClassifier(prompts=path, in_features=1024, out_features=1204, scaler=dict(train=0.01, val=0.007)).cuda()
        g = torch.Generator().manual_seed(21)
        n = 2 * (512 + 300 + 27)  # 2 images x (sampled RoIs + object boxes + blocks): SURVEY 8d-5
        x = torch.randn(n, 1024, generator=g).to(dtype)
        labels = torch.randint(0, 866, (n, ), generator=g)
        labels[::5] = 1203
        xs = x.cuda().requires_grad_(True)
        y = clf(xs)
        assert y.dtype == torch.float32
        loss = F.cross_entropy(y, labels.cuda())
        loss.backward()
        assert xs.grad.dtype == dtype

        w = clf._linear.weight.detach().cpu().requires_grad_(True)
        b = clf._linear.bias.detach().cpu().requires_grad_(True)
        bg = clf._bg_embedding.detach().cpu().requires_grad_(True)
        xr = x.float().requires_grad_(True)
        yr, _ = oc.vild_forward(xr, w, b, clf._embeddings.cpu(), bg, True, 866, 1203, 0.01, 0.007)
        loss_r = F.cross_entropy(yr, labels)
        loss_r.backward()
        errs = dict(loss=abs(float(loss) - float(loss_r)) / abs(float(loss_r)), dx=rel_err(xs.grad, xr.grad),
                    dw=rel_err(clf._linear.weight.grad, w.grad), db=rel_err(clf._linear.bias.grad, b.grad),
                    dbg=rel_err(clf._bg_embedding.grad, bg.grad))
        print(dtype, 'backward relative errors', {k: f'{v:.2e}' for k, v in errs.items()})
        assert errs['dx'] < (8e-3 if dtype == torch.bfloat16 else BWD_TOL), errs  # dx is rounded to bf16: 2^-9 per element
        assert max(v for k, v in errs.items() if k != 'dx') < BWD_TOL, errs
    finally:
        categories.Globals.categories = categories.coco
        categories.Globals.training = False


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16, torch.bfloat16])
def test_inference_fast_path(tmp_path, coco_globals, dtype):
    """Without gradients and without listeners on `_linear` the two modules run as ONE C-ABI call on prepared
    operands (oake_classifier_fwd); same values as the two-call path, which takes over as soon as a hook is
    registered; prepared operands follow in-place parameter updates."""
    path = make_prompts(tmp_path, sorted(categories.coco.all_))
    clf = C.Classifier(prompts=path, in_features=1024, out_features=66).cuda()
    x = torch.randn(1000, 1024, generator=torch.Generator().manual_seed(4)).to(dtype)

    def run(training):
        categories.Globals.training = training
        with torch.no_grad():
            assert clf._fast_path_ok(x.cuda())
            y = clf(x.cuda())
        want, _ = oc.classifier_forward(x.float(), clf._linear.weight.detach().cpu(), clf._linear.bias.detach().cpu(),
                                        clf._embeddings.cpu(), clf._bg_embedding.detach().cpu(), training, 48, 65, 50.0, 3.0)
        inf = torch.isinf(want)
        assert y.shape == (1000, 66) and torch.equal(torch.isinf(y.cpu()), inf)
        assert rel_err(y.cpu()[~inf], want[~inf]) < 3e-3
        return y

    y_eval = run(False)
    run(True)
    # in-place update of the weights (an optimiser step): the prepared copies must follow
    with torch.no_grad():
        clf._linear.weight.mul_(0.5)
        clf._bg_embedding.add_(0.1)
    run(False)
    # a listener on `_linear` (the todd distiller) switches to the two-call path, which feeds it
    got = []
    handle = clf._linear.register_forward_hook(lambda m, i, o: got.append(o))
    with torch.no_grad():
        assert not clf._fast_path_ok(x.cuda())
        y_hooked = clf(x.cuda())
    handle.remove()
    assert len(got) == 1 and got[0].shape == (1000, 512)
    with torch.no_grad():
        y_fast = clf(x.cuda())
    assert rel_err(y_fast, y_hooked) < 2e-3
    assert y_eval.shape == y_fast.shape
    # gradients wanted: never the fast path
    assert not clf._fast_path_ok(x.cuda().float().requires_grad_(True))


def test_frozen_background_and_no_input_grad(tmp_path, coco_globals):
    path = make_prompts(tmp_path, sorted(categories.coco.all_))
    clf = C.Classifier(prompts=path, in_features=256, out_features=66).cuda()
    clf._bg_embedding.requires_grad_(False)  # ObjectMixin (bbox_heads.py:52-55)
    x = torch.randn(50, 256).cuda()
    y = clf(x)
    y[:, -1] = float('-inf')
    F.cross_entropy(y, torch.randint(0, 65, (50, )).cuda()).backward()
    assert clf._bg_embedding.grad is None and clf._linear.weight.grad is not None
    assert torch.isfinite(clf._linear.weight.grad).all()
