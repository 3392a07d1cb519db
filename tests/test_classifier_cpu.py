"""CPU tier: classifier oracle sanity, category tables, registry surface, prompts handling."""
import json

import pytest
import torch

from oadp_b200.dp import categories
from oadp_b200.dp import classifiers as C
from oracle import classifier as oc


def make_prompts(tmp_path, names, seed=0, with_affine=True):
    g = torch.Generator().manual_seed(seed)
    emb = torch.nn.functional.normalize(torch.randn(len(names), 512, generator=g)) * 0.8  # norm < 1, like vild.pth
    d = dict(names=list(names), embeddings=emb)
    if with_affine:
        d.update(scaler=torch.tensor([50.0]), bias=torch.tensor([3.0]))
    path = tmp_path / 'prompts.pth'
    torch.save(d, path)
    return str(path), emb


def test_categories_and_globals(tmp_path):
    assert categories.coco.all_[:2] == ('person', 'bicycle') and categories.coco.all_[48] == 'airplane'
    ann = tmp_path / 'lvis.json'
    ann.write_text(json.dumps(dict(categories=[dict(id=3, name='c', frequency='r'), dict(id=1, name='a', frequency='f'),
                                               dict(id=2, name='b', frequency='c'), dict(id=4, name='d', frequency='f')])))
    lv = categories.Categories.from_lvis(str(ann))
    assert lv.bases == ('a', 'b', 'd') and lv.novels == ('c', ) and lv.num_all == 4
    with pytest.raises(TypeError):
        categories.Globals()


def test_construction_reorders_prompts_and_validates_out_features(tmp_path):
    categories.Globals.categories = categories.coco
    names = sorted(categories.coco.all_) + ['zzz_extra']
    path, emb = make_prompts(tmp_path, names)
    clf = C.LINEAR_LAYERS.build(dict(type='Classifier', prompts=path), in_features=1024, out_features=66)
    assert isinstance(clf._linear, torch.nn.Module) and clf._linear.weight.shape == (512, 1024)
    assert clf._bg_embedding.shape == (1, 512) and clf._scaler == 50.0 and clf._bias == 3.0
    want = emb[[names.index(n) for n in categories.coco.all_]]
    assert torch.equal(clf._embeddings, want)
    assert '_embeddings' not in clf.state_dict()  # non-persistent buffer (classifiers.py:47)
    assert clf.embeddings.shape == (66, 512)
    nobg = C.ViLDClassifier(prompts=path, in_features=256, out_features=65, scaler=dict(train=0.01, val=0.007))
    assert nobg._bg_embedding is None
    categories.Globals.training = True
    assert nobg.scaler == 0.01
    categories.Globals.training = False
    assert nobg.scaler == 0.007
    with pytest.raises(RuntimeError, match='64'):
        C.BaseClassifier(prompts=path, in_features=256, out_features=64)
    with pytest.raises(binding_error()):
        nobg(torch.zeros(2, 256))  # CPU tensors: no fallback


def binding_error():
    from oadp_b200 import binding
    return binding.OakeError


def test_oracle_semantics():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(7, 256, generator=g)
    w = torch.randn(512, 256, generator=g) * 0.05
    b = torch.randn(512, generator=g) * 0.1
    text = torch.randn(65, 512, generator=g) * 0.04
    bg = torch.randn(1, 512, generator=g)
    y, h = oc.base_forward(x, w, b, text, bg, True, 48, 65)
    assert y.shape == (7, 66) and torch.allclose(h.norm(dim=-1), torch.ones(7), atol=1e-5)
    assert torch.isinf(y[:, 48:65]).all() and torch.isfinite(y[:, :48]).all() and torch.isfinite(y[:, 65]).all()
    yv, _ = oc.base_forward(x, w, b, text, bg, False, 48, 65)
    assert torch.isfinite(yv).all()
    # bg row is normalised at use, text rows are used as stored
    assert torch.allclose(yv[:, 65], h @ torch.nn.functional.normalize(bg)[0], atol=1e-6)
    assert torch.allclose(yv[:, 3], h @ text[3], atol=1e-6)
    yc, _ = oc.classifier_forward(x, w, b, text, bg, False, 48, 65, 50.0, 3.0)
    assert torch.allclose(yc, yv * 50 - 3)
    yl, _ = oc.vild_forward(x, w, b, text, None, True, 48, 65)
    assert yl.shape == (7, 65) and torch.allclose(yl[:, :48], (h @ text[:48].T) / 0.007, rtol=1e-5)
    assert torch.isinf(oc.object_head_logits(yv)[:, -1]).all()
