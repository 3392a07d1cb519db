"""Generates tests/golden/ref_golden.pt by EXECUTING THE REFERENCE'S OWN SOURCE FILES.

    python tests/golden/make_ref_golden.py            # needs /root/reference (build container only)

The reference (LutingWang/OADP) cannot be imported as a package here -- its third-party
dependencies (todd, clip, mmcv, mmdet, lvis, pycocotools) are neither installed nor vendored -- so
the individual modules below are loaded with importlib on top of the stand-ins in ``ref_stubs.py``:

    oadp/oake/base.py  oadp/oake/globals.py  oadp/oake/blocks.py  oadp/oake/objects.py
    oadp/base/globals_.py  oadp/dp/utils.py  oadp/dp/classifiers.py  oadp/base/losses.py
    oadp/dp/datasets.py (LoadCLIPFeatures)

Everything recorded in the fixture is produced by the reference's code: ``Dataset._partition`` /
``_partitions`` / ``_block`` / ``_bbox`` / ``_preprocess`` (blocks.py:40-109), ``_preprocess``
(globals.py:26-33), ``COCODataset._expand`` / ``_object`` / ``_mask`` / ``_preprocess``
(objects.py:76-186), ``Validator._build_model`` + ``Hooks`` (objects.py:198-314) driving
``model.visual(o, m)``, ``model.encode_image`` as called in globals.py:57 / blocks.py:129, and
``BaseClassifier`` / ``Classifier`` / ``ViLDClassifier`` (classifiers.py:19-112) with
``NormalizedLinear`` (utils.py:47-51).  The datasets are instantiated with ``__new__`` because
``torchvision.datasets.CocoDetection.__init__`` needs pycocotools; only constructor plumbing is
skipped.  Inputs are regenerated from seeds at test time (``ref_inputs``); outputs are stored.

/root/reference does not exist on the GPU box: nothing imports this file at test time except for
``ref_inputs`` / ``checksums`` (pure helpers, no reference access).
"""
from __future__ import annotations

import importlib.util
import os
import pathlib
import sys
import types

import numpy as np
import PIL.Image
import torch

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parents[1]
REF = pathlib.Path(os.environ.get('OADP_REFERENCE', '/root/reference'))
PKG = 'oadp_ref'

WEIGHT_SEED = 1234
IMAGES = ((640, 480, 11), (500, 375, 12), (427, 640, 13), (224, 300, 14))  # (W, H, seed)


# ----------------------------------------------------------------------------- shared with the tests
def synth_image(w: int, h: int, seed: int) -> np.ndarray:
    """Low-pass filtered noise, uint8 HWC (bicubic resampling has something to do)."""
    rng = np.random.default_rng(seed)
    base = rng.uniform(0, 255, size=(h // 8 + 2, w // 8 + 2, 3)).astype(np.float32)
    img = PIL.Image.fromarray(base.clip(0, 255).astype(np.uint8)).resize((w, h), PIL.Image.BICUBIC)
    arr = np.asarray(img).astype(np.float32) + rng.normal(0, 6, size=(h, w, 3))
    return arr.clip(0, 255).astype(np.uint8)


def synth_proposals(w: int, h: int, n: int, seed: int) -> np.ndarray:
    """(n, 5) float32 xyxy + score: in-image boxes, boxes hanging over every border, a box larger
    than the image after expansion, and two degenerate (< 4 px) boxes for the min_wh filter."""
    rng = np.random.default_rng(seed)
    side = np.exp(rng.uniform(np.log(8), np.log(min(w, h)), size=n))
    aspect = np.exp(rng.uniform(np.log(1 / 3), np.log(3), size=n))
    bw, bh = side * np.sqrt(aspect), side / np.sqrt(aspect)
    cx, cy = rng.uniform(0, w, size=n), rng.uniform(0, h, size=n)
    x1, y1 = np.clip(cx - bw / 2, 0, w - 1), np.clip(cy - bh / 2, 0, h - 1)
    x2, y2 = np.clip(cx + bw / 2, x1 + 1, w), np.clip(cy + bh / 2, y1 + 1, h)
    boxes = np.stack([x1, y1, x2, y2], 1)
    boxes[0] = (0.0, 0.0, 30.5, 20.25)                # corner
    boxes[1] = (w - 40.0, h - 25.0, w, h)             # opposite corner
    boxes[2] = (5.0, 5.0, w - 5.0, h - 5.0)           # expansion exceeds the image
    boxes[3] = (100.0, 100.0, 103.0, 150.0)           # w < 4: filtered
    boxes[4] = (50.0, 60.0, 54.0, 64.0)               # exactly 4 x 4: kept (inclusive)
    boxes[5] = (200.0, 10.0, 260.0, 13.5)             # h < 4: filtered
    score = np.sort(rng.uniform(0, 1, size=n))[::-1]
    return np.concatenate([boxes, score[:, None]], 1).astype(np.float32)


def checksums(t: torch.Tensor) -> torch.Tensor:
    """Per-row (sum, sum of squares, 4 probes) in float64: pins a (N, 3, 224, 224) batch in 48 B/row."""
    flat = t.reshape(t.shape[0], -1).double()
    probes = flat[:, [0, 12345, 77777, flat.shape[1] - 1]]
    return torch.cat([flat.sum(1, keepdim=True), (flat * flat).sum(1, keepdim=True), probes], 1)


def ref_inputs():
    images = [synth_image(w, h, seed) for w, h, seed in IMAGES]
    proposals = [synth_proposals(w, h, 14, seed + 100) for w, h, seed in IMAGES]
    return images, proposals


def loss_inputs():
    g = torch.Generator().manual_seed(77)
    probs = torch.sigmoid(torch.randn(11, 65, generator=g) * 2)
    probs[0, 0], probs[0, 1] = 0.0, 1.0  # the eps / clip clamps
    targets = torch.rand(11, 65, generator=g) > 0.9
    student = torch.nn.functional.normalize(torch.randn(9, 512, generator=g), dim=-1)
    teacher = torch.nn.functional.normalize(torch.randn(9, 512, generator=g), dim=-1).half().float()
    return probs, targets, student, teacher


def feature_store_inputs():
    """Per-image records in the layout OAKE writes (fp16), plus mmdet-style `results` dicts."""
    g = torch.Generator().manual_seed(55)
    rng = np.random.default_rng(55)

    def boxes(m):
        lt = torch.rand(m, 2, generator=g) * 400
        return torch.cat([lt, lt + 2 + torch.rand(m, 2, generator=g) * 200], 1).half()

    records, results = {}, {}
    for i in range(3):
        key = f'{500 + i:012d}'
        nb, no = 5 + 3 * i, 9 + 2 * i
        ob = boxes(no)
        ob[1] = torch.tensor([10.0, 10.0, 13.0, 40.0]).half()  # w < 4: dropped by the re-applied min_wh filter
        records[key] = dict(globals=torch.randn(1, 512, generator=g).half()[0],
                            blocks=dict(embeddings=torch.randn(nb, 512, generator=g).half(), bboxes=boxes(nb)),
                            objects=dict(embeddings=torch.randn(no, 512, generator=g).half(), bboxes=ob,
                                         objectness=torch.rand(no, 1, generator=g).half()))
        xy = rng.uniform(0, 400, size=(6, 2)).astype(np.float32)
        results[key] = dict(img_info=dict(id=500 + i), bbox_fields=['gt_bboxes'],
                            gt_bboxes=np.concatenate([xy, xy + rng.uniform(5, 200, size=(6, 2)).astype(np.float32)], 1),
                            gt_labels=np.array([3, 5, 6, 9, 0, 2]))  # 6 and 9 are pseudo labels (>= num_all = 6)
    return records, results


def recall_inputs():
    g = torch.Generator().manual_seed(31)
    logits = torch.randn(54, 65, generator=g)
    targets = torch.rand(54, 65, generator=g) > 0.93
    targets[:, 7] = False  # a label that never occurs
    return logits, targets


def classifier_inputs():
    g = torch.Generator().manual_seed(99)
    names = [f'cat{i:02d}' for i in range(9)]
    prompts = dict(names=names, embeddings=torch.randn(9, 512, generator=g) * 0.05,
                   scaler=torch.tensor([4.5]), bias=torch.tensor([0.25]))
    bases, novels = ('cat03', 'cat07', 'cat01', 'cat08'), ('cat00', 'cat05')
    x = torch.randn(7, 48, generator=g)
    weight = torch.randn(512, 48, generator=g) * 0.1
    bias = torch.randn(512, generator=g) * 0.1
    bg = torch.randn(1, 512, generator=g)
    return prompts, bases, novels, x, weight, bias, bg


# ------------------------------------------------------------------------------ reference loading
def _load_reference_modules(with_heads: bool = False):
    sys.path.insert(0, str(HERE))
    import ref_stubs
    ref_stubs.install()
    for name in (PKG, f'{PKG}.oake', f'{PKG}.base', f'{PKG}.dp'):
        mod = types.ModuleType(name)
        mod.__path__ = []  # a package whose __init__ is never run
        sys.modules[name] = mod

    def load(mod_name: str, rel: str):
        spec = importlib.util.spec_from_file_location(f'{PKG}.{mod_name}', REF / rel)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        return mod

    globals_ = load('base.globals_', 'oadp/base/globals_.py')
    sys.modules[f'{PKG}.base'].Globals = globals_.Globals
    m = dict(ref_stubs=ref_stubs, globals_=globals_)
    m['base'] = load('oake.base', 'oadp/oake/base.py')
    m['globals'] = load('oake.globals', 'oadp/oake/globals.py')
    m['blocks'] = load('oake.blocks', 'oadp/oake/blocks.py')
    m['objects'] = load('oake.objects', 'oadp/oake/objects.py')
    m['utils'] = load('dp.utils', 'oadp/dp/utils.py')
    m['classifiers'] = load('dp.classifiers', 'oadp/dp/classifiers.py')
    m['losses'] = load('base.losses', 'oadp/base/losses.py')
    base = sys.modules[f'{PKG}.base']
    base.coco, base.lvis = globals_.coco, globals_.lvis  # `from ..base import Globals, coco, lvis`
    m['datasets'] = load('dp.datasets', 'oadp/dp/datasets.py')
    if with_heads:  # make_ref_heads_golden.py
        sys.modules[f'{PKG}.base'].globals_ = globals_
        m['bbox_heads'] = load('dp.bbox_heads', 'oadp/dp/bbox_heads.py')
        m['roi_heads'] = load('dp.roi_heads', 'oadp/dp/roi_heads.py')
    return m


def main() -> None:
    sys.path.insert(0, str(ROOT))
    from oracle import vit  # only for the seeded weights (init_visual_params): no oracle arithmetic below
    from torchvision.datasets.vision import StandardTransform
    torch.set_num_threads(os.cpu_count() or 1)
    m = _load_reference_modules()
    stubs = m['ref_stubs']
    params = vit.init_visual_params(WEIGHT_SEED)
    stubs.set_clip_weights(params)
    images, proposals = ref_inputs()
    out = dict(weight_seed=WEIGHT_SEED, images=IMAGES, reference_files=[
        'oadp/oake/base.py', 'oadp/oake/globals.py', 'oadp/oake/blocks.py', 'oadp/oake/objects.py',
        'oadp/base/globals_.py', 'oadp/dp/utils.py', 'oadp/dp/classifiers.py', 'oadp/base/losses.py',
        'oadp/dp/datasets.py'
    ])

    # ---- blocks.py: _partition over every length a COCO image can have
    bds = m['blocks'].Dataset.__new__(m['blocks'].Dataset)
    bds._r, bds._s, bds._rescale = 224, 112, 1.5  # the constructor defaults (blocks.py:30-37)
    out['partition'] = {n: bds._partition(n) for n in range(200, 1401)}

    # ---- globals.py / blocks.py: model + transform exactly as their Validators build them
    g_model, g_pre = m['globals'].Validator._build_model()
    b_model, b_pre = m['blocks'].Validator._build_model()
    gds = m['globals'].Dataset.__new__(m['globals'].Dataset)
    gds.transforms = StandardTransform(g_pre, None)
    bds.transforms = StandardTransform(b_pre, None)
    out['globals'], out['blocks'] = [], []
    with torch.no_grad():
        for arr in images:
            pil = PIL.Image.fromarray(arr)
            gb = gds._preprocess(1, pathlib.Path('x.pth'), pil)
            g_img = gb.image.unsqueeze(0)  # globals.py:54
            g_raw = g_model.encode_image(g_img)  # globals.py:57
            emb = torch.nn.functional.normalize(g_raw)  # globals.py:58
            out['globals'].append(dict(pixels=checksums(g_img), raw=g_raw[0], embedding=emb.squeeze(0).half()))
            bb = bds._preprocess(1, pathlib.Path('x.pth'), pil)
            raw = b_model.encode_image(bb.blocks[:6])  # first six crops keep the fixture small and fast
            out['blocks'].append(dict(n=bb.blocks.shape[0], bboxes=bb.bboxes, bboxes_half=bb.bboxes.half(),
                                      pixels=checksums(bb.blocks), raw6=raw))

    # ---- objects.py: surgery + hooks + dataset
    o_model, o_pre = m['objects'].Validator._build_model()
    visual = o_model.visual
    out['objects_model'] = dict(grid=visual.grid, stride=tuple(visual.conv1.stride),
                                padding=tuple(visual.conv1.padding),
                                positional_embedding=visual.positional_embedding.detach().clone())
    ods = m['objects'].COCODataset.__new__(m['objects'].COCODataset)
    ods._grid = visual.grid
    ods._expand_mode = m['objects'].ExpandMode['ADAPTIVE']
    ods.transforms = StandardTransform(o_pre, None)
    out['objects'] = []
    with torch.no_grad():
        for arr, prop in zip(images, proposals):
            pil = PIL.Image.fromarray(arr)
            ods._proposals = {1: torch.tensor(prop, dtype=torch.float32)}
            ob = ods._preprocess(1, pathlib.Path('x.pth'), pil)
            boxes = stubs._BBoxesXYXY(ob.bboxes)
            expanded = ods._expand(boxes, torch.tensor(pil.size)).to_tensor()
            raw = o_model.visual(ob.objects.type(o_model.dtype), ob.masks.type(o_model.dtype))  # objects.py:328-330
            out['objects'].append(dict(bboxes=ob.bboxes, objectness=ob.objectness, expanded=expanded,
                                       masks=ob.masks.to(torch.uint8), pixels=checksums(ob.objects), raw=raw,
                                       embeddings=torch.nn.functional.normalize(raw).half()))

    # ---- classifiers.py
    prompts, bases, novels, x, weight, bias, bg = classifier_inputs()
    G = m['globals_']
    G.Globals.categories = G.Categories(bases=bases, novels=novels)
    ppath = HERE / '_ref_prompts.tmp.pth'
    torch.save(prompts, ppath)
    C = m['classifiers']
    res = {}
    try:
        for with_bg in (False, True):
            k = 6 + int(with_bg)
            heads = dict(base=C.BaseClassifier(prompts=str(ppath), in_features=48, out_features=k),
                         classifier=C.Classifier(prompts=str(ppath), in_features=48, out_features=k),
                         vild=C.ViLDClassifier(prompts=str(ppath), in_features=48, out_features=k,
                                               scaler=dict(train=0.01, val=0.007)),
                         vild_default=C.ViLDClassifier(prompts=str(ppath), in_features=48, out_features=k))
            for name, head in heads.items():
                with torch.no_grad():
                    head._linear.weight.copy_(weight)
                    head._linear.bias.copy_(bias)
                    if with_bg:
                        head._bg_embedding.copy_(bg)
                for training in (False, True):
                    G.Globals.training = training
                    with torch.no_grad():
                        res[(name, with_bg, training)] = dict(logits=head(x.clone()), hooked=head._linear(x.clone()))
        try:
            C.BaseClassifier(prompts=str(ppath), in_features=48, out_features=5)
            res['bad_out_features'] = None
        except RuntimeError as e:
            res['bad_out_features'] = str(e)
    finally:
        ppath.unlink()
    out['classifier'] = res

    # ---- losses.py: AsymmetricLoss (block / global heads) and RKDLoss (block relations), value + gradient
    L = m['losses']
    probs, targets, student, teacher = loss_inputs()
    lres = {}
    for name, kw in (('asl_configs', dict(gamma_neg=4, gamma_pos=0)), ('asl_defaults', dict()),
                     ('asl_sum_w16', dict(gamma_neg=4, gamma_pos=0, reduction='sum', weight=16.0))):
        x = probs.clone().requires_grad_(True)
        loss = L.AsymmetricLoss(**kw)(x, targets)
        loss.backward()
        lres[name] = dict(loss=loss.detach(), grad=x.grad.clone())
    for name, kw in (('rkd', dict()), ('rkd_w8', dict(weight=8.0))):
        sx = student.clone().requires_grad_(True)
        loss = L.RKDLoss(**kw)(sx, teacher)
        loss.backward()
        lres[name] = dict(loss=loss.detach(), grad=sx.grad.clone())
    out['losses'] = lres

    # ---- utils.py: MultilabelTopKRecall (sklearn macro recall over the labels that occur)
    rres = {}
    for k in (1, 5, 20):
        rres[k] = m['utils'].MultilabelTopKRecall(k=k)(*recall_inputs())
    empty_logits, empty_targets = recall_inputs()
    rres['no_positive'] = m['utils'].MultilabelTopKRecall(k=5)(empty_logits, torch.zeros_like(empty_targets))
    out['recall'] = rres

    # ---- datasets.py: LoadCLIPFeatures (the consumer of the OAKE files) on in-memory access layers
    import todd
    records, results = feature_store_inputs()
    ALR = todd.datasets.AccessLayerRegistry
    for task in ('globals', 'blocks', 'objects'):
        ALR.stores[f'coco/oake/{task}/train'] = {k: v[task] for k, v in records.items()}
    G.Globals.categories = G.Categories(bases=('a', 'b', 'c', 'd'), novels=('e', 'f'))  # num_all = 6
    step = m['datasets'].LoadCLIPFeatures(default=dict(type='PthAccessLayer', data_root='unused'),
                                          globals_=dict(task_name='coco/oake/globals/train'),
                                          blocks=dict(task_name='coco/oake/blocks/train'),
                                          objects=dict(task_name='coco/oake/objects/train'))
    fres = {}
    for key, res in results.items():
        r = step(dict(res, bbox_fields=list(res['bbox_fields'])))
        fres[key] = {k: r[k] for k in ('clip_global', 'clip_blocks', 'block_bboxes', 'block_labels', 'clip_objects',
                                       'object_bboxes', 'bbox_fields')}
    out['load_clip_features'] = fres

    torch.save(out, HERE / 'ref_golden.pt')
    size = (HERE / 'ref_golden.pt').stat().st_size
    print(f'wrote ref_golden.pt ({size / 1024:.0f} KiB): {len(out["partition"])} partitions, '
          f'{sum(b["n"] for b in out["blocks"])} block crops, {sum(o["raw"].shape[0] for o in out["objects"])} object crops')


if __name__ == '__main__':
    main()
