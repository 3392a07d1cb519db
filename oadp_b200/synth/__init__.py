"""Seeded synthetic COCO-shaped inputs: there is no dataset / checkpoint offline.

All generators are pure numpy with `default_rng(seed)` so that tests, bench.py and the CPU oracle
see identical bytes.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

# (W, H) mix of COCO train2017-like shapes
COCO_SIZES: Tuple[Tuple[int, int], ...] = ((640, 480), (640, 427), (480, 640), (427, 640), (500, 375), (640, 640),
                                           (612, 612), (640, 426))


def image(width: int, height: int, seed: int) -> np.ndarray:
    """uint8 HWC RGB low-pass-filtered noise: smooth enough that bicubic resampling matters, with
    full dynamic range.  Built from a coarse random grid upsampled bilinearly plus fine noise."""
    rng = np.random.default_rng(seed)
    gh, gw = height // 16 + 2, width // 16 + 2
    coarse = rng.random((gh, gw, 3))
    ys = np.linspace(0, gh - 1.001, height)
    xs = np.linspace(0, gw - 1.001, width)
    y0, x0 = ys.astype(int), xs.astype(int)
    fy, fx = (ys - y0)[:, None, None], (xs - x0)[None, :, None]
    a = coarse[y0][:, x0]
    b = coarse[y0][:, x0 + 1]
    c = coarse[y0 + 1][:, x0]
    d = coarse[y0 + 1][:, x0 + 1]
    smooth = (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy
    fine = rng.random((height, width, 3))
    out = 0.8 * smooth + 0.2 * fine
    out = (out - out.min()) / (out.max() - out.min())
    # C-contiguous HWC like a decoded image (the fancy indexing above leaves a permuted memory order,
    # which would make every staging copy a strided gather)
    return np.ascontiguousarray((out * 255.0 + 0.5).astype(np.uint8))


def images(n: int, seed: int = 0, sizes=COCO_SIZES) -> List[np.ndarray]:
    return [image(*sizes[i % len(sizes)], seed=seed * 100003 + i) for i in range(n)]


def proposals(width: int, height: int, n: int, seed: int, degenerate_frac: float = 0.02) -> np.ndarray:
    """(n,5) f32 xyxy + objectness: side log-uniform in [8, min(W,H)], aspect log-uniform in
    [1/3, 3], centres uniform, clipped to the image, scores sorted descending, ~2 % boxes thinner
    than 4 px to exercise the min_wh filter.  The first box is always valid."""
    rng = np.random.default_rng(seed)
    m = min(width, height)
    side = np.exp(rng.uniform(np.log(8.0), np.log(float(m)), n))
    aspect = np.exp(rng.uniform(np.log(1 / 3), np.log(3.0), n))
    w = side * np.sqrt(aspect)
    h = side / np.sqrt(aspect)
    deg = rng.random(n) < degenerate_frac
    deg[0] = False
    w = np.where(deg, rng.uniform(0.5, 3.9, n), w)
    cx = rng.uniform(0, width, n)
    cy = rng.uniform(0, height, n)
    x1 = np.clip(cx - w / 2, 0, width)
    x2 = np.clip(cx + w / 2, 0, width)
    y1 = np.clip(cy - h / 2, 0, height)
    y2 = np.clip(cy + h / 2, 0, height)
    score = np.sort(rng.random(n))[::-1]
    out = np.stack([x1, y1, x2, y2, score], axis=1).astype(np.float32)
    # keep the guaranteed-valid first box comfortably valid after clipping
    out[0, :4] = np.array([width * 0.25, height * 0.25, width * 0.75, height * 0.75], dtype=np.float32)
    return out


def visual_params(seed: int = 0, layers: int = 12) -> Dict[str, 'np.ndarray']:
    """Seeded ViT-B/32 visual-tower weights in the OpenAI state-dict layout, as torch tensors.

    Same recipe as oracle.vit.init_visual_params (kept separate: the product never imports the
    oracle): CLIP init scales, non-trivial LayerNorm affines and biases."""
    import torch
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    w = 768
    scale = w**-0.5
    proj_std = (w**-0.5) * ((2 * layers)**-0.5)
    fc_std = (2 * w)**-0.5
    p = {}
    p['conv1.weight'] = rn(w, 3, 32, 32, std=(3 * 32 * 32)**-0.5)
    p['class_embedding'] = rn(w, std=scale)
    p['positional_embedding'] = rn(50, w, std=scale)
    for name in ('ln_pre', 'ln_post'):
        p[f'{name}.weight'] = 1.0 + rn(w, std=0.1)
        p[f'{name}.bias'] = rn(w, std=0.1)
    for i in range(layers):
        pre = f'transformer.resblocks.{i}.'
        for name in ('ln_1', 'ln_2'):
            p[pre + f'{name}.weight'] = 1.0 + rn(w, std=0.1)
            p[pre + f'{name}.bias'] = rn(w, std=0.1)
        p[pre + 'attn.in_proj_weight'] = rn(3 * w, w, std=scale)
        p[pre + 'attn.in_proj_bias'] = rn(3 * w, std=0.02)
        p[pre + 'attn.out_proj.weight'] = rn(w, w, std=proj_std)
        p[pre + 'attn.out_proj.bias'] = rn(w, std=0.02)
        p[pre + 'mlp.c_fc.weight'] = rn(4 * w, w, std=fc_std)
        p[pre + 'mlp.c_fc.bias'] = rn(4 * w, std=0.02)
        p[pre + 'mlp.c_proj.weight'] = rn(w, 4 * w, std=proj_std)
        p[pre + 'mlp.c_proj.bias'] = rn(w, std=0.02)
    p['proj'] = rn(w, 512, std=scale)
    return p


def _write_image(job) -> None:
    import PIL.Image
    path, w, h, seed, fmt, i = job
    arr = image(w, h, seed)
    if fmt == 'jpg':
        PIL.Image.fromarray(arr).save(path, quality=(95, 85, 75)[i % 3], subsampling=(2, 0, 1)[i % 3])
    else:
        PIL.Image.fromarray(arr).save(path)


def write_coco_dataset(root, n_images: int, seed: int = 0, n_proposals: int = 40, sizes=COCO_SIZES,
                       first_id: int = 101, fmt: str = 'png', workers: int = 0) -> dict:
    """Materialises a tiny COCO-format dataset (lossless PNG images -- or, with fmt='jpg', JPEG files
    of mixed quality / chroma sampling as COCO's are --, instances json, proposal pickle in image-id
    order) plus an OAKE config for each task.  Returns the paths."""
    import json
    import pathlib
    import pickle

    root = pathlib.Path(root)
    (root / 'images').mkdir(parents=True, exist_ok=True)
    infos, props = [], []
    jobs = []
    for i in range(n_images):
        w, h = sizes[i % len(sizes)]
        id_ = first_id + 7 * i
        name = f'{id_:012d}.{fmt}'
        jobs.append((str(root / 'images' / name), w, h, seed * 100003 + i, fmt, i))
        infos.append(dict(id=id_, file_name=name, width=w, height=h))
        props.append(proposals(w, h, n_proposals, seed=seed * 7919 + i))
    if workers > 1:
        import multiprocessing
        with multiprocessing.get_context('fork').Pool(workers) as pool:
            pool.map(_write_image, jobs, chunksize=16)
    else:
        for job in jobs:
            _write_image(job)
    ann = root / 'instances.json'
    ann.write_text(json.dumps(dict(images=infos, annotations=[], categories=[])))
    pkl = root / 'proposals.pkl'
    with open(pkl, 'wb') as f:
        pickle.dump(props, f)
    cfgs = {}
    for task in ('globals', 'blocks', 'objects'):
        ds = dict(root=str(root / 'images'), annFile=str(ann))
        extra = ''
        if task == 'objects':
            ds.update(type='COCODataset', proposal_file=str(pkl), proposal_sorted=True)
            extra = 'mini_batch_size = 512\n'
        body = ''
        for split in ('train', 'val'):
            d = dict(ds, output_dir=str(root / 'oake' / task / split))
            body += f'{split} = dict(dataloader=dict(dataset=dict({", ".join(f"{k}={v!r}" for k, v in d.items())}), num_workers=2))\n'
        cfg = root / f'{task}.py'
        cfg.write_text(body + 'log = dict(interval=2)\n' + extra)
        cfgs[task] = str(cfg)
    return dict(root=str(root), ann=str(ann), proposals=str(pkl), configs=cfgs, ids=[i['id'] for i in infos])
