"""Micro-benchmarks of the SURVEY 8f rows built so far, next to the way the reference does them:
  * f-2 ViLD ensemble scoring: one `oake_vild_ensemble` launch vs the reference's tensor lines
    (oadp/dp/roi_heads.py:93-112) in PyTorch eager on the same GPU; achieved GB/s against 12 B/element.
  * f-1 feature store: random-order reads of per-image object records (300 x 512 fp16 + boxes) from the
    reference's file-per-key layout (`torch.load` of a pickle) vs the packed, memory-mapped shards.
One JSON line per case."""
import json
import pathlib
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from oadp_b200 import store  # noqa: E402
from oadp_b200.dp import roi_heads  # noqa: E402


def reference_ensemble(bbox_logits, object_logits, lam):
    bbox_scores = bbox_logits.softmax(-1)**lam
    object_scores = object_logits.softmax(-1)**(1 - lam)
    cls_score = bbox_scores * object_scores
    cls_score[:, -1] = 1 - cls_score[:, :-1].sum(-1)
    return cls_score.log()


def time_gpu(fn, iters=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3  # us


def bench_ensemble(n, num_bases, num_all):
    k1 = num_all + 1
    g = torch.Generator(device='cuda').manual_seed(0)
    bbox = torch.randn(n, k1, device='cuda', generator=g) * 4
    obj = torch.randn(n, k1, device='cuda', generator=g) * 4
    obj[:, -1] = float('-inf')
    lam = roi_heads.ensemble_lambda(num_bases, num_all, device='cuda')
    ours = time_gpu(lambda: roi_heads.vild_ensemble(bbox, obj, lam))
    ref = time_gpu(lambda: reference_ensemble(bbox, obj, lam))
    err = float((roi_heads.vild_ensemble(bbox, obj, lam) - reference_ensemble(bbox, obj, lam)).abs().max())
    algo = 12 * n * k1
    return dict(row='f-2 vild_ensemble', n=n, k1=k1, us=ours, us_reference_eager=ref, speedup=ref / ours,
                algorithmic_bytes=algo, gbs=algo / ours / 1e3, max_abs_diff_vs_eager=err)


def bench_store(n_keys=2000, n_obj=300):
    g = torch.Generator().manual_seed(0)
    vals = {store.key_of(i): dict(embeddings=torch.randn(n_obj, 512, generator=g).half(),
                                  bboxes=(torch.rand(n_obj, 4, generator=g) * 600).half(),
                                  objectness=torch.rand(n_obj, 1, generator=g).half()) for i in range(n_keys)}
    order = np.random.default_rng(0).permutation(n_keys)
    keys = [store.key_of(int(i)) for i in order]
    out = {}
    with tempfile.TemporaryDirectory() as d:
        t0 = time.perf_counter()
        pth = store.PthStore(d, 'pth')
        for k, v in vals.items():
            pth[k] = v
        t_w_pth = time.perf_counter() - t0
        t0 = time.perf_counter()
        with store.PackedWriter(d, 'packed') as w:
            for k, v in vals.items():
                w.add(k, v)
        t_w_packed = time.perf_counter() - t0
        for name, s in (('pth', store.PthStore(d, 'pth')), ('packed', store.PackedStore(d, 'packed'))):
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                acc = 0.0
                for k in keys:
                    acc += float(s[k]['embeddings'][0, 0])  # touches the payload
                best = min(best, time.perf_counter() - t0)
            out[name] = best
    mb = n_keys * n_obj * (512 + 5) * 2 / 1e6
    return dict(row='f-1 feature store (objects records, warm page cache)', keys=n_keys, crops_per_key=n_obj, payload_mb=mb,
                write_s_pth=t_w_pth, write_s_packed=t_w_packed, read_us_per_key_pth=out['pth'] / n_keys * 1e6,
                read_us_per_key_packed=out['packed'] / n_keys * 1e6, read_speedup=out['pth'] / out['packed'])


if __name__ == '__main__':
    for args in ((1000, 48, 65), (1000, 866, 1203), (8000, 866, 1203)):
        print(json.dumps(bench_ensemble(*args)), flush=True)
    print(json.dumps(bench_store()), flush=True)
