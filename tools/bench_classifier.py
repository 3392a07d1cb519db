"""Cosine-classifier micro-benchmark: forward and forward+backward latency, and the
achieved HBM GB/s against the algorithmic byte count, at the reference's shapes."""
import json
import pathlib
import sys
import tempfile

import torch
import torch.nn.functional as F

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from oadp_b200.dp import categories  # noqa: E402
from oadp_b200.dp import classifiers as C  # noqa: E402


def run(n, in_f, num_all, num_bases, with_bg, train):
    names = [f'c{i:04d}' for i in range(num_all)]
    categories.Globals.categories = categories.Categories(names[:num_bases], names[num_bases:])
    categories.Globals.training = train
    with tempfile.TemporaryDirectory() as d:
        path = f'{d}/p.pth'
        torch.save(dict(names=names, embeddings=F.normalize(torch.randn(num_all, 512)) * 0.8,
                        scaler=torch.tensor([50.0]), bias=torch.tensor([3.0])), path)
        clf = C.Classifier(prompts=path, in_features=in_f, out_features=num_all + int(with_bg)).cuda()
    x = torch.randn(n, in_f, device='cuda', requires_grad=train)
    labels = torch.randint(0, num_bases, (n, ), device='cuda')

    def step():
        y = clf(x)
        if train:
            F.cross_entropy(y, labels).backward()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 50
    a.record()
    for _ in range(iters):
        step()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / iters * 1e3
    k = num_all + int(with_bg)
    algo_bytes = n * in_f * 2 + in_f * 512 * 2 + k * 512 * 2 + n * 512 * 2 + n * k * 4
    flops = 2 * n * (in_f * 512 + 512 * k)
    return dict(n=n, in_features=in_f, k=k, mode='fwd+bwd' if train else 'fwd', us=us,
                fwd_algorithmic_bytes=algo_bytes, fwd_gbs_if_fwd_only=None if train else algo_bytes / us / 1e3,
                gflops=flops / 1e9)


if __name__ == '__main__':
    for args in ((1000, 1024, 65, 48, True, False), (1000, 1024, 1203, 866, True, False),
                 (2 * (512 + 300 + 27), 1024, 1203, 866, True, True), (2, 256, 65, 48, False, False)):
        print(json.dumps(run(*args)))
