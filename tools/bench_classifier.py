"""Cosine-classifier micro-benchmark (BASELINE configs 4-5): inference latency at N = 1000 RoIs
(in = 1024, K = 66 / 1204) through the fused call, forward + backward at the LVIS training shape with bf16 inputs,
each next to the same arithmetic in PyTorch eager fp16 / bf16 on the same GPU, with the achieved GB/s against the
algorithmic byte count of BASELINE.md section 3.  `bench.py --workload classifier` prints the same lines."""
import json
import pathlib
import sys
import tempfile

import torch
import torch.nn.functional as F

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))


def timed(step, iters=100, warm=10):
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        step()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def eager_head(w, b, text, bg, alpha, shift, lo, hi, dtype):
    """classifiers.py:49-68,82-83 + utils.py:47-51 in PyTorch eager, parameters in `dtype`."""
    w, b, text = w.to(dtype), b.to(dtype), text.to(dtype)

    def fwd(x):
        h = F.normalize(F.linear(x, w, b))
        e = torch.cat([text, F.normalize(bg.to(dtype))]) if bg is not None else text
        y = h @ e.T
        if hi > lo:
            y[:, lo:hi] = float('-inf')
        return y * alpha - shift

    return fwd


def run(n, in_f, num_all, num_bases, train, dtype):
    from oadp_b200.dp import categories
    from oadp_b200.dp import classifiers as C
    names = [f'c{i:04d}' for i in range(num_all)]
    categories.Globals.categories = categories.Categories(names[:num_bases], names[num_bases:])
    categories.Globals.training = train
    with tempfile.TemporaryDirectory() as d:
        path = f'{d}/p.pth'
        torch.save(dict(names=names, embeddings=F.normalize(torch.randn(num_all, 512)) * 0.8,
                        scaler=torch.tensor([50.0]), bias=torch.tensor([3.0])), path)
        clf = C.Classifier(prompts=path, in_features=in_f, out_features=num_all + 1).cuda()
    x = torch.randn(n, in_f, device='cuda').to(dtype).requires_grad_(train)
    labels = torch.randint(0, num_bases, (n, ), device='cuda')
    lo, hi = (num_bases, num_all) if train else (0, 0)
    ref = eager_head(clf._linear.weight.detach(), clf._linear.bias.detach(), clf._embeddings, clf._bg_embedding.detach(),
                     50.0, 3.0, lo, hi, torch.float16 if dtype == torch.float32 else dtype)
    xe = x.detach().to(torch.float16 if dtype == torch.float32 else dtype).requires_grad_(train)

    def ours():
        if train:
            F.cross_entropy(clf(x), labels).backward()
        else:
            with torch.no_grad():
                clf(x)

    def eager():
        if train:
            F.cross_entropy(ref(xe).float(), labels).backward()
        else:
            with torch.no_grad():
                ref(xe)

    us, us_eager = timed(ours), timed(eager)
    k = num_all + 1
    algo_bytes = n * in_f * 2 + in_f * 512 * 2 + k * 512 * 2 + n * 512 * 2 + n * k * 4
    categories.Globals.training = False
    return dict(workload='classifier', n=n, in_features=in_f, k=k, mode='fwd+bwd' if train else 'fwd (fused call)',
                x_dtype=str(dtype).replace('torch.', ''), us=round(us, 2), eager_us=round(us_eager, 2),
                speedup_vs_eager=round(us_eager / us, 2), fwd_algorithmic_bytes=algo_bytes,
                gbs=None if train else round(algo_bytes / us / 1e3, 1), gflop=round(2 * n * (in_f * 512 + 512 * k) / 1e9, 3))


CASES = ((1000, 1024, 65, 48, False, torch.float16), (1000, 1024, 65, 48, False, torch.float32),
         (1000, 1024, 1203, 866, False, torch.float16), (2 * (512 + 300 + 27), 1024, 1203, 866, True, torch.bfloat16),
         (2 * (512 + 300 + 27), 1024, 1203, 866, True, torch.float32))

if __name__ == '__main__':
    for args in CASES:
        print(json.dumps(run(*args)), flush=True)
