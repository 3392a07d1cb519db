"""Developer probe: device-resident encode throughput + per-kernel-class event profile."""
import argparse
import json
import sys
import pathlib

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from oadp_b200 import binding  # noqa: E402
from oadp_b200.model import OakeEngine  # noqa: E402

GFLOP = {0: 8.818, 1: 33.552}


def init_params(seed=0, layers=12):
    g = torch.Generator().manual_seed(seed)
    w = 768
    p = {'conv1.weight': torch.randn(w, 3, 32, 32, generator=g) * 0.02,
         'class_embedding': torch.randn(w, generator=g) * 0.036,
         'positional_embedding': torch.randn(50, w, generator=g) * 0.036,
         'proj': torch.randn(w, 512, generator=g) * 0.036}
    for n in ('ln_pre', 'ln_post'):
        p[n + '.weight'] = torch.ones(w)
        p[n + '.bias'] = torch.zeros(w)
    for i in range(layers):
        pre = f'transformer.resblocks.{i}.'
        for n in ('ln_1', 'ln_2'):
            p[pre + n + '.weight'] = torch.ones(w)
            p[pre + n + '.bias'] = torch.zeros(w)
        p[pre + 'attn.in_proj_weight'] = torch.randn(3 * w, w, generator=g) * 0.036
        p[pre + 'attn.in_proj_bias'] = torch.zeros(3 * w)
        p[pre + 'attn.out_proj.weight'] = torch.randn(w, w, generator=g) * 0.007
        p[pre + 'attn.out_proj.bias'] = torch.zeros(w)
        p[pre + 'mlp.c_fc.weight'] = torch.randn(4 * w, w, generator=g) * 0.025
        p[pre + 'mlp.c_fc.bias'] = torch.zeros(4 * w)
        p[pre + 'mlp.c_proj.weight'] = torch.randn(w, 4 * w, generator=g) * 0.007
        p[pre + 'mlp.c_proj.bias'] = torch.zeros(w)
    return p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--variant', type=int, default=0)
    ap.add_argument('--batch', type=int, default=1024)
    ap.add_argument('--iters', type=int, default=5)
    args = ap.parse_args()
    eng = OakeEngine(init_params(), 'cuda')
    eng.MAX_CROPS = {0: args.batch, 1: args.batch}
    px = torch.randn(args.batch, 3, 224, 224, device='cuda')
    masks = (torch.rand(args.batch, 1, 14, 14, device='cuda') > 0.5).float() if args.variant else None
    for _ in range(3):
        eng.encode_pixels(px, masks, args.variant)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.iters):
        eng.encode_pixels(px, masks, args.variant)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.iters
    cps = args.batch / ms * 1e3
    tf = cps * GFLOP[args.variant] / 1e3
    print(json.dumps(dict(variant=args.variant, batch=args.batch, ms=ms, crops_per_s=cps, tflops=tf,
                          frac_of_1382=tf / 1382)))
    eng.profile(True)
    for _ in range(2):
        eng.encode_pixels(px, masks, args.variant)
    prof = eng.profile_collect()
    tot = sum(v['ms'] for v in prof.values())
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms']):
        if v['launches']:
            tfl = v['flops'] / (v['ms'] * 1e-3) / 1e12 if v['ms'] > 0 else 0
            print(f"{k:18s} {v['ms']/2:9.3f} ms  {100*v['ms']/tot:5.1f}%  launches {v['launches']//2:4d}  {tfl:8.1f} TFLOP/s")


if __name__ == '__main__':
    main()
