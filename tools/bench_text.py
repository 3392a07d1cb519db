"""Micro-benchmark of the text tower (second half): the prompt-builder workload of
oadp/prompts/vild.py:56-72 -- 74 templates x 1 217 category names, one (1217, L) token batch per template --
on the GPU against the fp32 oracle on the host cores (bounded sample).  Prints JSON lines.
Synthetic token ids (the BPE vocabulary is not available offline), seeded random weights.
(Developer measurement, not product code: like bench.py's cpu_baseline leg it uses the oracle as the timed CPU baseline
and as the checker of the sample; nothing under oadp_b200/ imports it -- tests/test_abi.py.)"""
import json
import os
import pathlib
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from oadp_b200.text import OakeTextModel  # noqa: E402
from oracle import text as otext  # noqa: E402


def main():
    templates, names, length = 74, 1217, 16
    p = otext.init_text_params(0)
    model = OakeTextModel(p, 'cuda')
    batches = [otext.synthetic_tokens(names, length, seed=s).cuda() for s in range(templates)]
    for _ in range(2):
        model.encode_text(batches[0])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    total = None
    for tokens in batches:
        e = F.normalize(model.encode_text(tokens))
        total = e if total is None else total + e
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    rows = templates * names * length
    flops = rows * 12 * 2 * (512 * 1536 + 512 * 512 + 2 * 512 * 2048)
    print(json.dumps(dict(what='prompts_gpu', templates=templates, names=names, length=length, ms=round(ms, 2),
                          prompts_per_s=round(templates * names / ms * 1e3), gemm_tflops=round(flops / ms / 1e9, 1))),
          flush=True)
    torch.set_num_threads(os.cpu_count() or 1)
    sample = batches[0][:256].cpu().long()
    t = time.perf_counter()
    want = otext.encode_text(p, sample)
    dt = time.perf_counter() - t
    got = model.encode_text(sample).cpu()
    cos = float((1 - F.cosine_similarity(got, want, dim=-1)).max())
    print(json.dumps(dict(what='prompts_cpu_oracle', threads=torch.get_num_threads(), sample_prompts=256,
                          prompts_per_s=round(256 / dt, 1), worst_1_minus_cos_vs_gpu=cos)), flush=True)


if __name__ == '__main__':
    main()
