#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -c "import torch; torch.zeros(1).cuda()"
python -m oadp_b200.build > gpurun_out/build.log 2>&1
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "== quick bench"; timeout 300 python tools/quick_bench.py --variant 1 --batch 478 --iters 10 2>&1 | grep -E "variant|gemm_|attn_"
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_ab.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks'])
for k,v in d['roofline']['classes'].items(): print(f"{k:22s} {v['ms_per_step']:8.3f} ms {v['tflops']:8.1f}")
PY
