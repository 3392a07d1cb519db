#!/bin/bash
# Round-2 bench checkpoint: the default line (both arms), the per-workload lines and the classifier line.
mkdir -p gpurun_out
export OAKE_ALLOW_RANDOM_WEIGHTS=1
timeout 900 python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name())"
python -m oadp_b200.build > gpurun_out/build.log 2>&1
echo "== bench (default)"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 400 gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2>/dev/null; tail -c 300 gpurun_out/r2_bench_ref.json
rm -f gpurun_out/r2_bench_workloads.jsonl
for w in globals blocks objects; do timeout 600 python bench.py --workload $w --steps 10 --no-cpu-baseline --no-library-baseline 2>/dev/null | tail -1 >> gpurun_out/r2_bench_workloads.jsonl; done
timeout 600 python bench.py --workload globals --images 512 --steps 10 --no-cpu-baseline --no-library-baseline 2>/dev/null | tail -1 >> gpurun_out/r2_bench_workloads.jsonl
timeout 600 python bench.py --workload blocks --images 64 --steps 10 --no-cpu-baseline --no-library-baseline 2>/dev/null | tail -1 >> gpurun_out/r2_bench_workloads.jsonl
timeout 600 python bench.py --workload classifier 2>/dev/null | tail -1 >> gpurun_out/r2_bench_workloads.jsonl
python - <<'PY'
import json
for ln in open('gpurun_out/r2_bench_workloads.jsonl'):
    d = json.loads(ln)
    print(d['config'].get('workload'), d['config'].get('images_per_step_per_gpu'), 'value', round(d['value']), 'e2e', round(d.get('e2e', {}).get('value', 0)) if d.get('e2e') else None)
PY
