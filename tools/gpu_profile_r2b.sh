#!/bin/bash
# Second profiling pass of round 2 (after the front-end rework): launch list of one bench step, the NVTX-filtered
# capture of the dominant GEMM class that feeds roofline.traffic, full captures of the tower and front-end kernels.
mkdir -p gpurun_out
export OAKE_ALLOW_RANDOM_WEIGHTS=1
timeout 900 python -c "import torch; torch.zeros(1).cuda()"
python -m oadp_b200.build > gpurun_out/build.log 2>&1
echo "== bench (reference numbers for the captures below)"
timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_for_profile.json 2> gpurun_out/bench_for_profile.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_for_profile.json').read().strip().splitlines()[-1])
r = d['roofline']
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'], 2), 'dominant', r['kernel'], 'flops/launch', r['flops_per_launch'])
open('gpurun_out/dominant.txt', 'w').write(f"{r['kernel'].split('(')[1].strip(')')} {r['flops_per_launch']}\n")
PY
read CLS FLOPS < gpurun_out/dominant.txt
echo "== launch list (one timed step)"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/launches_bench.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv "bench.py --steps 1 --warmup 3 (default oake workload, 8 images)" > gpurun_out/launch_list_summary.csv; head -24 gpurun_out/launch_list_summary.csv
echo "== ncu --set full, NVTX range $CLS (every launch of one step: warm-up skipped by launch count)"
timeout 1500 ncu --set full --clock-control none --nvtx --nvtx-include "$CLS/" -s 252 -c 84 -f -o gpurun_out/prof_$CLS \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/ncu_$CLS.log 2>&1; tail -2 gpurun_out/ncu_$CLS.log | cut -c1-200
python tools/ncu_traffic.py gpurun_out/prof_$CLS.ncu-rep $CLS $FLOPS gpurun_out/ncu_traffic.json; cat gpurun_out/ncu_traffic.json
echo "== ncu --set full: the objects tower's kernels at B = 478 (patch GEMM over the block matrix, assemble, one block)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attention_cs|gemm_tcgen05|assemble|blockcol" -s 0 -c 9 -f -o gpurun_out/prof_tower \
  python tools/quick_bench.py --variant 1 --batch 478 --iters 1 > gpurun_out/ncu_tower.log 2>&1; tail -1 gpurun_out/ncu_tower.log
python tools/ncu_summary.py gpurun_out/prof_tower.ncu-rep > gpurun_out/ncu_tower_summary.txt; cat gpurun_out/ncu_tower_summary.txt
echo "== ncu --set full: front end of the objects workload (prepare / fast / big resize kernels writing the block matrix)"
timeout 900 ncu --set full --clock-control none -k regex:"resize_" -s 24 -c 6 -f -o gpurun_out/prof_frontend \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-baseline --workload objects > gpurun_out/ncu_frontend.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_frontend.ncu-rep > gpurun_out/ncu_frontend_summary.txt; cat gpurun_out/ncu_frontend_summary.txt
rm -f gpurun_out/prof_*.ncu-rep  # the reports exceed the 64 MiB transfer limit: only the summaries above travel back
echo done
