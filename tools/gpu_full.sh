#!/bin/bash
# Round checkpoint: whole GPU test tier, the bench (both arms), launch list + full ncu captures.
mkdir -p gpurun_out
ts() { date +%H:%M:%S; }
timeout 900 python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name())"
python -m oadp_b200.build > gpurun_out/build.log 2>&1
echo "$(ts) == gpu tests"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/t_gpu.log
echo "$(ts) == smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "$(ts) == bench"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json
echo "$(ts) == bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.json
if [ -n "$PROFILE" ]; then
echo "$(ts) == launch list"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --images 2 > gpurun_out/launches_bench.log 2>&1
echo "$(ts) == ncu full: attention + gemm"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attention_cs|gemm_tcgen05" -s 6 -c 5 -f -o gpurun_out/prof_tower \
  python tools/quick_bench.py --variant 1 --batch 478 --iters 1 > gpurun_out/ncu_tower.log 2>&1; tail -2 gpurun_out/ncu_tower.log
fi
echo "$(ts) done"
