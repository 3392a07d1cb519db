#!/bin/bash
# Third profiling pass of round 2 (after the attention side warps): launch list of one bench step + full capture of the
# attention kernel and the GEMMs around it.
mkdir -p gpurun_out
export OAKE_ALLOW_RANDOM_WEIGHTS=1
timeout 900 python -c "import torch; torch.zeros(1).cuda()"
python -m oadp_b200.build > gpurun_out/build.log 2>&1
echo "== launch list (one timed step)"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/launches_bench.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv "bench.py --steps 1 --warmup 3 (default oake workload, 8 images)" > gpurun_out/launch_list_summary.csv; head -30 gpurun_out/launch_list_summary.csv
echo "== ncu --set full: attention (objects, B = 478) and its neighbours"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attention_cs|gemm_tcgen05" -s 8 -c 6 -f -o gpurun_out/prof_tower \
  python tools/quick_bench.py --variant 1 --batch 478 --iters 1 > gpurun_out/ncu_tower.log 2>&1; tail -1 gpurun_out/ncu_tower.log
python tools/ncu_summary.py gpurun_out/prof_tower.ncu-rep > gpurun_out/ncu_tower_summary.txt; cat gpurun_out/ncu_tower_summary.txt
ncu -i gpurun_out/prof_tower.ncu-rep --page raw --csv --kernel-name regex:attention_cs 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]; r=rows[2]
for m in ['sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__average_warp_latency_issue_stalled_long_scoreboard.pct','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio']:
    if m in hdr: print(m, r[hdr.index(m)])
" | tee gpurun_out/ncu_attn_pipes.txt
rm -f gpurun_out/prof_*.ncu-rep
echo done
