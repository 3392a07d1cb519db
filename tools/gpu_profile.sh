#!/bin/bash
# Profiling pass for profiles/: launch list of one bench run + full captures of the top kernels.
mkdir -p gpurun_out
timeout 900 python -c "import torch; torch.zeros(1).cuda()"
python -m oadp_b200.build > gpurun_out/build.log 2>&1
echo "== launch list (bench.py, 1 timed step)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --images 2 > gpurun_out/launches_bench.log 2>&1
tail -2 gpurun_out/launches_bench.log | cut -c1-300
echo "== full capture: GEMM instantiations (objects tower)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 5 -c 4 -f -o gpurun_out/prof_gemm \
  python tools/quick_bench.py --variant 1 --batch 478 --iters 1 > gpurun_out/ncu_gemm.log 2>&1; tail -2 gpurun_out/ncu_gemm.log
echo "== full capture: attention + front end"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attention_kernel|resize_u8|im2col_u8" -s 0 -c 6 -f -o gpurun_out/prof_misc \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --images 2 --workload objects > gpurun_out/ncu_misc.log 2>&1; tail -2 gpurun_out/ncu_misc.log | cut -c1-200
echo done
