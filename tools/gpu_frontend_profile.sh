#!/bin/bash
# Front-end profile: launch list (durations) of the resize / matrix kernels of one default step, then one full
# capture of each.
mkdir -p gpurun_out
export OAKE_ALLOW_RANDOM_WEIGHTS=1
timeout 900 python -c "import torch; torch.zeros(1).cuda()"
python -m oadp_b200.build > gpurun_out/build.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"resize_|blockcol|im2col|object_masks" --csv \
  --log-file gpurun_out/fe_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/fe_launches.log 2>&1
python tools/launch_summary.py gpurun_out/fe_launches.csv "front-end kernels, bench.py --steps 1 --warmup 3" | tee gpurun_out/fe_launch_summary.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"resize_fast|resize_prepare|blockcol_u8" -s 12 -c 6 -f -o gpurun_out/prof_fe \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-baseline --workload objects > gpurun_out/ncu_fe.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_fe.ncu-rep > gpurun_out/ncu_fe_summary.txt; cat gpurun_out/ncu_fe_summary.txt
ls -la gpurun_out/prof_fe.ncu-rep
echo done
