#!/bin/bash
# One gpurun call: GPU test tier, quick bench (+ optional ncu).  Each stage has its own timeout so a
# hung kernel cannot eat the whole lease.
mkdir -p gpurun_out
ts() { date +%H:%M:%S; }
echo "$(ts) warm-up import"; timeout 900 python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name())"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -m oadp_b200.build > gpurun_out/build.log 2>&1
echo "$(ts) == gpu tests"; timeout 900 python -m pytest tests -m gpu -q -x -s 2>&1 | tail -40 | tee gpurun_out/t_gpu.log
if [ -z "$NOBENCH" ]; then
echo "$(ts) == quick bench"; timeout 300 python tools/quick_bench.py --variant 0 --batch 1894 2>&1 | tail -20 | tee gpurun_out/qb0.log
timeout 300 python tools/quick_bench.py --variant 1 --batch 478 2>&1 | tail -20 | tee gpurun_out/qb1.log
fi
if [ -n "$NCU" ]; then
echo "$(ts) == ncu"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:$NCU -s ${NCU_SKIP:-5} -c ${NCU_COUNT:-4} -f -o gpurun_out/prof_$NCU python tools/quick_bench.py --variant ${NCU_VARIANT:-0} --batch ${NCU_BATCH:-1894} --iters 1 > gpurun_out/ncu.log 2>&1; tail -3 gpurun_out/ncu.log
fi
if [ -n "$EXTRA" ]; then echo "$(ts) == extra"; eval "$EXTRA"; fi
echo "$(ts) done"
