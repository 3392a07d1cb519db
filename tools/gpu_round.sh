#!/bin/bash
# One gpurun call: kernel tests, encoder parity, quick bench (+ optional ncu).  Each stage has its own
# timeout so a hung kernel cannot eat the whole lease.
mkdir -p gpurun_out
ts() { date +%H:%M:%S; }
echo "$(ts) warm-up import"; timeout 900 python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name())"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -m oadp_b200.build > gpurun_out/build.log 2>&1
echo "$(ts) == kernel tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x 2>&1 | tail -25 | tee gpurun_out/t_kern.log
echo "$(ts) == encoder tests"; timeout 600 python -m pytest tests/test_gpu_encoder.py -q -s 2>&1 | tail -30 | tee gpurun_out/t_enc.log
echo "$(ts) == quick bench"; timeout 300 python tools/quick_bench.py --variant 0 --batch 1894 2>&1 | tail -20 | tee gpurun_out/qb0.log
timeout 300 python tools/quick_bench.py --variant 1 --batch 478 2>&1 | tail -20 | tee gpurun_out/qb1.log
if [ -n "$NCU" ]; then
echo "$(ts) == ncu"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 5 -c 4 -f -o gpurun_out/prof_gemm python tools/quick_bench.py --variant 0 --batch 1894 --iters 1 > gpurun_out/ncu.log 2>&1; tail -3 gpurun_out/ncu.log
fi
echo "$(ts) done"
