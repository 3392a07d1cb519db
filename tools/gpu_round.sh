#!/bin/bash
# One gpurun call: kernel tests, encoder parity, quick bench.  Each stage has its own timeout so a
# hung kernel cannot eat the whole lease.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -m oadp_b200.build > gpurun_out/build.log 2>&1
echo "== gemm tests"; timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k gemm 2>&1 | tail -25 | tee gpurun_out/t_gemm.log
echo "== other kernel tests"; timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "not gemm" 2>&1 | tail -25 | tee gpurun_out/t_kern.log
echo "== encoder tests"; timeout 600 python -m pytest tests/test_gpu_encoder.py -q -s 2>&1 | tail -30 | tee gpurun_out/t_enc.log
echo "== quick bench"; timeout 300 python tools/quick_bench.py --variant 0 --batch 1024 2>&1 | tail -20 | tee gpurun_out/qb0.log
timeout 300 python tools/quick_bench.py --variant 1 --batch 256 2>&1 | tail -20 | tee gpurun_out/qb1.log
