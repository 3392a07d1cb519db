#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -c "import torch; torch.zeros(1).cuda()"
python -m oadp_b200.build > gpurun_out/build.log 2>&1
echo "== tests"; timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_encoder.py -q -x 2>&1 | tail -5
echo "== ncu attention_pp"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_pp -s 3 -c 1 -f -o gpurun_out/prof_attn_pp python tools/quick_bench.py --variant 1 --batch 478 --iters 1 > gpurun_out/ncu_attn_pp.log 2>&1; tail -3 gpurun_out/ncu_attn_pp.log
