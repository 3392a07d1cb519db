#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -c "import torch; torch.zeros(1).cuda()"
python -m oadp_b200.build > gpurun_out/build.log 2>&1
echo "== attention tests (120 s cap)"
timeout 120 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_edges.py -q -x -k attention 2>&1 | tail -15 | tee gpurun_out/t_attn.log
if grep -q "passed" gpurun_out/t_attn.log && ! grep -q "failed" gpurun_out/t_attn.log; then
  echo "== encoder / golden tests"; timeout 300 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_ref_golden.py tests/test_gpu_edges.py -q -x 2>&1 | tail -4
  for mode in rs cs; do echo "== OAKE_ATTN=$mode"; OAKE_ATTN=$mode timeout 300 python tools/quick_bench.py --variant 1 --batch 478 --iters 10 2>&1 | grep -E "variant|attn_"; done
  if [ -n "$NCU" ]; then
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_cs -s 3 -c 1 -f -o gpurun_out/prof_attn_cs python tools/quick_bench.py --variant 1 --batch 478 --iters 1 > gpurun_out/ncu_attn_cs.log 2>&1; tail -2 gpurun_out/ncu_attn_cs.log
  fi
else
  echo "attention tests did not pass"; nvidia-smi | head -15
fi
