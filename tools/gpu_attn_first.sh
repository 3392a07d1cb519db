#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -c "import torch; torch.zeros(1).cuda()"
python -m oadp_b200.build > gpurun_out/build.log 2>&1
echo "== attention tests, tcgen05 (180 s cap)"
timeout 180 python -m pytest tests/test_gpu_kernels.py -q -x -k attention 2>&1 | tail -25 | tee gpurun_out/t_attn.log
if grep -q "passed" gpurun_out/t_attn.log && ! grep -q "failed" gpurun_out/t_attn.log; then
  bash tools/gpu_round.sh
else
  echo "tcgen05 attention did not pass"; nvidia-smi | head -12
fi
