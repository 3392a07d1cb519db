#!/bin/bash
# Risky-kernel protocol: a short, separately timed GEMM test first; the full round only if it passes.
mkdir -p gpurun_out
timeout 900 python -c "import torch; torch.zeros(1).cuda()"
python -m oadp_b200.build > gpurun_out/build.log 2>&1
echo "== gemm tests (180 s cap)"
timeout 180 python -m pytest tests/test_gpu_kernels.py -q -x -k gemm 2>&1 | tail -15 | tee gpurun_out/t_gemm.log
if grep -q "passed" gpurun_out/t_gemm.log && ! grep -q "failed" gpurun_out/t_gemm.log; then
  bash tools/gpu_round.sh
else
  echo "GEMM tests did not pass: skipping the rest"; nvidia-smi | head -15
fi
