#!/bin/bash
# A/B of attention build flags on one box: FLAGSETS="a|b|c" (each a set of -D flags), quick_bench at B = 478.
mkdir -p gpurun_out
timeout 900 python -c "import torch; torch.zeros(1).cuda()"
IFS='|' read -ra SETS <<< "${FLAGSETS}"
for F in "${SETS[@]}"; do
  export OAKE_NVCC_FLAGS="$F"
  python -m oadp_b200.build > gpurun_out/build_ab.log 2>&1 || tail -5 gpurun_out/build_ab.log
  echo "== flags: [$F]"
  for i in 1 2; do timeout 300 python tools/quick_bench.py --variant 1 --batch 478 --iters 10 2>&1 | grep -E "attn_main"; done
  if [ -n "$TESTS" ]; then timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_edges.py tests/test_gpu_encoder.py tests/test_gpu_ref_golden.py -m gpu -q -x 2>&1 | tail -2; fi
done
nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader
