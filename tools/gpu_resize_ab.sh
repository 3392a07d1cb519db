#!/bin/bash
# A/B of resize build flags on one box (FLAGSETS="a|b"), tools/bench_resize.py on an otherwise idle GPU.
mkdir -p gpurun_out
export OAKE_ALLOW_RANDOM_WEIGHTS=1
timeout 900 python -c "import torch; torch.zeros(1).cuda()"
IFS='|' read -ra SETS <<< "${FLAGSETS}"
for F in "${SETS[@]}"; do
  export OAKE_NVCC_FLAGS="$F"
  python -m oadp_b200.build > gpurun_out/build_ab.log 2>&1 || tail -5 gpurun_out/build_ab.log
  echo "== flags: [$F]"
  for i in 1 2; do timeout 300 python tools/bench_resize.py; done
done
