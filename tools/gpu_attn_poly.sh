#!/bin/bash
# A/B on one box: share of the exponentials on the FMA pipe (OAKE_ATTN_POLY = every n-th pair; 0 = none)
mkdir -p gpurun_out
timeout 900 python -c "import torch; torch.zeros(1).cuda()"
for P in ${POLYS:-0 4 3 2}; do
  export OAKE_NVCC_FLAGS="-DOAKE_ATTN_POLY=$P"
  python -m oadp_b200.build > gpurun_out/build_$P.log 2>&1 || tail -5 gpurun_out/build_$P.log
  for mode in ${MODES:-rs cs}; do echo "== POLY=$P OAKE_ATTN=$mode"; OAKE_ATTN=$mode timeout 300 python tools/quick_bench.py --variant 1 --batch 478 --iters 10 2>&1 | grep -E "attn_main"; done
  if [ "$P" != "0" ]; then timeout 200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_edges.py tests/test_gpu_encoder.py -q -x 2>&1 | tail -2; fi
done
