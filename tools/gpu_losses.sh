#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -c "import torch; torch.zeros(1).cuda()"
python -m oadp_b200.build > gpurun_out/build.log 2>&1
echo "== losses / ensemble / attention tests"; timeout 600 python -m pytest tests/test_gpu_losses.py tests/test_gpu_ensemble.py tests/test_gpu_kernels.py tests/test_gpu_edges.py -q -x 2>&1 | tail -15
echo "== synccheck (attention, immediate barrier ids)"; timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention_main" > gpurun_out/synccheck.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/synccheck.log | head -4; grep -m2 -A3 "Barrier error" gpurun_out/synccheck.log
