#!/bin/bash
# compute-sanitizer passes over the kernel tests (small shapes).
mkdir -p gpurun_out
timeout 900 python -c "import torch; torch.zeros(1).cuda()"
python -m oadp_b200.build > gpurun_out/build.log 2>&1
echo "== ensemble + kernels (plain)"; timeout 300 python -m pytest tests/test_gpu_ensemble.py tests/test_gpu_kernels.py -q -x 2>&1 | tail -2
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_ensemble.py -q -x -k "attention or gemm_tcgen05_plain or layernorm_fold or im2col or ensemble" > gpurun_out/memcheck.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/memcheck.log | head -10
echo "== synccheck"; timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention" > gpurun_out/synccheck.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Barrier|divergent" gpurun_out/synccheck.log | head -10
echo "== memcheck: pipeline (front end + tower, 2 layers)"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck_smoke.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|smoke ok|Invalid" gpurun_out/memcheck_smoke.log | head -5
timeout 120 python tools/bench_next_rows.py 2>&1 | head -3
