#!/bin/bash
# First GPU run of the text tower (written at the end of round 1 without GPU time): parity, then timing.
mkdir -p gpurun_out
timeout 900 python -c "import torch; torch.zeros(1).cuda()"
python -m oadp_b200.build > gpurun_out/build.log 2>&1
echo "== text tower parity (120 s cap)"
timeout 300 python -m pytest tests/test_gpu_text.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/t_text.log
if grep -q "passed" gpurun_out/t_text.log && ! grep -q "failed\|error" gpurun_out/t_text.log; then
  echo "== memcheck"; timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_text.py -m gpu -q -x -k "1-16-5 or build_prompts" > gpurun_out/text_memcheck.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/text_memcheck.log | head -4
  echo "== bench"; timeout 300 python tools/bench_text.py 2>&1 | tee gpurun_out/bench_text.jsonl | tail -4
fi
