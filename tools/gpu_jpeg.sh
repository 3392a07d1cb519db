#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -c "import torch; torch.zeros(1).cuda()"
python -m oadp_b200.build > gpurun_out/build.log 2>&1
echo "== jpeg tests"
timeout 300 python -m pytest tests/test_gpu_jpeg.py -q -x 2>&1 | tail -25 | tee gpurun_out/t_jpeg.log
if grep -q "passed" gpurun_out/t_jpeg.log && ! grep -q "failed\|error" gpurun_out/t_jpeg.log; then
  echo "== bench"
  timeout 500 python tools/bench_jpeg.py 2>&1 | tee gpurun_out/bench_jpeg.jsonl | tail -14
fi
