#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -c "import torch; torch.zeros(1).cuda()"
python -m oadp_b200.build > gpurun_out/build.log 2>&1
timeout 200 python -m pytest tests/test_gpu_jpeg.py tests/test_gpu_frontend.py -q -x 2>&1 | tail -3 | tee gpurun_out/t_jpeg.log
timeout 300 python tools/bench_jpeg.py --e2e-only 2>&1 | tee gpurun_out/bench_jpeg_e2e.jsonl | tail -8
