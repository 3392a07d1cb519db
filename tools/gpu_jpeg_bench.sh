#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -c "import torch; torch.zeros(1).cuda()"
python -m oadp_b200.build > gpurun_out/build.log 2>&1
timeout 600 python tools/bench_jpeg.py 2>&1 | tee gpurun_out/bench_jpeg.jsonl | tail -14
if [ -n "$NCU" ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jpeg -c 30 --csv --log-file gpurun_out/jpeg_launches.csv python tools/bench_jpeg.py --decode-only > gpurun_out/ncu_jpeg.log 2>&1
  tail -3 gpurun_out/ncu_jpeg.log
fi
