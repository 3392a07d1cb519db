#!/bin/bash
# Last check of a round: whole GPU tier + smoke + both bench arms, then the JPEG micro-benchmark.
bash tools/gpu_full.sh
echo "== jpeg bench"; timeout 400 python tools/bench_jpeg.py 2>&1 | tee gpurun_out/bench_jpeg.jsonl | tail -12
