#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -c "import torch; torch.zeros(1).cuda()"
python -m oadp_b200.build > gpurun_out/build.log 2>&1
echo "== jpeg tests"
timeout 300 python -m pytest tests/test_gpu_jpeg.py tests/test_gpu_edges.py tests/test_gpu_frontend.py tests/test_gpu_cli.py -q -x 2>&1 | tail -6 | tee gpurun_out/t_jpeg.log
echo "== bench"
timeout 500 python tools/bench_jpeg.py 2>&1 | tee gpurun_out/bench_jpeg.jsonl | tail -12
echo "== ncu full: entropy + colour at batch 256"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"jpeg_entropy_par|jpeg_colour|jpeg_idct" -s 33 -c 3 -f -o gpurun_out/prof_jpeg python tools/bench_jpeg.py --decode-only > gpurun_out/ncu_jpeg_full.log 2>&1; tail -2 gpurun_out/ncu_jpeg_full.log
