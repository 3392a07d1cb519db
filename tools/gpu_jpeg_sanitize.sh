#!/bin/bash
# compute-sanitizer passes over the JPEG kernels + per-kernel launch times of a 256-file batch.
mkdir -p gpurun_out
timeout 900 python -c "import torch; torch.zeros(1).cuda()"
python -m oadp_b200.build > gpurun_out/build.log 2>&1
K="pillow or coco_sized or damaged"
echo "== memcheck"; timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_jpeg.py -q -x -k "$K" > gpurun_out/jpeg_memcheck.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/jpeg_memcheck.log | head -8
echo "== synccheck"; timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_jpeg.py -q -x -k "$K" > gpurun_out/jpeg_synccheck.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Barrier|ivergent" gpurun_out/jpeg_synccheck.log | head -8
echo "== racecheck (shared memory)"; timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_jpeg.py -q -x -k "coco_sized" > gpurun_out/jpeg_racecheck.log 2>&1; echo "rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/jpeg_racecheck.log | head -8
echo "== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jpeg --csv --log-file gpurun_out/jpeg_launches.csv python tools/bench_jpeg.py --decode-only > gpurun_out/ncu_jpeg.log 2>&1; tail -2 gpurun_out/ncu_jpeg.log; wc -l gpurun_out/jpeg_launches.csv
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
