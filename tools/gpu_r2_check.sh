#!/bin/bash
# Round-2 checkpoint: whole GPU test tier, smoke, the bench (library baseline included), optional extras.
mkdir -p gpurun_out
export OAKE_ALLOW_RANDOM_WEIGHTS=1
ts() { date +%H:%M:%S; }
timeout 900 python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name())"
python -m oadp_b200.build > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
if [ -z "$NOTESTS" ]; then
echo "$(ts) == gpu tests"; timeout 1200 python -m pytest tests -m gpu -q -x ${PYTEST_ARGS} 2>&1 | tail -15 | tee gpurun_out/t_gpu.log
echo "$(ts) == smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
fi
if [ -z "$NOBENCH" ]; then
echo "$(ts) == bench"; timeout 900 python bench.py ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
fi
if [ -n "$EXTRA" ]; then echo "$(ts) == extra"; eval "$EXTRA"; fi
echo "$(ts) done"
