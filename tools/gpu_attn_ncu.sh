#!/bin/bash
# Full ncu capture (source-level stall samples) of one attention launch of the objects tower.
mkdir -p gpurun_out
timeout 900 python -c "import torch; torch.zeros(1).cuda()"
python -m oadp_b200.build > gpurun_out/build.log 2>&1
for mode in ${MODES:-rs}; do
OAKE_ATTN=$mode timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_cs -s 3 -c 1 -f -o gpurun_out/prof_attn_$mode python tools/quick_bench.py --variant 1 --batch 478 --iters 1 > gpurun_out/ncu_attn_$mode.log 2>&1; tail -2 gpurun_out/ncu_attn_$mode.log
ncu -i gpurun_out/prof_attn_$mode.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/attn_${mode}_source.csv 2>/dev/null
ncu -i gpurun_out/prof_attn_$mode.ncu-rep --page raw --csv > gpurun_out/attn_${mode}_raw.csv 2>/dev/null
done
ls -la gpurun_out | head -20
