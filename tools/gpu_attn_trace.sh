#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -c "import torch; torch.zeros(1).cuda()"
for exp in ${EXPS:-0}; do for mode in ${MODES:-rs}; do echo "== OAKE_ATTN=$mode EXP=$exp"; EXP=$exp OAKE_ATTN=$mode timeout 600 python tools/attn_trace.py 2>&1 | tail -12 | tee gpurun_out/attn_trace_${mode}_$exp.txt; done; done
