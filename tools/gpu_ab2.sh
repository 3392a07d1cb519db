#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -c "import torch; torch.zeros(1).cuda()"
python -m oadp_b200.build > gpurun_out/build.log 2>&1
echo "== gpu tests (kernels, encoder)"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_encoder.py tests/test_gpu_ref_golden.py -q -x 2>&1 | tail -3
echo "== quick bench"; timeout 300 python tools/quick_bench.py --variant 1 --batch 478 --iters 10 2>&1 | grep -E "variant|gemm_|attn_main"
timeout 300 python tools/quick_bench.py --variant 0 --batch 1894 --iters 10 2>&1 | grep -E "variant|gemm_|attn_main"
echo "== ncu gemm (qkv, out, fc1, fc2)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 5 -c 4 -f -o gpurun_out/prof_gemm python tools/quick_bench.py --variant 1 --batch 478 --iters 1 > gpurun_out/ncu_gemm.log 2>&1; tail -1 gpurun_out/ncu_gemm.log
echo "== ncu resize"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resize_u8 -s 1 -c 1 -f -o gpurun_out/prof_resize python bench.py --steps 1 --warmup 1 --no-cpu-baseline --images 2 --workload objects > gpurun_out/ncu_resize.log 2>&1; tail -1 gpurun_out/ncu_resize.log
