timeout 900 python -c "import torch; torch.zeros(1).cuda()"
python -m oadp_b200.build > /dev/null 2>&1
echo "== tests (tc attention for T197)"; timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_encoder.py -q -x 2>&1 | tail -3
for mode in tc mma; do echo "== OAKE_ATTN=$mode"; OAKE_ATTN=$mode timeout 300 python tools/quick_bench.py --variant 1 --batch 478 --iters 10 2>&1 | grep -E "variant|attn_"; OAKE_ATTN=$mode timeout 300 python tools/quick_bench.py --variant 0 --batch 1894 --iters 10 2>&1 | grep -E "variant|attn_"; done
