timeout 900 python -c "import torch; torch.zeros(1).cuda()"
for i in 1 2; do for cg in 1 2; do echo "== CG=$cg run $i"; OAKE_GEMM_CTA_GROUP=$cg timeout 300 python tools/quick_bench.py --variant 1 --batch 478 --iters 10 2>&1 | grep -E "variant|gemm_(fc1|fc2|qkv|out) "; done; done
