#!/bin/bash
# 2-GPU checks: torchrun bench, the CLI test tier, and a 2-rank `oadp.oake.objects` run with packed output.
mkdir -p gpurun_out
export OAKE_ALLOW_RANDOM_WEIGHTS=1  # no CLIP checkpoint offline: the CLI needs the explicit opt-in
timeout 900 python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.device_count())"
python -m oadp_b200.build > gpurun_out/build.log 2>&1
echo "== 2-GPU bench (torchrun)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; grep "^{" gpurun_out/bench_2gpu.json | cut -c1-260; grep -i "warn\|error" gpurun_out/bench_2gpu.err | head -5
if [ -n "$REFARM" ]; then
echo "== 2-GPU reference arm (rank 0 only)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_2gpu.json 2> gpurun_out/bench_ref_2gpu.err; tail -c 300 gpurun_out/bench_ref_2gpu.json
fi
echo "== CLI tests"; timeout 600 python -m pytest tests/test_gpu_cli.py -q -x 2>&1 | tail -3
echo "== 2-rank CLI (objects, packed store)"
timeout 600 python - <<'PY'
import pathlib, subprocess, sys, tempfile
sys.path.insert(0, '.')
from oadp_b200 import synth, store
root = pathlib.Path(tempfile.mkdtemp())
ds = synth.write_coco_dataset(root, 6, seed=3, n_proposals=30)
cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
       '--master-port', '29513', '-m', 'oadp.oake.objects', 't', ds['configs']['objects'], '--override', '.store:packed']
r = subprocess.run(cmd, capture_output=True, text=True)
print(r.stdout[-400:], r.stderr[-400:])
for split in ('val', 'train'):
    d = pathlib.Path(ds['root']) / 'oake' / 'objects' / split
    s = store.PackedStore(str(d))
    print(split, len(s), sorted(p.name for p in d.glob('*.idx.json')))
    assert sorted(s) == [f'{i:012d}' for i in ds['ids']]
print('2-rank packed CLI ok')
PY
