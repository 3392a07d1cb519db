"""The reference-facing CLI at scale: `torchrun --nproc_per_node=N -m
oadp.oake.objects` on a synthetic ON-DISK COCO-shaped dataset of JPEG files (300 proposals per image), once
with the B200-native options (`.decode:gpu .store:packed`) and once the reference's way (`pillow` decode on
DataLoader-style host threads, one `.pth` per image), for N = 1, 2, 4, 8 ranks of one node.  Every run ends with
the NCCL collation of the manifest (`.collate:True`).  Prints one JSON line per (N, mode): images/s and crops/s of
the timed split (max over ranks of the split's wall time), the per-rank spread, and the limiter estimate.

    python tools/bench_cli_scaling.py --images 4096 --gpus 1 2            # on a 2-GPU box
    python tools/bench_cli_scaling.py --images 4096 --gpus 4 8            # on an 8-GPU box
"""
import argparse
import json
import os
import pathlib
import re
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--images', type=int, default=4096)
    ap.add_argument('--proposals', type=int, default=300)
    ap.add_argument('--gpus', type=int, nargs='+', default=[1, 2])
    ap.add_argument('--modes', nargs='+', default=['native', 'reference'])
    ap.add_argument('--workers', type=int, default=2, help='host decode threads per rank (configs/oake/base.py: 2)')
    ap.add_argument('--out', default=str(ROOT / 'gpurun_out' / 'cli_scaling.jsonl'))
    args = ap.parse_args()
    from oadp_b200 import synth
    import torch
    root = pathlib.Path(tempfile.mkdtemp(prefix='oake_cli_'))
    t0 = time.perf_counter()
    ds = synth.write_coco_dataset(root, args.images, seed=9, n_proposals=args.proposals, fmt='jpg',
                                  workers=min(os.cpu_count() or 1, 64))
    # the timed split is `val` (all images); `train` is the first 8 images only (main() always runs both)
    ann = json.loads(pathlib.Path(ds['ann']).read_text())
    small = root / 'instances_small.json'
    small.write_text(json.dumps(dict(images=ann['images'][:8], annotations=[], categories=[])))
    cfg = pathlib.Path(ds['configs']['objects'])
    text = cfg.read_text().splitlines()
    text = [ln.replace(ds['ann'], str(small)) if ln.startswith('train') else ln for ln in text]
    text = [ln.replace('num_workers=2', f'num_workers={args.workers}') for ln in text]
    cfg.write_text('\n'.join(text) + '\nlog = dict(interval=100000)\n')
    nbytes = sum(f.stat().st_size for f in (root / 'images').iterdir())
    print(json.dumps(dict(what='dataset', images=args.images, proposals_per_image=args.proposals, jpeg_mb=round(nbytes / 1e6, 1),
                          seconds=round(time.perf_counter() - t0, 1), host_cores=os.cpu_count(), gpus=torch.cuda.device_count())),
          flush=True)
    out_lines = []
    env = dict(os.environ, OAKE_ALLOW_RANDOM_WEIGHTS='1')
    env.pop('DRY_RUN', None)
    for n in args.gpus:
        if n > torch.cuda.device_count():
            print(json.dumps(dict(what='skipped', n=n, reason=f'only {torch.cuda.device_count()} GPUs on this box')), flush=True)
            continue
        for mode in args.modes:
            out_dir = root / f'out_{mode}_{n}'
            ov = ['.collate:True', f'.val.dataloader.dataset.output_dir::{out_dir}/val',
                  f'.train.dataloader.dataset.output_dir::{out_dir}/train']
            ov += ['.decode:gpu', '.store:packed'] if mode == 'native' else ['.decode:pillow', '.store:pth']
            cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}', '--master-addr',
                   '127.0.0.1', '--master-port', str(29600 + n), '-m', 'oadp.oake.objects', 'scale', str(cfg), '--override'] + ov
            t0 = time.perf_counter()
            r = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=str(ROOT))
            wall = time.perf_counter() - t0
            timings = [json.loads(m) for m in re.findall(r'\[oake-timing\] (\{.*\})', r.stdout + r.stderr)]
            val = [t for t in timings if t['output_dir'].endswith('/val')]
            if r.returncode != 0 or len(val) != n:
                print(json.dumps(dict(what='failed', n=n, mode=mode, rc=r.returncode, tail=(r.stdout + r.stderr)[-800:])), flush=True)
                continue
            import torch as _t
            manifest = _t.load(out_dir / 'val' / 'manifest.pth')
            secs = [t['seconds'] for t in val]
            line = dict(what='cli_objects', n_gpus=n, mode=mode, decode=val[0]['decode'], store=val[0]['store'],
                        images=sum(t['images'] for t in val), crops=sum(t['crops'] for t in val),
                        seconds_max_rank=max(secs), seconds_min_rank=min(secs), images_per_s=round(sum(t['images'] for t in val) / max(secs), 1),
                        crops_per_s=round(sum(t['crops'] for t in val) / max(secs), 1), wall_s_incl_startup=round(wall, 1),
                        manifest_ids=int(manifest['ids'].numel()), manifest_rows=int(manifest['rows'].sum()),
                        host_decode_threads_per_rank=args.workers)
            assert line['manifest_ids'] == args.images and line['manifest_rows'] == line['crops'], line
            out_lines.append(line)
            print(json.dumps(line), flush=True)
            shutil.rmtree(out_dir, ignore_errors=True)
    pathlib.Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    with open(args.out, 'a') as f:
        for ln in out_lines:
            f.write(json.dumps(ln) + '\n')
    shutil.rmtree(root, ignore_errors=True)


if __name__ == '__main__':
    main()
