#!/usr/bin/env python
"""OAKE crops/sec benchmark (BASELINE.json metric) -- one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload oake|globals|blocks|objects|classifier] [--images I]

A step = one pass of the OAKE hot path over one batch of I synthetic COCO-shaped images per GPU:
`oake` runs what the north star names -- globals (1 crop/image) + blocks (pyramid grid, 17..39
crops/image) + objects (300 proposals/image, ~294 survive the min_wh filter) -- on the same images.

  value     whole-job crops/s, inputs (uint8 images + descriptors) already resident in HBM; timed
            with CUDA events around exactly K steps, max over ranks.
  e2e       the same metric through the public API (`OakePipeline.encode_*`) with HOST images:
            pinned staging, H2D, kernels, D2H of the fp16 embeddings all inside the timed region.
  roofline  dominant kernel class (tcgen05 GEMM, c_fc instantiation): algorithmic FLOPs per launch /
            mean launch duration from CUDA events on the launch stream, against MEASURED_PEAKS.json.
  cpu_baseline / --impl reference: the CPU oracle (oracle/: PIL + torch fp32, all host threads),
            the only stand-in for the reference's CPU path that can run here, on a
            bounded sample of the same workload.
Multi-GPU: one process per GPU (torchrun), images sharded by rank, no data-path collective
(the reference's OAKE issues none either, oadp/oake/base.py:84-88); weak scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib
import statistics
import subprocess
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

GFLOP_T50 = 8.818  # algorithmic GFLOP per crop, T=50
GFLOP_T197 = 33.552  # T=197 + side stream with shared K/V
WEIGHT_SEED = 1234
PROPOSALS_PER_IMAGE = 300
# roofline.traffic: dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel class, averaged
# over the launches of ONE step of this same command under `ncu --set full` restricted to the class's NVTX range
# (tools/gpu_profile.sh -> tools/ncu_traffic.py -> profiles/ncu_traffic.json).  Used only if the capture saw the
# same average rows per launch as this run (same step mix); otherwise null.
NCU_TRAFFIC_FILE = ROOT / 'profiles' / 'ncu_traffic.json'


def ncu_traffic(kernel_class: str, flops_per_launch: float):
    try:
        ent = json.loads(NCU_TRAFFIC_FILE.read_text())[kernel_class]
        if abs(ent['flops_per_launch'] - flops_per_launch) <= 0.01 * flops_per_launch:
            return float(ent['dram_bytes_per_launch'])
    except Exception:
        pass
    return None


def peaks():
    f = ROOT / 'MEASURED_PEAKS.json'
    if f.exists():
        d = json.loads(f.read_text())
        return dict(tflops_sustained=d['bf16_tflops_sustained'], tflops_burst=d['bf16_tflops'], hbm_gbs=d['hbm_gbs'],
                    source='measured')
    return dict(tflops_sustained=1400.0, tflops_burst=1590.0, hbm_gbs=6650.0, source='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int) -> None:
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(gpu_index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(',')]))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['no samples'])
        sm = [float(r[1]) for r in rows]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[5 + i].lower().startswith('active') for r in rows)]
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=float(rows[0][2]), reasons=reasons,
                    power_w_max=max(float(r[3]) for r in rows), samples=len(rows))


def make_inputs(n_images: int, rank: int):
    from oadp_b200 import synth
    imgs = synth.images(n_images, seed=1000 + rank)
    props = [synth.proposals(im.shape[1], im.shape[0], PROPOSALS_PER_IMAGE, seed=7000 + 97 * rank + i)
             for i, im in enumerate(imgs)]
    return imgs, props


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the CPU oracle on a bounded sample of the same workload
# ------------------------------------------------------------------------------------------------
def cpu_reference(workload: str, n_images: int, steps: int, warmup: int, sample_objects: int = 96):
    import PIL.Image
    import torch
    from oracle import frontend as ofe
    from oracle import vit
    torch.set_num_threads(os.cpu_count() or 1)
    imgs, props = make_inputs(n_images, 0)
    counts = workload_counts(imgs, props)
    p = vit.init_visual_params(WEIGHT_SEED)
    p197 = vit.objects_surgery(p)
    img = PIL.Image.fromarray(imgs[0])
    prop = torch.from_numpy(props[0][:sample_objects + 4])

    def one_step():
        t = {}
        n = {}
        t0 = time.perf_counter()
        px = ofe.globals_preprocess(img).unsqueeze(0)  # B=1 as globals.py:54
        vit.normalize_half(vit.encode_image(p, px))
        t['globals'], n['globals'] = time.perf_counter() - t0, 1
        t0 = time.perf_counter()
        b = ofe.blocks_preprocess(img)  # whole image batch as blocks.py:126-129
        vit.normalize_half(vit.encode_image(p, b.blocks))
        t['blocks'], n['blocks'] = time.perf_counter() - t0, b.blocks.shape[0]
        t0 = time.perf_counter()
        o = ofe.objects_preprocess(img, prop)
        vit.normalize_half(vit.encode_objects(p197, o.objects, o.masks))
        t['objects'], n['objects'] = time.perf_counter() - t0, o.objects.shape[0]
        return t, n

    kinds = ['globals', 'blocks', 'objects'] if workload == 'oake' else [workload]
    for _ in range(warmup):
        one_step()
    per_crop = {k: [] for k in kinds}
    t_all0 = time.perf_counter()
    for _ in range(steps):
        t, n = one_step()
        for k in kinds:
            per_crop[k].append(t[k] / n[k])
    wall = time.perf_counter() - t_all0
    sec_per_crop = {k: statistics.median(v) for k, v in per_crop.items()}
    total_crops = sum(counts[k] for k in kinds)
    total_sec = sum(counts[k] * sec_per_crop[k] for k in kinds)
    value = total_crops / total_sec
    sample = (f'1 image 640x480 per step: 1 global (B=1) + {n["blocks"]} blocks (one batch) + {n["objects"]} '
              f'objects (one batch), PIL front end in-process; per-kind sec/crop (median of {steps}) re-weighted to '
              f'the step mix {counts}')
    return dict(value=value, sec_per_crop=sec_per_crop, sample=sample, cores=torch.get_num_threads(), wall_s=wall,
                ms_per_step=wall / max(steps, 1) * 1e3)


def workload_counts(imgs, props):
    from oadp_b200 import frontend
    blocks = sum(1 + len(frontend.blocks_plan(im.shape[1], im.shape[0]).cells) for im in imgs)
    objects = sum(frontend.objects_plan(p, (im.shape[1], im.shape[0])).bboxes.shape[0] for im, p in zip(imgs, props))
    return dict(globals=len(imgs), blocks=blocks, objects=objects)



# ------------------------------------------------------------------------------------------------
# library_baseline: the strongest existing GPU implementation on the same box
# ------------------------------------------------------------------------------------------------
def library_baseline(counts, kinds, steps: int, warmup: int, dev):
    """The oracle tower (oracle/vit.py, unchanged code) in fp16 on THIS GPU under PyTorch eager: cuDNN
    conv, cuBLAS linears, ATen LayerNorm / softmax -- what the reference's `model.half().cuda()` runs --
    and the same with the attention product routed through `F.scaled_dot_product_attention` (flash /
    cuDNN kernels), the better of the two reported.  A baseline leg like `cpu_baseline`: the only other
    place this file executes `oracle/`.  Inputs are synthetic CLIP-normalised pixel tensors of the step's
    shapes (the tower's time does not depend on pixel values); all crops of a kind travel as ONE batch
    (objects in chunks of 512, objects.py:323-327) -- more favourable to the library than the reference's
    own B=1 / per-image batches.  `device`: pixels resident in HBM as fp16.  `e2e`: pinned fp32 pixels
    (what the reference's DataLoader hands over, 602 KB per crop) -> H2D -> tower -> fp16 rows -> host."""
    import torch
    import torch.nn.functional as F
    from oracle import vit
    p16 = {k: v.to(dev, torch.float16) for k, v in vit.init_visual_params(WEIGHT_SEED).items()}
    p197 = vit.objects_surgery(p16)
    g = torch.Generator(device='cpu').manual_seed(5)
    eager_attend = vit._attend

    def sdpa_attend(q, k, v, bias=None):
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=None if bias is None else bias.to(q.dtype).expand(
            q.shape[0], q.shape[1], q.shape[2], k.shape[2]))
        b, h, nq, dh = o.shape
        return o.transpose(1, 2).reshape(b, nq, h * dh)

    # `vit.encode_image` / `vit.encode_objects` with their `.float()` casts removed -- the same block, side
    # stream and head functions of oracle/vit.py, driven in the parameters' dtype (fp16)
    @torch.no_grad()
    def encode_image(px):
        x = vit.embed(p16, px, stride=vit.PATCH, padding=0)
        for i in range(vit.num_layers(p16)):
            x = vit._block(x, p16, i)
        return vit.head(p16, x[:, 0])

    @torch.no_grad()
    def encode_objects(px, masks):
        x = vit.embed(p197, px, stride=vit.PATCH // 2, padding=(vit.PATCH - 1) // 2)
        bias = vit.mask_to_bias(masks)[:, None, None, :].to(x.dtype)
        y = x[:, :1]
        for i in range(vit.num_layers(p197)):
            pre = f'transformer.resblocks.{i}.'
            z = vit._ln(torch.cat([x[:, 1:], y], dim=1), p197, pre + 'ln_1')
            qkv = z @ p197[pre + 'attn.in_proj_weight'].T + p197[pre + 'attn.in_proj_bias']
            q, k, v = qkv.split(vit.WIDTH, dim=-1)
            o = vit._attend(vit._heads(q[:, -1:]), vit._heads(k), vit._heads(v), bias)
            y = y + o @ p197[pre + 'attn.out_proj.weight'].T + p197[pre + 'attn.out_proj.bias']
            y = y + vit._mlp(vit._ln(y, p197, pre + 'ln_2'), p197, pre)
            x = vit._block(x, p197, i)
        return vit.head(p197, y[:, 0])

    def run(kind, px, masks):
        if kind == 'objects':
            out = [vit.normalize_half(encode_objects(px[s:s + 512], masks[s:s + 512]))
                   for s in range(0, px.shape[0], 512)]
            return torch.cat(out)
        return vit.normalize_half(encode_image(px))

    def timed(fn, n):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e-3 / n

    out = {}
    try:
        for kind in kinds:
            n = counts[kind]
            host = torch.randn(n, 3, 224, 224, generator=g).pin_memory()
            masks_host = (torch.rand(n, 1, 14, 14, generator=g) < 0.5).float().pin_memory()
            px = host.to(dev, torch.float16)
            masks = masks_host.to(dev, torch.float16)

            def device_fn():
                return run(kind, px, masks)

            def e2e_fn():
                x = host.to(dev, non_blocking=True).half()
                m = masks_host.to(dev, non_blocking=True).half()
                return run(kind, x, m).cpu()

            res = {}
            for name, attend in (('eager', eager_attend), ('sdpa', sdpa_attend)):
                vit._attend = attend
                n_steps = max(2, min(steps, 5))
                res[name] = dict(device=n / timed(device_fn, n_steps), e2e=n / timed(e2e_fn, n_steps))
            best = max(res, key=lambda k: res[k]['device'])
            out[kind] = dict(crops=n, device=res[best]['device'], e2e=max(r['e2e'] for r in res.values()), best=best,
                             eager=res['eager'], sdpa=res['sdpa'])
            del host, masks_host, px, masks
            torch.cuda.empty_cache()
    finally:
        vit._attend = eager_attend
    return out

# ------------------------------------------------------------------------------------------------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='oake', choices=['oake', 'globals', 'blocks', 'objects', 'classifier'])
    ap.add_argument('--images', type=int, default=8, help='images per step per GPU')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-library-baseline', action='store_true')
    args = ap.parse_args()
    warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    pk = peaks()
    cfg = dict(workload=(f'{args.workload}: globals+blocks+objects over the same images' if args.workload == 'oake'
                         else args.workload), images_per_step_per_gpu=args.images,
               proposals_per_image=PROPOSALS_PER_IMAGE, image_sizes='COCO-train2017-like mix (SURVEY 8d)',
               tower='CLIP ViT-B/32 visual, 224^2, seeded random weights', parallelism=f'image-sharded x{world}')

    if args.impl == 'reference':
        if rank != 0:
            return
        steps = max(1, min(args.steps, 3))
        ref = cpu_reference(args.workload, args.images, steps, min(args.warmup, 1))
        line = dict(metric='OAKE crops/sec (ViT-B/32, 224^2, synthetic COCO proposals)', impl='reference',
                    value=ref['value'], unit='crops/s', n_gpus=args.gpus, steps=steps, warmup=min(args.warmup, 1),
                    ms_per_step=ref['ms_per_step'], higher_is_better=True, scaling='weak', vs_baseline=None,
                    dtype='f32', data='synthetic', config=cfg,
                    cpu_baseline=dict(value=ref['value'], unit='crops/s', cores=ref['cores'], kind='port',
                                      sample=ref['sample'], sec_per_crop=ref['sec_per_crop']),
                    e2e=dict(value=ref['value'], unit='crops/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl ours needs a CUDA device: the OAKE hot path has no CPU fallback')
    torch.cuda.set_device(local_rank % torch.cuda.device_count())
    if args.workload == 'classifier':
        # BASELINE configs 4-5 in isolation (mmdet is absent): the cosine classifier at the test-time
        # shape (N = 1000 RoIs, in = 1024, K = 66 / 1204) and the LVIS training shape (fwd + bwd, bf16 inputs), each
        # next to PyTorch eager on the same GPU.  value = RoIs/s of the first case (OV-COCO inference, fp16 RoI features).
        if rank != 0:
            return
        from oadp_b200 import build
        build.build()
        sys.path.insert(0, str(ROOT / 'tools'))
        import bench_classifier
        cases = [bench_classifier.run(*c) for c in bench_classifier.CASES]
        head = cases[0]
        print(json.dumps(dict(metric='cosine classifier RoIs/sec (N=1000, in=1024, K=66, inference, fused call)',
                              value=head['n'] / (head['us'] * 1e-6), unit='RoIs/s', n_gpus=1, steps=100, warmup=10,
                              ms_per_step=head['us'] * 1e-3, higher_is_better=True, scaling='weak', vs_baseline=None,
                              dtype='f16', data='synthetic', config=dict(workload='classifier', cases='BASELINE configs 4-5 '
                                                                         'in isolation (SURVEY 8d-4/5)'),
                              roofline=dict(bound='hbm', achieved=head['gbs'], peak=pk['hbm_gbs'], unit='GB/s',
                                            frac=head['gbs'] / pk['hbm_gbs'], traffic=None,
                                            note='launch / latency bound at this size: 4.45 MB of algorithmic traffic is '
                                                 '0.7 us of HBM time (BASELINE.md section 3)'),
                              cases=cases)))
        return
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', torch.cuda.current_device()))
    from oadp_b200 import build, synth
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    from oadp_b200.model import OakeEngine
    from oadp_b200.pipeline import OakePipeline

    dev = torch.device('cuda', torch.cuda.current_device())
    engine = OakeEngine(synth.visual_params(WEIGHT_SEED), dev)
    imgs, props = make_inputs(args.images, rank)
    counts = workload_counts(imgs, props)
    kinds = ['globals', 'blocks', 'objects'] if args.workload == 'oake' else [args.workload]
    crops_per_step = sum(counts[k] for k in kinds)
    gflop_per_step = sum(counts[k] * (GFLOP_T197 if k == 'objects' else GFLOP_T50) for k in kinds)
    cfg['crops_per_step_per_gpu'] = {k: counts[k] for k in kinds}
    cfg['l2_policy'] = ('no flush needed: every step streams > 2 GB of activations through HBM, far above the '
                        '126 MB L2')

    # one pipeline per kind so that each keeps its inputs resident in HBM
    pipes = {k: OakePipeline(engine) for k in kinds}

    def plan(k):
        if k == 'globals':
            return pipes[k].plan_globals(imgs)
        if k == 'blocks':
            return pipes[k].plan_blocks(imgs)[0]
        return pipes[k].plan_objects(imgs, props)[0]

    jobs = {}
    for k in kinds:
        jobs[k] = pipes[k].stage(*plan(k))
        pipes[k].upload(jobs[k])
    torch.cuda.synchronize()

    def device_step():
        return [pipes[k].launch(jobs[k]) for k in kinds]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------------------------------------------------------- value: device-resident
    for _ in range(warmup):
        device_step()
    barrier()
    launches0 = engine.launch_count() + sum(p.frontend_launches for p in pipes.values())
    sampler = ClockSampler(torch.cuda.current_device()) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        device_step()
    ev1.record()
    barrier()
    t_wall1 = time.perf_counter()
    ms = ev0.elapsed_time(ev1)
    launches = engine.launch_count() + sum(p.frontend_launches for p in pipes.values()) - launches0
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = crops_per_step * world * args.steps / (ms * 1e-3)

    # ---------------------------------------------------------------- e2e: public API, host buffers
    # `submit_*` is the asynchronous form of `encode_*` that the validators' loop uses: the host
    # stages and uploads batch i+1 while the GPU works on batch i; every step still does its own
    # pinned staging copy, H2D of the raw images + descriptors, all kernels and the D2H of its
    # embeddings, and every result is materialised on the host before the clock stops.
    def api_submit():
        out = []
        for k in kinds:
            if k == 'globals':
                out.append(pipes[k].submit_globals(imgs))
            elif k == 'blocks':
                out.append(pipes[k].submit_blocks(imgs))
            else:
                out.append(pipes[k].submit_objects(imgs, props))
        return out

    def api_loop(n):
        pend = None
        for _ in range(n):
            new = api_submit()
            if pend is not None:
                for t_ in pend:
                    t_.result()
            pend = new
        for t_ in pend:
            t_.result()

    h2d = d2h = 0
    api_loop(2)
    for k in kinds:
        h2d += pipes[k].h2d_bytes
        d2h += pipes[k].d2h_bytes
    barrier()
    e2e_steps = max(3, args.steps)
    t0 = time.perf_counter()
    api_loop(e2e_steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = crops_per_step * world * e2e_steps / float(t.item())
    # the device-resident loop below re-uses the slot the last submission filled
    for k in kinds:
        jobs[k] = pipes[k].stage(*plan(k))
        pipes[k].upload(jobs[k])
    torch.cuda.synchronize()

    # ---------------------------------------------------------------- roofline: per-class CUDA events
    engine.profile(True)
    for pp in pipes.values():
        pp.profile_frontend = True
    for _ in range(2):
        device_step()
    prof = engine.profile_collect()
    engine.profile(False)
    prof['frontend_im2col'] = prof.pop('frontend')
    for pp in pipes.values():
        pp.profile_frontend = False
        for name, v in pp.collect_frontend_profile().items():
            d = prof.setdefault('frontend_' + name, dict(ms=0.0, flops=0.0, launches=0))
            d['ms'] += v['ms']
            d['launches'] += v['launches']
    gemm = {k: v for k, v in prof.items() if k.startswith('gemm_') and v['launches']}
    dom = max(gemm, key=lambda k: gemm[k]['ms']) if gemm else None
    tot_ms = sum(v['ms'] for v in prof.values()) or 1.0
    roof = None
    if dom:
        d = gemm[dom]
        achieved = d['flops'] / (d['ms'] * 1e-3) / 1e12
        all_ach = sum(v['flops'] for v in gemm.values()) / (sum(v['ms'] for v in gemm.values()) * 1e-3) / 1e12
        roof = dict(bound='tensor', kernel=f'gemm_tcgen05_kernel ({dom})', achieved=achieved, peak=pk['tflops_sustained'],
                    unit='TFLOP/s', frac=achieved / pk['tflops_sustained'],
                    traffic=ncu_traffic(dom, d['flops'] / d['launches']),
                    peak_source=f"{pk['source']} bf16 cuBLAS, sustained (kernel timed inside a long step)",
                    flops_per_launch=d['flops'] / d['launches'], us_per_launch=d['ms'] / d['launches'] * 1e3,
                    share_of_step=d['ms'] / tot_ms,
                    all_gemm=dict(achieved=all_ach, frac=all_ach / pk['tflops_sustained'],
                                  share_of_step=sum(v['ms'] for v in gemm.values()) / tot_ms),
                    step=dict(achieved=value / world * (gflop_per_step / crops_per_step) / 1e3,
                              frac=value / world * (gflop_per_step / crops_per_step) / 1e3 / pk['tflops_sustained'],
                              note='whole step: crops/s x algorithmic GFLOP/crop (8.818 T50, 33.552 T197)'),
                    classes={k: dict(ms_per_step=v['ms'] / 2, share=v['ms'] / tot_ms,
                                     tflops=(v['flops'] / (v['ms'] * 1e-3) / 1e12) if v['flops'] and v['ms'] else 0.0)
                             for k, v in prof.items() if v['launches']})

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------------------------------------------------------- library baseline (N=1, rank 0)
    library = None
    if not args.no_library_baseline and world == 1:
        def ours_kind(k, n):
            for _ in range(2):
                pipes[k].launch(jobs[k])
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                pipes[k].launch(jobs[k])
            b.record()
            torch.cuda.synchronize()
            dev_cps = counts[k] * n / (a.elapsed_time(b) * 1e-3)
            sub = dict(globals=lambda: pipes[k].submit_globals(imgs), blocks=lambda: pipes[k].submit_blocks(imgs),
                       objects=lambda: pipes[k].submit_objects(imgs, props))[k]
            sub().result()
            t0_ = time.perf_counter()
            pend = None
            for _ in range(n):
                new = sub()
                if pend is not None:
                    pend.result()
                pend = new
            pend.result()
            e2e_cps = counts[k] * n / (time.perf_counter() - t0_)
            jobs[k] = pipes[k].stage(*plan(k))
            pipes[k].upload(jobs[k])
            torch.cuda.synchronize()
            return dev_cps, e2e_cps

        lib = library_baseline(counts, kinds, args.steps, 2, dev)
        per = {}
        for k in kinds:
            o_dev, o_e2e = ours_kind(k, max(3, min(args.steps, 10)))
            per[k] = dict(crops_per_step=counts[k], ours=o_dev, library=lib[k]['device'], ratio=o_dev / lib[k]['device'],
                          ours_e2e=o_e2e, library_e2e=lib[k]['e2e'], ratio_e2e=o_e2e / lib[k]['e2e'],
                          library_best=lib[k]['best'], library_eager=lib[k]['eager'], library_sdpa=lib[k]['sdpa'])
        lib_step_s = sum(counts[k] / lib[k]['device'] for k in kinds)
        lib_e2e_s = sum(counts[k] / lib[k]['e2e'] for k in kinds)
        library = dict(kind='torch %s eager fp16 on the same GPU: oracle/vit.py blocks (cuDNN conv, cuBLAS linears, '
                            'ATen LayerNorm/softmax or SDPA), all crops of a kind in one batch' % torch.__version__,
                       value=crops_per_step / lib_step_s, unit='crops/s', e2e=crops_per_step / lib_e2e_s,
                       ratio=value / (crops_per_step / lib_step_s), ratio_e2e=e2e_value / (crops_per_step / lib_e2e_s),
                       front_end='none: pixel tensors are given (the reference prepares them with PIL on the host); '
                                 'ours includes crop + resize + normalise from the uint8 images',
                       per_workload=per)

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        ref = cpu_reference(args.workload, args.images, 3, 1)
        cpu = dict(value=ref['value'], unit='crops/s', cores=ref['cores'], kind='port', sample=ref['sample'],
                   sec_per_crop=ref['sec_per_crop'])

    line = dict(metric='OAKE crops/sec (ViT-B/32, 224^2, synthetic COCO proposals)', value=value, unit='crops/s',
                n_gpus=world, steps=args.steps, warmup=warmup, ms_per_step=ms / args.steps, higher_is_better=True,
                scaling='weak', vs_baseline=None, dtype='f16', data='synthetic', config=cfg, roofline=roof,
                cpu_baseline=cpu, library_baseline=library,
                e2e=dict(value=e2e_value, unit='crops/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                         steps=e2e_steps),
                gpu_launches=launches, clocks=clocks)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
