"""Image-level OAKE pipeline: uint8 images (+ proposals) in, the tensors the reference stores out.

This is the public API the `oadp.oake.*` validators and `bench.py` call.  Per batch of images it
  1. copies the raw uint8 images and a few KB of job descriptors to the GPU (one pinned staging
     buffer each, async H2D); an image may also arrive as a still-compressed `jpeg.JpegSource`, in
     which case the file travels instead and `oake_jpeg_decode` produces the pixels on the GPU, on a
     side stream, while the previous batch is still in the tower,
  2. runs the Pillow-exact resize kernel for every crop / pyramid level (`oake_resize_u8`),
  3. runs the tower on the uint8 crops in SM-friendly chunks (`oake_encode_crops_u8`),
  4. copies the fp16 embeddings back.
Results per image have exactly the reference's layouts:
  globals: f16 (512,)                                     oadp/oake/globals.py:57-59
  blocks : {'embeddings': f16 (Nb,512), 'bboxes': f16 (Nb,4)}              blocks.py:125-134
  objects: {'embeddings': f16 (No,512), 'bboxes': f16 (No,4), 'objectness': f16 (No,1)}  objects.py:316-338
"""
from __future__ import annotations

import os

import ctypes as C
import functools
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import binding, frontend
from . import jpeg as oake_jpeg
from .model import OUT_DIM, OakeEngine

CROP_BYTES = frontend.SIZE * frontend.SIZE * 3


def _align(v: int, a: int = 256) -> int:
    return (v + a - 1) // a * a


_STAGE_POOL = None
_STAGE_WORKERS = 8
# OAKE_FUSED_FRONTEND=0: resize into uint8 crops + a separate matrix kernel, as before (A/B measurements)
_FUSED_FRONTEND = os.environ.get('OAKE_FUSED_FRONTEND', '1') != '0'


def _stage_pool():
    global _STAGE_POOL
    if _STAGE_POOL is None:
        import concurrent.futures
        import os
        _STAGE_POOL = concurrent.futures.ThreadPoolExecutor(max_workers=min(_STAGE_WORKERS, os.cpu_count() or 1),
                                                            thread_name_prefix='oake-jpeg-stage')
    return _STAGE_POOL


class _BlocksTemplate:
    """Everything `plan_blocks` needs for one image size, with offsets relative to the image's slots."""
    __slots__ = ('plan', 'level_rel', 'level_jobs', 'pyramid_bytes', 'glob_job', 'crops', 'cell_level', 'bboxes_half')


@functools.lru_cache(maxsize=4096)
def _blocks_template(w: int, h: int) -> _BlocksTemplate:
    t = _BlocksTemplate()
    t.plan = plan = frontend.blocks_plan(w, h)
    t.bboxes_half = None
    rel, off, jobs = [0], 0, []
    for lv in range(1, len(plan.levels)):
        (lw, lh), (pw, ph) = plan.levels[lv], plan.levels[lv - 1]
        rel.append(off)
        jobs.append(frontend.level_job(0, pw, ph, 0, lw, lh))
        off += _align(lw * lh * 3)
    t.level_rel = np.asarray(rel, dtype=np.int64)
    t.level_jobs = jobs
    t.pyramid_bytes = off
    t.glob_job = frontend.crop_jobs(0, w, h, np.array([[0, 0, w, h]], dtype=np.int64), 0)
    c = np.zeros(1 + len(plan.cells), dtype=frontend.CROP_SRC)
    c['pitch_px'][0] = frontend.SIZE
    for i, (lv, x, y) in enumerate(plan.cells):
        lw = plan.levels[lv][0]
        c['off'][1 + i] = (y * lw + x) * 3  # inside its level
        c['pitch_px'][1 + i] = lw
    t.crops = c
    t.cell_level = np.asarray([lv for lv, _, _ in plan.cells], dtype=np.int64)
    return t


def _bboxes_half(plan) -> torch.Tensor:
    t = _blocks_template(*plan.levels[0]) if plan.levels else None
    if t is None or t.plan is not plan:
        return torch.tensor(plan.bboxes, dtype=torch.float32).half()
    if t.bboxes_half is None:
        t.bboxes_half = torch.tensor(plan.bboxes, dtype=torch.float32).half()
    return t.bboxes_half


class _Staging:
    """Growable pinned host buffer + matching device buffer."""

    def __init__(self, device: torch.device) -> None:
        self.device = device
        self.host: Optional[torch.Tensor] = None
        self.dev: Optional[torch.Tensor] = None

    def reserve(self, host_bytes: int, dev_bytes: Optional[int] = None) -> bool:
        """-> True when the device buffer was (re)allocated by this call."""
        dev_bytes = host_bytes if dev_bytes is None else dev_bytes
        if self.host is None or self.host.numel() < host_bytes:
            self.host = torch.empty(max(host_bytes, 1 << 20), dtype=torch.uint8, pin_memory=True)
        if self.dev is None or self.dev.numel() < dev_bytes:
            self.dev = None
            self.dev = torch.empty(max(dev_bytes, 1 << 20), dtype=torch.uint8, device=self.device)
            return True
        return False

    def upload(self, nbytes: int) -> None:
        self.dev[:nbytes].copy_(self.host[:nbytes], non_blocking=True)

    def upload_ranges(self, ranges: Sequence[Tuple[int, int]]) -> int:
        """Copies only [lo, hi) byte ranges (adjacent ones merged); -> bytes sent."""
        merged: List[List[int]] = []
        for lo, hi in sorted(ranges):
            if merged and lo <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], hi)
            else:
                merged.append([lo, hi])
        for lo, hi in merged:
            self.dev[lo:hi].copy_(self.host[lo:hi], non_blocking=True)
        return sum(hi - lo for lo, hi in merged)


class _Slot:
    """One in-flight batch: its staging buffers, its result buffer and the event that ends it."""

    def __init__(self, device: torch.device) -> None:
        self.arena = _Staging(device)  # images | pyramid levels | resized crops
        self.meta = _Staging(device)  # resize jobs | crop descriptors | fg | box
        self.jpeg = _Staging(device)  # descriptors | entropy-coded streams (host + device)
        self.jpeg_scratch = _Staging(device)  # coefficients | component planes (device only)
        self.jpeg_status: Optional[torch.Tensor] = None  # int32 per compressed image
        self.jpeg_status_host: Optional[torch.Tensor] = None
        self.jpeg_count = 0
        self.jpeg_done = torch.cuda.Event()
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.err_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.out_host: Optional[torch.Tensor] = None
        self.event = torch.cuda.Event()
        self.busy = False
        self.ticket: Optional['Pending'] = None  # the uncollected batch that lives in this slot


class Pending:
    """Handle of a submitted batch; `result()` waits for the GPU and returns what `encode_*` returns.

    The batch lives in one of the pipeline's two slots until it is collected.  Submitting a third batch
    while two are outstanding collects the oldest one first (its result -- or its error -- is then kept
    in the handle), so a handle stays valid however late `result()` is called."""

    def __init__(self, pipe: 'OakePipeline', slot: _Slot, host: torch.Tensor, finish) -> None:
        self._pipe, self._slot, self._host, self._finish = pipe, slot, host, finish
        self._jpeg_count = slot.jpeg_count
        self._done = None
        self._error: Optional[Exception] = None
        self._collected = False

    def collect(self) -> None:
        """Waits for the GPU side and moves the result out of the slot's buffers (idempotent)."""
        if self._collected:
            return
        self._collected = True
        slot = self._slot
        slot.event.synchronize()
        slot.busy = False
        slot.ticket = None
        if int(slot.err_host.item()) != 0:
            slot.err.zero_()
            self._error = binding.OakeError('oake_resize_u8: a crop exceeded the resize kernel limits')
            return
        if self._jpeg_count:
            bad = slot.jpeg_status_host[:self._jpeg_count].nonzero().flatten().tolist()
            if bad:
                self._error = binding.OakeError(f'oake_jpeg_decode: damaged or truncated entropy-coded data in '
                                                f'compressed image(s) {bad} of the batch')
                return
        # `finish` copies what it keeps out of the slot's pinned buffer: one tensor with its OWN storage per
        # stored item (torch.save writes the whole storage behind a view)
        self._done = self._finish(self._host)

    def result(self):
        self.collect()
        if self._error is not None:
            raise self._error
        return self._done


class OakePipeline:
    """Two slots: while the GPU works on one batch the host stages (and uploads) the next one."""

    def __init__(self, engine: OakeEngine) -> None:
        self.engine = engine
        self.lib = engine.lib
        self.device = engine.device
        self._slots = [_Slot(self.device), _Slot(self.device)]
        self._cur = 0
        self.h2d_bytes = 0  # of the last call
        self.d2h_bytes = 0
        self.frontend_launches = 0  # resize / mask kernels launched so far
        self.profile_frontend = False  # bench roofline pass: CUDA events around the resize / mask kernels
        self.frontend_events = []  # (name, start, stop)
        self._decode_stream = torch.cuda.Stream(self.device)  # JPEG decode of batch k+1 beside the tower of batch k

    # the slot being filled
    _arena = property(lambda self: self._slots[self._cur].arena)
    _meta = property(lambda self: self._slots[self._cur].meta)
    _err = property(lambda self: self._slots[self._cur].err)
    _slot = property(lambda self: self._slots[self._cur])

    # ------------------------------------------------------------------------------ internals
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _place_images(self, images: Sequence[np.ndarray]) -> Tuple[List[int], int]:
        offs, off = [], 0
        for im in images:
            if im.dtype != np.uint8 or im.ndim != 3 or im.shape[2] != 3:
                raise ValueError('images must be uint8 HWC RGB arrays (or jpeg.JpegSource)')
            offs.append(off)
            off += _align(im.shape[0] * im.shape[1] * 3)
        return offs, off

    def _submit(self, plan_args: tuple, finish) -> Pending:
        """stage -> H2D -> kernels -> D2H, all asynchronous on the current stream."""
        nxt = self._cur ^ 1
        slot = self._slots[nxt]
        if slot.ticket is not None:  # its previous batch has not been collected yet: do that now
            slot.ticket.collect()
        self._cur = nxt
        job = self.stage(*plan_args)
        self.upload(job)
        out = self.launch(job)
        n = out.shape[0]
        if slot.out_host is None or slot.out_host.shape[0] < n:
            slot.out_host = torch.empty(max(n, 4096), OUT_DIM, dtype=torch.float16, pin_memory=True)
        host = slot.out_host[:n]
        host.copy_(out, non_blocking=True)
        slot.err_host.copy_(slot.err, non_blocking=True)
        slot.event.record(torch.cuda.current_stream(self.device))
        slot.busy = True
        self.d2h_bytes = n * OUT_DIM * 2 + 4
        if slot.jpeg_count:
            slot.jpeg_status_host[:slot.jpeg_count].copy_(slot.jpeg_status[:slot.jpeg_count], non_blocking=True)
            slot.event.record(torch.cuda.current_stream(self.device))
            self.d2h_bytes += 4 * slot.jpeg_count
        slot.ticket = Pending(self, slot, host, finish)
        return slot.ticket

    def _run(self, *plan_args) -> torch.Tensor:
        """Synchronous form: returns the HOST fp16 (n_crops, 512) tensor."""
        return self._submit(plan_args, lambda emb: emb.clone()).result()

    # The three phases are public so that bench.py can time the device-resident part alone.
    def stage(self, images, img_offs, img_bytes, arena_bytes, stages, crops, variant, fg=None, box=None) -> dict:
        """Host only: lays the images and descriptors out in the pinned staging buffers."""
        n = crops.shape[0]
        parts, off, stage_offs = [], 0, []
        for jobs in stages:
            stage_offs.append(off)
            parts.append((off, jobs.view(np.uint8).reshape(-1)))
            off = _align(off + jobs.nbytes)
        crops_off = off
        parts.append((off, crops.view(np.uint8).reshape(-1)))
        off = _align(off + crops.nbytes)
        fg_off = box_off = masks_off = 0
        if variant == binding.VARIANT_T197:
            fg_off = off
            parts.append((off, np.ascontiguousarray(fg, dtype=np.float32).view(np.uint8).reshape(-1)))
            off = _align(off + n * 16)
            box_off = off
            parts.append((off, np.ascontiguousarray(box, dtype=np.float32).view(np.uint8).reshape(-1)))
            off = _align(off + n * 16)
            masks_off = off  # device only
        meta_host_bytes = off
        meta_dev_bytes = off + (n * 196 * 4 if variant == binding.VARIANT_T197 else 0)
        self._meta.reserve(meta_host_bytes, meta_dev_bytes)
        fresh = self._arena.reserve(img_bytes, arena_bytes)
        mh = self._meta.host.numpy()
        for o, b in parts:
            mh[o:o + b.size] = b
        ah = self._arena.host.numpy()
        compressed, raw_ranges, raw = [], [], []
        for im, o in zip(images, img_offs):
            if isinstance(im, oake_jpeg.JpegSource):
                compressed.append((im, o))
            else:
                raw.append((im, o))
                raw_ranges.append((o, _align(o + im.size)))

        def copy_range(lo: int) -> None:  # numpy releases the GIL inside the copy loop
            for im, o in raw[lo:lo + per]:
                ah[o:o + im.size] = im.reshape(-1)

        # ~0.9 MB per COCO-sized image: one thread moves ~10 GB/s, so a large batch (globals: hundreds of images
        # per call) is copied into the pinned arena by a few threads, in contiguous runs
        workers = _STAGE_WORKERS if len(raw) >= 4 * _STAGE_WORKERS else 1
        per = (len(raw) + workers - 1) // workers if raw else 1
        if workers > 1:
            list(_stage_pool().map(copy_range, range(0, len(raw), per)))
        else:
            copy_range(0)
        jpeg_job = self._stage_jpeg(compressed, fresh) if compressed else None
        self._slot.jpeg_count = len(compressed)
        # crops that are exactly the outputs of one resize stage (globals, objects) never exist as uint8: the resize
        # kernel writes the tower's front-end matrix itself, chunk by chunk (oake_resize_to_patches)
        fused = _FUSED_FRONTEND and frontend.fused_stage(stages, crops)
        return dict(n=n, variant=variant, img_bytes=img_bytes, meta_host_bytes=meta_host_bytes, jpeg=jpeg_job, fused=fused,
                    raw_images=len(images) - len(compressed), raw_ranges=raw_ranges if compressed else None,
                    stages=[(j.size, frontend.max_tiles(j), o) for j, o in zip(stages, stage_offs)],
                    crops_off=crops_off, fg_off=fg_off, box_off=box_off, masks_off=masks_off)

    def _stage_jpeg(self, compressed, fresh: bool) -> dict:
        """Host only: descriptors (rebased onto this slot's arenas) and entropy-coded streams into pinned memory."""
        slot, lib = self._slot, self.lib
        n, db = len(compressed), oake_jpeg.desc_bytes()
        # layout from the parsed headers alone, so that the images can be staged independently: stream i at
        # stream_off[i] (its upper bound reserved), scratch of image i at scratch_off[i]
        stream_off, scratch_off = [], []
        off, scr = _align(n * db), 0
        for src, _ in compressed:
            stream_off.append(off)
            scratch_off.append(scr)
            off += _align(src.stream_bound, 16)
            scr += _align(src.scratch_bytes)
        total = off
        fresh |= slot.jpeg.reserve(total)
        base = slot.jpeg.host.data_ptr()

        def stage_one(i: int) -> None:
            # entropy-coded segment without its stuffing bytes -> pinned memory; descriptor i rebased
            src, out_off = compressed[i]
            scratch, written = C.c_uint64(scratch_off[i]), C.c_uint64(0)
            binding.check(lib.oake_jpeg_stage(src.desc, src.data, len(src.data), base + stream_off[i], stream_off[i],
                                              out_off, C.byref(scratch), base + i * db, C.byref(written)))

        def stage_range(lo: int) -> None:
            for i in range(lo, min(lo + per, n)):
                stage_one(i)

        # a memcpy-like pass per file; ctypes drops the GIL, so a few threads share a large batch -- in
        # a handful of contiguous chunks, not one task per file (every task costs GIL hand-offs)
        workers = _STAGE_WORKERS if n >= 4 * _STAGE_WORKERS else 1
        per = (n + workers - 1) // workers
        if workers > 1:
            list(_stage_pool().map(stage_range, range(0, n, per)))
        else:
            stage_range(0)
        scratch = C.c_uint64(scr)
        fresh |= slot.jpeg_scratch.reserve(0, int(scratch.value))
        if slot.jpeg_status is None or slot.jpeg_status.numel() < n:
            slot.jpeg_status = torch.zeros(max(n, 256), dtype=torch.int32, device=self.device)
            slot.jpeg_status_host = torch.zeros(max(n, 256), dtype=torch.int32).pin_memory()
            fresh = True
        return dict(n=n, bytes=total, fresh=fresh)

    def upload(self, job: dict) -> None:
        self.h2d_bytes = job['meta_host_bytes']
        if job['raw_images']:  # (a batch of compressed files only has nothing to send here)
            if job.get('raw_ranges'):
                # Mixed batch (`jpeg.load` hands progressive / CMYK / non-JPEG files over as Pillow-decoded
                # arrays): the decode stream writes the other images' pixels into this same arena, unordered
                # against the main stream, so only the byte ranges of the raw images may be copied here --
                # a whole-arena copy could land after the decode and bury it under stale staging bytes.
                self.h2d_bytes += self._arena.upload_ranges(job['raw_ranges'])
            else:
                self._arena.upload(job['img_bytes'])
                self.h2d_bytes += job['img_bytes']
        if job['meta_host_bytes']:
            self._meta.upload(job['meta_host_bytes'])
        jj = job['jpeg']
        if jj is not None:
            # Side stream: the files go up and are decoded while the main stream is still busy with the
            # previous batch.  The slot's buffers are idle by now (`_submit` waited for its last batch);
            # only memory the caching allocator handed out just now may still be in use by work queued
            # on the main stream, so a fresh allocation makes the side stream wait for that work once.
            slot, main = self._slot, torch.cuda.current_stream(self.device)
            if jj['fresh']:
                self._decode_stream.wait_stream(main)
            with torch.cuda.stream(self._decode_stream):
                slot.jpeg.upload(jj['bytes'])
                binding.check(self.lib.oake_jpeg_decode(
                    slot.jpeg.dev.data_ptr(), slot.jpeg.host.data_ptr(), slot.jpeg.dev.data_ptr(), jj['n'],
                    slot.jpeg_scratch.dev.data_ptr(), slot.arena.dev.data_ptr(), slot.jpeg_status.data_ptr(),
                    self._decode_stream.cuda_stream))
                slot.jpeg_done.record(self._decode_stream)
            main.wait_event(slot.jpeg_done)
            self.h2d_bytes += jj['bytes']
            self.frontend_launches += 3

    def launch(self, job: dict) -> torch.Tensor:
        """Device only: resize stages, masks, tower.  Returns the DEVICE fp16 (n, 512) tensor."""
        n, variant = job['n'], job['variant']
        st = self._stream()
        arena_ptr = self._arena.dev.data_ptr()
        meta_ptr = self._meta.dev.data_ptr()
        def timed(name, fn):
            if not self.profile_frontend:
                fn()
                return
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            self.frontend_events.append((name, a, b))

        fused = job.get('fused', False)
        for count, tiles, o in job['stages']:
            if count and not fused:
                timed('resize_u8', lambda: binding.check(self.lib.oake_resize_u8(
                    arena_ptr, arena_ptr, meta_ptr + o, count, tiles, self._err.data_ptr(), st)))
                self.frontend_launches += 3  # prepare, FAST tiles, BIG tiles
        masks_ptr = None
        if variant == binding.VARIANT_T197 and n:
            masks_ptr = meta_ptr + job['masks_off']
            timed('object_masks', lambda: binding.check(self.lib.oake_object_masks(
                meta_ptr + job['fg_off'], meta_ptr + job['box_off'], n, masks_ptr, st)))
            self.frontend_launches += 1
        out = torch.empty(n, OUT_DIM, dtype=torch.float16, device=self.device)
        step = self.engine.MAX_CROPS[variant]
        if n:
            ws = self.engine._workspace(min(n, step), variant)
            for s in range(0, n, step):
                b = min(step, n - s)
                if fused:
                    jobs_ptr = meta_ptr + job['stages'][0][2] + s * frontend.RESIZE_JOB.itemsize
                    timed('resize_u8', lambda: binding.check(self.lib.oake_resize_to_patches(
                        self.engine._handle, arena_ptr, jobs_ptr, b, variant, ws.data_ptr(), ws.numel(),
                        self._err.data_ptr(), st)))  # (its three kernels are counted by the handle)
                    binding.check(self.lib.oake_encode_patches(
                        self.engine._handle, b, variant, (masks_ptr + s * 196 * 4) if masks_ptr else None,
                        out[s:].data_ptr(), None, ws.data_ptr(), ws.numel(), st))
                    continue
                binding.check(self.lib.oake_encode_crops_u8(
                    self.engine._handle, arena_ptr, meta_ptr + job['crops_off'] + s * frontend.CROP_SRC.itemsize, b,
                    variant, (masks_ptr + s * 196 * 4) if masks_ptr else None, out[s:].data_ptr(), None,
                    ws.data_ptr(), ws.numel(), st))
        return out

    def collect_frontend_profile(self) -> Dict[str, Dict[str, float]]:
        torch.cuda.synchronize(self.device)
        out: Dict[str, Dict[str, float]] = {}
        for name, a, b in self.frontend_events:
            d = out.setdefault(name, dict(ms=0.0, flops=0.0, launches=0))
            d['ms'] += a.elapsed_time(b)
            d['launches'] += 1
        self.frontend_events = []
        return out

    # ------------------------------------------------------------------------------ public API
    # `encode_*` block until the result is on the host; `submit_*` return a `Pending` at once so that
    # the caller can prepare / submit the next batch while the GPU works (validators, bench e2e).
    def encode_globals(self, images: Sequence[np.ndarray]) -> List[torch.Tensor]:
        return self.submit_globals(images).result()

    def submit_globals(self, images: Sequence[np.ndarray]) -> Pending:
        n = len(images)
        return self._submit(self.plan_globals(images), lambda emb: [emb[i].clone() for i in range(n)])

    def plan_globals(self, images: Sequence[np.ndarray]) -> tuple:
        offs, img_bytes = self._place_images(images)
        n = len(images)
        if n:
            # one CLIP-transform job per image, built for the whole batch at once
            wh = np.array([(im.shape[1], im.shape[0]) for im in images], dtype=np.int64)
            boxes = np.concatenate([np.zeros((n, 2), np.int64), wh], axis=1)
            jobs = frontend.crop_jobs(np.asarray(offs, dtype=np.int64), wh[:, 0], wh[:, 1], boxes, img_bytes)
        else:
            jobs = np.zeros(0, frontend.RESIZE_JOB)
        crops = np.zeros(n, dtype=frontend.CROP_SRC)
        crops['off'] = jobs['dst_off']
        crops['pitch_px'] = frontend.SIZE
        return (images, offs, img_bytes, img_bytes + n * CROP_BYTES, [jobs], crops, binding.VARIANT_T50)

    def encode_blocks(self, images: Sequence[np.ndarray]) -> List[Dict[str, torch.Tensor]]:
        return self.submit_blocks(images).result()

    def submit_blocks(self, images: Sequence[np.ndarray]) -> Pending:
        args, plans, counts = self.plan_blocks(images)

        def finish(emb: torch.Tensor):
            out, s = [], 0
            for plan, n in zip(plans, counts):
                out.append(dict(embeddings=emb[s:s + n].clone(), bboxes=_bboxes_half(plan).clone()))
                s += n
            return out

        return self._submit(args, finish)

    def plan_blocks(self, images: Sequence[np.ndarray]):
        offs, img_bytes = self._place_images(images)
        tpls = [_blocks_template(im.shape[1], im.shape[0]) for im in images]
        plans = [t.plan for t in tpls]
        counts = [t.crops.shape[0] for t in tpls]
        # arena: images | global crops (packed) | per image: its pyramid levels 1..
        glob_off = img_bytes
        off = img_bytes + len(images) * CROP_BYTES
        n_stages = max([len(t.level_jobs) for t in tpls] + [0])
        stage_jobs: List[List[np.ndarray]] = [[] for _ in range(max(n_stages, 1))]
        crops_all = []
        for k, (o, t) in enumerate(zip(offs, tpls)):
            level_off = t.level_rel + off  # absolute arena offset of every level ...
            level_off[0] = o  # ... level 0 being the image itself
            g = t.glob_job.copy()
            g['src_off'] = o
            g['dst_off'] = glob_off + k * CROP_BYTES
            stage_jobs[0].append(g)
            for lv, job in enumerate(t.level_jobs, start=1):  # level lv is resized from level lv - 1
                j = job.copy()
                j['src_off'] = level_off[lv - 1]
                j['dst_off'] = level_off[lv]
                stage_jobs[lv - 1].append(j)
            c = t.crops.copy()
            c['off'][0] = glob_off + k * CROP_BYTES
            c['off'][1:] += level_off[t.cell_level]
            crops_all.append(c)
            off += t.pyramid_bytes
        stages = [np.concatenate(s) if s else np.zeros(0, frontend.RESIZE_JOB) for s in stage_jobs]
        crops = np.concatenate(crops_all) if crops_all else np.zeros(0, frontend.CROP_SRC)
        return (images, offs, img_bytes, off, stages, crops, binding.VARIANT_T50), plans, counts

    def encode_objects(self, images: Sequence[np.ndarray], proposals: Sequence[np.ndarray], dry_run: bool = False,
                       expand_mode: str = 'ADAPTIVE') -> List[Dict[str, torch.Tensor]]:
        return self.submit_objects(images, proposals, dry_run, expand_mode).result()

    def submit_objects(self, images: Sequence[np.ndarray], proposals: Sequence[np.ndarray], dry_run: bool = False,
                       expand_mode: str = 'ADAPTIVE') -> Pending:
        args, plans = self.plan_objects(images, proposals, dry_run, expand_mode)

        def finish(emb: torch.Tensor):
            out, s = [], 0
            for plan in plans:
                n = plan.bboxes.shape[0]
                out.append(dict(embeddings=emb[s:s + n].clone(), bboxes=torch.from_numpy(plan.bboxes).half(),
                                objectness=torch.from_numpy(plan.objectness).half()))
                s += n
            return out

        return self._submit(args, finish)

    def plan_objects(self, images: Sequence[np.ndarray], proposals: Sequence[np.ndarray], dry_run: bool = False,
                     expand_mode: str = 'ADAPTIVE'):
        offs, img_bytes = self._place_images(images)
        plans = [frontend.objects_plan(p, (im.shape[1], im.shape[0]), dry_run, expand_mode)
                 for im, p in zip(images, proposals)]
        total = sum(p.bboxes.shape[0] for p in plans)
        jobs, s = [], 0
        for im, o, plan in zip(images, offs, plans):
            h, w = im.shape[:2]
            jobs.append(frontend.crop_jobs(o, w, h, plan.boxes_int, img_bytes + s * CROP_BYTES))
            s += plan.bboxes.shape[0]
        jobs = np.concatenate(jobs) if jobs else np.zeros(0, frontend.RESIZE_JOB)
        crops = np.zeros(total, dtype=frontend.CROP_SRC)
        crops['off'] = jobs['dst_off']
        crops['pitch_px'] = frontend.SIZE
        fg = np.concatenate([p.foregrounds for p in plans]) if plans else np.zeros((0, 4), np.float32)
        box = np.concatenate([p.expanded for p in plans]) if plans else np.zeros((0, 4), np.float32)
        return (images, offs, img_bytes, img_bytes + total * CROP_BYTES, [jobs], crops, binding.VARIANT_T197, fg,
                box), plans

    def decode_jpegs(self, sources: Sequence['oake_jpeg.JpegSource']) -> List[np.ndarray]:
        """Runs only the JPEG decode; returns the uint8 HWC pixels of each file (tests, tools)."""
        offs, img_bytes = self._place_images(sources)
        self._cur ^= 1
        slot = self._slot
        if slot.ticket is not None:
            slot.ticket.collect()
        fresh = slot.arena.reserve(1, img_bytes)
        jj = self._stage_jpeg(list(zip(sources, offs)), fresh)
        slot.jpeg_count = 0
        self.upload(dict(meta_host_bytes=0, raw_images=0, img_bytes=img_bytes, jpeg=jj))
        torch.cuda.synchronize(self.device)
        status = slot.jpeg_status[:len(sources)].cpu()
        if bool(status.any()):
            raise binding.OakeError(f'oake_jpeg_decode: damaged data in {status.nonzero().flatten().tolist()}')
        arena = slot.arena.dev
        return [arena[o:o + s.size].cpu().numpy().reshape(s.shape) for s, o in zip(sources, offs)]

    # --------------------------------------------------------------- test hooks (uint8 crops)
    def debug_crops_u8(self, images: Sequence[np.ndarray], jobs: np.ndarray) -> np.ndarray:
        """Runs only the resize stage; returns the packed (n,224,224,3) uint8 crops (tests)."""
        offs, img_bytes = self._place_images(images)
        n = jobs.shape[0]
        hi = int(jobs['dst_off'].max()) + CROP_BYTES if n else img_bytes
        self._arena.reserve(img_bytes, max(hi, img_bytes))
        self._meta.reserve(_align(jobs.nbytes))
        ah = self._arena.host.numpy()
        for im, o in zip(images, offs):
            ah[o:o + im.size] = im.reshape(-1)
        self._meta.host.numpy()[:jobs.nbytes] = jobs.view(np.uint8).reshape(-1)
        self._arena.upload(img_bytes)
        self._meta.upload(jobs.nbytes)
        p = self._arena.dev.data_ptr()
        binding.check(self.lib.oake_resize_u8(p, p, self._meta.dev.data_ptr(), n, frontend.max_tiles(jobs),
                                              self._err.data_ptr(), self._stream()))
        torch.cuda.synchronize(self.device)
        if int(self._err.item()) != 0:
            self._err.zero_()
            raise binding.OakeError('oake_resize_u8: a crop exceeded the resize kernel limits')
        lo = int(jobs['dst_off'].min()) if n else 0
        return self._arena.dev[lo:lo + n * CROP_BYTES].cpu().numpy().reshape(n, 224, 224, 3)
