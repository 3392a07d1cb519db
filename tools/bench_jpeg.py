"""Micro-benchmark of the JPEG row: GPU decode of COCO-sized files against Pillow on
the host cores, and the blocks task end to end from compressed files with either decoder.
Prints one JSON line per measurement.  Synthetic files (PIL-encoded low-pass noise, q90 4:2:0)."""
import concurrent.futures
import io
import json
import os
import pathlib
import sys
import time

import numpy as np
import PIL.Image
import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from oadp_b200 import jpeg as oake_jpeg
from oadp_b200 import synth
from oadp_b200.model import OakeModel
from oadp_b200.pipeline import OakePipeline


def make_files(n):
    files = []
    for i in range(n):
        w, h = synth.COCO_SIZES[i % len(synth.COCO_SIZES)]
        buf = io.BytesIO()
        PIL.Image.fromarray(synth.image(w, h, 900 + i % 32)).save(buf, 'JPEG', quality=90)
        files.append(buf.getvalue())
    return files


def pil_decode(data):
    return np.asarray(PIL.Image.open(io.BytesIO(data)).convert('RGB'), dtype=np.uint8)


def main():
    dev = torch.device('cuda', 0)
    pipe = OakePipeline(OakeModel(synth.visual_params(0), dev).engine)
    files = make_files(1024)
    mb = sum(map(len, files)) / 1e6
    print(json.dumps(dict(what='corpus', files=len(files), compressed_mb=round(mb, 1),
                          pixels_mb=round(sum(oake_jpeg.parse(f).size for f in files) / 1e6, 1))), flush=True)

    # -- decode only, device time (H2D of the files + the three kernels), by batch size
    for n in (() if '--e2e-only' in sys.argv else (1, 8, 64, 256, 1024)):
        sources = [oake_jpeg.parse(f) for f in files[:n]]
        offs, img_bytes = pipe._place_images(sources)
        slot = pipe._slot
        fresh = slot.arena.reserve(1, img_bytes)
        jj = pipe._stage_jpeg(list(zip(sources, offs)), fresh)
        job = dict(meta_host_bytes=0, raw_images=0, img_bytes=img_bytes, jpeg=jj)
        for _ in range(2):
            pipe.upload(job)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 5
        a.record()
        for _ in range(iters):
            pipe.upload(job)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / iters
        print(json.dumps(dict(what='gpu_decode', batch=n, ms=round(ms, 3), images_per_s=round(n / ms * 1e3, 1),
                              compressed_mb_per_s=round(sum(len(f) for f in files[:n]) / 1e6 / ms * 1e3, 1),
                              pixel_gb_per_s=round(sum(s.size for s in sources) / 1e9 / ms * 1e3, 2))), flush=True)

    if '--decode-only' in sys.argv:
        return

    # -- Pillow on the host: one thread, and every core
    cores = os.cpu_count() or 1
    t = time.perf_counter()
    for f in files[:128]:
        pil_decode(f)
    one = 128 / (time.perf_counter() - t)
    cores = os.cpu_count() or 1
    with concurrent.futures.ThreadPoolExecutor(cores) as pool:
        list(pool.map(pil_decode, files[:64]))
        t = time.perf_counter()
        list(pool.map(pil_decode, files))
        many = len(files) / (time.perf_counter() - t)
    print(json.dumps(dict(what='pillow_decode', images_per_s_1_thread=round(one, 1), threads=cores,
                          images_per_s_all_threads=round(many, 1))), flush=True)
    t = time.perf_counter()
    for f in files[:256]:
        oake_jpeg.parse(f)
    print(json.dumps(dict(what='host_parse', images_per_s_1_thread=round(256 / (time.perf_counter() - t), 1))), flush=True)

    # -- blocks and globals tasks end to end from compressed files (file bytes in host memory -> fp16 rows on
    #    the host); the decoder is either the GPU one (host: header parse only) or Pillow on every host thread
    def run(task, decoder, batch_images, workers=cores):
        submit = pipe.submit_blocks if task == 'blocks' else pipe.submit_globals
        count = (lambda res: sum(r['embeddings'].shape[0] for r in res)) if task == 'blocks' else len
        with concurrent.futures.ThreadPoolExecutor(workers) as pool:
            crops, in_flight = 0, None
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for s in range(0, len(files), batch_images):
                batch = list(pool.map(decoder, files[s:s + batch_images]))
                ticket = submit(batch)
                if in_flight is not None:
                    crops += count(in_flight.result())
                in_flight = ticket
            crops += count(in_flight.result())
            dt = time.perf_counter() - t0
        return len(files) / dt, crops / dt

    for task, batch_images in (('blocks', 64), ('globals', 256)):
        for name, decoder in (('gpu', oake_jpeg.parse), ('pillow', pil_decode)):
            run(task, decoder, batch_images)
            ips, cps = run(task, decoder, batch_images)
            print(json.dumps(dict(what=f'{task}_e2e_from_files', decoder=name, host_threads=cores,
                                  batch_images=batch_images, images_per_s=round(ips, 1), crops_per_s=round(cps, 1))),
                  flush=True)


if __name__ == '__main__':
    main()
