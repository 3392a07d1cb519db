"""Developer probe: the objects workload's resize stage alone (8 images x 300 proposals), fused into the front-end
matrix (the product path) and as uint8 crops, CUDA events over repeated launches on an otherwise idle GPU."""
import sys
import pathlib

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from oadp_b200 import binding, frontend, synth  # noqa: E402
from oadp_b200.model import OakeEngine  # noqa: E402
from oadp_b200.pipeline import OakePipeline  # noqa: E402


def main():
    eng = OakeEngine(synth.visual_params(0, layers=1), 'cuda')
    pipe = OakePipeline(eng)
    imgs = synth.images(8, seed=0)
    props = [synth.proposals(im.shape[1], im.shape[0], 300, seed=i) for i, im in enumerate(imgs)]
    args, _ = pipe.plan_objects(imgs, props)
    job = pipe.stage(*args)
    pipe.upload(job)
    torch.cuda.synchronize()
    lib, st = pipe.lib, pipe._stream()
    n, (count, tiles, off) = job['n'], job['stages'][0]
    arena, meta = pipe._arena.dev.data_ptr(), pipe._meta.dev.data_ptr()
    step = eng.MAX_CROPS[binding.VARIANT_T197]
    ws = eng._workspace(min(n, step), binding.VARIANT_T197)

    def fused():
        for s in range(0, n, step):
            b = min(step, n - s)
            binding.check(lib.oake_resize_to_patches(eng._handle, arena, meta + off + s * frontend.RESIZE_JOB.itemsize, b,
                                                     binding.VARIANT_T197, ws.data_ptr(), ws.numel(), pipe._err.data_ptr(), st))

    def u8():
        binding.check(lib.oake_resize_u8(arena, arena, meta + off, count, tiles, pipe._err.data_ptr(), st))

    for name, fn in (('fused (matrix)', fused), ('uint8 crops', u8)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            fn()
        b.record()
        torch.cuda.synchronize()
        print(f'{name:16s} {a.elapsed_time(b) / 20:.3f} ms per {n} crops')


if __name__ == '__main__':
    main()
