"""Developer probe: what the objects side row costs the 197-token attention kernel (same B, with and without it)."""
import sys
import pathlib

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from oadp_b200 import binding  # noqa: E402


def main():
    lib = binding.load()
    B, P = 478, 196
    dt = torch.float16 if lib.oake_act_dtype().decode() == 'f16' else torch.bfloat16
    g = torch.Generator(device='cuda').manual_seed(0)
    st = torch.cuda.current_stream().cuda_stream

    def run(side, mask_kind):
        R = B * (P + 1) + (B if side else 0)
        qkv = (torch.randn(R, 2304, device='cuda', generator=g) * 1.0).to(dt)
        out = torch.zeros(R, 768, device='cuda', dtype=dt)
        mask = None
        if side:
            mask = (torch.rand(B, P, device='cuda', generator=g) > 0.5).float()
            if mask_kind == 'soft':
                mask = mask * 0.7

        def call():
            if side:
                binding.check(lib.oake_test_attention_side(qkv.data_ptr(), mask.data_ptr(), out.data_ptr(), B, P, 0, st))
            else:
                binding.check(lib.oake_test_attention_main(qkv.data_ptr(), out.data_ptr(), B, P, st))

        for _ in range(3):
            call()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            call()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / 20 * 1e3

    for _ in range(2):
        print(f'no side row {run(False, None):7.1f} us | side row, 0/1 mask {run(True, "bits"):7.1f} us | '
              f'side row, soft mask {run(True, "soft"):7.1f} us   (B = {B})')


if __name__ == '__main__':
    main()
