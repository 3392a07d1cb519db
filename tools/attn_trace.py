"""Developer probe: hand-over timeline of the persistent attention kernel (block 0, first 32 tiles).
Needs a library built with -DOAKE_ATTN_TRACE (tools/gpu_attn_trace.sh does that on the GPU box)."""
import ctypes as C
import pathlib
import sys

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from oadp_b200 import build  # noqa: E402

import os
build.build(force=True, extra_flags=['-DOAKE_ATTN_TRACE'] + ([f"-DOAKE_ATTN_EXP={os.environ['EXP']}"] if os.environ.get('EXP') else []))
from oadp_b200 import binding  # noqa: E402
from oadp_b200.model import OakeEngine  # noqa: E402
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent))
from quick_bench import init_params  # noqa: E402

eng = OakeEngine(init_params(layers=2), 'cuda')
B = 478
eng.MAX_CROPS = {0: B, 1: B}
px = torch.randn(B, 3, 224, 224, device='cuda')
masks = (torch.rand(B, 1, 14, 14, device='cuda') > 0.5).float()
for _ in range(3):
    eng.encode_pixels(px, masks, 1)
torch.cuda.synchronize()
lib = binding.load()
buf = (C.c_longlong * (8 * 32 * 4))()
lib.oake_debug_attn_trace.restype = C.c_int
assert lib.oake_debug_attn_trace(buf) == 0
v = list(buf)
get = lambda role, k, ev: v[(role * 32 + k) * 4 + ev]
t0 = min(x for x in v if x > 0)
print('tile k = 2n + t; cycles relative to the first event.  sm0 = softmax warp 0 (lo, q0); sm6 = warp 6 (hi, q2)')
print(f'{"k":>3} | {"sm0 s_full":>10} {"regs":>7} {"max":>7} {"p_ready":>8} | {"sm6 s_full":>10} {"regs":>7} {"max":>7} {"p_ready":>8} | '
      f'{"mma p_rdy":>9} {"PV iss":>7} {"o_free":>7} {"S iss":>7} | {"drn o_full":>10} {"o_free":>7}')
for k in range(12, 22):
    r = lambda role, ev: get(role, k, ev) - t0 if get(role, k, ev) else -1
    print(f'{k:3d} | {r(0, 0):10d} {r(0, 1):7d} {r(0, 2):7d} {r(0, 3):8d} | {r(4, 0):10d} {r(4, 1):7d} {r(4, 2):7d} {r(4, 3):8d} | '
          f'{r(1, 0):9d} {r(1, 1):7d} {r(1, 2):7d} {r(1, 3):7d} | {r(2, 0):10d} {r(2, 1):7d}')

print('prefetch state at tile start (which of slots 0..3 were written = value of `have` & 3 ... ) sm0 / sm6:')
for k in range(12, 22):
    print(k, [ev for ev in range(4) if get(5, k, ev)], [ev for ev in range(4) if get(6, k, ev)])
