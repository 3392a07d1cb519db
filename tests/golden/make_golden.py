"""Regenerates tests/golden/vit_golden.pt from the CPU oracle (and HF CLIP for the T=50 tower).

    python tests/golden/make_golden.py

The fixture pins: seeded weights (oracle.vit.init_visual_params(seed)) + seeded inputs -> fp32
outputs of (a) oracle T=50, (b) HuggingFace CLIPVisionModelWithProjection on the same weights,
(c) oracle T=197 + side stream, (d) the hook-driven second restatement of (c).
Inputs are regenerated from the seeds at test time, only the outputs are stored.
"""
import pathlib
import sys

import torch

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import hooks_ref, vit  # noqa: E402

WEIGHT_SEED = 1234
INPUT_SEED = 4321
B = 4


def golden_inputs(b: int = B, seed: int = INPUT_SEED):
    g = torch.Generator().manual_seed(seed)
    pixels = torch.randn(b, 3, 224, 224, generator=g) * 1.2
    masks = (torch.rand(b, 1, 14, 14, generator=g) > 0.6).float()
    masks[0] = 0.0  # nothing masked
    masks[1] = 1.0  # everything masked: -100 on every patch, y only sees itself
    return pixels, masks


def main() -> None:
    torch.set_num_threads(8)
    p = vit.init_visual_params(WEIGHT_SEED)
    pixels, masks = golden_inputs()
    t50 = vit.encode_image(p, pixels)
    with torch.no_grad():
        hf = vit.build_hf_model(p)(pixel_values=pixels).image_embeds
    p197 = vit.objects_surgery(p)
    t197 = vit.encode_objects(p197, pixels, masks)
    t197_hooks = hooks_ref.HookedVisual(p197)(pixels, masks)
    print('t50 vs hf  max-abs', (t50 - hf).abs().max().item())
    print('t197 vs hooks max-abs', (t197 - t197_hooks).abs().max().item())
    out = dict(weight_seed=WEIGHT_SEED, input_seed=INPUT_SEED, batch=B, t50=t50, t50_hf=hf, t197=t197,
               t197_hooks=t197_hooks, pixels_checksum=pixels.double().sum().item(),
               masks_checksum=masks.double().sum().item(), torch_version=str(torch.__version__))
    torch.save(out, pathlib.Path(__file__).with_name('vit_golden.pt'))


if __name__ == '__main__':
    main()
