"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- the checker for the GPU JPEG decode.

For this row the oracle is not a restatement but the reference's own decoder: oadp/oake/base.py:53
loads every image with torchvision's `CocoDetection._load_image`, i.e.
`PIL.Image.open(path).convert('RGB')`, and Pillow (libjpeg-turbo) is present both in this container
and on the GPU box.  Parity bar: bit-exact uint8 pixels.
"""
import io

import numpy as np
import PIL.Image


def decode(data: bytes) -> np.ndarray:
    """uint8 HWC RGB pixels of one image file, as the reference's loader produces them."""
    return np.asarray(PIL.Image.open(io.BytesIO(data)).convert('RGB'), dtype=np.uint8)


def corpus(seed: int = 0):
    """Seeded JPEG files covering what the decoder has to get right: qualities 5..100, 4:4:4 / 4:2:2 /
    4:2:0 / grayscale, sizes that are and are not multiples of the MCU, restart intervals, optimised
    (image-specific) Huffman tables.  Yields (label, file bytes)."""
    import itertools
    rng = np.random.default_rng(seed)

    def picture(w, h):
        coarse = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2, 3), dtype=np.uint8)
        smooth = np.asarray(PIL.Image.fromarray(coarse).resize((w, h), PIL.Image.BICUBIC)).astype(np.int16)
        return PIL.Image.fromarray(np.clip(smooth + rng.integers(-20, 21, (h, w, 3)), 0, 255).astype(np.uint8))

    sizes = [(64, 48), (67, 45), (17, 9), (33, 130), (8, 8), (5, 3), (320, 213)]
    for (w, h), quality, sub, gray, rst in itertools.product(sizes, (30, 75, 95, 100), (0, 1, 2), (False, True), (0, 3)):
        im = picture(w, h)
        if gray:
            im = im.convert('L')
        kw = dict(quality=quality, subsampling=sub)
        if rst:
            kw['restart_marker_blocks'] = rst
        buf = io.BytesIO()
        im.save(buf, 'JPEG', **kw)
        yield f'{w}x{h} q{quality} sub{sub} gray{int(gray)} rst{rst}', buf.getvalue()
    for kw in (dict(optimize=True, quality=92), dict(quality=5), dict(quality=10, subsampling=1),
               dict(optimize=True, quality=98, subsampling=0), dict(quality=85, restart_marker_rows=1)):
        buf = io.BytesIO()
        picture(333, 217).save(buf, 'JPEG', **kw)
        yield f'333x217 {kw}', buf.getvalue()


def outside_envelope(seed: int = 0):
    """Files the GPU decoder must hand back to Pillow: progressive, CMYK."""
    rng = np.random.default_rng(seed)
    im = PIL.Image.fromarray(rng.integers(0, 256, (40, 56, 3), dtype=np.uint8))
    for label, image, kw in (('progressive', im, dict(progressive=True)), ('cmyk', im.convert('CMYK'), {})):
        buf = io.BytesIO()
        image.save(buf, 'JPEG', **kw)
        yield label, buf.getvalue()
