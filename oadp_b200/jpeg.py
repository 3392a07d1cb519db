"""Compressed JPEG files as pipeline inputs: the host only parses the headers; the
entropy decode, IDCT and colour conversion run on the GPU (`oake_jpeg_decode`) and leave the pixels
in the arena the resize kernel reads -- bit-identical to `PIL.Image.open(f).convert('RGB')`, the
reference's loader (oadp/oake/base.py:53, torchvision CocoDetection._load_image).

A `JpegSource` quacks like the uint8 HWC array it will become (`shape`, `dtype`, `ndim`, `size`), so
`OakePipeline.encode_*` / `submit_*` accept it wherever they accept a decoded image.  Files outside
the GPU decoder's envelope (progressive, CMYK, ... see include/oake_b200.h) are decoded with Pillow on
the host exactly as the reference does -- `load()` returns a plain array for those.
"""
from __future__ import annotations

import ctypes as C
import io
import struct
from typing import Optional, Union

import numpy as np

from . import binding

UNSUPPORTED = 2
# Pillow warns above Image.MAX_IMAGE_PIXELS (89 478 485) and raises above twice that; a header claiming more
# than the warning threshold is not worth a GPU arena
MAX_PIXELS = 89_478_485


def desc_bytes() -> int:
    return int(binding.load().oake_jpeg_desc_bytes())


class JpegSource:
    """One parsed, still compressed JPEG file."""
    __slots__ = ('data', 'desc', 'shape', 'scratch_bytes', 'stream_bound')
    dtype = np.dtype(np.uint8)
    ndim = 3

    def __init__(self, data: bytes, desc: bytes) -> None:
        self.data = data
        self.desc = desc  # parsed oake_jpeg_desc with relative offsets
        width, height = struct.unpack_from('<II', desc, 0)
        self.shape = (height, width, 3)
        self.scratch_bytes = _scratch_bytes(desc)
        self.stream_bound = int(binding.load().oake_jpeg_stream_bound(desc))  # what oake_jpeg_stage writes at most

    @property
    def size(self) -> int:
        return self.shape[0] * self.shape[1] * 3


def _scratch_bytes(desc: bytes) -> int:
    # uint32 x 10, then scan_off, scan_len, out_off, scratch_bytes (uint64 each): include/oake_b200.h
    return struct.unpack_from('<Q', desc, 40 + 3 * 8)[0]


def parse(data: bytes) -> Optional[JpegSource]:
    """-> JpegSource, or None when the GPU decoder does not cover this kind of file.  Raises
    `binding.OakeError` for a file that is not a well-formed JPEG."""
    lib = binding.load()
    buf = C.create_string_buffer(desc_bytes())
    rc = lib.oake_jpeg_parse(data, len(data), buf)
    if rc == UNSUPPORTED:
        return None
    binding.check(rc)
    src = JpegSource(data, buf.raw)
    if src.shape[0] * src.shape[1] > MAX_PIXELS:
        return None  # left to Pillow, whose decompression-bomb check then speaks (as in the reference)
    return src


def load(path) -> Union[JpegSource, np.ndarray]:
    """What the OAKE datasets hand to the pipeline for one image file: the compressed file when the
    GPU can decode it, else the pixels from Pillow (any other format, or a JPEG flavour outside the
    envelope)."""
    with open(path, 'rb') as f:
        data = f.read()
    if data[:2] == b'\xff\xd8':
        src = parse(data)
        if src is not None:
            return src
    import PIL.Image
    return np.asarray(PIL.Image.open(io.BytesIO(data)).convert('RGB'), dtype=np.uint8)
