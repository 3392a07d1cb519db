"""CPU tier for the JPEG row.

1. The per-item functions the CUDA kernels wrap (oadp_b200/csrc/jpeg_core.cuh: entropy decode, islow
   IDCT, fancy upsampling, YCbCr->RGB) compiled with g++ into a throw-away harness
   (tests/jpeg_host_harness.cpp) and held bit-exact to Pillow -- the reference's decoder,
   oadp/oake/base.py:53 -- over oracle.jpeg.corpus().
2. The host half of the C-ABI (`oake_jpeg_parse` / `oake_jpeg_place`, no GPU involved) through
   liboake_b200.so: geometry, the unsupported envelope, malformed input, offset rebasing.
"""
import ctypes
import io
import pathlib
import struct
import subprocess

import numpy as np
import PIL.Image
import pytest

from oadp_b200 import binding
from oadp_b200 import jpeg as oake_jpeg
from oracle import jpeg as ojpeg

ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.fixture(scope='module')
def harness(tmp_path_factory):
    so = tmp_path_factory.mktemp('jpeg_harness') / 'harness.so'
    subprocess.run(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-Wno-unknown-pragmas', '-o', str(so),
                    str(ROOT / 'tests' / 'jpeg_host_harness.cpp')], check=True)
    lib = ctypes.CDLL(str(so))
    lib.harness_decode.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int),
                                   ctypes.POINTER(ctypes.c_int)]

    def decode(data: bytes):
        w, h = ctypes.c_int(), ctypes.c_int()
        rc = lib.harness_decode(data, len(data), None, w, h)
        if rc:
            return rc, None
        out = np.zeros((h.value, w.value, 3), np.uint8)
        return lib.harness_decode(data, len(data), out.ctypes.data, w, h), out

    return decode


def test_kernel_arithmetic_matches_pillow_bit_for_bit(harness):
    n = 0
    for label, data in ojpeg.corpus(0):
        rc, got = harness(data)
        assert rc == 0, label
        assert np.array_equal(got, ojpeg.decode(data)), label
        n += 1
    assert n > 300


def test_damaged_entropy_data_is_reported(harness):
    _, data = next(iter(ojpeg.corpus(1)))
    rc, _ = harness(data[:len(data) * 2 // 3])
    assert rc == 3
    for label, data in ojpeg.outside_envelope():
        assert harness(data)[0] == oake_jpeg.UNSUPPORTED, label


def test_parse_geometry_and_envelope(lib):
    im = PIL.Image.fromarray(np.random.default_rng(0).integers(0, 256, (45, 67, 3), dtype=np.uint8))
    for sub, (hmax, vmax) in ((0, (1, 1)), (1, (2, 1)), (2, (2, 2))):
        buf = io.BytesIO()
        im.save(buf, 'JPEG', quality=80, subsampling=sub)
        src = oake_jpeg.parse(buf.getvalue())
        assert src is not None and src.shape == (45, 67, 3) and src.size == 45 * 67 * 3
        width, height, ncomp, h, v, mx, my = struct.unpack_from('<7I', src.desc, 0)
        assert (width, height, ncomp, h, v) == (67, 45, 3, hmax, vmax)
        assert (mx, my) == (-(-67 // (8 * hmax)), -(-45 // (8 * vmax)))
        # coefficients (128 B / block) + planes (64 B / block), each component 256-byte aligned
        blocks = mx * my * (hmax * vmax + 2)
        assert blocks * 192 <= src.scratch_bytes <= blocks * 192 + 6 * 256
    for label, data in ojpeg.outside_envelope():
        assert oake_jpeg.parse(data) is None, label
    with pytest.raises(binding.OakeError, match='SOI'):
        oake_jpeg.parse(b'not a jpeg at all')
    buf = io.BytesIO()
    im.save(buf, 'JPEG')
    with pytest.raises(binding.OakeError):
        oake_jpeg.parse(buf.getvalue()[:100])  # ends inside the tables


def test_place_rebases_offsets(lib):
    buf = io.BytesIO()
    PIL.Image.fromarray(np.zeros((16, 16, 3), np.uint8)).save(buf, 'JPEG')
    src = oake_jpeg.parse(buf.getvalue())
    raw = ctypes.create_string_buffer(src.desc, len(src.desc))
    scan_off0, scan_len0 = struct.unpack_from('<2Q', src.desc, 40)
    scratch = ctypes.c_uint64(1000)  # not aligned on purpose
    assert lib.oake_jpeg_place(raw, 4096, 123, ctypes.byref(scratch)) == 0
    scan_off, scan_len, out_off, scratch_bytes = struct.unpack_from('<4Q', raw.raw, 40)
    assert (scan_off, scan_len, out_off) == (scan_off0 + 4096, scan_len0, 123)
    assert scratch.value == 1024 + scratch_bytes and scratch_bytes == src.scratch_bytes


def test_load_falls_back_to_pillow_outside_the_envelope(lib, tmp_path):
    arr = np.random.default_rng(3).integers(0, 256, (30, 40, 3), dtype=np.uint8)
    PIL.Image.fromarray(arr).save(tmp_path / 'a.png')
    got = oake_jpeg.load(tmp_path / 'a.png')
    assert isinstance(got, np.ndarray) and np.array_equal(got, arr)
    PIL.Image.fromarray(arr).save(tmp_path / 'p.jpg', progressive=True)
    got = oake_jpeg.load(tmp_path / 'p.jpg')
    assert isinstance(got, np.ndarray) and np.array_equal(got, ojpeg.decode((tmp_path / 'p.jpg').read_bytes()))
    PIL.Image.fromarray(arr).save(tmp_path / 'b.jpg')
    assert isinstance(oake_jpeg.load(tmp_path / 'b.jpg'), oake_jpeg.JpegSource)
