"""CPU tier for the JPEG row.

1. The per-item functions the CUDA kernels wrap (oadp_b200/csrc/jpeg_core.cuh: entropy decode, islow
   IDCT, fancy upsampling, YCbCr->RGB) compiled with g++ into a throw-away harness
   (tests/jpeg_host_harness.cpp) and held bit-exact to Pillow -- the reference's decoder,
   oadp/oake/base.py:53 -- over oracle.jpeg.corpus().
2. The host half of the C-ABI (`oake_jpeg_parse` / `oake_jpeg_stage`, no GPU involved) through
   liboake_b200.so: geometry, the unsupported envelope, malformed input, offset rebasing.
"""
import ctypes
import io
import os
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
                                   ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.POINTER(ctypes.c_int)]

    def decode(data: bytes, parallel: bool = False):
        """-> (rc, pixels, rounds the parallel decode needed to synchronise)"""
        w, h, rounds = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(-1)
        rc = lib.harness_decode(data, len(data), None, w, h, 0, None)
        if rc:
            return rc, None, -1
        out = np.zeros((h.value, w.value, 3), np.uint8)
        rc = lib.harness_decode(data, len(data), out.ctypes.data, w, h, int(parallel), rounds)
        return rc, out, rounds.value

    return decode


@pytest.mark.parametrize('parallel', [False, True], ids=['serial', 'subsequence-parallel'])
def test_kernel_arithmetic_matches_pillow_bit_for_bit(harness, parallel):
    n = 0
    for label, data in ojpeg.corpus(0):
        rc, got, _ = harness(data, parallel)
        assert rc == 0, label
        assert np.array_equal(got, ojpeg.decode(data)), label
        n += 1
    assert n > 300


def test_parallel_decode_synchronises_in_a_few_rounds(harness):
    """COCO-sized files: hundreds to thousands of subsequences, yet the entry states settle in a handful of rounds."""
    from oadp_b200 import synth
    for i, (quality, sub) in enumerate(((90, 2), (75, 0), (98, 1), (30, 2))):
        w, h = synth.COCO_SIZES[i]
        buf = io.BytesIO()
        PIL.Image.fromarray(synth.image(w, h, 11 + i)).save(buf, 'JPEG', quality=quality, subsampling=sub)
        data = buf.getvalue()
        rc, got, rounds = harness(data, True)
        assert rc == 0 and np.array_equal(got, ojpeg.decode(data))
        assert len(data) * 8 // 1024 > 20 and 1 <= rounds <= 24, (len(data), rounds)


def test_damaged_entropy_data_is_reported(harness):
    _, data = next(iter(ojpeg.corpus(1)))
    for parallel in (False, True):
        assert harness(data[:len(data) * 2 // 3], parallel)[0] == 3
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


def test_stage_strips_stuffing_and_rebases_offsets(lib):
    rng = np.random.default_rng(5)
    buf = io.BytesIO()
    PIL.Image.fromarray(rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)).save(buf, 'JPEG', quality=100,
                                                                                  restart_marker_blocks=2)
    data = buf.getvalue()
    src = oake_jpeg.parse(data)
    scan_off0, scan_len0 = struct.unpack_from('<2Q', src.desc, 40)
    assert scan_off0 + scan_len0 == len(data)
    assert lib.oake_jpeg_stream_bound(src.desc) == src.stream_bound
    dst = ctypes.create_string_buffer(b'\xaa' * (src.stream_bound + 8), src.stream_bound + 8)
    placed = ctypes.create_string_buffer(len(src.desc))
    scratch, written = ctypes.c_uint64(1000), ctypes.c_uint64(0)  # scratch offset not aligned on purpose
    assert lib.oake_jpeg_stage(src.desc, data, len(data), dst, 4096, 123, ctypes.byref(scratch), placed,
                               ctypes.byref(written)) == 0
    scan_off, scan_len, out_off, scratch_bytes = struct.unpack_from('<4Q', placed.raw, 40)
    assert (scan_off, out_off) == (4096, 123)
    assert scratch.value == 1024 + scratch_bytes and scratch_bytes == src.scratch_bytes
    # the clean stream: stuffed zeros gone, RSTn markers kept, EOI cut, zero padding, nothing beyond
    scan = data[scan_off0:]
    assert scan.count(b'\xff\x00') > 0
    want = scan[:scan.rindex(b'\xff\xd9')].replace(b'\xff\x00', b'\xff')
    assert scan_len == len(want) and dst.raw[:scan_len] == want and b'\xff\xd0' in want
    # ... then the restart table: where each of the 16 MCUs / 2 = 8 intervals starts in the clean stream
    mcus_x, mcus_y, interval = struct.unpack_from('<3I', placed.raw, 20)
    count = struct.unpack_from('<I', placed.raw, 80)[0]
    assert (mcus_x * mcus_y, interval, count) == (16, 2, 8)
    table_at = (scan_len + 3) // 4 * 4 + 16
    assert written.value == table_at + 4 * count <= src.stream_bound
    assert dst.raw[scan_len:table_at] == bytes(table_at - scan_len)
    starts = list(struct.unpack_from('<8I', dst.raw, table_at))
    # (markers can only be told from data in the raw scan, where a data byte 0xFF is followed by 0x00)
    markers, i, clean = [], 0, 0
    while i < len(scan) - 1:
        if scan[i] == 0xFF and scan[i + 1] == 0x00:
            i, clean = i + 2, clean + 1
        elif scan[i] == 0xFF and 0xD0 <= scan[i + 1] <= 0xD7:
            i, clean = i + 2, clean + 2
            markers.append(clean)
        elif scan[i] == 0xFF:
            break
        else:
            i, clean = i + 1, clean + 1
    assert starts == [0] + markers and len(markers) == 7
    # a file that lost one of its markers: the interval nobody found is flagged, the decode reports it
    assert dst.raw[written.value:] == b'\xaa' * (len(dst.raw) - written.value)
    assert lib.oake_jpeg_stage(src.desc, data, len(data) - 1, dst, 0, 0, ctypes.byref(scratch), placed,
                               ctypes.byref(written)) != 0  # not the file the descriptor came from


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


def _mutations(seed: int, per_file: int):
    """Valid files and damaged variants of them: flipped bits, truncation, scribbled headers, stray
    markers in the scan, absurd frame sizes."""
    rng = np.random.default_rng(seed)
    base = []
    for (w, h), quality, sub, rst in (((64, 48), 90, 2, 0), ((67, 45), 75, 0, 0), ((33, 130), 95, 1, 3),
                                      ((17, 9), 30, 2, 0), ((120, 80), 85, 2, 5)):
        coarse = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2, 3), dtype=np.uint8)
        im = PIL.Image.fromarray(coarse).resize((w, h), PIL.Image.BICUBIC)
        kw = dict(quality=quality, subsampling=sub)
        if rst:
            kw['restart_marker_blocks'] = rst
        for image in (im, im.convert('L')):
            buf = io.BytesIO()
            image.save(buf, 'JPEG', **kw)
            base.append(buf.getvalue())
    for data in base:
        yield data
        for _ in range(per_file):
            m = bytearray(data)
            kind = int(rng.integers(0, 5))
            if kind == 0:
                for _ in range(int(rng.integers(1, 8))):
                    m[int(rng.integers(0, len(m)))] ^= 1 << int(rng.integers(0, 8))
            elif kind == 1:
                m = m[:int(rng.integers(0, len(m)))]
            elif kind == 2:
                for _ in range(int(rng.integers(1, 6))):
                    m[int(rng.integers(0, min(len(m), 700)))] = int(rng.integers(0, 256))
            elif kind == 3:
                i = int(rng.integers(len(m) // 2, len(m)))
                m[i:i] = bytes([0xFF, int(rng.integers(0, 256))])
            else:
                j = bytes(m).find(b'\xff\xc0')
                if j > 0:
                    m[j + 5:j + 9] = struct.pack('>HH', int(rng.integers(0, 3000)), int(rng.integers(0, 3000)))
            yield bytes(m)


def test_damaged_files_never_touch_memory_out_of_bounds(tmp_path):
    """The parser, the staging pass and the device-side decode functions under ASAN + UBSAN on ~600 damaged
    files (the GPU kernels run the same functions on whatever the dataset directory holds)."""
    exe = tmp_path / 'jpeg_fuzz'
    build = subprocess.run(['g++', '-O1', '-g', '-std=c++17', '-fsanitize=address,undefined', '-fno-omit-frame-pointer',
                            '-Wno-unknown-pragmas', '-o', str(exe), str(ROOT / 'tests' / 'jpeg_fuzz_main.cpp')],
                           capture_output=True, text=True)
    if build.returncode != 0 and 'sanitize' in build.stderr:
        pytest.skip('this g++ has no sanitizer runtime')
    assert build.returncode == 0, build.stderr
    corpus = tmp_path / 'corpus.bin'
    n = 0
    with open(corpus, 'wb') as f:
        for data in _mutations(1, 60):
            f.write(struct.pack('<I', len(data)))
            f.write(data)
            n += 1
    run = subprocess.run([str(exe), str(corpus)], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, ASAN_OPTIONS='detect_leaks=0'))
    assert run.returncode == 0, run.stderr[-2000:]
    assert 'runtime error' not in run.stderr and 'AddressSanitizer' not in run.stderr, run.stderr[-2000:]
    assert run.stdout.startswith(f'records {n} decoded ')


def test_missing_restart_marker_is_reported(harness):
    rng = np.random.default_rng(6)
    buf = io.BytesIO()
    PIL.Image.fromarray(rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)).save(buf, 'JPEG', quality=90,
                                                                                  restart_marker_blocks=2)
    data = buf.getvalue()
    assert harness(data)[0] == 0
    i = data.index(b'\xff\xd3')
    assert harness(data[:i] + data[i + 2:])[0] == 3  # RST3 cut out: intervals no longer line up


def test_absurd_frame_size_is_left_to_pillow(lib):
    buf = io.BytesIO()
    PIL.Image.fromarray(np.zeros((16, 16, 3), np.uint8)).save(buf, 'JPEG')
    data = bytearray(buf.getvalue())
    j = bytes(data).find(b'\xff\xc0')
    data[j + 5:j + 9] = struct.pack('>HH', 60000, 60000)  # 3.6 gigapixels in the header, 16 x 16 of data
    assert oake_jpeg.parse(bytes(data)) is None
