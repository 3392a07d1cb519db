"""GPU tier for the JPEG row: `oake_jpeg_decode` through the C-ABI against Pillow, the
reference's own decoder (oadp/oake/base.py:53) -- bit-exact uint8 pixels -- and the compressed-input
path of the pipeline / the `.decode:gpu` CLI option against the decoded-input path (identical files)."""
import io
import pathlib

import numpy as np
import PIL.Image
import pytest
import torch

from oadp_b200 import binding, synth
from oadp_b200 import jpeg as oake_jpeg
from oracle import jpeg as ojpeg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def pipe(lib):
    from oadp_b200.model import OakeModel
    from oadp_b200.pipeline import OakePipeline
    return OakePipeline(OakeModel(synth.visual_params(0), 'cuda').engine)


def test_decode_matches_pillow_bit_for_bit(pipe):
    labels, sources, want = [], [], []
    for label, data in ojpeg.corpus(0):
        src = oake_jpeg.parse(data)
        assert src is not None, label
        labels.append(label)
        sources.append(src)
        want.append(ojpeg.decode(data))
    got = pipe.decode_jpegs(sources)  # one batch: every size / sampling / restart mix side by side
    for label, g, w in zip(labels, got, want):
        assert g.shape == w.shape and np.array_equal(g, w), label
    # batch composition must not matter
    again = pipe.decode_jpegs(sources[5:6])
    assert np.array_equal(again[0], want[5])


def test_coco_sized_files(pipe):
    sources, want = [], []
    for i, (w, h) in enumerate(synth.COCO_SIZES):
        buf = io.BytesIO()
        PIL.Image.fromarray(synth.image(w, h, 50 + i)).save(buf, 'JPEG', quality=(95, 85, 75)[i % 3],
                                                            subsampling=(2, 0, 1)[i % 3])
        sources.append(oake_jpeg.parse(buf.getvalue()))
        want.append(ojpeg.decode(buf.getvalue()))
    for g, w in zip(pipe.decode_jpegs(sources), want):
        assert np.array_equal(g, w)


def test_damaged_file_fails_loudly(pipe):
    _, data = next(iter(ojpeg.corpus(1)))
    src = oake_jpeg.parse(data)
    cut = oake_jpeg.parse(data[:len(data) * 2 // 3])  # headers intact, entropy-coded data ends early
    with pytest.raises(binding.OakeError, match='damaged'):
        pipe.decode_jpegs([src, cut])
    with pytest.raises(binding.OakeError, match=r'\[1\]'):
        pipe.encode_globals([src, cut])
    assert len(pipe.encode_globals([src])) == 1  # the pipeline is still usable


def test_pipeline_takes_compressed_and_decoded_images_alike(pipe):
    files = []
    for i, (w, h) in enumerate(synth.COCO_SIZES[:4]):
        buf = io.BytesIO()
        PIL.Image.fromarray(synth.image(w, h, 70 + i)).save(buf, 'JPEG', quality=90, subsampling=(2, 0)[i % 2])
        files.append(buf.getvalue())
    decoded = [ojpeg.decode(f) for f in files]
    compressed = [oake_jpeg.parse(f) for f in files]
    mixed = [compressed[0], decoded[1], compressed[2], decoded[3]]
    ref = pipe.encode_blocks(decoded)
    for batch in (compressed, mixed):
        for a, b in zip(pipe.encode_blocks(batch), ref):
            assert torch.equal(a['embeddings'], b['embeddings']) and torch.equal(a['bboxes'], b['bboxes'])
    g_ref = pipe.encode_globals(decoded)
    for a, b in zip(pipe.encode_globals(compressed), g_ref):
        assert torch.equal(a, b)
    props = [synth.proposals(d.shape[1], d.shape[0], 30, seed=i) for i, d in enumerate(decoded)]
    o_ref = pipe.encode_objects(decoded, props)
    for a, b in zip(pipe.encode_objects(compressed, props), o_ref):
        assert torch.equal(a['embeddings'], b['embeddings'])
    # back-to-back submissions (two slots in flight, decode on the side stream)
    # -- and more submissions than slots before anything is collected: handles stay valid
    tickets = [pipe.submit_blocks(compressed[i:i + 2]) for i in (0, 2)] + [pipe.submit_blocks(mixed[1:3])]
    outs = [t.result() for t in reversed(tickets)][::-1]
    for a, b in zip(outs[0] + outs[1] + outs[2], ref + ref[1:3]):
        assert torch.equal(a['embeddings'], b['embeddings'])


def test_mixed_batch_survives_stale_staging_bytes_and_a_busy_main_stream(pipe):
    """A batch mixing decoded arrays with compressed files: the decode stream writes the compressed images'
    pixels into the arena while the main stream is still busy with earlier work.  The main stream must not
    copy the staging bytes of those images over them (only the raw images' ranges travel)."""
    files = []
    for i, (w, h) in enumerate(synth.COCO_SIZES[:4]):
        buf = io.BytesIO()
        PIL.Image.fromarray(synth.image(w, h, 90 + i)).save(buf, 'JPEG', quality=90)
        files.append(buf.getvalue())
    decoded = [ojpeg.decode(f) for f in files]
    compressed = [oake_jpeg.parse(f) for f in files]
    ref = pipe.encode_globals(decoded)
    for _ in range(2):  # both slots
        pipe._cur ^= 1
        slot = pipe._slots[pipe._cur]
        if slot.ticket is not None:
            slot.ticket.collect()
        if slot.arena.host is not None:
            slot.arena.host.fill_(0xA5)  # poison: whatever a whole-arena copy would carry
    busy = torch.randn(8192, 8192, device='cuda')
    for _ in range(20):  # ~100 ms of queued main-stream work ahead of the next submission
        busy = (busy @ busy).clamp_(-1, 1)
    mixed = [compressed[0], decoded[1], compressed[2], decoded[3]]
    for a, b in zip(pipe.encode_globals(mixed), ref):
        assert torch.equal(a, b)


def test_cli_decode_gpu_writes_the_same_files(tmp_path_factory, lib, monkeypatch):
    monkeypatch.delenv('DRY_RUN', raising=False)
    monkeypatch.setenv('OAKE_ALLOW_RANDOM_WEIGHTS', '1')
    monkeypatch.delenv('OAKE_CLIP_WEIGHTS', raising=False)
    import oadp.oake.blocks as cli_blocks
    ds = synth.write_coco_dataset(tmp_path_factory.mktemp('coco_jpg'), 5, seed=4, n_proposals=10, fmt='jpg')
    root = pathlib.Path(ds['root'])
    out = {}
    for mode in ('pillow', 'gpu'):
        cli_blocks.Validator.main(['t', ds['configs']['blocks'], '--override', f'.decode:{mode}',
                                   f'.val.dataloader.dataset.output_dir::{root}/{mode}/val',
                                   f'.train.dataloader.dataset.output_dir::{root}/{mode}/train'])
        out[mode] = {f.name: torch.load(f) for f in sorted((root / mode / 'val').glob('*.pth'))}
    assert len(out['gpu']) == 5 and out['gpu'].keys() == out['pillow'].keys()
    for k, v in out['gpu'].items():
        assert torch.equal(v['embeddings'], out['pillow'][k]['embeddings'])
        assert torch.equal(v['bboxes'], out['pillow'][k]['bboxes'])
