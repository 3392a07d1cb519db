"""GPU tier: the sm_100a pipeline (uint8 image in, fp16 embeddings out, through the C ABI) against
tests/golden/ref_golden.pt, i.e. against outputs of the reference's own source files
(tests/golden/make_ref_golden.py).  Bar: 1 - cos < 1e-3 per crop (north star); bboxes / objectness
bit-exact in the stored fp16 format."""
import pathlib
import sys

import pytest
import torch
import torch.nn.functional as F

from oadp_b200 import synth
from oadp_b200.model import OakeEngine
from oadp_b200.pipeline import OakePipeline

sys.path.insert(0, str(pathlib.Path(__file__).parent / 'golden'))
import make_ref_golden as mk  # noqa: E402  (helpers only)

pytestmark = pytest.mark.gpu

REF = torch.load(pathlib.Path(__file__).parent / 'golden' / 'ref_golden.pt', weights_only=False)
COS_BAR = 1e-3


@pytest.fixture(scope='module')
def pipe(lib):
    return OakePipeline(OakeEngine(synth.visual_params(REF['weight_seed']), 'cuda'))


def worst(got, want):
    return float((1 - F.cosine_similarity(got.float().cpu(), want.float(), dim=-1)).max())


def test_globals_vs_reference(pipe):
    images, _ = mk.ref_inputs()
    got = pipe.encode_globals(images)
    for g, want in zip(got, REF['globals']):
        assert g.dtype == torch.float16 and g.shape == (512, )
        assert worst(g[None], want['embedding'][None]) < COS_BAR


def test_blocks_vs_reference(pipe):
    images, _ = mk.ref_inputs()
    got = pipe.encode_blocks(images)
    for g, want in zip(got, REF['blocks']):
        assert g['embeddings'].shape == (want['n'], 512)
        assert torch.equal(g['bboxes'].cpu(), want['bboxes_half'])
        ref6 = F.normalize(want['raw6'], dim=-1)
        assert worst(g['embeddings'][:6], ref6) < COS_BAR


def test_objects_vs_reference_hooks(pipe):
    """The whole objects path -- min_wh filter, expansion, crop, masks, stride-16 tower, hook-driven
    side stream -- against model.visual(o, m) of the reference (objects.py:157-186, 198-338)."""
    images, proposals = mk.ref_inputs()
    got = pipe.encode_objects(images, proposals)
    for g, want in zip(got, REF['objects']):
        assert torch.equal(g['bboxes'].cpu(), want['bboxes'].half())
        assert torch.equal(g['objectness'].cpu(), want['objectness'].half())
        assert g['embeddings'].shape == want['embeddings'].shape
        assert worst(g['embeddings'], want['embeddings']) < COS_BAR
