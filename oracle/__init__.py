"""CPU oracle for the OAKE hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU (torch fp32 / numpy / PIL) restatement of the algorithms the
reference runs on the OAKE path and in the cosine classifier.  It is the *checker*:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``oadp_b200/`` imports it and
the product path fails loudly when the CUDA extension is missing.

Parity status: **pinned against outputs of the reference's own source files** where the
arithmetic lives in the reference repository, and against an independent implementation where it
lives in an un-vendored dependency.  LutingWang/OADP ships no tests, golden vectors or fixtures
 and cannot be imported as a package here (``clip``, ``todd``, ``mmcv``,
``mmdet`` are absent, no network), so ``tests/golden/make_ref_golden.py`` loads the individual
reference modules (oadp/oake/{base,globals,blocks,objects}.py, oadp/dp/{classifiers,utils}.py,
oadp/base/globals_.py) with importlib on stand-ins for the third-party packages
(``tests/golden/ref_stubs.py``) and records what they compute in ``tests/golden/ref_golden.pt``;
``tests/test_ref_golden.py`` holds this oracle to it:

* ``frontend.partition`` / ``blocks_preprocess`` / ``objects_preprocess`` (block grid, bboxes,
  min_wh filter, adaptive expansion, PIL crops, 14x14 masks): bit-exact.
* ``vit.objects_surgery`` + ``vit.encode_objects`` against ``Validator._build_model`` + ``Hooks``
  (objects.py:198-314) driving ``model.visual(o, m)``: max-abs < 3e-5 at full depth.
* ``vit.encode_image`` against ``model.encode_image`` as called by globals.py:57 / blocks.py:129.
* ``classifier`` against ``BaseClassifier`` / ``Classifier`` / ``ViLDClassifier``.
* ``text.encode_text`` (the CLIP text tower behind oadp/prompts/vild.py:56-72) against HuggingFace
  ``CLIPTextModelWithProjection`` under ``text.to_hf_state_dict`` (``tests/test_oracle_text.py``).
* ``jpeg.decode`` IS the reference's loader (``PIL.Image.open(f).convert('RGB')``, base.py:53):
  Pillow runs here and on the GPU box, so that row needs no restatement.

What remains a restatement is the ``clip`` package itself (LutingWang/CLIP, an unpinned fork of
openai/CLIP, README.md:44 of the reference): its module tree is restated once in ``ref_stubs.py``
(driven by the reference's hooks) and once, functionally, in ``vit.py``; the two are also checked
against a third, independent implementation:

* ``vit.encode_image`` (T=50) against HuggingFace ``CLIPVisionModelWithProjection`` under the
  weight mapping in ``vit.to_hf_state_dict``.
* ``vit.encode_objects`` against ``hooks_ref.HookedVisual`` (hook-driven, ``nn.MultiheadAttention``).
* ``frontend`` uses PIL + the published torchvision transform semantics directly, i.e. the same
  library code the reference's DataLoader workers execute.

Unseen third-party behaviour (the fork's positional-embedding interpolation
mode, todd ``BBoxes.indices`` strictness) is an explicit, documented choice shared by the stubs and
the oracle.
"""
