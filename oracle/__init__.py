"""CPU oracle for the OAKE hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU (torch fp32 / numpy / PIL) restatement of the algorithms the
reference runs on the OAKE path and in the cosine classifier.  It is the *checker*:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``oadp_b200/`` imports it and
the product path fails loudly when the CUDA extension is missing.

Parity status: **parity unpinned by the reference** -- LutingWang/OADP ships no tests,
golden vectors or fixtures and its own implementation cannot be
imported here (``clip``, ``todd``, ``mmcv``, ``mmdet`` are absent, no network).  The
encoder arithmetic lives in the un-vendored ``clip`` dependency (LutingWang/CLIP, an
unpinned fork of openai/CLIP, README.md:44 of the reference).  What pins this oracle
instead:

* ``vit.encode_image`` (T=50) is cross-checked against HuggingFace
  ``CLIPVisionModelWithProjection`` (an independent implementation of the same
  published architecture) under the weight mapping in ``vit.to_hf_state_dict``.
* ``vit.encode_objects`` (T=197 + masked CLS side stream) is cross-checked against
  ``hooks_ref.HookedVisual``, a second, structurally different restatement that keeps
  the reference's hook-based control flow (oadp/oake/objects.py:198-314) on top of
  ``torch.nn.MultiheadAttention``.
* ``frontend`` uses PIL + the published torchvision transform semantics directly, i.e.
  the same library code the reference's DataLoader workers execute.
"""
