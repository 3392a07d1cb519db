"""`LoadCLIPFeatures` in the `PIPELINES` registry -- counterpart of oadp/dp/datasets.py:137-214.

The pipeline step that feeds the cached OAKE features to the detector.  The implementation lives with the
feature store it reads (`oadp_b200.store.LoadCLIPFeatures`: file-per-key or packed shards behind the same
`Mapping[key]` contract); this module registers it under the name the reference's dataset configs use
(configs/dp/datasets/ov_coco.py:23-32 `dict(type='LoadCLIPFeatures', default=..., globals_=..., ...)`) --
in mmdet's `PIPELINES` when mmdet is importable, else in `oadp_b200.registry.PIPELINES`.  The evaluators
(`OV_COCO`, `OV_LVIS`: pycocotools / lvis) are out of scope.
"""
from __future__ import annotations

from ..registry import PIPELINES
from ..store import LoadCLIPFeatures

__all__ = ['LoadCLIPFeatures']

if 'LoadCLIPFeatures' not in getattr(PIPELINES, 'module_dict', {}):
    PIPELINES.register_module(name='LoadCLIPFeatures', module=LoadCLIPFeatures)
