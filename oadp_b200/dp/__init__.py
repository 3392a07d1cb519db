"""Host-side mirror of the reference's `oadp.dp` package on the sm_100a kernels: the cosine classifiers
(oadp/dp/classifiers.py, utils.py:47-51), the bbox / RoI heads that call them (bbox_heads.py, roi_heads.py),
the `LoadCLIPFeatures` pipeline step (datasets.py:137-214), the distillation losses (oadp/base/losses.py) and
the category tables / global state they read (oadp/base/globals_.py).  The detector (`detectors.py`), the
evaluators and the runners stay mmdet's."""
from .categories import Categories, Globals, coco  # noqa: F401
from .classifiers import BaseClassifier, Classifier, NormalizedLinear, ViLDClassifier  # noqa: F401
from .losses import AsymmetricLoss, L1Loss, MSELoss, RKDLoss  # noqa: F401
from .utils import MultilabelTopKRecall  # noqa: F401
from .bbox_heads import (BlockMixin, ObjectMixin, Shared2FCBlockBBoxHead,  # noqa: F401
                         Shared4Conv1FCObjectBBoxHead)
from .roi_heads import OADPRoIHead, ViLDEnsembleRoIHead, ensemble_lambda, vild_ensemble  # noqa: F401
from .datasets import LoadCLIPFeatures  # noqa: F401
