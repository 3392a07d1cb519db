"""Host-side mirror of the classifier surface of the reference's `oadp.dp` package
(oadp/dp/classifiers.py, oadp/dp/utils.py:47-51, oadp/base/globals_.py) on the sm_100a kernels."""
from .categories import Categories, Globals, coco  # noqa: F401
from .classifiers import BaseClassifier, Classifier, NormalizedLinear, ViLDClassifier  # noqa: F401
