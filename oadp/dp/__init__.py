"""Import-path alias of the reference's `oadp.dp` package (oadp/dp/__init__.py:1-6): the classifier, head,
pipeline and utility names resolve to the B200-native implementation in `oadp_b200.dp`."""
from .bbox_heads import *  # noqa: F401,F403
from .classifiers import *  # noqa: F401,F403
from .datasets import *  # noqa: F401,F403
from .roi_heads import *  # noqa: F401,F403
from .utils import *  # noqa: F401,F403
