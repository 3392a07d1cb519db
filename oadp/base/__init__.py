"""Import-path alias of the reference's `oadp.base` package (oadp/base/__init__.py:1-3)."""
from .globals_ import *  # noqa: F401,F403
from .losses import *  # noqa: F401,F403
