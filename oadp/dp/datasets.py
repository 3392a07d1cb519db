"""`oadp.dp.datasets` -- same import path as the reference (oadp/dp/datasets.py)."""
from oadp_b200.dp.datasets import *  # noqa: F401,F403
from oadp_b200.dp import datasets as _impl

__all__ = list(getattr(_impl, '__all__', [n for n in dir(_impl) if not n.startswith('_')]))
