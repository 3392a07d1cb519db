"""`oadp.dp.utils` -- same import path as the reference (oadp/dp/utils.py)."""
from oadp_b200.dp.utils import *  # noqa: F401,F403
from oadp_b200.dp import utils as _impl

__all__ = list(getattr(_impl, '__all__', [n for n in dir(_impl) if not n.startswith('_')]))
