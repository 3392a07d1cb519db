"""`oadp.dp.roi_heads` -- same import path as the reference (oadp/dp/roi_heads.py)."""
from oadp_b200.dp.roi_heads import *  # noqa: F401,F403
from oadp_b200.dp import roi_heads as _impl

__all__ = list(getattr(_impl, '__all__', [n for n in dir(_impl) if not n.startswith('_')]))
