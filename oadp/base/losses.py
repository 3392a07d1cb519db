"""`oadp.base.losses` -- `AsymmetricLoss`, `RKDLoss` in the todd-style `LossRegistry` (oadp/base/losses.py:10,68)."""
from oadp_b200.dp.losses import AsymmetricLoss, RKDLoss  # noqa: F401

__all__ = ['AsymmetricLoss', 'RKDLoss']
