"""Import-path alias: the reference's module paths (`python -m oadp.oake.globals ...`, `oadp.dp.classifiers`,
`oadp.base.Globals`) resolve to the B200-native implementation in `oadp_b200` (oadp/__init__.py:1-9)."""
from . import base, dp, oake  # noqa: F401
