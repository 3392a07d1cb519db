"""`oadp.base.globals_` -- category tables and the global state the classifiers read (oadp/base/globals_.py)."""
from oadp_b200.dp.categories import Categories, Globals, coco  # noqa: F401

__all__ = ['Categories', 'Globals', 'coco']
