"""Category tables and the process-global state the classifier reads at call time -- the part of
oadp/base/globals_.py the hot path depends on (classifiers.py:34-41,62-67,108).

`Globals.categories` / `Globals.training` are plain class attributes, read at *call* time exactly as
in the reference.  The OV-COCO 48/17 split is the public split of Bansal et
al.; the LVIS 866/337 split is derived from an LVIS annotation file (`Categories.from_lvis`): bases =
frequent + common, novels = rare, each in category-id order -- there is no hard-coded copy of the
1203 names here.
"""
from __future__ import annotations

import json
from typing import Iterable, Tuple


class Categories:

    def __init__(self, bases: Iterable[str], novels: Iterable[str]) -> None:
        self._bases = tuple(bases)
        self._novels = tuple(novels)

    bases = property(lambda self: self._bases)
    novels = property(lambda self: self._novels)

    @property
    def all_(self) -> Tuple[str, ...]:
        return self._bases + self._novels

    num_bases = property(lambda self: len(self._bases))
    num_novels = property(lambda self: len(self._novels))
    num_all = property(lambda self: len(self._bases) + len(self._novels))

    @classmethod
    def from_lvis(cls, annotation_file: str) -> 'Categories':
        with open(annotation_file) as f:
            cats = sorted(json.load(f)['categories'], key=lambda c: c['id'])
        return cls((c['name'] for c in cats if c['frequency'] in ('f', 'c')),
                   (c['name'] for c in cats if c['frequency'] == 'r'))


class Globals:
    """Entry point for global state (not instantiable, like the reference's NonInstantiableMeta)."""
    categories: Categories
    training: bool = False

    def __new__(cls, *args, **kwargs):
        raise TypeError('Globals is a namespace')


_COCO_48 = ('person bicycle car motorcycle train truck boat bench bird horse sheep bear zebra giraffe backpack '
            'handbag suitcase frisbee skis kite surfboard bottle fork spoon bowl banana apple sandwich orange '
            'broccoli carrot pizza donut chair bed toilet tv laptop mouse remote microwave oven toaster '
            'refrigerator book clock vase toothbrush').split()
_COCO_17 = 'airplane bus cat dog cow elephant umbrella tie snowboard skateboard cup knife cake couch keyboard sink scissors'.split()
coco = Categories(_COCO_48, _COCO_17)
assert coco.num_bases == 48 and coco.num_novels == 17
