"""CPU restatement of the OAKE host front end (PIL + torchvision) -- TEST INFRASTRUCTURE.

Parity unpinned by the reference (no tests upstream).  The pixel arithmetic is done by the
very libraries the reference's DataLoader workers call (Pillow ``crop``/``resize`` and
torchvision ``Resize/CenterCrop/ToTensor/Normalize``), so this file only restates the
*geometry* around them:

  * CLIP preprocess: openai/CLIP ``clip.py::_transform(224)``.
  * globals: oadp/oake/globals.py:26-33.
  * blocks:  oadp/oake/blocks.py:40-109 (partition, pyramid, bbox quirk of the first row).
  * objects: oadp/oake/objects.py:76-186 (min_wh filter, ADAPTIVE expand, crop, mask).

``todd.BBoxes*`` (todd_ai 0.3.0, not installed) is replaced by explicit fp32 tensor math;
the choices that cannot be seen from the reference are marked ``[unseen]``.
"""
from __future__ import annotations

import itertools
import math
from typing import List, NamedTuple, Tuple

import PIL.Image
import torch
import torch.nn.functional as F
import torchvision.transforms as T

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def clip_transform(n_px: int = 224) -> T.Compose:
    """openai/CLIP ``_transform``: bicubic short-side resize, centre crop, normalise."""
    return T.Compose([
        T.Resize(n_px, interpolation=T.InterpolationMode.BICUBIC),
        T.CenterCrop(n_px),
        lambda im: im.convert('RGB'),
        T.ToTensor(),
        T.Normalize(CLIP_MEAN, CLIP_STD),
    ])


_TRANSFORM = clip_transform()


# ------------------------------------------------------------------------------ globals
def globals_preprocess(image: PIL.Image.Image) -> torch.Tensor:
    """globals.py:32 -> (3,224,224) f32."""
    return _TRANSFORM(image)


# ------------------------------------------------------------------------------- blocks
def partition(length: int, r: int = 224, s: int = 112) -> List[int]:
    """blocks.py:40-52: start offsets of r-wide windows with stride <= s covering length."""
    if length < r:
        return []
    starts = [0]
    if length == r:
        return starts
    n = (length - r - 1) // s + 1
    q, rem = divmod(length - r, n)
    for i in range(n):
        starts.append(starts[-1] + q + (1 if i < rem else 0))
    return starts


class BlocksBatch(NamedTuple):
    blocks: torch.Tensor  # (Nb,3,224,224) f32
    bboxes: torch.Tensor  # (Nb,4) f32 -- row 0 is (x0,y0,side,side), the rest xyxy (quirk)


def blocks_preprocess(image: PIL.Image.Image, r: int = 224, s: int = 112,
                      rescale: float = 1.5) -> BlocksBatch:
    """blocks.py:89-109 with the pyramid of blocks.py:54-77."""
    crops = [_TRANSFORM(image)]
    w, h = image.size
    bboxes = [((w - h) / 2, 0, h, h) if w > h else (0, (h - w) / 2, w, w)]
    scale = 1.0
    level = image
    while True:
        lw, lh = level.size
        cells = list(itertools.product(partition(lw, r, s), partition(lh, r, s)))
        if not cells:
            break
        for x, y in cells:  # x outer, y inner
            crops.append(_TRANSFORM(level.crop((x, y, x + r, y + r))))
            x1, y1, side = x * scale, y * scale, r * scale
            bboxes.append((x1, y1, x1 + side, y1 + side))
        level = level.resize((int(lw / rescale), int(lh / rescale)))  # PIL default BICUBIC
        scale *= rescale
    return BlocksBatch(torch.stack(crops), torch.tensor(bboxes))


# ------------------------------------------------------------------------------ objects
def min_wh_indices(xyxy: torch.Tensor, min_wh: Tuple[float, float] = (4, 4)) -> torch.Tensor:
    """todd ``BBoxes.indices(min_wh=...)`` [unseen]: inclusive >= on width and height."""
    wh = xyxy[:, 2:] - xyxy[:, :2]
    return (wh >= torch.tensor(min_wh, dtype=wh.dtype)).all(-1)


def expand_adaptive(xyxy: torch.Tensor, image_wh: torch.Tensor, mode: str = 'ADAPTIVE') -> torch.Tensor:
    """objects.py:76-114.  ExpandMode.ADAPTIVE (:94-99): square of side sqrt(8*area); ExpandMode.CONSTANT
    (:92-93): side 224; either pushed inside the image when it fits.  fp32 throughout; cxcywh -> xyxy as
    c -/+ wh/2 [unseen]."""
    lt, rb = xyxy[:, :2], xyxy[:, 2:]
    wh = rb - lt
    area = wh[:, 0] * wh[:, 1]
    if mode == 'CONSTANT':
        side = torch.full((xyxy.shape[0], 1), 224, dtype=xyxy.dtype)
    else:
        assert mode == 'ADAPTIVE', mode
        side = torch.sqrt(area * 8).unsqueeze(-1)
    center = (lt + rb) / 2
    swh = torch.cat([side, side], dim=-1)
    e_lt = center - swh / 2
    e_rb = center + swh / 2
    image_wh = image_wh.to(xyxy.dtype)
    offset = torch.zeros_like(e_lt)
    offset = torch.where(e_lt >= 0, offset, -e_lt)
    offset = torch.where(e_rb <= image_wh, offset, image_wh - e_rb)
    offset = torch.where(swh <= image_wh, offset, torch.tensor(0.0))
    center = center + offset
    return torch.cat([center - swh / 2, center + swh / 2], dim=-1)


def object_mask(foreground: Tuple[float, ...], obj: Tuple[float, ...], grid: int = 14) -> torch.Tensor:
    """objects.py:129-155: 1 = background, nearest-resampled to (1,1,grid,grid)."""
    x = torch.arange(obj[2] - obj[0])
    wm = ((foreground[0] <= x) & (x <= foreground[2])).reshape(1, -1)
    y = torch.arange(obj[3] - obj[1])
    hm = ((foreground[1] <= y) & (y <= foreground[3])).reshape(-1, 1)
    mask = ~(wm & hm)
    return F.interpolate(mask[None, None].float(), size=(grid, grid), mode='nearest')


class ObjectsBatch(NamedTuple):
    objects: torch.Tensor  # (No,3,224,224) f32
    bboxes: torch.Tensor  # (No,4) f32 -- the FILTERED ORIGINAL proposals (objects.py:183)
    objectness: torch.Tensor  # (No,1) f32
    masks: torch.Tensor  # (No,1,14,14) f32
    expanded: torch.Tensor  # (No,4) f32 -- the square crops actually cut (for tests)


def objects_preprocess(image: PIL.Image.Image, proposals: torch.Tensor, grid: int = 14,
                       dry_run: bool = False, expand_mode: str = 'ADAPTIVE') -> ObjectsBatch:
    """objects.py:157-186.  ``proposals`` is (N,5) f32: xyxy + objectness."""
    proposals = proposals.float()
    boxes, objectness = proposals.split((4, 1), dim=-1)
    keep = min_wh_indices(boxes)
    if dry_run:
        keep[5:] = False
    boxes, objectness = boxes[keep], objectness[keep]
    expanded = expand_adaptive(boxes, torch.tensor(image.size), expand_mode)
    lt2 = expanded[:, :2].repeat(1, 2)
    foregrounds = boxes - lt2
    crops, masks = [], []
    for fg, bb in zip(foregrounds.tolist(), expanded.tolist()):
        crops.append(_TRANSFORM(image.crop(tuple(bb))))  # PIL rounds half-to-even, pads 0
        masks.append(object_mask(tuple(fg), tuple(bb), grid))
    return ObjectsBatch(torch.stack(crops), boxes, objectness, torch.cat(masks), expanded)


def crops_per_image_blocks(w: int, h: int, r: int = 224, s: int = 112, rescale: float = 1.5) -> int:
    n = 1
    while True:
        c = len(partition(w, r, s)) * len(partition(h, r, s))
        if c == 0:
            return n
        n += c
        w, h = int(w / rescale), int(h / rescale)


assert crops_per_image_blocks(640, 480) == 27 and crops_per_image_blocks(640, 640) == 39, \
    'SURVEY 8(a) probe of blocks.py:40-77'
assert math.isclose(CLIP_STD[0], 0.26862954)
