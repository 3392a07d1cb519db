"""Host geometry of the OAKE front end + the device arena that feeds liboake_b200's resize / encode.

The reference cuts and resizes every crop with PIL inside DataLoader workers
(oadp/oake/globals.py:26-33, blocks.py:40-109, objects.py:76-186).  Here the host only does the
*integer geometry* (which rectangle, which output size, which centre window -- vectorised numpy,
microseconds per image); the pixels never leave the GPU: one uint8 upload per image, then
`oake_resize_u8` (Pillow-exact bicubic) and `oake_encode_crops_u8`.

Geometry rules reproduced:
  * torchvision Resize(224): short side -> 224, long side -> int(224 * long / short); CenterCrop
    offsets int(round((n - 224) / 2.0)) (round half to even).
  * PIL `Image.crop(box)` rounds float coordinates half-to-even and zero-pads outside the image.
  * blocks: `_partition` starts, pyramid `int(w / 1.5)`, bbox quirk of the first row (xywh-like).
  * objects: min_wh (4,4) filter (inclusive [unseen in the reference: todd.BBoxes.indices]),
    ADAPTIVE expand = square of side sqrt(8 * area) pushed inside the image when it fits, fp32.
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional, Sequence, Tuple

import numpy as np

SIZE = 224

RESIZE_JOB = np.dtype([
    ('src_off', '<i8'), ('dst_off', '<i8'), ('src_w', '<i4'), ('src_h', '<i4'), ('src_pitch_px', '<i4'),
    ('box_x0', '<i4'), ('box_y0', '<i4'), ('box_w', '<i4'), ('box_h', '<i4'), ('out_w', '<i4'), ('out_h', '<i4'),
    ('win_x', '<i4'), ('win_y', '<i4'), ('win_w', '<i4'), ('win_h', '<i4'), ('dst_pitch_px', '<i4'),
])
CROP_SRC = np.dtype([('off', '<i8'), ('pitch_px', '<i4'), ('reserved', '<i4')])
assert RESIZE_JOB.itemsize == 72 and CROP_SRC.itemsize == 16

MAX_SCALE = 11.0  # limit of the resize kernel's tap / row buffers (include/oake_b200.h)


# ----------------------------------------------------------------------------------- CLIP transform
def clip_resize_window(cw: np.ndarray, ch: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """Resize(224) + CenterCrop(224) of a (cw x ch) crop -> (out_w, out_h, win_x, win_y)."""
    cw = np.asarray(cw, dtype=np.int64)
    ch = np.asarray(ch, dtype=np.int64)
    w_short = cw <= ch
    long_ = np.where(w_short, ch, cw)
    short = np.where(w_short, cw, ch)
    new_long = np.trunc((SIZE * long_) / short).astype(np.int64)  # int(size * long / short)
    out_w = np.where(w_short, SIZE, new_long)
    out_h = np.where(w_short, new_long, SIZE)
    win_x = np.rint((out_w - SIZE) / 2.0).astype(np.int64)  # int(round(.)): half to even
    win_y = np.rint((out_h - SIZE) / 2.0).astype(np.int64)
    return out_w, out_h, win_x, win_y


# ------------------------------------------------------------------------------------------ blocks
def partition(length: int, r: int = SIZE, s: int = 112) -> List[int]:
    """Window starts so that r-wide windows with stride <= s tile [0, length) (blocks.py:40-52)."""
    if length < r:
        return []
    if length == r:
        return [0]
    n = (length - r - 1) // s + 1
    q, rem = divmod(length - r, n)
    starts = [0]
    for i in range(n):
        starts.append(starts[-1] + q + (i < rem))
    return starts


@dataclasses.dataclass
class BlocksPlan:
    levels: List[Tuple[int, int]]  # (w, h) of pyramid level 0, 1, ... (level 0 = the image)
    cells: List[Tuple[int, int, int]]  # (level, x, y) of every 224x224 block, reference order
    bboxes: np.ndarray  # (1 + len(cells), 4) float64 rows as the reference stores them


def blocks_plan(w: int, h: int, r: int = SIZE, s: int = 112, rescale: float = 1.5) -> BlocksPlan:
    """blocks.py:54-109: global crop first, then x-outer / y-inner blocks per pyramid level."""
    bboxes = [((w - h) / 2, 0, h, h) if w > h else (0, (h - w) / 2, w, w)]  # sic: not xyxy
    levels, cells = [], []
    scale, lw, lh, level = 1.0, w, h, 0
    while True:
        xs, ys = partition(lw, r, s), partition(lh, r, s)
        if not xs or not ys:
            break
        levels.append((lw, lh))
        for x in xs:
            for y in ys:
                cells.append((level, x, y))
                x1, y1, side = x * scale, y * scale, r * scale
                bboxes.append((x1, y1, x1 + side, y1 + side))
        lw, lh = int(lw / rescale), int(lh / rescale)
        scale *= rescale
        level += 1
    return BlocksPlan(levels, cells, np.asarray(bboxes, dtype=np.float64))


# ----------------------------------------------------------------------------------------- objects
@dataclasses.dataclass
class ObjectsPlan:
    bboxes: np.ndarray  # (No,4) f32 filtered ORIGINAL proposals (what the reference stores)
    objectness: np.ndarray  # (No,1) f32
    expanded: np.ndarray  # (No,4) f32 square crops, float coordinates
    foregrounds: np.ndarray  # (No,4) f32 proposals relative to the crop origin
    boxes_int: np.ndarray  # (No,4) i64 PIL-rounded crop rectangles


def expand_adaptive(xyxy: np.ndarray, image_wh: Tuple[int, int], mode: str = 'ADAPTIVE') -> np.ndarray:
    """objects.py:76-114, fp32 arithmetic throughout.  ExpandMode.ADAPTIVE: square of side sqrt(8 w h)
    (:94-99); ExpandMode.CONSTANT: side 224 (:92-93).  The other two modes cannot run in the reference
     and are refused by the dataset."""
    xyxy = xyxy.astype(np.float32)
    lt, rb = xyxy[:, :2], xyxy[:, 2:]
    wh = rb - lt
    if mode == 'CONSTANT':
        side = np.full((xyxy.shape[0], 1), 224, dtype=np.float32)
    elif mode == 'ADAPTIVE':
        side = np.sqrt(wh[:, 0] * wh[:, 1] * np.float32(8))[:, None]
    else:
        raise ValueError(f'expand mode {mode}')
    center = (lt + rb) / np.float32(2)
    swh = np.concatenate([side, side], axis=1)
    half = swh / np.float32(2)
    e_lt, e_rb = center - half, center + half
    iwh = np.asarray(image_wh, dtype=np.float32)[None, :]
    offset = np.zeros_like(e_lt)
    offset = np.where(e_lt >= 0, offset, -e_lt)
    offset = np.where(e_rb <= iwh, offset, iwh - e_rb)
    offset = np.where(swh <= iwh, offset, np.float32(0))
    center = center + offset
    return np.concatenate([center - half, center + half], axis=1).astype(np.float32)


def objects_plan(proposals: np.ndarray, image_wh: Tuple[int, int], dry_run: bool = False,
                 expand_mode: str = 'ADAPTIVE') -> ObjectsPlan:
    """objects.py:157-186 without the pixels."""
    proposals = np.asarray(proposals, dtype=np.float32).reshape(-1, 5)
    boxes, objectness = proposals[:, :4], proposals[:, 4:]
    wh = boxes[:, 2:] - boxes[:, :2]
    keep = (wh >= np.float32(4)).all(axis=1)
    if dry_run:
        keep[5:] = False
    boxes, objectness = boxes[keep], objectness[keep]
    expanded = expand_adaptive(boxes, image_wh, expand_mode)
    foregrounds = boxes - np.tile(expanded[:, :2], (1, 2))
    boxes_int = np.rint(expanded.astype(np.float64)).astype(np.int64)  # PIL crop: int(round(x))
    return ObjectsPlan(boxes, objectness, expanded, foregrounds.astype(np.float32), boxes_int)


# ------------------------------------------------------------------------------------- job builders
def crop_jobs(src_off: int, src_w: int, src_h: int, boxes_int: np.ndarray, dst_off0: int) -> np.ndarray:
    """CLIP-transform jobs for integer rectangles of one image; outputs packed 224x224x3 at dst_off0."""
    n = boxes_int.shape[0]
    jobs = np.zeros(n, dtype=RESIZE_JOB)
    if n == 0:
        return jobs
    cw = boxes_int[:, 2] - boxes_int[:, 0]
    ch = boxes_int[:, 3] - boxes_int[:, 1]
    if (cw <= 0).any() or (ch <= 0).any():
        raise ValueError('degenerate crop rectangle')
    out_w, out_h, win_x, win_y = clip_resize_window(cw, ch)
    if (cw / out_w).max() > MAX_SCALE or (ch / out_h).max() > MAX_SCALE:
        raise ValueError(f'crop is more than {MAX_SCALE}x larger than its 224 px target: not supported by the '
                         'GPU resize kernel')
    jobs['src_off'] = src_off
    jobs['dst_off'] = dst_off0 + np.arange(n, dtype=np.int64) * (SIZE * SIZE * 3)
    jobs['src_w'], jobs['src_h'], jobs['src_pitch_px'] = src_w, src_h, src_w
    jobs['box_x0'], jobs['box_y0'], jobs['box_w'], jobs['box_h'] = boxes_int[:, 0], boxes_int[:, 1], cw, ch
    jobs['out_w'], jobs['out_h'] = out_w, out_h
    jobs['win_x'], jobs['win_y'], jobs['win_w'], jobs['win_h'] = win_x, win_y, SIZE, SIZE
    jobs['dst_pitch_px'] = SIZE
    return jobs


def level_job(src_off: int, src_w: int, src_h: int, dst_off: int, dst_w: int, dst_h: int) -> np.ndarray:
    """Whole-image `image.resize((dst_w, dst_h))` (blocks pyramid step, PIL default BICUBIC)."""
    job = np.zeros(1, dtype=RESIZE_JOB)
    job['src_off'], job['dst_off'] = src_off, dst_off
    job['src_w'], job['src_h'], job['src_pitch_px'] = src_w, src_h, src_w
    job['box_x0'], job['box_y0'], job['box_w'], job['box_h'] = 0, 0, src_w, src_h
    job['out_w'], job['out_h'] = dst_w, dst_h
    job['win_x'], job['win_y'], job['win_w'], job['win_h'] = 0, 0, dst_w, dst_h
    job['dst_pitch_px'] = dst_w
    return job


def fused_stage(stages, crops: np.ndarray) -> bool:
    """True when the crops are exactly the outputs of ONE resize stage, each a whole 224 x 224 window (the globals and
    objects tasks): such crops never exist as uint8 -- `oake_resize_to_patches` writes the tower's front-end matrix.
    The blocks task (pyramid stages, crops = windows into the levels) is not."""
    if len(stages) != 1 or crops.shape[0] == 0 or stages[0].size != crops.shape[0]:
        return False
    jobs = stages[0]
    return bool((jobs['win_w'] == SIZE).all() and (jobs['win_h'] == SIZE).all()
                and np.array_equal(jobs['dst_off'], crops['off']) and (crops['pitch_px'] == SIZE).all())


def max_tiles(jobs: np.ndarray) -> int:
    if jobs.size == 0:
        return 0
    return int((((jobs['win_w'] + 31) // 32) * ((jobs['win_h'] + 31) // 32)).max())
