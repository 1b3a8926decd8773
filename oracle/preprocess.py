"""CPU oracle: YOLO box -> crop -> cv2 INTER_LINEAR resize -> Normalize -> CHW.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates, per sample, what the
reference does on DataLoader workers:

* ``AnnotatedYOLODataset.bbox_xywhn2xyxy``      nkb_classification/dataset.py:414-421
* ``check_boxes_sizes_annotation``              nkb_classification/dataset.py:433-434
* ``AnnotatedYOLODataset.__getitem__`` slice    nkb_classification/dataset.py:398-409
* ``Transforms.__call__``                       nkb_classification/dataset.py:96-102
* pipeline ops named by the configs             configs/singletask_config.py:203-219,
  (A.Resize | A.LongestMaxSize + A.PadIfNeeded, configs/multitask_config.py:130-140,
  A.Normalize, ToTensorV2)                      metrics/det_cls_val.py:86-109
* ``Evaluator.classify_crops`` box conversion   metrics/det_cls_val.py:228-239

* train-time ops of configs/singletask_config.py:172-194 (A.HorizontalFlip, A.VerticalFlip,
  A.RandomBrightnessContrast, A.CoarseDropout) with GIVEN per-sample parameters: ``augment_u8``

Two implementations of the 8-bit bilinear resize live here so the oracle checks
itself: ``resize_cv2`` calls OpenCV (what the reference executes, through
albumentations), ``resize_int`` is an independent integer restatement of
OpenCV's fixed-point path (11-bit coefficients, >>4 / >>16 / +2>>2 rounding).
tests/test_oracle_preprocess.py requires them to agree bit-for-bit.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np

INTER_RESIZE_COEF_BITS = 11
INTER_RESIZE_COEF_SCALE = 1 << INTER_RESIZE_COEF_BITS  # 2048


# --------------------------------------------------------------------------
# a1: YOLO normalised xywh -> integer xyxy  (dataset.py:414-421, :433-434)
# --------------------------------------------------------------------------
def bbox_xywhn2xyxy(x_center, y_center, width, height, image_size):
    """float64 arithmetic, Python ``int()`` truncation toward zero, clip to the
    frame -- exactly the reference's sequence (dataset.py:414-421)."""
    image_height, image_width = image_size
    x_min = int(np.clip(int((x_center - width / 2) * image_width), 0, image_width))
    y_min = int(np.clip(int((y_center - height / 2) * image_height), 0, image_height))
    x_max = int(np.clip(int((x_center + width / 2) * image_width), 0, image_width))
    y_max = int(np.clip(int((y_center + height / 2) * image_height), 0, image_height))
    return x_min, y_min, x_max, y_max


def box_is_kept(x_min, y_min, x_max, y_max, min_box_size=5):
    """dataset.py:433-434."""
    return (x_max - x_min >= min_box_size) and (y_max - y_min >= min_box_size)


def parse_yolo_label_lines(lines, image_size, min_box_size=5):
    """dataset.py:350-359: ``cls xc yc w h`` text rows -> [(box, label)]."""
    out = []
    for line in lines:
        parts = line.split()
        if not parts:
            continue
        label = int(parts[0])
        xc, yc, w, h = tuple(map(float, parts[1:5]))
        box = bbox_xywhn2xyxy(xc, yc, w, h, image_size)
        if not box_is_kept(*box, min_box_size=min_box_size):
            continue
        out.append((box, label))
    return out


def detector_boxes_to_int(boxes_xyxyn: np.ndarray, img_h: int, img_w: int) -> np.ndarray:
    """metrics/det_cls_val.py:231-236: normalised xyxy * (W,H) then astype(int)."""
    b = np.array(boxes_xyxyn, dtype=np.float64, copy=True)
    b[:, [0, 2]] *= img_w
    b[:, [1, 3]] *= img_h
    return b.astype(int)


# --------------------------------------------------------------------------
# cv2 8-bit INTER_LINEAR, integer restatement
# --------------------------------------------------------------------------
def axis_tables(dsize: int, ssize: int, horizontal: bool):
    """Per-destination-index source index and the two 11-bit coefficients.

    Follows OpenCV's ``resize`` set-up loop (imgproc/src/resize.cpp, the
    ``INTER_LINEAR`` branch): scale = 1/(dsize/ssize) in double, the coordinate
    is evaluated in double then rounded to float32, the fraction is a float32
    subtraction, the horizontal clamp zeroes the fraction, the vertical one
    does not, coefficients = saturate_cast<short>(c * 2048) (round half even).
    """
    scale = 1.0 / (float(dsize) / float(ssize))
    d = np.arange(dsize, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if horizontal:
        lo = s < 0
        f[lo] = 0.0
        s[lo] = 0
        hi = s >= ssize - 1
        f[hi] = 0.0
        s[hi] = ssize - 1
    one = np.float32(1.0)
    k = np.float32(INTER_RESIZE_COEF_SCALE)
    c0 = np.rint((one - f) * k).astype(np.int32)
    c1 = np.rint(f * k).astype(np.int32)
    return s, c0, c1


def resize_int(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """Bit-exact restatement of ``cv2.resize(img, (dw, dh), INTER_LINEAR)`` for
    uint8 HxWxC input.  Vectorised numpy, int64 intermediates."""
    assert img.dtype == np.uint8 and img.ndim == 3
    sh, sw, _ = img.shape
    if sh == dh and sw == dw:
        return img.copy()
    sx, ax0, ax1 = axis_tables(dw, sw, horizontal=True)
    sy, by0, by1 = axis_tables(dh, sh, horizontal=False)
    x0 = sx
    x1 = np.minimum(sx + 1, sw - 1)
    r0 = np.clip(sy, 0, sh - 1)
    r1 = np.clip(sy + 1, 0, sh - 1)
    src = img.astype(np.int64)
    # horizontal pass on every source row that is referenced
    H = src[:, x0, :] * ax0[None, :, None] + src[:, x1, :] * ax1[None, :, None]
    H4 = H >> 4
    t0 = (by0[:, None, None] * H4[r0]) >> 16
    t1 = (by1[:, None, None] * H4[r1]) >> 16
    out = (t0 + t1 + 2) >> 2
    assert out.min() >= 0 and out.max() <= 255
    return out.astype(np.uint8)


def resize_cv2(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """What the reference executes: albumentations' ``F.resize`` ->
    ``cv2.resize(img, dsize=(w, h), interpolation=cv2.INTER_LINEAR)``; it
    short-circuits when the size already matches."""
    import cv2

    sh, sw = img.shape[:2]
    if sh == dh and sw == dw:
        return img
    return cv2.resize(img, dsize=(dw, dh), interpolation=cv2.INTER_LINEAR)


# --------------------------------------------------------------------------
# pipeline description (the deterministic val / inference subset)
# --------------------------------------------------------------------------
MODE_STRETCH = 0    # A.Resize(S_h, S_w)
MODE_LETTERBOX = 1  # A.LongestMaxSize(S) + A.PadIfNeeded(S_h, S_w, BORDER_CONSTANT, value)


@dataclass
class Plan:
    mode: int = MODE_STRETCH
    out_h: int = 224
    out_w: int = 224
    max_size: int = 224                    # LongestMaxSize (letterbox only)
    pad_value: Tuple[int, int, int] = (0, 0, 0)
    mean: Tuple[float, float, float] = (0.485, 0.456, 0.406)
    std: Tuple[float, float, float] = (0.229, 0.224, 0.225)
    max_pixel_value: float = 255.0
    channel_swap: bool = False             # True: source is BGR (cv2.imread), emit RGB


def py3round(number: float) -> int:
    """albumentations 1.x ``py3round`` == Python-3 round-half-even."""
    if abs(round(number) - number) == 0.5:
        return int(2.0 * round(number / 2.0))
    return int(round(number))


def letterbox_geometry(h: int, w: int, max_size: int, out_h: int, out_w: int):
    """(new_h, new_w, top, left) for LongestMaxSize(max_size) followed by a
    centred PadIfNeeded(out_h, out_w)."""
    scale = max_size / float(max(w, h))
    if scale != 1.0:
        new_h, new_w = py3round(h * scale), py3round(w * scale)
    else:
        new_h, new_w = h, w
    if new_h < out_h:
        top = int((out_h - new_h) / 2.0)
    else:
        top = 0
    if new_w < out_w:
        left = int((out_w - new_w) / 2.0)
    else:
        left = 0
    return new_h, new_w, top, left


def normalize_constants(mean: Sequence[float], std: Sequence[float], max_pixel_value: float = 255.0):
    """albumentations 1.x ``F.normalize``: all in float32.
    m = f32(mean)*f32(maxv);  d = 1/(f32(std)*f32(maxv))."""
    m = np.array(mean, dtype=np.float32)
    m *= np.float32(max_pixel_value)
    s = np.array(std, dtype=np.float32)
    s *= np.float32(max_pixel_value)
    d = np.reciprocal(s, dtype=np.float32)
    return m, d


def normalize_f32(img_u8: np.ndarray, mean, std, max_pixel_value=255.0) -> np.ndarray:
    """fp32 ``(x - m) * d`` as two separately rounded ops."""
    m, d = normalize_constants(mean, std, max_pixel_value)
    x = img_u8.astype(np.float32)
    x -= m
    x *= d
    return x


def resize_stage_u8(crop: np.ndarray, plan: Plan, impl: str = "int") -> np.ndarray:
    """uint8 HWC crop -> uint8 HWC (out_h, out_w, 3) after Resize or
    LongestMaxSize + PadIfNeeded(BORDER_CONSTANT)."""
    rs = resize_int if impl == "int" else resize_cv2
    if plan.channel_swap:
        crop = crop[:, :, ::-1]
    crop = np.ascontiguousarray(crop)  # dataset.py:102  np.array(img)
    h, w = crop.shape[:2]
    if plan.mode == MODE_STRETCH:
        return np.ascontiguousarray(rs(crop, plan.out_w, plan.out_h))
    new_h, new_w, top, left = letterbox_geometry(h, w, plan.max_size, plan.out_h, plan.out_w)
    img = rs(crop, new_w, new_h)
    oh, ow = max(plan.out_h, new_h), max(plan.out_w, new_w)
    canvas = np.empty((oh, ow, 3), dtype=np.uint8)
    canvas[:, :] = np.array(plan.pad_value, dtype=np.uint8)
    canvas[top : top + new_h, left : left + new_w] = img
    return canvas


@dataclass
class AugSample:
    """Parameters one Compose call would have drawn for one sample (albumentations 1.3.x semantics)."""
    hflip: bool = False
    vflip: bool = False
    bc: bool = False                 # RandomBrightnessContrast applied
    alpha: float = 1.0               # 1 + contrast draw
    beta: float = 0.0                # brightness draw (fraction of max_value: brightness_by_max=True)
    holes: Sequence[Tuple[int, int, int, int]] = ()   # CoarseDropout (x1, y1, x2, y2), ends exclusive
    fill: Tuple[int, int, int] = (0, 0, 0)
    hsv: Optional[Tuple[float, float, float]] = None  # HueSaturationValue (hue, sat, val) shifts; None = not applied


def brightness_contrast_lut(alpha: float, beta: float) -> np.ndarray:
    """albumentations 1.3 ``_brightness_contrast_adjust_uint`` with beta_by_max=True: a float32 ramp times
    alpha (python float -> float32 multiply), plus beta * 255 (python double product, added in float32), clipped to
    [0, 255] and truncated to uint8.  The ``!= 1`` / ``!= 0`` short-cuts of the original are exact no-ops."""
    lut = np.arange(0, 256).astype("float32")
    if alpha != 1:
        lut *= alpha
    if beta != 0:
        lut += beta * 255
    return np.clip(lut, 0, 255).astype(np.uint8)


def shift_hsv_u8(img: np.ndarray, hue_shift: float, sat_shift: float, val_shift: float) -> np.ndarray:
    """albumentations 1.3 ``shift_hsv`` / ``_shift_hsv_uint8`` for a 3-channel uint8 RGB image, verbatim:
    untouched for all-zero shifts, else cv2 RGB2HSV, one cv2.LUT per non-zero shift (int16 ramp + float shift,
    ``mod 180`` / ``clip``, truncated), cv2 HSV2RGB."""
    import cv2

    if hue_shift == 0 and sat_shift == 0 and val_shift == 0:
        return img
    dtype = img.dtype
    img = cv2.cvtColor(np.ascontiguousarray(img), cv2.COLOR_RGB2HSV)
    hue, sat, val = cv2.split(img)
    if hue_shift != 0:
        lut_hue = np.arange(0, 256, dtype=np.int16)
        lut_hue = np.mod(lut_hue + hue_shift, 180).astype(dtype)
        hue = cv2.LUT(hue, lut_hue)
    if sat_shift != 0:
        lut_sat = np.arange(0, 256, dtype=np.int16)
        lut_sat = np.clip(lut_sat + sat_shift, 0, 255).astype(dtype)
        sat = cv2.LUT(sat, lut_sat)
    if val_shift != 0:
        lut_val = np.arange(0, 256, dtype=np.int16)
        lut_val = np.clip(lut_val + val_shift, 0, 255).astype(dtype)
        val = cv2.LUT(val, lut_val)
    img = cv2.merge((hue, sat, val)).astype(dtype)
    return cv2.cvtColor(img, cv2.COLOR_HSV2RGB)


def rgb2hsv_int(rgb: np.ndarray) -> np.ndarray:
    """Independent integer restatement of OpenCV's 8-bit RGB2HSV (hsv_shift = 12, hrange 180), uint8 [..., 3]."""
    x = np.asarray(rgb).astype(np.int64)
    r, g, b = x[..., 0], x[..., 1], x[..., 2]
    i = np.arange(1, 256, dtype=np.float64)
    sdiv = np.concatenate([[0], np.rint((255 << 12) / (1.0 * i))]).astype(np.int64)
    hdiv = np.concatenate([[0], np.rint((180 << 12) / (6.0 * i))]).astype(np.int64)
    v = np.maximum(np.maximum(r, g), b)
    diff = v - np.minimum(np.minimum(r, g), b)
    s = (diff * sdiv[v] + (1 << 11)) >> 12
    h = np.where(v == r, g - b, np.where(v == g, b - r + 2 * diff, r - g + 4 * diff))
    h = (h * hdiv[diff] + (1 << 11)) >> 12
    h = np.where(h < 0, h + 180, h)
    return np.stack([h, s, v], -1).astype(np.uint8)


def hsv2rgb_f32(hsv: np.ndarray, trunc: bool = False) -> np.ndarray:
    """Restatement of OpenCV's 8-bit HSV2RGB (float path, hue < 180): s, v scaled by 1/255, h by 6/180, sector =
    trunc, tab1 = v(1 - s), tab2 = v * fma(-s, f, 1), tab3 = v * fma(-s, 1 - f, 1), then x * 255 -> uint8 by
    round-to-nearest-even (OpenCV's scalar code: the tail of each row) or, ``trunc=True``, by truncation (its
    vectorised body: the first (W / lanes) * lanes pixels of each row).  The two fma's are how OpenCV's build
    contracts ``1 - s*f``; formulas and both roundings were found by exhaustive comparison with cv2 4.13 over all
    180 * 2^16 inputs (tests/test_transforms.py repeats it).  fma is emulated in float64 (exact product, one extra
    rounding that the exhaustive check shows never matters on this domain)."""
    f32 = np.float32
    x = np.asarray(hsv)
    h, s, v = (x[..., k].astype(np.float32) for k in range(3))
    s = s * f32(1.0 / 255.0)
    v = v * f32(1.0 / 255.0)
    h = h * (f32(6.0) / f32(180.0))
    pre = np.trunc(h)
    fr = (h - pre).astype(np.float32)
    sector = pre.astype(np.int64) % 6

    def fma(a, b_, c):
        return (a.astype(np.float64) * b_.astype(np.float64) + c.astype(np.float64)).astype(np.float32)

    one = np.ones_like(s)
    t0 = v
    t1 = (v * (one - s)).astype(np.float32)
    t2 = (v * fma(-s, fr, one)).astype(np.float32)
    t3 = (v * fma(-s, (one - fr).astype(np.float32), one)).astype(np.float32)
    tab = np.stack([t0, t1, t2, t3], 0)
    sd = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]])
    pick = lambda k: np.take_along_axis(tab, sd[sector, k][None], 0)[0]
    bb, gg, rr = pick(0), pick(1), pick(2)
    rnd = np.trunc if trunc else np.rint
    sat = lambda t: np.clip(rnd((t * f32(255.0)).astype(np.float32)), 0, 255).astype(np.uint8)
    return np.stack([sat(rr), sat(gg), sat(bb)], -1)


def augment_u8(img: np.ndarray, a: "AugSample") -> np.ndarray:
    """The deterministic part of the train pipeline on the resized / padded uint8 HWC image, in Compose order:
    HorizontalFlip (cv2.flip(img, 1)), VerticalFlip (cv2.flip(img, 0)), RandomBrightnessContrast (cv2.LUT),
    HueSaturationValue (shift_hsv_u8), CoarseDropout (``img[y1:y2, x1:x2] = fill_value`` per hole)."""
    import cv2

    out = np.ascontiguousarray(img)
    if a.hflip:
        out = cv2.flip(out, 1)
    if a.vflip:
        out = cv2.flip(out, 0)
    if a.bc:
        out = cv2.LUT(out, brightness_contrast_lut(a.alpha, a.beta))
    if a.hsv is not None:
        out = shift_hsv_u8(out, *a.hsv)
    if len(a.holes):
        out = out.copy()
        for x1, y1, x2, y2 in a.holes:
            out[y1:y2, x1:x2] = a.fill
    return out


def preprocess_crop(frame: np.ndarray, box, plan: Plan, impl: str = "int", aug: Optional["AugSample"] = None):
    """One sample: returns (u8 HWC resized/padded[/augmented], f32 CHW normalised)."""
    x0, y0, x1, y1 = [int(v) for v in box]
    crop = frame[y0:y1, x0:x1]  # dataset.py:404
    u8 = resize_stage_u8(crop, plan, impl)
    if aug is not None:
        u8 = augment_u8(u8, aug)
    f = normalize_f32(u8, plan.mean, plan.std, plan.max_pixel_value)
    chw = np.ascontiguousarray(f.transpose(2, 0, 1))  # ToTensorV2
    return u8, chw


def preprocess_batch(frames, boxes, frame_idx, plan: Plan, impl: str = "int", augs=None):
    """default_collate of B samples: ([B,H,W,3] u8, [B,3,H,W] f32)."""
    u8s, chws = [], []
    for i, (b, fi) in enumerate(zip(boxes, frame_idx)):
        u8, chw = preprocess_crop(frames[int(fi)], b, plan, impl, None if augs is None else augs[i])
        u8s.append(u8)
        chws.append(chw)
    return np.stack(u8s), np.stack(chws)


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16, returned as uint16 bit patterns."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    rounding = ((u >> 16) & 1) + 0x7FFF
    return ((u + rounding) >> 16).astype(np.uint16)
