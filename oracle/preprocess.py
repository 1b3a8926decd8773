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


def augment_u8(img: np.ndarray, a: "AugSample") -> np.ndarray:
    """The deterministic part of the train pipeline on the resized / padded uint8 HWC image, in Compose order:
    HorizontalFlip (cv2.flip(img, 1)), VerticalFlip (cv2.flip(img, 0)), RandomBrightnessContrast (cv2.LUT),
    CoarseDropout (``img[y1:y2, x1:x2] = fill_value`` per hole)."""
    import cv2

    out = np.ascontiguousarray(img)
    if a.hflip:
        out = cv2.flip(out, 1)
    if a.vflip:
        out = cv2.flip(out, 0)
    if a.bc:
        out = cv2.LUT(out, brightness_contrast_lut(a.alpha, a.beta))
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
