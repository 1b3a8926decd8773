"""Pipeline descriptions and their compilation to one K1 launch.

The reference's configs hold live ``albumentations.Compose`` objects
(configs/singletask_config.py:162-219) that ``Transforms.__call__`` runs per
sample on DataLoader workers (nkb_classification/dataset.py:89-102).  Here a
pipeline is *compiled*: the deterministic val / inference subset

    Resize(h, w) | LongestMaxSize(s) + PadIfNeeded(h, w, BORDER_CONSTANT, value)
    Normalize(mean, std, max_pixel_value)
    ToTensorV2()

becomes a :class:`PreprocessPlan` that one fused kernel executes for the whole
batch.  Real albumentations objects are accepted when that package is present
(duck-typed on class name + public attributes); the small classes below carry
the same names / keyword arguments so a config can be written without it.
Anything else (random train-time augmentations, other border modes or
interpolations) raises ``NotImplementedError`` -- there is no silent fallback.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, List, Optional, Sequence, Tuple

import numpy as np

BORDER_CONSTANT = 0  # cv2.BORDER_CONSTANT
INTER_LINEAR = 1     # cv2.INTER_LINEAR

MODE_STRETCH, MODE_LETTERBOX = 0, 1


class _Op:
    def __repr__(self):
        kv = ", ".join(f"{k}={v!r}" for k, v in self.__dict__.items())
        return f"{type(self).__name__}({kv})"


class Resize(_Op):
    def __init__(self, height: int, width: int, interpolation: int = INTER_LINEAR, always_apply=False, p=1):
        self.height, self.width, self.interpolation = int(height), int(width), interpolation


class LongestMaxSize(_Op):
    def __init__(self, max_size: int = 1024, interpolation: int = INTER_LINEAR, always_apply=False, p=1):
        self.max_size, self.interpolation = max_size, interpolation


class PadIfNeeded(_Op):
    def __init__(self, min_height: int = 1024, min_width: int = 1024, border_mode: int = 4, value=None,
                 always_apply=False, p=1.0, position="center"):
        self.min_height, self.min_width = int(min_height), int(min_width)
        self.border_mode, self.value, self.position = border_mode, value, position


class Normalize(_Op):
    def __init__(self, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), max_pixel_value=255.0,
                 always_apply=False, p=1.0):
        self.mean, self.std, self.max_pixel_value = tuple(mean), tuple(std), float(max_pixel_value)


class ToTensorV2(_Op):
    def __init__(self, transpose_mask=False, always_apply=True, p=1.0):
        pass


class Compose(_Op):
    def __init__(self, transforms: Sequence[Any], **kwargs):
        self.transforms = list(transforms)


@dataclass(frozen=True)
class PreprocessPlan:
    """Everything K1 needs besides the pixels and the boxes."""
    mode: int
    out_h: int
    out_w: int
    max_size: int
    pad_value: Tuple[int, int, int]
    mean255: Tuple[float, float, float]   # f32(mean) * f32(max_pixel_value), as float32 values
    denom: Tuple[float, float, float]     # 1 / (f32(std) * f32(max_pixel_value))
    channel_swap: bool = False

    def with_channel_swap(self, swap: bool) -> "PreprocessPlan":
        return PreprocessPlan(self.mode, self.out_h, self.out_w, self.max_size, self.pad_value, self.mean255,
                              self.denom, bool(swap))


def normalize_constants(mean, std, max_pixel_value=255.0):
    """albumentations 1.x ``F.normalize`` constants, float32 arithmetic throughout."""
    m = np.array(mean, dtype=np.float32)
    m *= np.float32(max_pixel_value)
    s = np.array(std, dtype=np.float32)
    s *= np.float32(max_pixel_value)
    d = np.reciprocal(s, dtype=np.float32)
    if m.size == 1:
        m, d = np.repeat(m, 3), np.repeat(d, 3)
    return m, d


def _kind(op) -> str:
    return type(op).__name__


def _as_int_triplet(value) -> Tuple[int, int, int]:
    if value is None:
        return (0, 0, 0)
    if np.isscalar(value):
        v = int(value)
        return (v, v, v)
    v = [int(x) for x in value]
    if len(v) != 3:
        raise NotImplementedError(f"PadIfNeeded value {value!r}: need a scalar or 3 channels")
    return (v[0], v[1], v[2])


def compile_pipeline(pipeline, channel_swap: bool = False) -> PreprocessPlan:
    """``A.Compose([...])`` (or a plain list of ops) -> PreprocessPlan."""
    ops = list(getattr(pipeline, "transforms", pipeline))
    resize = lms = pad = norm = None
    saw_tensor = False
    for op in ops:
        k = _kind(op)
        if saw_tensor:
            raise NotImplementedError(f"{k} after ToTensorV2 is not supported")
        if k == "Resize" and resize is None and lms is None:
            resize = op
        elif k == "LongestMaxSize" and lms is None and resize is None:
            lms = op
        elif k == "PadIfNeeded" and pad is None and lms is not None:
            pad = op
        elif k == "Normalize" and norm is None:
            norm = op
        elif k == "ToTensorV2":
            saw_tensor = True
        else:
            raise NotImplementedError(
                f"pipeline op {op!r} is outside the fused deterministic subset "
                "{Resize | LongestMaxSize+PadIfNeeded, Normalize, ToTensorV2}"
            )
        if norm is not None and k in ("Resize", "LongestMaxSize", "PadIfNeeded"):
            raise NotImplementedError("geometry ops must precede Normalize")
    if norm is None:
        raise NotImplementedError("pipeline must contain Normalize (fp32 output contract)")
    if not saw_tensor:
        raise NotImplementedError("pipeline must end with ToTensorV2 (CHW tensor contract)")
    for op in (resize, lms):
        if op is not None and getattr(op, "interpolation", INTER_LINEAR) != INTER_LINEAR:
            raise NotImplementedError("only cv2.INTER_LINEAR is implemented")
    m, d = normalize_constants(norm.mean, norm.std, getattr(norm, "max_pixel_value", 255.0))
    mean255 = tuple(float(x) for x in m)
    denom = tuple(float(x) for x in d)
    if resize is not None:
        return PreprocessPlan(MODE_STRETCH, int(resize.height), int(resize.width), 0, (0, 0, 0), mean255, denom,
                              bool(channel_swap))
    if lms is None or pad is None:
        raise NotImplementedError("need Resize, or LongestMaxSize followed by PadIfNeeded (fixed output size)")
    if getattr(pad, "border_mode", None) != BORDER_CONSTANT:
        raise NotImplementedError("PadIfNeeded: only border_mode=cv2.BORDER_CONSTANT is implemented")
    pos = getattr(pad, "position", "center")
    if getattr(pos, "value", pos) != "center":
        raise NotImplementedError("PadIfNeeded: only position='center' is implemented")
    max_size = lms.max_size
    if not isinstance(max_size, (int, np.integer)):
        raise NotImplementedError("LongestMaxSize: a single integer max_size is required")
    out_h, out_w = int(pad.min_height), int(pad.min_width)
    if max_size > min(out_h, out_w):
        raise NotImplementedError("LongestMaxSize larger than PadIfNeeded would give variable output sizes")
    return PreprocessPlan(MODE_LETTERBOX, out_h, out_w, int(max_size), _as_int_triplet(getattr(pad, "value", None)),
                          mean255, denom, bool(channel_swap))
