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

Train-time pipelines (configs/singletask_config.py:162-201) may add, between the
geometry and Normalize, the ops whose arithmetic K1 fuses: ``HorizontalFlip``,
``VerticalFlip``, ``RandomBrightnessContrast`` (brightness_by_max),
``HueSaturationValue`` and ``CoarseDropout`` (last) -- i.e. the whole train
pipeline of configs/singletask_config.py and configs/multitask_config.py.  Their per-sample parameters are drawn on the host by
:func:`draw_augmentations` (same distributions and draw order as albumentations
1.3.x with Python's ``random``) and shipped to the kernel as small arrays.
Anything else (MotionBlur, fog / rain / shadow, other border
modes or interpolations) raises ``NotImplementedError`` -- no silent fallback.
"""
from __future__ import annotations

import dataclasses
import random as _random
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


class HorizontalFlip(_Op):
    def __init__(self, always_apply=False, p=0.5):
        self.p, self.always_apply = float(p), bool(always_apply)


class VerticalFlip(_Op):
    def __init__(self, always_apply=False, p=0.5):
        self.p, self.always_apply = float(p), bool(always_apply)


def _to_tuple(v, low=None):
    """albumentations 1.3 ``to_tuple``: a scalar x becomes (-x, x); a pair is kept AS GIVEN (not sorted)."""
    if isinstance(v, (tuple, list)):
        if len(v) != 2:
            raise ValueError(f"limit {v!r} must be a scalar or a pair")
        return (float(v[0]), float(v[1]))
    return (-float(v), float(v)) if low is None else (float(low), float(v))


class RandomBrightnessContrast(_Op):
    def __init__(self, brightness_limit=0.2, contrast_limit=0.2, brightness_by_max=True, always_apply=False, p=0.5):
        self.brightness_limit, self.contrast_limit = _to_tuple(brightness_limit), _to_tuple(contrast_limit)
        self.brightness_by_max, self.p, self.always_apply = bool(brightness_by_max), float(p), bool(always_apply)


class HueSaturationValue(_Op):
    def __init__(self, hue_shift_limit=20, sat_shift_limit=30, val_shift_limit=20, always_apply=False, p=0.5):
        self.hue_shift_limit = _to_tuple(hue_shift_limit)
        self.sat_shift_limit = _to_tuple(sat_shift_limit)
        self.val_shift_limit = _to_tuple(val_shift_limit)
        self.p, self.always_apply = float(p), bool(always_apply)


_CV2_HSV_LANES: Optional[int] = None


def cv2_hsv_simd_lanes() -> int:
    """How many pixels per step cv2's vectorised 8-bit HSV2RGB body handles on THIS host (16 / 32 / 64; 0 if it has no
    vector body).  OpenCV truncates ``x * 255`` there but rounds to nearest in the scalar tail of each row, so
    ``A.HueSaturationValue`` output depends on it; K1 reproduces the host it runs next to.  Probed once from cv2
    itself (a 1 x 256 strip against the same pixels fed one per row); 32 (AVX2) when cv2 is not importable."""
    global _CV2_HSV_LANES
    if _CV2_HSV_LANES is None:
        try:
            import cv2
            rng = np.random.default_rng(0)
            strip = np.stack([rng.integers(0, 180, 255), rng.integers(1, 256, 255), rng.integers(1, 256, 255)],
                             -1).astype(np.uint8).reshape(1, 255, 3)
            wide = cv2.cvtColor(strip, cv2.COLOR_HSV2RGB)[0]
            tall = cv2.cvtColor(strip.reshape(255, 1, 3), cv2.COLOR_HSV2RGB)[:, 0]
            diff = np.nonzero((wide != tall).any(-1))[0]
            lanes = 0
            if diff.size:   # the body covers columns [0, (255 // lanes) * lanes): the largest power of two explaining it
                for cand in (64, 32, 16, 8):
                    if diff.max() < (255 // cand) * cand and diff.max() >= (255 // cand) * cand - cand:
                        lanes = cand
                        break
            _CV2_HSV_LANES = lanes
        except Exception:
            _CV2_HSV_LANES = 32
    return _CV2_HSV_LANES


def hsv_luts(hue_shift: float, sat_shift: float, val_shift: float) -> np.ndarray:
    """uint8 [3, 256]: the hue / sat / val tables of albumentations 1.3 ``_shift_hsv_uint8`` (an int16 ramp plus the
    float shift, ``mod 180`` for hue / ``clip(0, 255)`` for the others, truncated to uint8); a zero shift leaves the
    identity, as the original skips that LUT."""
    ramp = np.arange(0, 256, dtype=np.int16)
    out = np.tile(np.arange(256, dtype=np.uint8), (3, 1))
    if hue_shift != 0:
        out[0] = np.mod(ramp + hue_shift, 180).astype(np.uint8)
    if sat_shift != 0:
        out[1] = np.clip(ramp + sat_shift, 0, 255).astype(np.uint8)
    if val_shift != 0:
        out[2] = np.clip(ramp + val_shift, 0, 255).astype(np.uint8)
    return out


class CoarseDropout(_Op):
    def __init__(self, max_holes=8, max_height=8, max_width=8, min_holes=None, min_height=None, min_width=None,
                 fill_value=0, mask_fill_value=None, always_apply=False, p=0.5):
        self.max_holes, self.max_height, self.max_width = max_holes, max_height, max_width
        self.min_holes = min_holes if min_holes is not None else max_holes
        self.min_height = min_height if min_height is not None else max_height
        self.min_width = min_width if min_width is not None else max_width
        self.fill_value, self.mask_fill_value = fill_value, mask_fill_value
        self.p, self.always_apply = float(p), bool(always_apply)
        if not 0 < self.min_holes <= self.max_holes:
            raise ValueError(f"Invalid combination of min_holes and max_holes. Got: {[min_holes, max_holes]}")


class ToTensorV2(_Op):
    def __init__(self, transpose_mask=False, always_apply=True, p=1.0):
        pass


class Compose(_Op):
    def __init__(self, transforms: Sequence[Any], **kwargs):
        self.transforms = list(transforms)


MAX_HOLES = 16   # K1_AUG_MAX_HOLES
AUG_HFLIP, AUG_VFLIP, AUG_BC, AUG_HSV = 1, 2, 4, 8


@dataclass(frozen=True)
class AugmentSpec:
    """The random train-time ops K1 fuses, in the order the pipeline lists them (= the order their parameters are
    drawn in).  ``order`` holds op names out of {"HorizontalFlip", "VerticalFlip", "RandomBrightnessContrast",
    "HueSaturationValue", "CoarseDropout"}."""
    order: Tuple[str, ...] = ()
    hflip_p: float = 0.0
    vflip_p: float = 0.0
    bc_p: float = 0.0
    brightness_limit: Tuple[float, float] = (0.0, 0.0)
    contrast_limit: Tuple[float, float] = (0.0, 0.0)
    hsv_p: float = 0.0
    hue_limit: Tuple[float, float] = (0.0, 0.0)
    sat_limit: Tuple[float, float] = (0.0, 0.0)
    val_limit: Tuple[float, float] = (0.0, 0.0)
    cd_p: float = 0.0
    holes: Tuple[int, int] = (1, 1)                    # min_holes, max_holes
    hole_h: Tuple[Any, Any] = (8, 8)                   # min_height, max_height (both int, or both float fractions)
    hole_w: Tuple[Any, Any] = (8, 8)
    fill: Tuple[int, int, int] = (0, 0, 0)             # as stored into a uint8 image


@dataclass
class AugmentBatch:
    """Per-sample parameters of one batch, host side (numpy); ``ops.preprocess_crops`` uploads them."""
    flags: np.ndarray       # int32 [n]
    alpha: np.ndarray       # float32 [n]
    beta: np.ndarray        # float32 [n] = f32(brightness * 255)
    holes: np.ndarray       # int32 [n, max_holes, 4]
    fill: Tuple[int, int, int]
    brightness: np.ndarray  # float64 [n], the raw draw (the oracle re-derives beta * 255 from it)
    hsv_shift: Optional[np.ndarray] = None   # float64 [n, 3] raw hue / sat / val draws (None: op not in the pipeline)
    hsv_lut: Optional[np.ndarray] = None     # uint8 [n, 3, 256] the tables built from them
    hsv_trunc_cols: int = 0                  # output columns that take cv2's vectorised (truncating) HSV2RGB rounding

    @property
    def max_holes(self) -> int:
        return int(self.holes.shape[1])


def _fill_hsv_luts(flags: np.ndarray, hsv_shift: np.ndarray, hsv_lut: np.ndarray) -> None:
    """hsv_luts() for every flagged sample at once (same arithmetic: int16 ramp + float64 shift); a table is only
    computed for the samples whose shift on that channel is non-zero (albumentations keeps the identity otherwise)."""
    on = np.nonzero(flags & AUG_HSV)[0]
    if not on.size:
        return
    ramp = np.arange(0, 256, dtype=np.int16)[None, :]
    for c in range(3):
        rows = on[hsv_shift[on, c] != 0]
        if not rows.size:
            continue
        t = ramp + hsv_shift[rows, c:c + 1]                    # float64 [rows, 256]
        if c == 0:
            np.mod(t, 180, out=t)
        else:
            np.clip(t, 0, 255, out=t)
        hsv_lut[rows, c] = t.astype(np.uint8)


def draw_augmentations_fast(spec: AugmentSpec, n: int, out_h: int, out_w: int, gen: np.random.Generator,
                            host_luts: bool = False) -> AugmentBatch:
    """The same distributions as :func:`draw_augmentations`, drawn for all ``n`` samples at once with a numpy generator
    (op by op instead of sample by sample): ~1 ms per 4096 samples instead of ~90 ms, which is what keeps a training
    loop with the reference's train pipeline from being bound by the host.  Per op and sample: ``random() < p``, then
    the op's parameters -- ``uniform(a, b)`` as ``a + (b - a) * random()`` (so reversed limits behave as in Python),
    ``randint`` inclusive on both ends.  The STREAM differs from the per-sample draw (and from albumentations', which
    cannot be pinned here anyway); the arithmetic applied to a drawn parameter set does not.  The hue / sat / val tables
    are left to the device (``hsv_lut`` None: ``ops.upload_augment`` builds them from ``hsv_shift`` with
    nkbk_build_hsv_luts) unless ``host_luts``."""
    max_holes = max(1, spec.holes[1]) if "CoarseDropout" in spec.order else 1
    flags = np.zeros(n, dtype=np.int32)
    alpha = np.ones(n, dtype=np.float32)
    beta = np.zeros(n, dtype=np.float32)
    bright = np.zeros(n, dtype=np.float64)
    holes = np.zeros((n, max_holes, 4), dtype=np.int32)
    has_hsv = "HueSaturationValue" in spec.order
    hsv_shift = np.zeros((n, 3), dtype=np.float64) if has_hsv else None
    hsv_lut = np.tile(np.arange(256, dtype=np.uint8), (n, 3, 1)) if (has_hsv and host_luts) else None

    def uni(lim):
        return lim[0] + (lim[1] - lim[0]) * gen.random(n)

    for name in spec.order:
        if name == "HorizontalFlip":
            flags |= np.where(gen.random(n) < spec.hflip_p, AUG_HFLIP, 0).astype(np.int32)
        elif name == "VerticalFlip":
            flags |= np.where(gen.random(n) < spec.vflip_p, AUG_VFLIP, 0).astype(np.int32)
        elif name == "RandomBrightnessContrast":
            on = gen.random(n) < spec.bc_p
            a, b = 1.0 + uni(spec.contrast_limit), 0.0 + uni(spec.brightness_limit)
            alpha = np.where(on, a.astype(np.float32), alpha).astype(np.float32)
            beta = np.where(on, (b * 255).astype(np.float32), beta).astype(np.float32)
            bright = np.where(on, b, bright)
            flags |= np.where(on, AUG_BC, 0).astype(np.int32)
        elif name == "HueSaturationValue":
            on = gen.random(n) < spec.hsv_p
            sh = np.stack([uni(spec.hue_limit), uni(spec.sat_limit), uni(spec.val_limit)], 1)
            hsv_shift[on] = sh[on]
            flags |= np.where(on & (sh != 0).any(1), AUG_HSV, 0).astype(np.int32)
        elif name == "CoarseDropout":
            on = gen.random(n) < spec.cd_p

            def rint_incl(lo, hi):     # randint(lo, hi) inclusive, hi an int or an array: floor(lo + u * (hi - lo + 1))
                return (lo + np.floor(gen.random(n) * (hi - lo + 1))).astype(np.int32)

            k = rint_incl(spec.holes[0], spec.holes[1])
            ints = all(isinstance(v, (int, np.integer)) for v in (*spec.hole_h, *spec.hole_w))
            for h in range(max_holes):
                if ints:
                    hh, hw = rint_incl(spec.hole_h[0], spec.hole_h[1]), rint_incl(spec.hole_w[0], spec.hole_w[1])
                else:
                    hh = (out_h * uni(spec.hole_h)).astype(np.int32)
                    hw = (out_w * uni(spec.hole_w)).astype(np.int32)
                y1, x1 = rint_incl(0, out_h - hh), rint_incl(0, out_w - hw)
                use = (on & (h < k)).astype(np.int32)            # holes beyond the drawn count stay (0, 0, 0, 0)
                holes[:, h, 0], holes[:, h, 1] = x1 * use, y1 * use
                holes[:, h, 2], holes[:, h, 3] = (x1 + hw) * use, (y1 + hh) * use
            flags |= np.where(on, k << 8, 0).astype(np.int32)
    if hsv_lut is not None:
        _fill_hsv_luts(flags, hsv_shift, hsv_lut)
    lanes = cv2_hsv_simd_lanes() if has_hsv else 0
    return AugmentBatch(flags, alpha, beta, holes, spec.fill, bright, hsv_shift, hsv_lut,
                        (out_w // lanes) * lanes if lanes else 0)


def draw_augmentations(spec: AugmentSpec, n: int, out_h: int, out_w: int, rng=None) -> AugmentBatch:
    """Draw what ``A.Compose.__call__`` would draw for ``n`` samples: per op, ``random() < p`` and then the op's
    ``get_params`` (albumentations 1.3.x: RandomBrightnessContrast draws contrast then brightness with ``uniform``;
    CoarseDropout draws ``randint(min_holes, max_holes)`` holes, each ``randint`` sizes (or ``int(size * uniform)``
    for fractional limits) and then ``y1``, ``x1``).  ``rng``: a ``random.Random`` (default: the module-level one,
    which is what albumentations itself uses)."""
    rng = rng or _random
    max_holes = max(1, spec.holes[1]) if "CoarseDropout" in spec.order else 1
    flags = np.zeros(n, dtype=np.int32)
    alpha = np.ones(n, dtype=np.float32)
    beta = np.zeros(n, dtype=np.float32)
    bright = np.zeros(n, dtype=np.float64)
    holes = np.zeros((n, max_holes, 4), dtype=np.int32)
    has_hsv = "HueSaturationValue" in spec.order
    hsv_shift = np.zeros((n, 3), dtype=np.float64) if has_hsv else None
    hsv_lut = np.tile(np.arange(256, dtype=np.uint8), (n, 3, 1)) if has_hsv else None
    for i in range(n):
        f = 0
        for name in spec.order:
            if name == "HorizontalFlip":
                if rng.random() < spec.hflip_p:
                    f |= AUG_HFLIP
            elif name == "VerticalFlip":
                if rng.random() < spec.vflip_p:
                    f |= AUG_VFLIP
            elif name == "RandomBrightnessContrast":
                if rng.random() < spec.bc_p:
                    a = 1.0 + rng.uniform(spec.contrast_limit[0], spec.contrast_limit[1])
                    b = 0.0 + rng.uniform(spec.brightness_limit[0], spec.brightness_limit[1])
                    f |= AUG_BC
                    alpha[i] = np.float32(a)          # `lut *= alpha`: python float enters a float32 multiply
                    beta[i] = np.float32(b * 255)     # `lut += beta * max_value`: double product, float32 add
                    bright[i] = b
            elif name == "HueSaturationValue":
                if rng.random() < spec.hsv_p:
                    hs = rng.uniform(spec.hue_limit[0], spec.hue_limit[1])
                    ss = rng.uniform(spec.sat_limit[0], spec.sat_limit[1])
                    vs = rng.uniform(spec.val_limit[0], spec.val_limit[1])
                    hsv_shift[i] = (hs, ss, vs)
                    if hs != 0 or ss != 0 or vs != 0:    # `shift_hsv` returns the image untouched for all-zero shifts
                        f |= AUG_HSV                     # (its tables are built for the whole batch below)
            elif name == "CoarseDropout":
                if rng.random() < spec.cd_p:
                    k = rng.randint(spec.holes[0], spec.holes[1])
                    for h in range(k):
                        if all(isinstance(v, (int, np.integer)) for v in (*spec.hole_h, *spec.hole_w)):
                            hh = rng.randint(spec.hole_h[0], spec.hole_h[1])
                            hw = rng.randint(spec.hole_w[0], spec.hole_w[1])
                        else:
                            hh = int(out_h * rng.uniform(spec.hole_h[0], spec.hole_h[1]))
                            hw = int(out_w * rng.uniform(spec.hole_w[0], spec.hole_w[1]))
                        y1 = rng.randint(0, out_h - hh)
                        x1 = rng.randint(0, out_w - hw)
                        holes[i, h] = (x1, y1, x1 + hw, y1 + hh)
                    f |= k << 8
        flags[i] = f
    if has_hsv:
        _fill_hsv_luts(flags, hsv_shift, hsv_lut)
    lanes = cv2_hsv_simd_lanes() if has_hsv else 0
    return AugmentBatch(flags, alpha, beta, holes, spec.fill, bright, hsv_shift, hsv_lut,
                        (out_w // lanes) * lanes if lanes else 0)


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
    augment: Optional[AugmentSpec] = None   # train-time ops fused into K1 (None: val / inference pipeline)

    def with_channel_swap(self, swap: bool) -> "PreprocessPlan":
        return dataclasses.replace(self, channel_swap=bool(swap))

    def draw(self, n: int, rng=None) -> Optional[AugmentBatch]:
        """Per-sample augmentation parameters for a batch of ``n`` (None for a val / inference pipeline).
        ``rng``: None -> all samples at once with a numpy generator seeded from Python's global ``random`` (so
        ``random.seed(...)`` still governs the run, as it governs albumentations); a ``numpy.random.Generator`` -> the
        same with that generator; a ``random.Random`` -> the per-sample walk in albumentations' call order
        (:func:`draw_augmentations`, ~90 ms per 4096 samples)."""
        if self.augment is None:
            return None
        if rng is None:
            rng = np.random.default_rng(_random.getrandbits(64))
        if isinstance(rng, np.random.Generator):
            return draw_augmentations_fast(self.augment, n, self.out_h, self.out_w, rng)
        return draw_augmentations(self.augment, n, self.out_h, self.out_w, rng)


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


_AUG_KINDS = ("HorizontalFlip", "VerticalFlip", "RandomBrightnessContrast", "HueSaturationValue", "CoarseDropout")


def _prob(op) -> float:
    return 1.0 if getattr(op, "always_apply", False) else float(getattr(op, "p", 0.5))


def _fill_as_u8(value) -> Tuple[int, int, int]:
    """What ``img[y1:y2, x1:x2] = fill_value`` stores into a uint8 image (floats truncate: [0, 0.5, 1] -> 0, 0, 1)."""
    v = np.asarray(value, dtype=np.float64).reshape(-1)
    if v.size == 1:
        v = np.repeat(v, 3)
    if v.size != 3:
        raise NotImplementedError(f"CoarseDropout fill_value {value!r}: need a scalar or 3 channels")
    if np.any(v < 0) or np.any(v >= 256):
        raise NotImplementedError(f"CoarseDropout fill_value {value!r} outside the uint8 range")
    return tuple(int(x) for x in v.astype(np.uint8))


def _compile_augment(aug_ops) -> Optional[AugmentSpec]:
    if not aug_ops:
        return None
    kw = {"order": tuple(_kind(o) for o in aug_ops)}
    for op in aug_ops:
        k = _kind(op)
        if k == "HorizontalFlip":
            kw["hflip_p"] = _prob(op)
        elif k == "VerticalFlip":
            kw["vflip_p"] = _prob(op)
        elif k == "RandomBrightnessContrast":
            if not getattr(op, "brightness_by_max", True):
                raise NotImplementedError("RandomBrightnessContrast(brightness_by_max=False) needs the image mean; "
                                          "only brightness_by_max=True is fused")
            kw["bc_p"] = _prob(op)
            kw["brightness_limit"] = _to_tuple(op.brightness_limit)
            kw["contrast_limit"] = _to_tuple(op.contrast_limit)
        elif k == "HueSaturationValue":
            kw["hsv_p"] = _prob(op)
            kw["hue_limit"] = _to_tuple(op.hue_shift_limit)
            kw["sat_limit"] = _to_tuple(op.sat_shift_limit)
            kw["val_limit"] = _to_tuple(op.val_shift_limit)
        elif k == "CoarseDropout":
            if getattr(op, "mask_fill_value", None) is not None:
                raise NotImplementedError("CoarseDropout(mask_fill_value=...) is not supported (no masks on this path)")
            if op.max_holes > MAX_HOLES:
                raise NotImplementedError(f"CoarseDropout: at most {MAX_HOLES} holes are fused, got {op.max_holes}")
            sizes = (op.min_height, op.max_height, op.min_width, op.max_width)
            ints = all(isinstance(v, (int, np.integer)) for v in sizes)
            floats = all(isinstance(v, float) for v in sizes)
            if not (ints or floats):   # the same check albumentations makes when it draws
                raise ValueError("Min width, max width, min height and max height should all either be ints or floats. "
                                 f"Got: {[type(v) for v in sizes]} respectively")
            kw["cd_p"] = _prob(op)
            kw["holes"] = (int(op.min_holes), int(op.max_holes))
            kw["hole_h"] = (op.min_height, op.max_height)
            kw["hole_w"] = (op.min_width, op.max_width)
            kw["fill"] = _fill_as_u8(op.fill_value)
    return AugmentSpec(**kw)


def compile_pipeline(pipeline, channel_swap: bool = False) -> PreprocessPlan:
    """``A.Compose([...])`` (or a plain list of ops) -> PreprocessPlan."""
    ops = list(getattr(pipeline, "transforms", pipeline))
    resize = lms = pad = norm = None
    saw_tensor = False
    aug_ops: List[Any] = []
    for op in ops:
        k = _kind(op)
        if saw_tensor:
            raise NotImplementedError(f"{k} after ToTensorV2 is not supported")
        if k in _AUG_KINDS:
            if norm is not None or (resize is None and (lms is None or pad is None)):
                raise NotImplementedError(f"{k} must sit between the geometry ops and Normalize")
            if any(_kind(a) == k for a in aug_ops):
                raise NotImplementedError(f"{k} listed twice")
            if k == "RandomBrightnessContrast" and any(_kind(a) == "HueSaturationValue" for a in aug_ops):
                raise NotImplementedError("RandomBrightnessContrast after HueSaturationValue: the fused kernel applies "
                                          "brightness/contrast first (the order of the reference's configs)")
            if any(_kind(a) == "CoarseDropout" for a in aug_ops):
                raise NotImplementedError("CoarseDropout must be the last fused augmentation (its fill is not "
                                          "flipped or re-coloured)")
            aug_ops.append(op)
            continue
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
                f"pipeline op {op!r} is outside the fused subset "
                "{Resize | LongestMaxSize+PadIfNeeded, [HorizontalFlip, VerticalFlip, RandomBrightnessContrast, "
                "HueSaturationValue, CoarseDropout], Normalize, ToTensorV2}"
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
    if float(getattr(norm, "max_pixel_value", 255.0)) != 255.0 and aug_ops:
        raise NotImplementedError("fused augmentations assume max_pixel_value=255 (uint8 images)")
    augment = _compile_augment(aug_ops)
    if resize is not None:
        return PreprocessPlan(MODE_STRETCH, int(resize.height), int(resize.width), 0, (0, 0, 0), mean255, denom,
                              bool(channel_swap), augment)
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
                          mean255, denom, bool(channel_swap), augment)
