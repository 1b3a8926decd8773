import numpy as np
import pytest

from nkb_classification_b200 import transforms as T
from oracle import preprocess as opre


def test_compile_resize():
    p = T.compile_pipeline(T.Compose([T.Resize(224, 200), T.Normalize(), T.ToTensorV2()]))
    assert (p.mode, p.out_h, p.out_w) == (T.MODE_STRETCH, 224, 200)
    m, d = opre.normalize_constants((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
    assert np.array_equal(np.array(p.mean255, np.float32), m) and np.array_equal(np.array(p.denom, np.float32), d)


def test_compile_letterbox_like_reference_configs():
    # configs/singletask_config.py:203-219
    p = T.compile_pipeline(T.Compose([
        T.LongestMaxSize(128, always_apply=True),
        T.PadIfNeeded(128, 128, always_apply=True, border_mode=T.BORDER_CONSTANT, value=0),
        T.Normalize(mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)),
        T.ToTensorV2(),
    ]))
    assert (p.mode, p.out_h, p.out_w, p.max_size, p.pad_value) == (T.MODE_LETTERBOX, 128, 128, 128, (0, 0, 0))


def test_duck_typed_albumentations_objects():
    class Resize:  # stands in for albumentations.Resize (same class name + public attrs)
        height, width, interpolation = 64, 32, 1

    class Normalize:
        mean, std, max_pixel_value = (0.5, 0.5, 0.5), (0.25, 0.25, 0.25), 255.0

    class ToTensorV2:
        pass

    class Compose:
        transforms = [Resize(), Normalize(), ToTensorV2()]

    p = T.compile_pipeline(Compose())
    assert (p.out_h, p.out_w) == (64, 32)


@pytest.mark.parametrize("ops", [
    [T.Resize(8, 8), T.ToTensorV2()],                                            # no Normalize
    [T.Resize(8, 8), T.Normalize()],                                             # no ToTensorV2
    [T.LongestMaxSize(8), T.Normalize(), T.ToTensorV2()],                        # variable output size
    [T.LongestMaxSize(8), T.PadIfNeeded(8, 8), T.Normalize(), T.ToTensorV2()],   # default BORDER_REFLECT_101
    [T.LongestMaxSize(16), T.PadIfNeeded(8, 8, border_mode=0), T.Normalize(), T.ToTensorV2()],
    [T.Resize(8, 8, interpolation=2), T.Normalize(), T.ToTensorV2()],            # cubic
    [type("HorizontalFlip", (), {})(), T.Resize(8, 8), T.Normalize(), T.ToTensorV2()],  # before the geometry
    [T.Resize(8, 8), type("RandomShadow", (), {})(), T.Normalize(), T.ToTensorV2()],  # not fused (stays on CPU)
    [T.Resize(8, 8), T.HueSaturationValue(), T.RandomBrightnessContrast(), T.Normalize(), T.ToTensorV2()],  # order
    [T.Resize(8, 8), type("MotionBlur", (), {})(), T.Normalize(), T.ToTensorV2()],
    [T.Resize(8, 8), T.Normalize(), T.HorizontalFlip(), T.ToTensorV2()],               # after Normalize
    [T.Resize(8, 8), T.CoarseDropout(), T.HorizontalFlip(), T.Normalize(), T.ToTensorV2()],  # dropout must be last
    [T.Resize(8, 8), T.VerticalFlip(), T.VerticalFlip(), T.Normalize(), T.ToTensorV2()],
    [T.Resize(8, 8), T.RandomBrightnessContrast(brightness_by_max=False), T.Normalize(), T.ToTensorV2()],
    [T.Resize(8, 8), T.CoarseDropout(max_holes=40), T.Normalize(), T.ToTensorV2()],
])
def test_unsupported_pipelines_raise(ops):
    with pytest.raises(NotImplementedError):
        T.compile_pipeline(ops)


# ---- train-time pipeline (configs/singletask_config.py:162-201 minus HueSaturationValue) ----
def reference_train_ops(size=128):
    return [
        T.LongestMaxSize(size, always_apply=True),
        T.PadIfNeeded(size, size, always_apply=True, border_mode=T.BORDER_CONSTANT, value=0),
        T.HorizontalFlip(p=0.5),
        T.VerticalFlip(p=0.5),
        T.RandomBrightnessContrast(brightness_limit=(-0.2, 0.2), contrast_limit=(0.1, -0.5), p=0.5),
        T.HueSaturationValue(hue_shift_limit=0, sat_shift_limit=10, val_shift_limit=50, p=0.5),
        T.CoarseDropout(max_holes=4, min_holes=1, max_height=0.2, min_height=0.05, max_width=0.2, min_width=0.05,
                        fill_value=[0, 0.5, 1], p=0.5),
        T.Normalize(mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)),
        T.ToTensorV2(),
    ]


def test_compile_train_pipeline_of_the_reference_config():
    """configs/singletask_config.py:162-201, every op of it."""
    p = T.compile_pipeline(T.Compose(reference_train_ops()))
    a = p.augment
    assert a.order == ("HorizontalFlip", "VerticalFlip", "RandomBrightnessContrast", "HueSaturationValue",
                       "CoarseDropout")
    assert (a.hflip_p, a.vflip_p, a.bc_p, a.hsv_p, a.cd_p) == (0.5, 0.5, 0.5, 0.5, 0.5)
    assert (a.hue_limit, a.sat_limit, a.val_limit) == ((0.0, 0.0), (-10.0, 10.0), (-50.0, 50.0))
    assert a.contrast_limit == (0.1, -0.5)          # kept as given, not sorted (albumentations 1.3 to_tuple)
    assert a.holes == (1, 4) and a.hole_h == (0.05, 0.2)
    assert a.fill == (0, 0, 1)                      # what `img[...] = [0, 0.5, 1]` stores into uint8
    assert T.compile_pipeline(T.Compose([T.Resize(8, 8), T.Normalize(), T.ToTensorV2()])).augment is None
    assert T.RandomBrightnessContrast(brightness_limit=0.3).brightness_limit == (-0.3, 0.3)
    with pytest.raises(ValueError):
        T.CoarseDropout(max_holes=2, min_holes=3)
    with pytest.raises(ValueError):                 # mixed int / float hole sizes: albumentations' own check
        T.compile_pipeline([T.Resize(8, 8), T.CoarseDropout(max_height=4, max_width=0.5), T.Normalize(), T.ToTensorV2()])


def test_draw_follows_compose_call_order():
    """Replaying the draws by hand with the same random.Random stream gives the same parameters:
    per op `random() < p`, then RandomBrightnessContrast: contrast, brightness; CoarseDropout: n, (h, w, y1, x1) * n."""
    import random
    p = T.compile_pipeline(T.Compose(reference_train_ops(64)))
    n = 200
    b = p.draw(n, random.Random(1234))
    r = random.Random(1234)
    for i in range(n):
        hf = r.random() < 0.5
        vf = r.random() < 0.5
        bc = r.random() < 0.5
        if bc:
            alpha = 1.0 + r.uniform(0.1, -0.5)
            beta = 0.0 + r.uniform(-0.2, 0.2)
            assert b.alpha[i] == np.float32(alpha) and b.beta[i] == np.float32(beta * 255)
            assert 0.5 <= alpha <= 1.1 and -0.2 <= beta <= 0.2
        hsv = r.random() < 0.5
        if hsv:
            shifts = (r.uniform(-0.0, 0.0), r.uniform(-10.0, 10.0), r.uniform(-50.0, 50.0))
            assert tuple(b.hsv_shift[i]) == shifts and shifts[0] == 0
            assert np.array_equal(b.hsv_lut[i], T.hsv_luts(*shifts))
            assert np.array_equal(b.hsv_lut[i][0], np.arange(256, dtype=np.uint8))     # zero hue shift: identity table
        holes = []
        if r.random() < 0.5:
            for _ in range(r.randint(1, 4)):
                hh, hw = int(64 * r.uniform(0.05, 0.2)), int(64 * r.uniform(0.05, 0.2))
                y1 = r.randint(0, 64 - hh)
                x1 = r.randint(0, 64 - hw)
                holes.append((x1, y1, x1 + hw, y1 + hh))
        assert b.flags[i] == (int(hf) | 2 * int(vf) | 4 * int(bc) | 8 * int(hsv) | len(holes) << 8)
        assert [tuple(h) for h in b.holes[i, :len(holes)]] == holes
    assert 60 < int((b.flags & 1).sum()) < 140 and 60 < int((b.flags >> 2 & 1).sum()) < 140
    assert T.compile_pipeline([T.Resize(8, 8), T.Normalize(), T.ToTensorV2()]).draw(5) is None


def test_brightness_contrast_lut_matches_oracle(nkbk_lib):
    """The per-pixel arithmetic K1 uses (compiled for the host) == the albumentations LUT restated in the oracle."""
    import random
    from nkb_classification_b200 import ops
    r = random.Random(5)
    cases = [(1.0, 0.0), (1.1, 0.2), (0.5, -0.2), (1.0, 0.1), (0.73, 0.0)]
    cases += [(1.0 + r.uniform(0.1, -0.5), r.uniform(-0.2, 0.2)) for _ in range(300)]
    cases += [(1.0 + r.uniform(-1.0, 1.0), r.uniform(-1.0, 1.0)) for _ in range(100)]
    for alpha, beta in cases:
        got = ops.debug_brightness_contrast_lut(np.float32(alpha), np.float32(beta * 255))
        assert np.array_equal(got, opre.brightness_contrast_lut(alpha, beta)), (alpha, beta)


def test_oracle_augment_is_cv2_and_numpy():
    import cv2
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, (20, 30, 3), dtype=np.uint8)
    a = opre.AugSample(hflip=True, vflip=True, bc=True, alpha=0.8, beta=0.1, holes=[(3, 4, 9, 11)], fill=(0, 0, 1))
    out = opre.augment_u8(img, a)
    exp = img[::-1, ::-1].copy()
    exp = cv2.LUT(exp, np.clip(np.arange(256, dtype=np.float32) * np.float32(0.8) + np.float32(0.1 * 255), 0, 255).astype(np.uint8))
    exp[4:11, 3:9] = (0, 0, 1)
    assert np.array_equal(out, exp)
    assert np.array_equal(opre.augment_u8(img, opre.AugSample()), img)


def test_hsv_restatements_match_cv2_exhaustively():
    """Pins the colour-space restatements to cv2 on their ENTIRE domains: RGB2HSV (fixed point) over all 2^24 RGB
    triples; HSV2RGB (float path with OpenCV's fma contraction) over all 180 * 2^16 HSV triples, in BOTH of cv2's
    roundings: one pixel per row exercises its scalar code (round to nearest even), 65536-pixel rows its vectorised
    body (truncation)."""
    import cv2
    g = np.arange(256, dtype=np.uint8)
    rgb = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 1, 3)
    assert np.array_equal(opre.rgb2hsv_int(rgb), cv2.cvtColor(rgb, cv2.COLOR_RGB2HSV))
    hsv = np.stack(np.meshgrid(np.arange(180, dtype=np.uint8), g, g, indexing="ij"), -1).reshape(180, 65536, 3)
    assert np.array_equal(opre.hsv2rgb_f32(hsv.reshape(-1, 1, 3)), cv2.cvtColor(hsv.reshape(-1, 1, 3), cv2.COLOR_HSV2RGB))
    if T.cv2_hsv_simd_lanes():
        assert np.array_equal(opre.hsv2rgb_f32(hsv, trunc=True), cv2.cvtColor(hsv, cv2.COLOR_HSV2RGB))


def test_hsv_pixel_function_matches_albumentations_oracle(nkbk_lib):
    """K1's HueSaturationValue pixel function (the same source compiled for the host) == albumentations'
    `_shift_hsv_uint8` executed verbatim with cv2, on every RGB triple, for shifts of the reference config's range
    and for hue shifts, in both roundings; all-zero shifts leave the image untouched (flag not set by the draw)."""
    from nkb_classification_b200 import ops
    g = np.arange(256, dtype=np.uint8)
    rgb = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(4096, 4096, 3)
    lanes = T.cv2_hsv_simd_lanes()
    assert lanes in (0, 16, 32, 64)
    for shifts in [(0.0, 7.3, -31.2), (12.5, -9.99, 49.9), (-170.25, 0.0, 0.0)]:
        lut = T.hsv_luts(*shifts)
        tall = rgb.reshape(-1, 1, 3)                       # one pixel per row: cv2's scalar rounding
        assert np.array_equal(ops.debug_hsv_shift(tall, lut, False), opre.shift_hsv_u8(tall, *shifts)), shifts
        if lanes:                                          # 4096-pixel rows: cv2's vectorised rounding everywhere
            assert np.array_equal(ops.debug_hsv_shift(rgb, lut, True), opre.shift_hsv_u8(rgb, *shifts)), shifts
    assert opre.shift_hsv_u8(rgb, 0, 0, 0) is rgb
    assert np.array_equal(T.hsv_luts(0, 0, 0), np.tile(g, (3, 1)))
    # a row that is not a multiple of the vector width: body truncates, tail rounds -- the rule K1 is given
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (9, 100, 3), dtype=np.uint8)
    lut = T.hsv_luts(3.3, 4.4, -20.2)
    tc = (100 // lanes) * lanes if lanes else 0
    emu = np.concatenate([ops.debug_hsv_shift(np.ascontiguousarray(img[:, :tc]), lut, True),
                          ops.debug_hsv_shift(np.ascontiguousarray(img[:, tc:]), lut, False)], 1)
    assert np.array_equal(emu, opre.shift_hsv_u8(img, 3.3, 4.4, -20.2))
    b = T.compile_pipeline(reference_train_ops(100)).draw(3, __import__("random").Random(0))
    assert b.hsv_trunc_cols == tc


def test_vectorised_draw_has_the_distributions_of_the_per_sample_draw():
    """`plan.draw(n)` (numpy, op by op) against `plan.draw(n, random.Random)` (albumentations' call order): same flag
    frequencies, parameter ranges and hole geometry; holes beyond the drawn count stay zero; no host tables."""
    import random
    p = T.compile_pipeline(T.Compose(reference_train_ops(64)))
    n = 40000
    random.seed(11)
    a = p.draw(n)
    b = p.draw(n, random.Random(12))
    assert a.hsv_lut is None and b.hsv_lut is not None
    for bit in range(4):
        fa, fb = ((a.flags >> bit) & 1).mean(), ((b.flags >> bit) & 1).mean()
        assert abs(fa - 0.5) < 0.02 and abs(fb - 0.5) < 0.02
    ka, kb = a.flags >> 8, b.flags >> 8
    assert ka.max() == 4 and set(np.unique(ka)) == {0, 1, 2, 3, 4} and abs(ka.mean() - kb.mean()) < 0.05
    on = (a.flags & 4) != 0
    assert 0.5 <= a.alpha[on].min() and a.alpha[on].max() <= 1.1 and np.all(a.alpha[~on] == 1) and np.all(a.beta[~on] == 0)
    assert np.allclose(a.beta[on], (a.brightness[on] * 255).astype(np.float32)) and np.abs(a.brightness).max() <= 0.2
    hs = (a.flags & 8) != 0
    assert np.all(a.hsv_shift[hs, 0] == 0) and np.abs(a.hsv_shift[:, 1]).max() <= 10 and np.abs(a.hsv_shift[:, 2]).max() <= 50
    w, h = a.holes[..., 2] - a.holes[..., 0], a.holes[..., 3] - a.holes[..., 1]
    assert np.array_equal((w > 0).sum(1), ka) and np.array_equal((h > 0).sum(1), ka)
    assert w[w > 0].min() >= int(64 * 0.05) and w.max() <= int(64 * 0.2) and a.holes[..., 2].max() <= 64 and a.holes.min() >= 0
    wb = b.holes[..., 2] - b.holes[..., 0]
    assert abs(w[w > 0].mean() - wb[wb > 0].mean()) < 0.2
