"""The oracle is pinned before it is trusted: integer restatement == cv2 == C
restatement == committed cv2-generated goldens (tests/golden/make_golden.py)."""
import json

import numpy as np
import pytest

from oracle import preprocess as opre
from oracle.cref import preprocess_batch_c


def _rand_cases(rng, n):
    for it in range(n):
        sh, sw = int(rng.integers(1, 400)), int(rng.integers(1, 400))
        dh, dw = int(rng.choice([224, 128, 17, 64])), int(rng.choice([224, 128, 33, 64]))
        k = it % 8
        if k == 0: sh, sw = 2 * dh, 2 * dw          # exact 2x (cv2 switches to INTER_AREA)
        if k == 1: sh, sw = dh, dw                  # identity
        if k == 2: sh, sw = 5, 5                    # min_box_size crop
        if k == 3: sh, sw = 3 * dh, 4 * dw          # integer ratios that stay bilinear
        if k == 4: sh, sw = dh - 1, dw + 1
        yield sh, sw, dh, dw


def test_resize_int_equals_cv2_random():
    rng = np.random.default_rng(0)
    for sh, sw, dh, dw in _rand_cases(rng, 160):
        img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        a = opre.resize_int(img, dw, dh)
        b = opre.resize_cv2(img, dw, dh)
        assert np.array_equal(a, b), (sh, sw, dh, dw)


def test_resize_int_equals_cv2_1080p_crop():
    rng = np.random.default_rng(1)
    frame = rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
    for box in [(0, 0, 1920, 1080), (100, 200, 580, 680), (1900, 1000, 1920, 1080), (7, 9, 271, 273)]:
        x0, y0, x1, y1 = box
        crop = frame[y0:y1, x0:x1]
        assert np.array_equal(opre.resize_int(np.ascontiguousarray(crop), 224, 224), opre.resize_cv2(crop, 224, 224))


@pytest.mark.parametrize("tag,plan", [
    ("stretch32x48", opre.Plan(out_h=32, out_w=48)),
    ("letterbox40", opre.Plan(mode=opre.MODE_LETTERBOX, out_h=40, out_w=40, max_size=40)),
    ("letterbox48x56pad", opre.Plan(mode=opre.MODE_LETTERBOX, out_h=48, out_w=56, max_size=44, pad_value=(7, 200, 33))),
    ("stretch64", opre.Plan(out_h=64, out_w=64)),
])
def test_oracles_match_cv2_golden(golden_dir, tag, plan):
    g = np.load(golden_dir / "pixels_golden.npz")
    idx = g[f"{tag}.idx"]
    boxes, fidx = g["boxes"][idx], g["frame_idx"][idx]
    for impl in ("int", "cv2"):
        u8, f32 = opre.preprocess_batch(g["frames"], boxes, fidx, plan, impl)
        assert np.array_equal(u8, g[f"{tag}.u8"]), impl
        assert np.array_equal(f32.view(np.uint32), g[f"{tag}.f32"].view(np.uint32)), impl
    cu8, cf32 = preprocess_batch_c(g["frames"], boxes, fidx, plan)
    assert np.array_equal(cu8, g[f"{tag}.u8"])
    assert np.array_equal(cf32.view(np.uint32), g[f"{tag}.f32"].view(np.uint32))


def test_channel_swap_is_bgr2rgb():
    rng = np.random.default_rng(2)
    frames = rng.integers(0, 256, (1, 50, 70, 3), dtype=np.uint8)
    boxes, fidx = [(3, 4, 60, 44)], [0]
    plan = opre.Plan(out_h=32, out_w=32)
    a_u8, a_f = opre.preprocess_batch(frames[..., ::-1], boxes, fidx, plan, "cv2")
    swapped = opre.Plan(out_h=32, out_w=32, channel_swap=True)
    b_u8, b_f = opre.preprocess_batch(frames, boxes, fidx, swapped, "int")
    c_u8, c_f = preprocess_batch_c(frames, boxes, fidx, swapped)
    assert np.array_equal(a_u8, b_u8) and np.array_equal(a_u8, c_u8)
    assert np.array_equal(a_f, b_f) and np.array_equal(a_f, c_f)


def test_yolo_box_conversion_golden(golden_dir):
    g = json.loads((golden_dir / "yolo_boxes_golden.json").read_text())
    size = tuple(g["image_size"])
    for line, exp in zip(g["lines"], g["expected"]):
        p = line.split()
        box = opre.bbox_xywhn2xyxy(*map(float, p[1:]), size)
        assert list(box) == exp[:4]
        assert opre.box_is_kept(*box) == bool(exp[5])
    kept = opre.parse_yolo_label_lines(g["lines"], size)
    assert len(kept) == sum(e[5] for e in g["expected"])


def test_normalize_is_two_rounded_ops_not_fma():
    m, d = opre.normalize_constants((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
    assert [float.hex(float(x)) for x in m] == ["0x1.eeb3340000000p+6", "0x1.d11eb80000000p+6", "0x1.9e1eb80000000p+6"]
    x = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, axis=2)
    y = opre.normalize_f32(x, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
    tv = (x.astype(np.float32) / np.float32(255) - np.array((0.485, 0.456, 0.406), np.float32)) / np.array((0.229, 0.224, 0.225), np.float32)
    assert not np.array_equal(y, tv)  # the torchvision-style formula is NOT what the reference computes


def test_bf16_rounding_helper():
    import torch
    x = np.random.default_rng(3).normal(size=1000).astype(np.float32)
    ref = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(opre.f32_to_bf16_bits(x), ref)
