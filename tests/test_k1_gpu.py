"""K1 parity on the GPU: CUDA (through the C ABI) vs the CPU oracle.
uint8 pixels bit-exact, fp32 tensors bit-exact (0 ulp; the bar is <= 1 ulp),
bf16 = RNE of the fp32 value."""
import numpy as np
import pytest
import torch

from oracle import preprocess as opre
from oracle.cref import preprocess_batch_c

pytestmark = pytest.mark.gpu

MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def make_plan(T, mode="stretch", out_h=224, out_w=224, max_size=None, pad=0, swap=False):
    if mode == "stretch":
        ops = [T.Resize(out_h, out_w)]
    else:
        ops = [T.LongestMaxSize(max_size or min(out_h, out_w)), T.PadIfNeeded(out_h, out_w, border_mode=0, value=pad)]
    return T.compile_pipeline(ops + [T.Normalize(MEAN, STD), T.ToTensorV2()], channel_swap=swap)


def oracle_plan(plan):
    return opre.Plan(mode=plan.mode, out_h=plan.out_h, out_w=plan.out_w, max_size=plan.max_size,
                     pad_value=plan.pad_value, mean=MEAN, std=STD, channel_swap=plan.channel_swap)


def run_k1(dev, frames, boxes, fidx, plan, dtype=torch.float32, want_u8=True, **kw):
    from nkb_classification_b200 import ops
    fr = torch.from_numpy(frames).to(dev) if isinstance(frames, np.ndarray) else frames
    bx = torch.from_numpy(np.ascontiguousarray(boxes, dtype=np.int32)).to(dev)
    fi = torch.from_numpy(np.ascontiguousarray(fidx, dtype=np.int32)).to(dev)
    n = len(fidx)
    u8 = torch.full((n, plan.out_h, plan.out_w, 3), 99, dtype=torch.uint8, device=dev) if want_u8 else None
    out = ops.preprocess_crops(fr, bx, fi, plan, out_dtype=dtype, out_u8=u8, **kw)
    torch.cuda.synchronize()
    return out, u8


def assert_same_f32(got: torch.Tensor, exp: np.ndarray):
    g = got.cpu().numpy()
    assert g.shape == exp.shape
    assert np.array_equal(g.view(np.uint32), exp.view(np.uint32)), f"max abs diff {np.abs(g - exp).max()}"


@pytest.mark.parametrize("tag,kw", [
    ("stretch32x48", dict(out_h=32, out_w=48)),
    ("letterbox40", dict(mode="letterbox", out_h=40, out_w=40)),
    ("letterbox48x56pad", dict(mode="letterbox", out_h=48, out_w=56, max_size=44, pad=(7, 200, 33))),
    ("stretch64", dict(out_h=64, out_w=64)),
])
def test_k1_matches_cv2_golden(cuda_device, golden_dir, tag, kw):
    from nkb_classification_b200 import transforms as T
    g = np.load(golden_dir / "pixels_golden.npz")
    idx = g[f"{tag}.idx"]
    plan = make_plan(T, **kw)
    out, u8 = run_k1(cuda_device, g["frames"], g["boxes"][idx], g["frame_idx"][idx], plan)
    assert np.array_equal(u8.cpu().numpy(), g[f"{tag}.u8"])
    assert_same_f32(out, g[f"{tag}.f32"])


def _random_boxes(rng, n, H, W, wmin=1, wmax=None, hmax=None):
    wmax, hmax = wmax or W, hmax or H
    boxes = []
    for _ in range(n):
        w, h = int(rng.integers(wmin, wmax + 1)), int(rng.integers(wmin, hmax + 1))
        x0, y0 = int(rng.integers(0, W - w + 1)), int(rng.integers(0, H - h + 1))
        boxes.append((x0, y0, x0 + w, y0 + h))
    return boxes


@pytest.mark.parametrize("kw", [
    dict(out_h=224, out_w=224),
    dict(out_h=224, out_w=224, swap=True),
    dict(out_h=128, out_w=128),
    dict(out_h=256, out_w=256),
    dict(out_h=100, out_w=60),
    dict(out_h=40, out_w=384),
    dict(mode="letterbox", out_h=224, out_w=224),
    dict(mode="letterbox", out_h=128, out_w=128, swap=True, pad=(1, 2, 250)),
    dict(mode="letterbox", out_h=96, out_w=160, max_size=90),
])
def test_k1_random_boxes_vs_oracle(cuda_device, kw):
    from nkb_classification_b200 import transforms as T
    rng = np.random.default_rng(11)
    frames = rng.integers(0, 256, (3, 270, 480, 3), dtype=np.uint8)
    boxes = _random_boxes(rng, 96, 270, 480, wmin=3 if kw.get("mode") == "letterbox" else 1)
    if kw.get("mode") == "letterbox":  # drop boxes whose letterboxed side rounds to 0 (the reference asserts there)
        ms = kw.get("max_size") or min(kw["out_h"], kw["out_w"])
        boxes = [b for b in boxes if min(opre.letterbox_geometry(b[3] - b[1], b[2] - b[0], ms, kw["out_h"], kw["out_w"])[:2]) >= 1]
    boxes += [(0, 0, 480, 270), (479, 269, 480, 270) if kw.get("mode") != "letterbox" else (470, 260, 480, 270),
              (0, 0, 2, 270), (470, 0, 480, 270), (0, 265, 480, 270)]
    fidx = [i % 3 for i in range(len(boxes))]
    plan = make_plan(T, **kw)
    out, u8 = run_k1(cuda_device, frames, boxes, fidx, plan)
    eu8, ef32 = preprocess_batch_c(frames, boxes, fidx, oracle_plan(plan))
    assert np.array_equal(u8.cpu().numpy(), eu8)
    assert_same_f32(out, ef32)
    # without the uint8 side output the TMA fast path runs where it applies: same bits
    out_fast, _ = run_k1(cuda_device, frames, boxes, fidx, plan, want_u8=False)
    assert_same_f32(out_fast, ef32)
    # bf16 output is the RNE rounding of the same fp32 values
    outb, _ = run_k1(cuda_device, frames, boxes, fidx, plan, dtype=torch.bfloat16, want_u8=False)
    got_bits = outb.view(torch.int16).cpu().numpy().view(np.uint16)
    assert np.array_equal(got_bits, opre.f32_to_bf16_bits(ef32).reshape(ef32.shape))


def test_k1_edge_boxes_1080p(cuda_device):
    """SURVEY 8d edge set: 5x5, 5x1080, full frame, x1 = W, exact 2x, identity, 223x225."""
    from nkb_classification_b200 import transforms as T
    rng = np.random.default_rng(12)
    frames = rng.integers(0, 256, (2, 1080, 1920, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:1080, 0:1920]
    frames[1, :, :, 0] = (xx + 2 * yy) % 256
    boxes = [(100, 100, 105, 105), (7, 0, 12, 1080), (0, 0, 1920, 1080), (1500, 300, 1920, 700), (10, 20, 458, 468),
             (33, 44, 257, 268), (1, 1, 224, 226), (0, 1075, 1920, 1080), (1915, 0, 1920, 1080), (640, 360, 1120, 840),
             (1919, 1079, 1920, 1080)]
    fidx = [i % 2 for i in range(len(boxes))]
    plan = make_plan(T)
    out, u8 = run_k1(cuda_device, frames, boxes, fidx, plan)
    eu8, ef32 = preprocess_batch_c(frames, boxes, fidx, oracle_plan(plan))
    assert np.array_equal(u8.cpu().numpy(), eu8)
    assert_same_f32(out, ef32)
    out_fast, _ = run_k1(cuda_device, frames, boxes, fidx, plan, want_u8=False)  # fast path + fix-up for wide boxes
    assert_same_f32(out_fast, ef32)
    # identity crop: the resized pixels are the source pixels
    assert np.array_equal(u8[5].cpu().numpy(), frames[1][44:268, 33:257])


def test_k1_bad_boxes_are_counted_and_padded(cuda_device):
    from nkb_classification_b200 import transforms as T
    rng = np.random.default_rng(13)
    frames = rng.integers(0, 256, (1, 64, 64, 3), dtype=np.uint8)
    boxes = [(0, 0, 10, 10), (5, 5, 5, 20), (10, 10, 70, 20), (-1, 0, 10, 10), (3, 9, 2, 12), (20, 20, 40, 40)]
    fidx = [0, 0, 0, 0, 0, 7]
    plan = make_plan(T, out_h=32, out_w=32)
    bad = torch.zeros(1, dtype=torch.int32, device=cuda_device)
    out, u8 = run_k1(cuda_device, frames, boxes, fidx, plan, bad_count=bad)
    assert int(bad.item()) == 5
    eu8, ef32 = preprocess_batch_c(frames, boxes[:1], fidx[:1], oracle_plan(plan))
    assert np.array_equal(u8[0].cpu().numpy(), eu8[0])
    m, d = opre.normalize_constants(MEAN, STD)
    pad = ((np.zeros(3, np.float32) - m) * d).astype(np.float32)
    for i in range(1, 6):
        assert np.array_equal(out[i].cpu().numpy(), np.broadcast_to(pad[:, None, None], (3, 32, 32)))


def test_k1_ragged_frames_with_pitch(cuda_device):
    """Frames of different sizes in one buffer, rows padded to an odd pitch (unaligned 3-byte pixels)."""
    from nkb_classification_b200 import ops, transforms as T
    rng = np.random.default_rng(14)
    shapes = [(50, 67, 211), (33, 90, 275), (80, 41, 130)]  # (H, W, pitch) -- pitches not multiples of 4
    bufs, desc, off, frames = [], [], 3, []                  # start at an odd offset
    flat = np.zeros(3 + sum(h * p for h, w, p in shapes) + 16, dtype=np.uint8)
    for h, w, p in shapes:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        frames.append(img)
        view = flat[off: off + h * p].reshape(h, p)
        view[:, : w * 3] = img.reshape(h, w * 3)
        desc.append((off, h, w, p))
        off += h * p
    boxes, fidx = [], []
    for fi, (h, w, p) in enumerate(shapes):
        for b in _random_boxes(rng, 12, h, w) + [(0, 0, w, h), (w - 1, 0, w, h), (w - 2, h - 2, w, h)]:
            boxes.append(b)
            fidx.append(fi)
    plan = make_plan(T, out_h=48, out_w=72)
    dev = cuda_device
    out_u8 = torch.empty((len(boxes), 48, 72, 3), dtype=torch.uint8, device=dev)
    out = ops.preprocess_crops(torch.from_numpy(flat).to(dev), torch.tensor(boxes, dtype=torch.int32, device=dev),
                               torch.tensor(fidx, dtype=torch.int32, device=dev), plan, out_u8=out_u8,
                               frame_desc=torch.tensor(desc, dtype=torch.int64, device=dev))
    torch.cuda.synchronize()
    # same call without the side output: frames are not 16-byte aligned, so the fix-up pass produces every crop
    out2 = ops.preprocess_crops(torch.from_numpy(flat).to(dev), torch.tensor(boxes, dtype=torch.int32, device=dev),
                                torch.tensor(fidx, dtype=torch.int32, device=dev), plan,
                                frame_desc=torch.tensor(desc, dtype=torch.int64, device=dev))
    torch.cuda.synchronize()
    assert torch.equal(out, out2)
    op = oracle_plan(plan)
    for i, (b, fi) in enumerate(zip(boxes, fidx)):
        eu8, ef = opre.preprocess_crop(frames[fi], b, op, "int")
        assert np.array_equal(out_u8[i].cpu().numpy(), eu8), (i, b, fi)
        assert np.array_equal(out[i].cpu().numpy().view(np.uint32), ef.view(np.uint32))


def test_k1_fast_path_mixed_alignment_and_widths(cuda_device):
    """One launch mixing crops the TMA kernel takes (aligned frames, narrow boxes) with crops it must leave to
    the fix-up pass (frame with a 16-byte-misaligned pitch, boxes wider than the ring): every crop exactly once."""
    from nkb_classification_b200 import ops, transforms as T
    rng = np.random.default_rng(15)
    dev = cuda_device
    shapes = [(300, 1600, 4800), (200, 333, 1000), (120, 640, 1920 + 16)]   # (H, W, pitch): aligned / misaligned / aligned+padded
    total = sum(h * p for h, w, p in shapes)
    flat = np.zeros(total, dtype=np.uint8)
    frames, desc, off = [], [], 0
    for h, w, p in shapes:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        frames.append(img)
        flat[off: off + h * p].reshape(h, p)[:, : w * 3] = img.reshape(h, w * 3)
        desc.append((off, h, w, p))
        off += h * p
    boxes, fidx = [], []
    for fi, (h, w, p) in enumerate(shapes):
        for b in _random_boxes(rng, 40, h, w, wmin=1) + [(0, 0, w, h), (w - 1, 0, w, h), (0, 0, min(w, 1500), h),
                                                         (max(0, w - 500), 0, w, h), (3, 3, 4, 4)]:
            boxes.append(b)
            fidx.append(fi)
    for kw in (dict(out_h=224, out_w=224), dict(out_h=64, out_w=160), dict(out_h=96, out_w=384)):
        plan = make_plan(T, **kw)
        args = (torch.from_numpy(flat).to(dev), torch.tensor(boxes, dtype=torch.int32, device=dev),
                torch.tensor(fidx, dtype=torch.int32, device=dev), plan)
        out = torch.full((len(boxes), 3, plan.out_h, plan.out_w), float("nan"), device=dev)
        ops.preprocess_crops(*args, out=out, frame_desc=torch.tensor(desc, dtype=torch.int64, device=dev))
        torch.cuda.synchronize()
        op = oracle_plan(plan)
        got = out.cpu().numpy()
        for fi in range(len(shapes)):
            sel = [i for i, f in enumerate(fidx) if f == fi]
            _, ef = preprocess_batch_c(frames[fi][None], [boxes[i] for i in sel], [0] * len(sel), op, want_u8=False)
            assert np.array_equal(got[sel].view(np.uint32), ef.view(np.uint32)), (kw, fi)


def test_k1_full_size_batch_properties(cuda_device):
    """BASELINE config 5 shape: 64 x 1080p frames, 64 boxes each -> 4096 crops of 224^2.
    Spot-check 96 crops against the oracle; determinism; identity-crop property on all frames."""
    from nkb_classification_b200 import transforms as T
    dev = cuda_device
    g = torch.Generator(device="cpu").manual_seed(1234)
    frames = torch.randint(0, 256, (64, 1080, 1920, 3), dtype=torch.uint8, generator=g)
    rng = np.random.default_rng(4321)
    boxes, fidx = [], []
    for f in range(64):
        for k in range(64):
            w, h = int(rng.integers(48, 481)), int(rng.integers(48, 481))
            x0, y0 = int(rng.integers(0, 1920 - w + 1)), int(rng.integers(0, 1080 - h + 1))
            if k == 0:
                w = h = 224  # identity crop in every frame
                x0, y0 = min(x0, 1920 - 224), min(y0, 1080 - 224)
            boxes.append((x0, y0, x0 + w, y0 + h))
            fidx.append(f)
    plan = make_plan(T)
    fr = frames.to(dev)
    out1, u8 = run_k1(dev, fr, boxes, fidx, plan)
    out2, _ = run_k1(dev, fr, boxes, fidx, plan, want_u8=False)
    assert torch.equal(out1, out2)
    pick = sorted(set(rng.integers(0, 4096, 96).tolist()) | {0, 4095})
    fn = frames.numpy()
    eu8, ef32 = preprocess_batch_c(fn, [boxes[i] for i in pick], [fidx[i] for i in pick], oracle_plan(plan))
    assert np.array_equal(u8[pick].cpu().numpy(), eu8)
    assert_same_f32(out1[pick], ef32)
    for f in range(0, 64, 7):
        x0, y0, x1, y1 = boxes[f * 64]
        assert np.array_equal(u8[f * 64].cpu().numpy(), fn[f][y0:y1, x0:x1])


@pytest.mark.parametrize("mode", ["stretch", "letterbox"])
def test_k1_full_size_letterbox_and_train_pipeline(cuda_device, mode):
    """BASELINE config 5 shape (4096 crops of 224^2 out of 64 x 1080p frames) through (a) the TMA kernel in the given
    geometry and (b) the reference's full train pipeline: 64 sampled crops of each against the oracle, bit for bit;
    size-independent properties on the whole batch: determinism, and `flags = 0` reproducing the val pipeline."""
    import random
    from nkb_classification_b200 import transforms as T
    dev = cuda_device
    g = torch.Generator(device="cpu").manual_seed(99)
    frames = torch.randint(0, 256, (64, 1080, 1920, 3), dtype=torch.uint8, generator=g)
    rng = np.random.default_rng(77)
    boxes, fidx = [], []
    for f in range(64):
        for _ in range(64):
            w, h = int(rng.integers(48, 481)), int(rng.integers(48, 481))
            x0, y0 = int(rng.integers(0, 1920 - w + 1)), int(rng.integers(0, 1080 - h + 1))
            boxes.append((x0, y0, x0 + w, y0 + h))
            fidx.append(f)
    fr, fn = frames.to(dev), frames.numpy()
    pick = sorted(set(rng.integers(0, 4096, 64).tolist()) | {0, 4095})
    val_plan = make_plan(T, mode=mode)
    out_val, _ = run_k1(dev, fr, boxes, fidx, val_plan, want_u8=False)            # TMA kernel, both geometries
    _, ev32 = preprocess_batch_c(fn, [boxes[i] for i in pick], [fidx[i] for i in pick], oracle_plan(val_plan))
    assert_same_f32(out_val[pick], ev32)
    geo = ([T.Resize(224, 224)] if mode == "stretch" else
           [T.LongestMaxSize(224), T.PadIfNeeded(224, 224, border_mode=0, value=0)])
    plan = T.compile_pipeline(geo + [
        T.HorizontalFlip(p=0.5), T.VerticalFlip(p=0.5),
        T.RandomBrightnessContrast(brightness_limit=(-0.2, 0.2), contrast_limit=(0.1, -0.5), p=0.5),
        T.HueSaturationValue(hue_shift_limit=0, sat_shift_limit=10, val_shift_limit=50, p=0.5),
        T.CoarseDropout(max_holes=4, min_holes=1, max_height=0.2, min_height=0.05, max_width=0.2, min_width=0.05,
                        fill_value=[0, 0.5, 1], p=0.5),
        T.Normalize(MEAN, STD), T.ToTensorV2()])
    batch = plan.draw(4096, random.Random(5))
    out1, _ = run_k1(dev, fr, boxes, fidx, plan, want_u8=False, aug=batch)
    out2, _ = run_k1(dev, fr, boxes, fidx, plan, want_u8=False, aug=batch)
    assert torch.equal(out1, out2)
    augs = _aug_samples(batch)
    _, ef32 = opre.preprocess_batch(fn, [boxes[i] for i in pick], [fidx[i] for i in pick], oracle_plan(plan),
                                    impl="cv2", augs=[augs[i] for i in pick])
    assert_same_f32(out1[pick], ef32)
    ident = plan.draw(4096, random.Random(6))
    ident.flags[:] = 0
    out0, _ = run_k1(dev, fr, boxes, fidx, plan, want_u8=False, aug=ident)
    assert torch.equal(out0, out_val)


def test_k1_empty_batch(cuda_device):
    from nkb_classification_b200 import ops, transforms as T
    plan = make_plan(T, out_h=16, out_w=16)
    out = ops.preprocess_crops(torch.zeros((1, 8, 8, 3), dtype=torch.uint8, device=cuda_device),
                               torch.zeros((0, 4), dtype=torch.int32, device=cuda_device),
                               torch.zeros((0,), dtype=torch.int32, device=cuda_device), plan)
    assert tuple(out.shape) == (0, 3, 16, 16)


@pytest.mark.parametrize("name,hw,kw", [
    ("cfg1 singletask: 224x224 ImageFolder, Resize(224) = identity + Normalize", (224, 224), dict(out_h=224, out_w=224)),
    ("cfg2 multitask: 256x256 -> Resize(224,224)", (256, 256), dict(out_h=224, out_w=224)),
    ("cfg2 multitask sample config: img_size=128 letterbox", (256, 256), dict(mode="letterbox", out_h=128, out_w=128)),
    ("non-square whole images, letterbox 224", (200, 333), dict(mode="letterbox", out_h=224, out_w=224)),
])
def test_k1_whole_image_configs(cuda_device, name, hw, kw):
    """BASELINE configs 1 and 2: the box is the whole image (AnnotatedSingletask/MultitaskDataset, ImageFolder)."""
    from nkb_classification_b200 import transforms as T
    rng = np.random.default_rng(31)
    H, W = hw
    frames = rng.integers(0, 256, (32, H, W, 3), dtype=np.uint8)
    boxes = [(0, 0, W, H)] * 32
    fidx = list(range(32))
    plan = make_plan(T, **kw)
    out, u8 = run_k1(cuda_device, frames, boxes, fidx, plan)
    eu8, ef32 = preprocess_batch_c(frames, boxes, fidx, oracle_plan(plan))
    assert np.array_equal(u8.cpu().numpy(), eu8), name
    assert_same_f32(out, ef32)
    out_fast, _ = run_k1(cuda_device, frames, boxes, fidx, plan, want_u8=False)
    assert_same_f32(out_fast, ef32)
    if hw == (224, 224):
        assert np.array_equal(u8.cpu().numpy(), frames)  # identity resize


@pytest.mark.parametrize("mode", ["stretch", "letterbox"])
def test_k1_fast_path_last_column_boxes(cuda_device, mode):
    """Boxes that start on the last column of a frame whose row offset is 16-byte aligned there (the 2-tap window is
    shifted left of the staged span): the TMA kernel must hand them to the direct-load routine.  Also 2-pixel boxes on
    the edge, in both geometries, through the fast path (no uint8 side output)."""
    from nkb_classification_b200 import transforms as T
    rng = np.random.default_rng(23)
    W = 17 + 16 * 3            # 65: column 64 -> byte offset 192, a multiple of 16
    frames = rng.integers(0, 256, (2, 40, W, 3), dtype=np.uint8)
    boxes = [(W - 1, 0, W, 40), (W - 1, 5, W, 6), (W - 2, 0, W, 40), (16, 3, 17, 30), (0, 0, W, 40), (48, 10, 49, 35)]
    fidx = [0, 1, 0, 1, 0, 1]
    plan = make_plan(T, mode=mode, out_h=64, out_w=64, pad=(9, 8, 7))
    # pitch must be 16-byte aligned for the TMA path: pack the frames with a padded pitch
    pitch = (W * 3 + 15) // 16 * 16
    flat = np.zeros((2, 40, pitch), dtype=np.uint8)
    flat[:, :, : W * 3] = frames.reshape(2, 40, W * 3)
    desc = torch.tensor([[f * 40 * pitch, 40, W, pitch] for f in range(2)], dtype=torch.int64, device=cuda_device)
    out, _ = run_k1(cuda_device, torch.from_numpy(flat.reshape(-1)).to(cuda_device), boxes, fidx, plan, want_u8=False,
                    frame_desc=desc)
    _, ef32 = preprocess_batch_c(frames, boxes, fidx, oracle_plan(plan))
    assert_same_f32(out, ef32)


def _aug_samples(batch):
    """transforms.AugmentBatch -> the oracle's per-sample parameter objects."""
    out = []
    for i in range(len(batch.flags)):
        f = int(batch.flags[i])
        k = f >> 8
        hsv = tuple(float(v) for v in batch.hsv_shift[i]) if (f & 8) else None
        out.append(opre.AugSample(hflip=bool(f & 1), vflip=bool(f & 2), bc=bool(f & 4), alpha=float(batch.alpha[i]),
                                  beta=float(batch.brightness[i]), holes=[tuple(int(v) for v in h) for h in batch.holes[i, :k]],
                                  fill=batch.fill, hsv=hsv))
    return out


@pytest.mark.parametrize("kw", [
    dict(mode="letterbox", out_h=128, out_w=128),                  # the reference train config (img_size 128)
    dict(mode="letterbox", out_h=96, out_w=160, max_size=90, pad=(3, 50, 200), swap=True),
    dict(out_h=224, out_w=224),                                    # A.Resize variant, full column tiles
    dict(out_h=100, out_w=60),                                     # partial column tile + flips
])
def test_k1_train_pipeline_augmentations_vs_oracle(cuda_device, kw):
    """flips + brightness/contrast LUT + HueSaturationValue + CoarseDropout (the whole train pipeline of
    configs/singletask_config.py:162-201) fused into K1 with GIVEN per-sample parameters == cv2.flip / cv2.LUT /
    cv2.cvtColor round trip / slice-assign on the oracle's resized image: uint8 bit-exact, fp32 bit-exact, bf16 = RNE."""
    import random
    from nkb_classification_b200 import transforms as T
    rng = np.random.default_rng(17)
    frames = rng.integers(0, 256, (3, 270, 480, 3), dtype=np.uint8)
    boxes = _random_boxes(rng, 120, 270, 480, wmin=6)
    if kw.get("mode") == "letterbox":
        ms = kw.get("max_size") or min(kw["out_h"], kw["out_w"])
        boxes = [b for b in boxes if min(opre.letterbox_geometry(b[3] - b[1], b[2] - b[0], ms, kw["out_h"], kw["out_w"])[:2]) >= 1]
    boxes += [(0, 0, 480, 270), (0, 0, 7, 270), (0, 260, 480, 270)]
    fidx = [i % 3 for i in range(len(boxes))]
    base = make_plan(T, **kw)
    geo = ([T.Resize(base.out_h, base.out_w)] if base.mode == T.MODE_STRETCH else
           [T.LongestMaxSize(base.max_size), T.PadIfNeeded(base.out_h, base.out_w, border_mode=0, value=base.pad_value)])
    plan = T.compile_pipeline(geo + [
        T.HorizontalFlip(p=0.5), T.VerticalFlip(p=0.5),
        T.RandomBrightnessContrast(brightness_limit=(-0.2, 0.2), contrast_limit=(0.1, -0.5), p=0.6),
        T.HueSaturationValue(hue_shift_limit=(0 if kw.get("mode") == "letterbox" else 20), sat_shift_limit=10,
                             val_shift_limit=50, p=0.6),
        T.CoarseDropout(max_holes=4, min_holes=1, max_height=0.2, min_height=0.05, max_width=0.2, min_width=0.05,
                        fill_value=[0, 0.5, 1], p=0.6),
        T.Normalize(MEAN, STD), T.ToTensorV2()], channel_swap=base.channel_swap)
    batch = plan.draw(len(boxes), random.Random(99))
    assert len({int(f) & 15 for f in batch.flags}) == 16, "the draw should cover every flip / LUT / HSV combination"
    out, u8 = run_k1(cuda_device, frames, boxes, fidx, plan, aug=batch)
    eu8, ef32 = opre.preprocess_batch(frames, boxes, fidx, oracle_plan(plan), impl="cv2", augs=_aug_samples(batch))
    assert np.array_equal(u8.cpu().numpy(), eu8)
    assert_same_f32(out, ef32)
    outb, _ = run_k1(cuda_device, frames, boxes, fidx, plan, dtype=torch.bfloat16, want_u8=False, aug=batch)
    assert np.array_equal(outb.view(torch.int16).cpu().numpy().view(np.uint16), opre.f32_to_bf16_bits(ef32).reshape(ef32.shape))
    # identity parameters reproduce the val pipeline exactly
    ident = plan.draw(len(boxes), random.Random(1))
    ident.flags[:] = 0
    out0, u80 = run_k1(cuda_device, frames, boxes, fidx, plan, aug=ident)
    ev8, ev32 = preprocess_batch_c(frames, boxes, fidx, oracle_plan(plan))
    assert np.array_equal(u80.cpu().numpy(), ev8)
    assert_same_f32(out0, ev32)
    # the plan refuses to run without its parameters (and a val plan refuses stray ones)
    with pytest.raises(ValueError):
        run_k1(cuda_device, frames, boxes, fidx, plan)
    with pytest.raises(ValueError):
        run_k1(cuda_device, frames, boxes, fidx, base, aug=batch)


def test_hsv_luts_built_on_the_device_equal_albumentations_tables(cuda_device):
    """nkbk_build_hsv_luts: the hue / sat / val tables from the per-sample shifts on the device == transforms.hsv_luts
    (albumentations' `_shift_hsv_uint8` tables: int16 ramp + float64 shift, mod 180 / clip, truncated), bit for bit;
    hue shifts of both signs, zero shifts (identity), samples without the HueSaturationValue flag (identity)."""
    import ctypes
    from nkb_classification_b200 import _lib, transforms as T
    rng = np.random.default_rng(5)
    n = 600
    sh = np.stack([rng.uniform(-200, 200, n), rng.uniform(-300, 300, n), rng.uniform(-300, 300, n)], 1)
    sh[::7, 0] = 0.0
    sh[::5, 1] = 0.0
    sh[::11] = np.round(sh[::11])            # integer shifts: exact ties of the mod / clip
    flags = np.where(rng.random(n) < 0.8, 8, 0).astype(np.int32) | rng.integers(0, 8, n).astype(np.int32)
    sd, fd = torch.from_numpy(sh).to(cuda_device), torch.from_numpy(flags).to(cuda_device)
    out = torch.empty((n, 3, 256), dtype=torch.uint8, device=cuda_device)
    _lib.check(_lib.lib().nkbk_build_hsv_luts(ctypes.c_void_p(sd.data_ptr()), ctypes.c_void_p(fd.data_ptr()), n,
                                             ctypes.c_void_p(out.data_ptr()),
                                             ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    got = out.cpu().numpy()
    ident = np.tile(np.arange(256, dtype=np.uint8), (3, 1))
    for i in range(n):
        exp = T.hsv_luts(*sh[i]) if flags[i] & 8 else ident
        assert np.array_equal(got[i], exp), (i, sh[i], flags[i])


def test_k1_train_pipeline_with_vectorised_draw_and_device_tables(cuda_device):
    """`plan.draw(n)` (all samples at once, tables left to the device) feeds K1 the same way the per-sample draw does:
    bit-exact against the oracle evaluated on the drawn parameters, through the TMA-ring kernel (no uint8 side output)
    and the direct-load kernel (with it)."""
    import random
    from nkb_classification_b200 import transforms as T
    rng = np.random.default_rng(29)
    frames = rng.integers(0, 256, (3, 270, 480, 3), dtype=np.uint8)
    boxes = _random_boxes(rng, 150, 270, 480, wmin=6) + [(0, 0, 480, 270)]
    fidx = [i % 3 for i in range(len(boxes))]
    base = make_plan(T, out_h=224, out_w=224)
    plan = T.compile_pipeline([T.Resize(224, 224), T.HorizontalFlip(p=0.5), T.VerticalFlip(p=0.5),
                               T.RandomBrightnessContrast(brightness_limit=(-0.2, 0.2), contrast_limit=(0.1, -0.5), p=0.5),
                               T.HueSaturationValue(hue_shift_limit=15, sat_shift_limit=10, val_shift_limit=50, p=0.5),
                               T.CoarseDropout(max_holes=4, min_holes=1, max_height=0.2, min_height=0.05, max_width=0.2,
                                               min_width=0.05, fill_value=[0, 0.5, 1], p=0.5),
                               T.Normalize(MEAN, STD), T.ToTensorV2()], channel_swap=base.channel_swap)
    random.seed(4)
    batch = plan.draw(len(boxes))
    assert batch.hsv_lut is None and batch.hsv_shift is not None          # tables are the device's job
    random.seed(4)
    again = plan.draw(len(boxes))
    assert np.array_equal(batch.flags, again.flags) and np.array_equal(batch.holes, again.holes)   # random.seed governs
    eu8, ef32 = opre.preprocess_batch(frames, boxes, fidx, oracle_plan(plan), impl="cv2", augs=_aug_samples(batch))
    out, u8 = run_k1(cuda_device, frames, boxes, fidx, plan, aug=batch)
    assert np.array_equal(u8.cpu().numpy(), eu8)
    assert_same_f32(out, ef32)
    out_fast, _ = run_k1(cuda_device, frames, boxes, fidx, plan, want_u8=False, aug=batch)
    assert_same_f32(out_fast, ef32)
