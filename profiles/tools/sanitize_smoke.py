"""Small pass over every kernel family for `compute-sanitizer --tool memcheck` (one tool per gpurun call)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from nkb_classification_b200 import hotpath, ops, transforms as T  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
# frames whose LAST row ends exactly at the end of the allocation: exercises the aligned-word / bulk-copy edge reads
H, W = 96, 160
frames = torch.from_numpy(rng.integers(0, 256, (3, H, W, 3), dtype=np.uint8)).to(dev)
boxes = [(0, 0, W, H), (W - 7, H - 9, W, H), (W - 1, 0, W, H), (0, H - 1, W, H), (5, 5, 60, 70), (100, 3, 159, 95)]
fidx = [2, 2, 2, 2, 0, 1]
bx = torch.tensor(boxes, dtype=torch.int32, device=dev)
fi = torch.tensor(fidx, dtype=torch.int32, device=dev)
norm = [T.Normalize(), T.ToTensorV2()]
for pipe, u8 in (([T.Resize(224, 224)], False), ([T.Resize(224, 224)], True), ([T.Resize(50, 70)], False),
                 ([T.LongestMaxSize(64), T.PadIfNeeded(64, 64, border_mode=0, value=3)], True)):
    plan = T.compile_pipeline(pipe + norm)
    out_u8 = torch.empty((len(fidx), plan.out_h, plan.out_w, 3), dtype=torch.uint8, device=dev) if u8 else None
    ops.preprocess_crops(frames, bx, fi, plan, out_u8=out_u8)
    ops.preprocess_crops(frames, bx, fi, plan, out_dtype=torch.bfloat16)
# train pipeline: augmentations fused into the TMA ring kernel (full column tiles) and into the direct-load kernel (partial
# tiles / uint8 side output); hue / sat / val tables built on the device; K1 as a programmatic dependent
import random  # noqa: E402
for pipe, u8 in (([T.Resize(224, 224)], False), ([T.Resize(50, 70)], True),
                 ([T.LongestMaxSize(64), T.PadIfNeeded(64, 64, border_mode=0, value=3)], False)):
    plan = T.compile_pipeline(pipe + [T.HorizontalFlip(p=0.5), T.VerticalFlip(p=0.5),
                                      T.RandomBrightnessContrast(brightness_limit=(-0.2, 0.2), contrast_limit=(0.1, -0.5), p=0.7),
                                      T.HueSaturationValue(hue_shift_limit=15, sat_shift_limit=10, val_shift_limit=50, p=0.7),
                                      T.CoarseDropout(max_holes=4, min_holes=1, max_height=0.2, min_height=0.05,
                                                      max_width=0.2, min_width=0.05, fill_value=[0, 0.5, 1], p=0.7)] + norm)
    out_u8 = torch.empty((len(fidx), plan.out_h, plan.out_w, 3), dtype=torch.uint8, device=dev) if u8 else None
    ops.preprocess_crops(frames, bx, fi, plan, out_u8=out_u8, aug=plan.draw(len(fidx)))                      # device tables
    ops.preprocess_crops(frames, bx, fi, plan, out_dtype=torch.bfloat16, aug=plan.draw(len(fidx), random.Random(3)))
# odd pitch -> fix-up pass
flat = torch.from_numpy(rng.integers(0, 256, 3 + 40 * 101, dtype=np.uint8)).to(dev)
desc = torch.tensor([[3, 40, 33, 101]], dtype=torch.int64, device=dev)
ops.preprocess_crops(flat, torch.tensor([[0, 0, 33, 40], [30, 35, 33, 40]], dtype=torch.int32, device=dev),
                     torch.zeros(2, dtype=torch.int32, device=dev), T.compile_pipeline([T.Resize(64, 64)] + norm),
                     frame_desc=desc)
g = torch.Generator().manual_seed(0)
for B, D, classes, dt in ((37, 64, (4, 7, 2), torch.float32), (300, 128, (40, 20), torch.bfloat16),
                          (129, 256, (10,), torch.bfloat16), (5, 260, (3, 70), torch.float32)):
    hp = hotpath.HotPath(T.compile_pipeline([T.Resize(32, 32)] + norm), classes, D, "FocalLoss", 1.0, device=dev)
    emb = torch.randn(B, D, generator=g).to(dev).to(dt)
    Wc = (torch.randn(sum(classes), D, generator=g) * 0.1).to(dev)
    b = torch.zeros(sum(classes), device=dev)
    labels = torch.stack([torch.randint(0, c, (B,), generator=g) for c in classes], 1).contiguous().to(dev)
    bufs = hp.heads_step(emb, Wc, b, labels, train=True, want_probs=True)
    ops.heads_demb(bufs, Wc, out_dtype=dt)
    hp.heads_step(emb, Wc, b, labels, train=False)
    ops.loss_fwd_bwd(bufs.logits, hp.seg, labels, 1, 2.0)
# forward-only heads with the warp-aggregated confusion counts (fp32 FFMA forward, tcgen05 forward), K5
for dt in (torch.float32, torch.bfloat16):
    B, D, classes = 300, 128, (10, 4)
    seg = [0, 10, 14]
    emb = torch.randn(B, D, generator=g).to(dev).to(dt)
    Wc = (torch.randn(14, D, generator=g) * 0.1).to(dev)
    labels = torch.stack([torch.randint(0, c, (B,), generator=g) for c in classes], 1).contiguous().to(dev)
    bufs = ops.HeadsBuffers(B, D, seg, dev, want_grads=False)
    cs = torch.zeros(ops.confusion_len(seg), dtype=torch.int64, device=dev)
    pred = torch.empty((B, 2), dtype=torch.int32, device=dev)
    ops.heads_fwd_loss_bwd(emb, Wc, torch.zeros(14, device=dev), labels, bufs, 0, 0.0, out_pred=pred, cm_step=cs)
    ops.roc_auc_counts(bufs.probs, seg, labels)
# K1 of the next batch as a programmatic dependent of the fused heads step
hp = hotpath.HotPath(T.compile_pipeline([T.Resize(224, 224)] + norm), (10,), 256, "CrossEntropyLoss", 0.0, device=dev)
emb = torch.randn(64, 256, generator=g).to(dev)
Wc, b = (torch.randn(10, 256, generator=g) * 0.1).to(dev), torch.zeros(10, device=dev)
labels = torch.randint(0, 10, (64, 1), generator=g).to(dev)
for _ in range(3):
    hp.heads_step(emb, Wc, b, labels, train=True)
    hp.mark_heads_done()
    hp.preprocess(frames, bx, fi, overlap_previous=True)
torch.cuda.synchronize()
print("SANITIZE_SMOKE_DONE")
