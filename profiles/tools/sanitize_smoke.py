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
torch.cuda.synchronize()
print("SANITIZE_SMOKE_DONE")
