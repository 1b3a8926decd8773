"""Synthetic inputs of the BASELINE.json configurations.

Frames are iid uint8 noise (seeded per rank), boxes are drawn in pixels,
written as YOLO ``cls xc yc w h`` text with 6 decimals and parsed back through
the a1 box path so truncation / clip / min-size semantics are exercised.
Used by bench.py, __graft_entry__.smoke() and the full-size tests.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

from .boxes import parse_yolo_label_lines


@dataclass(frozen=True)
class Workload:
    name: str
    frames: int            # frames per GPU per step
    frame_h: int
    frame_w: int
    boxes_per_frame: int
    out_size: int
    mode: str              # "stretch" | "letterbox"
    emb_dim: int
    classes: Tuple[int, ...]
    loss: str              # "FocalLoss" | "CrossEntropyLoss"
    gamma: float
    whole_image: bool = False   # box = full frame (configs 1, 2)
    train_aug: bool = False     # K1 with the reference's train-time augmentations fused (bench --train-aug)

    @property
    def crops(self) -> int:
        return self.frames * self.boxes_per_frame


WORKLOADS = {
    # BASELINE.json configs[4]: inference.py sweep shape, the crops/s headline: 64 x 1080p frames x 64 boxes
    # -> 4096 crops of 224^2 per GPU per step, resnet50-width single head, + loss/metric for the training leg.
    "cfg5_1080p_64x64": Workload("cfg5_1080p_64x64", 64, 1080, 1920, 64, 224, "stretch", 2048, (10,),
                                 "CrossEntropyLoss", 0.0),
    # configs[2]: yolo_dataset_config shape, ~20 boxes per 1080p frame, convnext_tiny width
    "cfg3_1080p_20": Workload("cfg3_1080p_20", 64, 1080, 1920, 20, 224, "stretch", 768, (3,), "CrossEntropyLoss", 0.0),
    # configs[1]: multitask_config, 256 x 256 whole images, 3 heads (4/7/2), focal gamma 1, batch 256
    "cfg2_multitask_256": Workload("cfg2_multitask_256", 256, 256, 256, 1, 224, "stretch", 1280, (4, 7, 2),
                                   "FocalLoss", 1.0, True),
    # configs[3]: vit_base multitask 5 heads, batch 1024 (heads only matter; crops are 224 whole images)
    "cfg4_vit_5heads": Workload("cfg4_vit_5heads", 1024, 224, 224, 1, 224, "stretch", 768, (2, 3, 4, 7, 14),
                                "FocalLoss", 1.0, True),
    # configs[0]: singletask resnet18, 224 x 224 ImageFolder, eval batch 32
    "cfg1_single_224": Workload("cfg1_single_224", 32, 224, 224, 1, 224, "stretch", 512, (10,), "CrossEntropyLoss",
                                0.0, True),
}
# one rank's share of configs[4] when its global batch of 4096 is split over 8 / 4 / 2 GPUs (strong scaling): lets the
# per-rank step of an N-GPU run be studied on one GPU (no exchange)
for _n in (2, 4, 8):
    WORKLOADS[f"cfg5_shard{_n}"] = Workload(f"cfg5_shard{_n}", 64 // _n, 1080, 1920, 64, 224, "stretch", 2048, (10,),
                                            "CrossEntropyLoss", 0.0)
DEFAULT_WORKLOAD = "cfg5_1080p_64x64"


def synth_boxes(wl: Workload, seed: int = 4321) -> Tuple[np.ndarray, np.ndarray]:
    """int32 [n,4] boxes and int32 [n] frame indices for one step on one GPU."""
    rng = np.random.default_rng(seed)
    H, W = wl.frame_h, wl.frame_w
    boxes, fidx = [], []
    for f in range(wl.frames):
        if wl.whole_image:
            boxes.append((0, 0, W, H))
            fidx.append(f)
            continue
        lines = []
        for _ in range(wl.boxes_per_frame):
            w, h = int(rng.integers(48, 481)), int(rng.integers(48, 481))
            w, h = min(w, W), min(h, H)
            x0, y0 = int(rng.integers(0, W - w + 1)), int(rng.integers(0, H - h + 1))
            lines.append(f"0 {(x0 + w / 2) / W:.6f} {(y0 + h / 2) / H:.6f} {w / W:.6f} {h / H:.6f}")
        for box, _ in parse_yolo_label_lines(lines, (H, W)):
            boxes.append(box)
            fidx.append(f)
    return np.asarray(boxes, dtype=np.int32).reshape(-1, 4), np.asarray(fidx, dtype=np.int32)


def k1_algorithmic_bytes(boxes: np.ndarray, out_h: int, out_w: int, out_elem_bytes: int, mode: str = "stretch",
                         max_size: int = 0) -> int:
    """per crop 3*min(w,2*dw)*min(h,2*dh) source bytes a 2-tap filter can touch (dw x dh = the
    resized size: the whole output for A.Resize, the letterboxed extent for LongestMaxSize + PadIfNeeded)
    + 3*out_h*out_w*sizeof(out) written + 16 bytes of box."""
    b = np.asarray(boxes, dtype=np.int64).reshape(-1, 4)
    bw, bh = b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]
    if mode == "letterbox":
        scale = (max_size or min(out_h, out_w)) / np.maximum(bw, bh).astype(np.float64)
        dw = np.where(scale != 1.0, np.rint(bw * scale), bw).astype(np.int64)
        dh = np.where(scale != 1.0, np.rint(bh * scale), bh).astype(np.int64)
    else:
        dw, dh = np.full_like(bw, out_w), np.full_like(bh, out_h)
    w = np.minimum(bw, 2 * dw)
    h = np.minimum(bh, 2 * dh)
    return int((3 * w * h).sum() + b.shape[0] * (3 * out_h * out_w * out_elem_bytes + 16))


def k2_algorithmic_bytes(B: int, D: int, NC: int, T: int, emb_elem_bytes: int, with_grads: bool = True) -> int:
    """Read emb once (+ once more for dW), labels, W; write logits/probs/dlogits and the reduce buffer."""
    rd = B * D * emb_elem_bytes * (2 if with_grads else 1) + B * T * 8 + NC * D * 4
    wr = B * NC * 4 * (3 if with_grads else 2) + (NC * D + NC + 2 * T) * 4
    return int(rd + wr)
