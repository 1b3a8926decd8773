"""The reference's CPU path for the hot path, restated with the libraries the
image has, timed as the ``cpu_baseline`` / ``--impl reference`` arm of bench.py.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/__init__.py).

What one "step" does, following the reference's sequence:
  * P worker processes (the reference's DataLoader workers, dataset.py:612-628;
    ``cv2.setNumThreads(0)`` as albumentations sets at import) each take a slice
    of the batch and, per crop: ``np.array(frame[y0:y1, x0:x1])`` (dataset.py:404,
    :102) -> ``cv2.resize(INTER_LINEAR)`` -> fp32 Normalize (cv2.subtract /
    cv2.multiply form) -> CHW transpose, stacked into the batch tensor in shared
    memory (default_collate inside the worker + shared-memory hand-off);
  * the main process runs the heads (``F.linear`` per task, model.py:114-116),
    the criterion (losses.py semantics via oracle.heads), backward, and the
    logger's per-iteration softmax / argmax / ``.tolist()`` (logging.py:261-281),
    fp32, ``torch.set_num_threads(P)``.
JPEG decode and the backbone are excluded on both sides (synthetic in-memory
RGB frames; the backbone is out of scope for this path).
cv2 + numpy stand in for albumentations, which is not installed in this image.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time
from typing import Optional, Sequence

import numpy as np

_G = {}


def _worker_init(frames_shape, frames_raw, out_shape, out_raw, mean255, denom, out_hw):
    import cv2

    cv2.setNumThreads(0)
    _G["frames"] = np.frombuffer(frames_raw, dtype=np.uint8).reshape(frames_shape)
    _G["out"] = np.frombuffer(out_raw, dtype=np.float32).reshape(out_shape)
    _G["m"], _G["d"], _G["hw"] = mean255, denom, out_hw


def _worker_chunk(args):
    import cv2

    lo, hi, boxes, fidx = args
    frames, out = _G["frames"], _G["out"]
    oh, ow = _G["hw"]
    m = np.broadcast_to(_G["m"], (oh, ow, 3)).astype(np.float32).copy()
    d = np.broadcast_to(_G["d"], (oh, ow, 3)).astype(np.float32).copy()
    for i in range(lo, hi):
        x0, y0, x1, y1 = boxes[i]
        crop = np.array(frames[fidx[i]][y0:y1, x0:x1])                                   # dataset.py:404, :102
        img = crop if crop.shape[:2] == (oh, ow) else cv2.resize(crop, dsize=(ow, oh), interpolation=cv2.INTER_LINEAR)
        f = img.astype(np.float32)
        f = cv2.subtract(f, m)
        f = cv2.multiply(f, d)
        out[i] = f.transpose(2, 0, 1)                                                     # ToTensorV2 + collate
    return hi - lo


class CpuReferencePath:
    """Persistent worker pool + shared buffers so pool start-up is outside the timed region."""

    def __init__(self, frames: np.ndarray, out_hw=(224, 224), mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225),
                 procs: Optional[int] = None, max_crops: int = 4096):
        from .preprocess import normalize_constants

        self.procs = procs or os.cpu_count() or 1
        self.out_hw = tuple(out_hw)
        self.max_crops = max_crops
        m, d = normalize_constants(mean, std)
        ctx = mp.get_context("fork")
        self._frames_raw = ctx.RawArray("B", int(frames.size))
        np.frombuffer(self._frames_raw, dtype=np.uint8)[:] = frames.reshape(-1)
        self.frames_shape = frames.shape
        self.out_shape = (max_crops, 3, out_hw[0], out_hw[1])
        self._out_raw = ctx.RawArray("f", int(np.prod(self.out_shape)))
        self.out = np.frombuffer(self._out_raw, dtype=np.float32).reshape(self.out_shape)
        self.pool = ctx.Pool(self.procs, initializer=_worker_init,
                             initargs=(self.frames_shape, self._frames_raw, self.out_shape, self._out_raw, m, d,
                                       self.out_hw)) if self.procs > 1 else None
        if self.pool is None:
            _worker_init(self.frames_shape, self._frames_raw, self.out_shape, self._out_raw, m, d, self.out_hw)

    def preprocess(self, boxes: np.ndarray, fidx: np.ndarray) -> np.ndarray:
        n = len(fidx)
        assert n <= self.max_crops
        if self.pool is None:
            _worker_chunk((0, n, boxes, fidx))
            return self.out[:n]
        per = max(1, (n + self.procs * 4 - 1) // (self.procs * 4))
        tasks = [(lo, min(lo + per, n), boxes, fidx) for lo in range(0, n, per)]
        done = sum(self.pool.map(_worker_chunk, tasks))
        assert done == n
        return self.out[:n]

    def heads_loss_metrics(self, emb, Ws, bs, labels, loss_type: str, gamma: float, train: bool = True):
        """Main-process part: heads + criterion (+ backward) + logger-style stats, torch CPU fp32."""
        import torch
        import torch.nn.functional as F

        from . import heads as oh

        torch.set_num_threads(self.procs)
        emb = emb.detach().clone().requires_grad_(False)
        Wp = [w.detach().clone().requires_grad_(train) for w in Ws]
        bp = [b.detach().clone().requires_grad_(train) for b in bs]
        preds = [F.linear(emb, w, b) for w, b in zip(Wp, bp)]
        losses = []
        for t, z in enumerate(preds):
            y = labels[:, t]
            losses.append(oh.focal_loss(z, y, None, gamma) if loss_type == "FocalLoss" else oh.cross_entropy(z, y))
        total = sum(losses)
        if train:
            total.backward()
        stats = []
        for t, z in enumerate(preds):   # logging.py:268-281
            gt = labels[:, t].cpu().numpy().tolist()
            conf = z.softmax(dim=-1, dtype=torch.float32).detach().cpu().numpy().tolist()
            pr = z.argmax(dim=-1).detach().cpu().numpy().tolist()
            stats.append((gt, conf, pr, losses[t].item()))
        return float(total.detach()), stats

    def step(self, boxes, fidx, emb, Ws, bs, labels, loss_type, gamma, train=True) -> float:
        import torch

        t0 = time.perf_counter()
        batch = self.preprocess(boxes, fidx)
        img = torch.from_numpy(batch)          # the collated [B,3,H,W] fp32 batch the engine would receive
        assert img.shape[0] == len(fidx)
        self.heads_loss_metrics(emb[: len(fidx)], Ws, bs, labels[: len(fidx)], loss_type, gamma, train)
        return time.perf_counter() - t0

    def close(self):
        if self.pool is not None:
            self.pool.terminate()
            self.pool.join()
            self.pool = None
