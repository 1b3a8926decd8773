"""The fused hot path as one object: K1 -> (backbone, external) -> K2 -> K3 -> K4.

This is the call a user (or engine.py's train/val loop) makes per batch; it
owns the reusable device buffers so a steady-state step allocates nothing.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib, ops
from ._lib import LOSS_CE, LOSS_FOCAL
from .parallel import Communicator
from .transforms import PreprocessPlan

LOSS_KINDS = {"CrossEntropyLoss": LOSS_CE, "FocalLoss": LOSS_FOCAL}


class HotPath:
    def __init__(self, plan: PreprocessPlan, classes_per_task: Sequence[int], emb_dim: int, loss_type: str = "FocalLoss",
                 gamma: float = 2.0, class_weight: Optional[torch.Tensor] = None, ignore_index: int = -100,
                 device="cuda:0", comm: Optional[Communicator] = None, out_dtype: torch.dtype = torch.float32,
                 transport: str = "peer"):
        if loss_type not in LOSS_KINDS:
            raise NotImplementedError(f"Unknown loss type in config: {loss_type}")  # losses.py:171
        self.plan = plan
        self.device = torch.device(device)
        self.seg = np.concatenate([[0], np.cumsum(list(classes_per_task))]).astype(int).tolist()
        self.T, self.NC, self.D = len(classes_per_task), self.seg[-1], int(emb_dim)
        self.loss_kind, self.gamma, self.ignore_index = LOSS_KINDS[loss_type], float(gamma), int(ignore_index)
        self.class_weight = None if class_weight is None else class_weight.to(self.device, torch.float32).contiguous()
        self.comm = comm or Communicator()
        if transport not in ("peer", "nccl"):
            raise ValueError(f"transport must be 'peer' or 'nccl', got {transport!r}")
        self.transport = transport   # exchange step: K4' (NVLink peer memory, fused finalize) or K4 (NCCL)
        self.out_dtype = out_dtype
        self.cm = torch.zeros(ops.confusion_len(self.seg), dtype=torch.int64, device=self.device)       # epoch totals
        self.cm_step = torch.zeros_like(self.cm)   # this step's counts: K3 -> all-reduce -> folded into cm by finalize
        self._bufs: Dict[tuple, ops.HeadsBuffers] = {}
        self._images: Dict[int, torch.Tensor] = {}
        self._pred: Dict[int, torch.Tensor] = {}
        self._desc = None
        self._heads_done: Optional[torch.cuda.Event] = None   # set by mark_heads_done() in overlapped loops
        # False: the heads step as its separate kernels (nkbk_heads_one_launch(0)) -- the better partner for a K1 that
        # runs on ANOTHER stream at a full batch per GPU: small CTAs slot in between K1's, and at world > 1 the wait for
        # the other ranks does not hold SMs
        self.one_launch = True

    def mark_heads_done(self):
        """Call right after heads_step in a loop that launches K1 with ``overlap_previous=True``."""
        if self._heads_done is None:
            self._heads_done = torch.cuda.Event()
        self._heads_done.record(torch.cuda.current_stream(self.device))

    # ---- K1 ----
    def preprocess(self, frames: torch.Tensor, boxes: torch.Tensor, frame_idx: torch.Tensor,
                   frame_desc: Optional[torch.Tensor] = None, aug=None, overlap_previous: bool = False) -> torch.Tensor:
        """``overlap_previous``: this K1 (batch i+1) may start while the fused heads step enqueued just before it on the
        same stream (batch i) is still running, on the SMs that step leaves free (nkbk_k1_overlap_previous: a
        programmatic dependent launch).  Safe here because this object owns K1's output buffer and the heads buffers."""
        n = int(frame_idx.numel())
        out = self._images.get(n)
        if out is None:
            out = torch.empty((n, 3, self.plan.out_h, self.plan.out_w), dtype=self.out_dtype, device=self.device)
            self._images = {n: out}
        if frame_desc is None:
            key = (tuple(frames.shape), frames.data_ptr())
            if self._desc is None or self._desc[0] != key[0]:
                self._desc = (key[0], ops.frame_descriptors(frames))
            frame_desc = self._desc[1]
        if aug is None and self.plan.augment is not None:
            aug = self.plan.draw(n)            # train-time pipeline: draw this batch's per-sample parameters
        if not overlap_previous:
            return ops.preprocess_crops(frames, boxes, frame_idx, self.plan, self.out_dtype, out=out,
                                        frame_desc=frame_desc, aug=aug)
        prev = _lib.lib().nkbk_k1_overlap_previous(1)
        try:
            return ops.preprocess_crops(frames, boxes, frame_idx, self.plan, self.out_dtype, out=out,
                                        frame_desc=frame_desc, aug=aug)
        finally:
            _lib.lib().nkbk_k1_overlap_previous(prev)

    # ---- K2 + K3 + K4 ----
    def _buffers(self, B: int, train: bool, want_probs: bool) -> ops.HeadsBuffers:
        key = (B, train, want_probs)
        b = self._bufs.get(key)
        if b is None:
            b = ops.HeadsBuffers(B, self.D, self.seg, self.device, want_logits=True, want_probs=want_probs,
                                 want_grads=train)
            self._bufs[key] = b
        return b

    def heads_step(self, emb: torch.Tensor, W_cat: torch.Tensor, b_cat: torch.Tensor,
                   labels: Optional[torch.Tensor], train: bool = True, want_probs: bool = False,
                   want_pred: bool = True, update_confusion: bool = True):
        """One step of K2 + K3 + K4.  Returns the HeadsBuffers holding (after finalize) the global-mean
        dW/db, logits, probs, ``.loss`` [T+1] and ``.pred`` [B,T]; the epoch confusion totals are in ``self.cm``
        (already summed over ranks)."""
        if not self.one_launch:
            prev = _lib.lib().nkbk_heads_one_launch(0)
            try:
                self.one_launch = True
                return self.heads_step(emb, W_cat, b_cat, labels, train, want_probs, want_pred, update_confusion)
            finally:
                self.one_launch = False
                _lib.lib().nkbk_heads_one_launch(prev)
        B = emb.shape[0]
        bufs = self._buffers(B, train, want_probs)
        do_cm = update_confusion and labels is not None
        pred = None
        if want_pred or do_cm:
            pred = self._pred.get(B)
            if pred is None:
                pred = torch.empty((B, self.T), dtype=torch.int32, device=self.device)
                self._pred = {B: pred}
        if self._heads_done is not None:
            # an overlapped K1 sits between two heads steps: make the order of the heads steps themselves explicit
            torch.cuda.current_stream(self.device).wait_event(self._heads_done)
        if train and (self.comm.world == 1 or self.transport == "peer"):
            # the whole step -- forward, loss, K3, dW / db, exchange (K4' as the kernel's epilogue), finalize -- as ONE
            # launch (nkbk_heads_train_step); the NCCL transport keeps the separate launches below
            peer = self.comm.world > 1 and self.comm.init_peer(self.device, bufs.reduce_buf.numel(), self.cm.numel())
            if self.comm.world == 1 or peer:
                ops.heads_train_step(emb, W_cat, b_cat, labels, bufs, self.loss_kind, self.gamma, self.class_weight,
                                     self.ignore_index, out_pred=pred, cm_total=self.cm if do_cm else None,
                                     cm_step=self.cm_step if do_cm else None, peer=peer)
                bufs.pred = pred
                return bufs
            self.transport = "nccl"
        # K2 with K3 fused into its epilogue: argmax + confusion counts while the logits are on chip
        ops.heads_fwd_loss_bwd(emb, W_cat, b_cat, labels, bufs, self.loss_kind, self.gamma, self.class_weight,
                               self.ignore_index, out_pred=pred, cm_step=self.cm_step if do_cm else None)
        # K4' (fused push / rank-ordered sum / finalize over NVLink peer memory) or K4 (NCCL) + finalize
        used = self.comm.exchange_finalize(bufs, self.cm if do_cm else None, self.cm_step if do_cm else None,
                                           self.transport)
        if self.comm.world > 1:
            self.transport = used
        bufs.pred = pred
        return bufs

    def reset_confusion(self):
        self.cm.zero_()

    def confusion_matrices(self):
        """Per-task int64 [C_t, C_t] numpy matrices (one D2H)."""
        flat = self.cm.cpu().numpy()
        out, off = [], 0
        for t in range(self.T):
            C = self.seg[t + 1] - self.seg[t]
            out.append(flat[off: off + C * C].reshape(C, C).copy())
            off += C * C
        return out
