"""Thin torch-tensor wrappers over the C ABI (include/nkbk.h).

PyTorch is plumbing here: it owns device memory and streams; every byte of
arithmetic on the hot path happens inside libnkbk.so.  All functions require
CUDA tensors and raise otherwise -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_int32, c_uint8
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import BF16, F32, LOSS_CE, LOSS_FOCAL, check, lib
from .transforms import PreprocessPlan

_DT = {torch.float32: F32, torch.bfloat16: BF16}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _need_cuda(name: str, t: torch.Tensor):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: nkb_classification_b200 has no CPU path")


def _seg_array(seg_offsets: Sequence[int]):
    seg = [int(x) for x in seg_offsets]
    return (c_int32 * len(seg))(*seg), len(seg) - 1, seg[-1]


# --------------------------------------------------------------------------
# K1
# --------------------------------------------------------------------------
def frame_descriptors(frames: torch.Tensor) -> torch.Tensor:
    """int64 [F,4] = {byte offset, height, width, pitch} for a contiguous uint8 [F,H,W,3] batch."""
    F, H, W, C = frames.shape
    assert C == 3
    d = torch.empty((F, 4), dtype=torch.int64)
    d[:, 0] = torch.arange(F, dtype=torch.int64) * (H * W * 3)
    d[:, 1] = H
    d[:, 2] = W
    d[:, 3] = W * 3
    return d.to(frames.device, non_blocking=True)


def preprocess_crops(
    frames: torch.Tensor,              # uint8 CUDA [F,H,W,3] contiguous, or a flat uint8 buffer with `frame_desc`
    boxes: torch.Tensor,               # int32 CUDA [n,4]  x0,y0,x1,y1
    frame_idx: torch.Tensor,           # int32 CUDA [n]
    plan: PreprocessPlan,
    out_dtype: torch.dtype = torch.float32,
    out: Optional[torch.Tensor] = None,
    out_u8: Optional[torch.Tensor] = None,
    frame_desc: Optional[torch.Tensor] = None,
    bad_count: Optional[torch.Tensor] = None,
    aug=None,                          # transforms.AugmentBatch (train-time pipeline) or None
) -> torch.Tensor:
    """One K1 launch: [n,3,out_h,out_w] normalised crops (see nkbk_preprocess_crops / nkbk_preprocess_crops_aug).
    ``aug``: the per-sample parameters ``plan.draw(n)`` returned; required iff the plan has fused augmentations."""
    _need_cuda("frames", frames)
    _need_cuda("boxes", boxes)
    _need_cuda("frame_idx", frame_idx)
    if frames.dtype != torch.uint8 or not frames.is_contiguous():
        raise ValueError("frames must be a contiguous uint8 tensor")
    if boxes.dtype != torch.int32 or frame_idx.dtype != torch.int32:
        raise ValueError("boxes and frame_idx must be int32")
    if not boxes.is_contiguous() or not frame_idx.is_contiguous():
        raise ValueError("boxes and frame_idx must be contiguous")
    n = int(frame_idx.numel())
    if boxes.numel() != 4 * n:
        raise ValueError(f"boxes has {boxes.numel()} elements for {n} crops")
    if frame_desc is None:
        if frames.dim() != 4 or frames.shape[-1] != 3:
            raise ValueError("frames must be [F,H,W,3] when frame_desc is not given")
        if frames.shape[2] < 2:
            raise ValueError("frames narrower than 2 pixels are not supported")
        frame_desc = frame_descriptors(frames)
    _need_cuda("frame_desc", frame_desc)
    n_frames = int(frame_desc.shape[0])
    if out_dtype not in _DT:
        raise ValueError(f"out_dtype {out_dtype} not supported (float32 | bfloat16)")
    if out is None:
        out = torch.empty((n, 3, plan.out_h, plan.out_w), dtype=out_dtype, device=frames.device)
    else:
        if out.dtype != out_dtype or not out.is_contiguous() or out.numel() != n * 3 * plan.out_h * plan.out_w:
            raise ValueError("out has the wrong dtype / shape / layout")
    if out_u8 is not None:
        if out_u8.dtype != torch.uint8 or not out_u8.is_contiguous() or out_u8.numel() != n * 3 * plan.out_h * plan.out_w:
            raise ValueError("out_u8 must be contiguous uint8 [n,out_h,out_w,3]")
    pad = (c_uint8 * 3)(*plan.pad_value)
    m = (c_float * 3)(*plan.mean255)
    d = (c_float * 3)(*plan.denom)
    if (aug is None) != (getattr(plan, "augment", None) is None):
        raise ValueError("`aug` must be given exactly when the plan has fused augmentations (aug = plan.draw(n))")
    if aug is None:
        rc = lib().nkbk_preprocess_crops(
            _ptr(frames), _ptr(frame_desc), n_frames, _ptr(boxes), _ptr(frame_idx), n, plan.mode, plan.out_h,
            plan.out_w, plan.max_size, pad, m, d, int(plan.channel_swap), _ptr(out), _DT[out_dtype], _ptr(out_u8),
            _ptr(bad_count), _stream(frames.device),
        )
        check(rc)
        return out
    if len(aug.flags) != n:
        raise ValueError(f"augmentation parameters for {len(aug.flags)} samples, batch has {n}")
    dev = frames.device
    a_flags, a_alpha, a_beta, a_holes, a_hsv = upload_augment(aug, dev)
    fill = (c_uint8 * 3)(*[int(v) for v in aug.fill])
    rc = lib().nkbk_preprocess_crops_aug(
        _ptr(frames), _ptr(frame_desc), n_frames, _ptr(boxes), _ptr(frame_idx), n, plan.mode, plan.out_h, plan.out_w,
        plan.max_size, pad, m, d, int(plan.channel_swap), _ptr(a_flags), _ptr(a_alpha), _ptr(a_beta), _ptr(a_holes),
        aug.max_holes, fill, _ptr(a_hsv), int(getattr(aug, "hsv_trunc_cols", 0)), _ptr(out), _DT[out_dtype],
        _ptr(out_u8), _ptr(bad_count), _stream(dev),
    )
    check(rc)
    return out


def upload_augment(aug, device):
    """Device copies of one batch's augmentation parameters (flags, alpha, beta*255, holes, HSV tables): 12 + 16 *
    max_holes (+ 768) bytes per sample, uploaded once per batch on the CURRENT stream and cached on the batch object
    (the loader calls this from its producer thread on the copy stream, so K1 never waits for it)."""
    dev = torch.device(device)
    cached = getattr(aug, "_dev", None)
    if cached is not None and cached[0] == dev:
        return cached[1]
    n = len(aug.flags)
    hsv = None
    if getattr(aug, "hsv_lut", None) is not None:
        if aug.hsv_lut.shape != (n, 3, 256) or aug.hsv_lut.dtype != np.uint8:
            raise ValueError("hsv_lut must be uint8 [n, 3, 256]")
        hsv = torch.from_numpy(np.ascontiguousarray(aug.hsv_lut)).to(dev, non_blocking=True)
    elif getattr(aug, "hsv_shift", None) is None and np.any(np.asarray(aug.flags) & 8):
        raise ValueError("a sample has the HueSaturationValue flag but the batch carries neither hsv_lut nor hsv_shift")
    flags_d = torch.from_numpy(np.ascontiguousarray(aug.flags, dtype=np.int32)).to(dev, non_blocking=True)
    if hsv is None and getattr(aug, "hsv_shift", None) is not None:
        # tables built on the device from the shift draws: 24 bytes per sample cross the bus instead of 768
        sh = torch.from_numpy(np.ascontiguousarray(aug.hsv_shift, dtype=np.float64).reshape(n, 3)).to(dev, non_blocking=True)
        hsv = torch.empty((n, 3, 256), dtype=torch.uint8, device=dev)
        check(lib().nkbk_build_hsv_luts(_ptr(sh), _ptr(flags_d), n, _ptr(hsv), _stream(dev)))
    t = (flags_d,
         torch.from_numpy(np.ascontiguousarray(aug.alpha, dtype=np.float32)).to(dev, non_blocking=True),
         torch.from_numpy(np.ascontiguousarray(aug.beta, dtype=np.float32)).to(dev, non_blocking=True),
         torch.from_numpy(np.ascontiguousarray(aug.holes, dtype=np.int32)).to(dev, non_blocking=True), hsv)
    aug._dev = (dev, t)
    return t


def debug_axis_table(dsize: int, ssize: int, horizontal: bool):
    """Host-only: the coefficient table the kernel prologue computes (no GPU needed)."""
    s = np.empty(dsize, dtype=np.int32)
    c0 = np.empty(dsize, dtype=np.int32)
    c1 = np.empty(dsize, dtype=np.int32)
    check(lib().nkbk_debug_axis_table(dsize, ssize, int(horizontal), s.ctypes.data, c0.ctypes.data, c1.ctypes.data))
    return s, c0, c1


def debug_brightness_contrast_lut(alpha: float, beta255: float) -> np.ndarray:
    """Host-only: the 256-entry brightness/contrast table K1 evaluates per pixel."""
    lut = np.empty(256, dtype=np.uint8)
    check(lib().nkbk_debug_brightness_contrast_lut(float(alpha), float(beta255), lut.ctypes.data))
    return lut


def debug_hsv_shift(rgb: np.ndarray, lut: np.ndarray, trunc: bool = False) -> np.ndarray:
    """Host-only: K1's HueSaturationValue pixel function on uint8 [..., 3] RGB pixels with a [3, 256] table;
    ``trunc``: cv2's vectorised rounding (truncate) instead of its scalar one (nearest even)."""
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    lut = np.ascontiguousarray(lut, dtype=np.uint8).reshape(768)
    out = np.empty_like(rgb)
    check(lib().nkbk_debug_hsv_shift(rgb.ctypes.data, rgb.size // 3, lut.ctypes.data, int(bool(trunc)), out.ctypes.data))
    return out


def debug_letterbox(h: int, w: int, max_size: int, out_h: int, out_w: int):
    o = np.empty(4, dtype=np.int32)
    check(lib().nkbk_debug_letterbox(h, w, max_size, out_h, out_w, o.ctypes.data))
    return tuple(int(x) for x in o)


# --------------------------------------------------------------------------
# K2
# --------------------------------------------------------------------------
def heads_reduce_buf_len(D: int, NC: int, T: int) -> int:
    return int(lib().nkbk_heads_reduce_buf_len(D, NC, T))


class HeadsBuffers:
    """Caller-owned scratch for K2, reused across steps (allocation-free steady state)."""

    def __init__(self, B: int, D: int, seg_offsets: Sequence[int], device, want_logits=True, want_probs=True,
                 want_grads=True):
        self.B, self.D = B, D
        self.seg = [int(x) for x in seg_offsets]
        self.T, self.NC = len(self.seg) - 1, self.seg[-1]
        f32 = dict(dtype=torch.float32, device=device)
        self.logits = torch.empty((B, self.NC), **f32) if want_logits else None
        self.probs = torch.empty((B, self.NC), **f32) if want_probs else None
        self.dlogits = torch.empty((B, self.NC), **f32) if want_grads else None
        self.reduce_buf = torch.empty(heads_reduce_buf_len(D, self.NC, self.T), **f32)
        ws = int(lib().nkbk_heads_workspace_bytes(B, D, self.NC, self.T))
        self.workspace = torch.empty(max(ws, 16), dtype=torch.uint8, device=device)
        self.loss = torch.empty(self.T + 1, **f32)

    # views into the reduce buffer
    def dW(self) -> torch.Tensor:
        return self.reduce_buf[: self.NC * self.D].view(self.NC, self.D)

    def db(self) -> torch.Tensor:
        return self.reduce_buf[self.NC * self.D : self.NC * self.D + self.NC]

    def loss_sum(self) -> torch.Tensor:
        o = self.NC * self.D + self.NC
        return self.reduce_buf[o : o + self.T]

    def denom(self) -> torch.Tensor:
        o = self.NC * self.D + self.NC + self.T
        return self.reduce_buf[o : o + self.T]


def heads_fwd_loss_bwd(
    emb: torch.Tensor, W_cat: torch.Tensor, b_cat: torch.Tensor, labels: Optional[torch.Tensor], bufs: HeadsBuffers,
    loss_kind: int = LOSS_FOCAL, gamma: float = 2.0, class_weight: Optional[torch.Tensor] = None,
    ignore_index: int = -100, out_pred: Optional[torch.Tensor] = None, cm_step: Optional[torch.Tensor] = None,
    weights_version: Optional[int] = None,
) -> HeadsBuffers:
    """Forward + loss terms + dlogits + unnormalised dW/db into ``bufs`` (see nkbk_heads_step).
    ``weights_version``: what identifies the current contents of ``W_cat`` (default: its torch version counter); lets
    the tcgen05 forward keep its bf16 copy of the weights across calls (nkbk_heads_weights_version).
    ``out_pred`` int32 [B,T] / ``cm_step`` int64 confusion counts (accumulated; needs labels): K3 fused into the
    forward epilogue -- same results as a separate :func:`argmax_confusion` on ``bufs.logits``, one launch fewer."""
    _need_cuda("emb", emb)
    _need_cuda("W_cat", W_cat)
    if emb.dtype not in _DT:
        raise ValueError(f"emb dtype {emb.dtype} not supported (float32 | bfloat16)")
    if not emb.is_contiguous() or emb.dim() != 2:
        raise ValueError("emb must be contiguous [B,D]")
    B, D = emb.shape
    if (B, D) != (bufs.B, bufs.D):
        raise ValueError(f"emb {tuple(emb.shape)} does not match buffers ({bufs.B},{bufs.D})")
    if W_cat.dtype != torch.float32 or b_cat.dtype != torch.float32:
        raise ValueError("W_cat / b_cat must be float32")
    if tuple(W_cat.shape) != (bufs.NC, D) or b_cat.numel() != bufs.NC or not W_cat.is_contiguous():
        raise ValueError("W_cat must be contiguous [NC,D] and b_cat [NC]")
    if labels is not None:
        _need_cuda("labels", labels)
        if labels.dtype != torch.int64 or not labels.is_contiguous() or labels.numel() != B * bufs.T:
            raise ValueError("labels must be contiguous int64 [B,T]")
    if class_weight is not None:
        if class_weight.dtype != torch.float32 or class_weight.numel() != bufs.NC or not class_weight.is_cuda:
            raise ValueError("class_weight must be CUDA float32 [NC]")
    seg, T, _ = _seg_array(bufs.seg)
    if out_pred is not None and (out_pred.dtype != torch.int32 or out_pred.numel() != B * T
                                 or not out_pred.is_contiguous() or not out_pred.is_cuda):
        raise ValueError("out_pred must be a contiguous CUDA int32 [B,T]")
    if cm_step is not None:
        if labels is None:
            raise ValueError("cm_step needs labels")
        if cm_step.dtype != torch.int64 or cm_step.numel() != confusion_len(bufs.seg) or not cm_step.is_cuda:
            raise ValueError("cm_step must be CUDA int64 of confusion_len(seg) elements")
    if emb.dtype == torch.bfloat16:
        v = W_cat._version if weights_version is None else int(weights_version)
        lib().nkbk_heads_weights_version(int(v) + 1)            # (never 0: 0 means "unknown")
    rc = lib().nkbk_heads_step(
        _ptr(emb), _DT[emb.dtype], B, D, _ptr(W_cat), _ptr(b_cat), seg, T, _ptr(labels), int(loss_kind), float(gamma),
        _ptr(class_weight), int(ignore_index), _ptr(bufs.logits), _ptr(bufs.probs), _ptr(bufs.dlogits),
        _ptr(bufs.reduce_buf), _ptr(out_pred), _ptr(cm_step), _ptr(bufs.workspace), bufs.workspace.numel(),
        _stream(emb.device),
    )
    check(rc)
    return bufs


def _check_heads_args(emb, W_cat, b_cat, labels, bufs, class_weight, out_pred, cm_step):
    _need_cuda("emb", emb)
    _need_cuda("W_cat", W_cat)
    if emb.dtype not in _DT:
        raise ValueError(f"emb dtype {emb.dtype} not supported (float32 | bfloat16)")
    if not emb.is_contiguous() or emb.dim() != 2:
        raise ValueError("emb must be contiguous [B,D]")
    B, D = emb.shape
    if (B, D) != (bufs.B, bufs.D):
        raise ValueError(f"emb {tuple(emb.shape)} does not match buffers ({bufs.B},{bufs.D})")
    if W_cat.dtype != torch.float32 or b_cat.dtype != torch.float32:
        raise ValueError("W_cat / b_cat must be float32")
    if tuple(W_cat.shape) != (bufs.NC, D) or b_cat.numel() != bufs.NC or not W_cat.is_contiguous():
        raise ValueError("W_cat must be contiguous [NC,D] and b_cat [NC]")
    if labels is not None:
        _need_cuda("labels", labels)
        if labels.dtype != torch.int64 or not labels.is_contiguous() or labels.numel() != B * bufs.T:
            raise ValueError("labels must be contiguous int64 [B,T]")
    if class_weight is not None:
        if class_weight.dtype != torch.float32 or class_weight.numel() != bufs.NC or not class_weight.is_cuda:
            raise ValueError("class_weight must be CUDA float32 [NC]")
    if out_pred is not None and (out_pred.dtype != torch.int32 or out_pred.numel() != B * bufs.T
                                 or not out_pred.is_contiguous() or not out_pred.is_cuda):
        raise ValueError("out_pred must be a contiguous CUDA int32 [B,T]")
    if cm_step is not None:
        if labels is None:
            raise ValueError("cm_step needs labels")
        if cm_step.dtype != torch.int64 or cm_step.numel() != confusion_len(bufs.seg) or not cm_step.is_cuda:
            raise ValueError("cm_step must be CUDA int64 of confusion_len(seg) elements")


def heads_train_step(
    emb: torch.Tensor, W_cat: torch.Tensor, b_cat: torch.Tensor, labels: Optional[torch.Tensor], bufs: HeadsBuffers,
    loss_kind: int = LOSS_FOCAL, gamma: float = 2.0, class_weight: Optional[torch.Tensor] = None,
    ignore_index: int = -100, out_pred: Optional[torch.Tensor] = None, cm_total: Optional[torch.Tensor] = None,
    cm_step: Optional[torch.Tensor] = None, peer: bool = False,
) -> HeadsBuffers:
    """The whole training step of the heads in one call -- forward, loss, K3, dW / db, the exchange step when
    ``peer`` (K4' over NVLink peer memory; needs ``Communicator.init_peer``) and the finalize -- ONE persistent kernel
    when the shape qualifies (nkbk_heads_train_step).  Afterwards ``bufs.dW()`` / ``bufs.db()`` are the (global) mean
    gradients, ``bufs.loss`` [T+1] the mean losses, ``cm_total`` has the step's (global) confusion counts folded in."""
    if bufs.dlogits is None:
        raise ValueError("heads_train_step needs buffers created with want_grads=True")
    _check_heads_args(emb, W_cat, b_cat, labels, bufs, class_weight, out_pred, cm_step)
    n_cm = 0
    if cm_total is not None or cm_step is not None:
        if cm_total is None or cm_step is None or cm_total.numel() != cm_step.numel() \
                or cm_total.dtype != torch.int64 or not cm_total.is_cuda:
            raise ValueError("cm_total / cm_step must both be CUDA int64 of equal length")
        n_cm = cm_total.numel()
    seg, T, _ = _seg_array(bufs.seg)
    B, D = emb.shape
    rc = lib().nkbk_heads_train_step(
        _ptr(emb), _DT[emb.dtype], B, D, _ptr(W_cat), _ptr(b_cat), seg, T, _ptr(labels), int(loss_kind), float(gamma),
        _ptr(class_weight), int(ignore_index), _ptr(bufs.logits), _ptr(bufs.probs), _ptr(bufs.dlogits),
        _ptr(bufs.reduce_buf), _ptr(out_pred), _ptr(cm_step), _ptr(cm_total), n_cm, _ptr(bufs.loss),
        _lib.EXCHANGE_PEER if peer else _lib.EXCHANGE_LOCAL, _ptr(bufs.workspace), bufs.workspace.numel(),
        _stream(emb.device),
    )
    check(rc)
    return bufs


def heads_last_path() -> int:
    """Bit mask of the kernels that served this thread's last heads call: _lib.PATH_FUSED | PATH_TC_FWD | PATH_FFMA_FWD."""
    return int(lib().nkbk_heads_last_path())


def heads_finalize(bufs: HeadsBuffers, cm_total: Optional[torch.Tensor] = None,
                   cm_step: Optional[torch.Tensor] = None) -> torch.Tensor:
    """reduce_buf sums -> mean gradients in place; returns fp32 [T+1] = per-task losses + total.
    With ``cm_total`` / ``cm_step``: cm_total += cm_step; cm_step = 0 in the same launch."""
    seg, T, _ = _seg_array(bufs.seg)
    n_cm = 0
    if cm_total is not None or cm_step is not None:
        if cm_total is None or cm_step is None or cm_total.numel() != cm_step.numel() \
                or cm_total.dtype != torch.int64 or cm_step.dtype != torch.int64:
            raise ValueError("cm_total / cm_step must both be int64 of equal length")
        n_cm = cm_total.numel()
    check(lib().nkbk_heads_finalize(_ptr(bufs.reduce_buf), bufs.D, seg, T, _ptr(bufs.loss), _ptr(cm_total),
                                    _ptr(cm_step), n_cm, _stream(bufs.reduce_buf.device)))
    return bufs.loss


def peer_allreduce_finalize(bufs: HeadsBuffers, cm_total: Optional[torch.Tensor] = None,
                            cm_step: Optional[torch.Tensor] = None) -> torch.Tensor:
    """K4': the cross-rank sum of ``reduce_buf`` (+ ``cm_step``) over NVLink peer memory fused with
    :func:`heads_finalize` -- one launch, same results contract (sums are over all ranks).  Requires
    ``Communicator.init_peer`` (nkbk_peer_init / nkbk_peer_connect)."""
    seg, T, _ = _seg_array(bufs.seg)
    n_cm = 0
    if cm_total is not None or cm_step is not None:
        if cm_total is None or cm_step is None or cm_total.numel() != cm_step.numel() \
                or cm_total.dtype != torch.int64 or cm_step.dtype != torch.int64:
            raise ValueError("cm_total / cm_step must both be int64 of equal length")
        n_cm = cm_total.numel()
    check(lib().nkbk_peer_allreduce_finalize(_ptr(bufs.reduce_buf), bufs.D, seg, T, _ptr(bufs.loss), _ptr(cm_total),
                                             _ptr(cm_step), n_cm, _stream(bufs.reduce_buf.device)))
    return bufs.loss


def heads_demb(bufs: HeadsBuffers, W_cat: torch.Tensor, out_dtype: torch.dtype = torch.float32,
               out: Optional[torch.Tensor] = None, task_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    if bufs.dlogits is None:
        raise ValueError("buffers were created without gradients")
    if out is None:
        out = torch.empty((bufs.B, bufs.D), dtype=out_dtype, device=W_cat.device)
    seg, T, _ = _seg_array(bufs.seg)
    if task_scale is not None and (task_scale.dtype != torch.float32 or task_scale.numel() != T or not task_scale.is_cuda):
        raise ValueError("task_scale must be CUDA float32 [T]")
    check(lib().nkbk_heads_demb(_ptr(bufs.dlogits), _ptr(bufs.reduce_buf), _ptr(W_cat), seg, T, bufs.B, bufs.D,
                                _ptr(task_scale), _ptr(out), _DT[out_dtype], _stream(W_cat.device)))
    return out


# --------------------------------------------------------------------------
# K3
# --------------------------------------------------------------------------
def confusion_len(seg_offsets: Sequence[int]) -> int:
    return int(sum((b - a) ** 2 for a, b in zip(seg_offsets[:-1], seg_offsets[1:])))


def argmax_confusion(logits: torch.Tensor, seg_offsets: Sequence[int], labels: Optional[torch.Tensor] = None,
                     cm: Optional[torch.Tensor] = None, out_pred: Optional[torch.Tensor] = None,
                     want_pred: bool = True) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """Per-task argmax (+ accumulate int64 confusion counts into ``cm``)."""
    _need_cuda("logits", logits)
    if logits.dtype not in _DT or logits.dim() != 2 or logits.stride(1) != 1:
        raise ValueError("logits must be [B,ld] float32 | bfloat16 with unit inner stride")
    B, ld = logits.shape[0], logits.stride(0) if logits.shape[0] > 1 else logits.shape[1]
    seg, T, NC = _seg_array(seg_offsets)
    if labels is not None:
        _need_cuda("labels", labels)
        if labels.dtype != torch.int64 or not labels.is_contiguous() or labels.numel() != B * T:
            raise ValueError("labels must be contiguous int64 [B,T]")
    if cm is not None:
        if cm.dtype != torch.int64 or not cm.is_cuda or cm.numel() != confusion_len(seg_offsets):
            raise ValueError("cm must be CUDA int64 with sum(C_t^2) entries")
    if want_pred and out_pred is None:
        out_pred = torch.empty((B, T), dtype=torch.int32, device=logits.device)
    check(lib().nkbk_argmax_confusion(_ptr(logits), _DT[logits.dtype], B, ld, seg, T, _ptr(labels), _ptr(out_pred),
                                      _ptr(cm), _stream(logits.device)))
    return out_pred, cm


# --------------------------------------------------------------------------
# loss on given logits (criterion(pred, true) of the reference API)
# --------------------------------------------------------------------------
def loss_fwd_bwd(logits: torch.Tensor, seg_offsets: Sequence[int], labels: torch.Tensor, loss_kind: int,
                 gamma: float = 2.0, class_weight: Optional[torch.Tensor] = None, ignore_index: int = -100,
                 want_probs: bool = False, want_grad: bool = True):
    """Returns (loss [T+1] fp32, dlogits [B,NC] fp32 | None, probs [B,NC] fp32 | None)."""
    _need_cuda("logits", logits)
    _need_cuda("labels", labels)
    if logits.dtype not in _DT or logits.dim() != 2 or logits.stride(1) != 1:
        raise ValueError("logits must be [B,ld] float32 | bfloat16 with unit inner stride")
    seg, T, NC = _seg_array(seg_offsets)
    B = logits.shape[0]
    ld = logits.stride(0) if B > 1 else logits.shape[1]
    if labels.dtype != torch.int64 or not labels.is_contiguous() or labels.numel() != B * T:
        raise ValueError("labels must be contiguous int64 [B,T]")
    if class_weight is not None and (class_weight.dtype != torch.float32 or class_weight.numel() != NC
                                     or not class_weight.is_cuda):
        raise ValueError("class_weight must be CUDA float32 [NC]")
    dev = logits.device
    loss = torch.empty(T + 1, dtype=torch.float32, device=dev)
    dl = torch.empty((B, NC), dtype=torch.float32, device=dev) if want_grad else None
    pr = torch.empty((B, NC), dtype=torch.float32, device=dev) if want_probs else None
    ws = torch.empty(max(16, int(lib().nkbk_loss_workspace_bytes(B, T))), dtype=torch.uint8, device=dev)
    check(lib().nkbk_loss_fwd_bwd(_ptr(logits), _DT[logits.dtype], B, ld, seg, T, _ptr(labels), int(loss_kind),
                                  float(gamma), _ptr(class_weight), int(ignore_index), _ptr(pr), _ptr(dl), _ptr(loss),
                                  _ptr(ws), ws.numel(), _stream(dev)))
    return loss, dl, pr


def loss_rows(logits: torch.Tensor, seg_offsets: Sequence[int], labels: torch.Tensor, loss_kind: int,
              gamma: float = 2.0, class_weight: Optional[torch.Tensor] = None, ignore_index: int = -100,
              want_grad: bool = True):
    """Unreduced loss (nkbk_loss_rows): returns (row_loss [B,T] fp32, dlogits [B,NC] fp32 | None), nothing divided."""
    _need_cuda("logits", logits)
    _need_cuda("labels", labels)
    if logits.dtype not in _DT or logits.dim() != 2 or logits.stride(1) != 1:
        raise ValueError("logits must be [B,ld] float32 | bfloat16 with unit inner stride")
    seg, T, NC = _seg_array(seg_offsets)
    B = logits.shape[0]
    ld = logits.stride(0) if B > 1 else logits.shape[1]
    if labels.dtype != torch.int64 or not labels.is_contiguous() or labels.numel() != B * T:
        raise ValueError("labels must be contiguous int64 [B,T]")
    if class_weight is not None and (class_weight.dtype != torch.float32 or class_weight.numel() != NC
                                     or not class_weight.is_cuda):
        raise ValueError("class_weight must be CUDA float32 [NC]")
    dev = logits.device
    rows = torch.empty((B, T), dtype=torch.float32, device=dev)
    dl = torch.empty((B, NC), dtype=torch.float32, device=dev) if want_grad else None
    ws = torch.empty(max(16, int(lib().nkbk_loss_workspace_bytes(B, T))), dtype=torch.uint8, device=dev)
    check(lib().nkbk_loss_rows(_ptr(logits), _DT[logits.dtype], B, ld, seg, T, _ptr(labels), int(loss_kind), float(gamma),
                               _ptr(class_weight), int(ignore_index), _ptr(rows), _ptr(dl), _ptr(ws), ws.numel(),
                               _stream(dev)))
    return rows, dl


# --------------------------------------------------------------------------
# K5
# --------------------------------------------------------------------------
def roc_auc_counts(probs: torch.Tensor, seg_offsets: Sequence[int], labels: torch.Tensor) -> torch.Tensor:
    """int64 [NC, 3] = (num2, P, Q) per class column: AUC = num2 / (2 P Q) (see nkbk_roc_auc_counts).
    ``probs`` fp32 [N, >= NC], ``labels`` int64 [N, T], both CUDA; one small D2H is left to the caller."""
    _need_cuda("probs", probs)
    _need_cuda("labels", labels)
    seg, T, NC = _seg_array(seg_offsets)
    if probs.dtype != torch.float32 or probs.dim() != 2 or probs.stride(1) != 1 or probs.shape[1] < NC:
        raise ValueError("probs must be float32 [N, >= NC] with unit column stride")
    N = int(probs.shape[0])
    labels = labels.reshape(N, T) if labels.numel() == N * T else labels
    if labels.dtype != torch.int64 or tuple(labels.shape) != (N, T) or not labels.is_contiguous():
        raise ValueError(f"labels must be contiguous int64 [{N}, {T}]")
    out = torch.empty((NC, 3), dtype=torch.int64, device=probs.device)
    nbytes = int(lib().nkbk_auc_workspace_bytes(N, NC))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=probs.device)
    check(lib().nkbk_roc_auc_counts(_ptr(probs), N, int(probs.stride(0)), seg, T, _ptr(labels), _ptr(out), _ptr(ws),
                                    nbytes, _stream(probs.device)))
    return out
