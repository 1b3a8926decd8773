"""ctypes shim over oracle/preprocess_ref.c (TEST INFRASTRUCTURE ONLY).

The C restatement is the fast CPU checker for full-size batches and the
single-thread ``cpu_baseline`` "port" arm of bench.py.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

from .preprocess import MODE_LETTERBOX, Plan, normalize_constants

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "liboracle_preprocess.so"
_lib = None


def build(force: bool = False) -> Path:
    if force or not _SO.exists() or _SO.stat().st_mtime < (_HERE / "preprocess_ref.c").stat().st_mtime:
        subprocess.check_call(["make", "-C", str(_HERE), "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(_SO))
        _lib.oracle_preprocess_batch.restype = ctypes.c_int
        _lib.oracle_preprocess_batch.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
        ]
    return _lib


def preprocess_batch_c(frames: np.ndarray, boxes, frame_idx, plan: Plan, want_u8=True, want_f32=True):
    """frames: uint8 [F, H, W, 3] contiguous.  Returns (u8 [n,H,W,3] | None, f32 [n,3,H,W] | None)."""
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    F, H, W, _ = frames.shape
    boxes = np.ascontiguousarray(boxes, dtype=np.int32).reshape(-1, 4)
    frame_idx = np.ascontiguousarray(frame_idx, dtype=np.int32)
    n = boxes.shape[0]
    desc = np.empty((F, 4), dtype=np.int64)
    desc[:, 0] = np.arange(F, dtype=np.int64) * (H * W * 3)
    desc[:, 1] = H
    desc[:, 2] = W
    desc[:, 3] = W * 3
    m, d = normalize_constants(plan.mean, plan.std, plan.max_pixel_value)
    pad = np.array(plan.pad_value, dtype=np.uint8)
    u8 = np.empty((n, plan.out_h, plan.out_w, 3), dtype=np.uint8) if want_u8 else None
    f32 = np.empty((n, 3, plan.out_h, plan.out_w), dtype=np.float32) if want_f32 else None
    rc = lib().oracle_preprocess_batch(
        frames.ctypes.data, desc.ctypes.data, boxes.ctypes.data, frame_idx.ctypes.data, n,
        int(plan.mode), plan.out_h, plan.out_w, plan.max_size if plan.mode == MODE_LETTERBOX else 0,
        pad.ctypes.data, m.ctypes.data, d.ctypes.data, int(plan.channel_swap),
        u8.ctypes.data if want_u8 else None, f32.ctypes.data if want_f32 else None,
    )
    if rc != 0:
        raise ValueError(f"oracle_preprocess_batch failed rc={rc}")
    return u8, f32
