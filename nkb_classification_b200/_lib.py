"""ctypes binding of libnkbk.so (include/nkbk.h).

The product path has no CPU fallback: if the library is missing or a call
fails, an exception is raised -- nothing here routes to another implementation.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint8, c_void_p
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["NKBK_LIB_PATH"]) if os.environ.get("NKBK_LIB_PATH") else _PKG / "libnkbk.so"  # (A/B builds)

NKBK_OK = 0
NKBK_E_ARG, NKBK_E_SHAPE, NKBK_E_CUDA, NKBK_E_NCCL, NKBK_E_UNSUPPORTED = -1, -2, -3, -4, -5
F32, BF16 = 0, 1
MODE_STRETCH, MODE_LETTERBOX = 0, 1
LOSS_CE, LOSS_FOCAL = 0, 1
EXCHANGE_LOCAL, EXCHANGE_PEER = 0, 1
PATH_FFMA_FWD, PATH_TC_FWD, PATH_FUSED = 1, 2, 4
UNIQUE_ID_BYTES = 128
IPC_HANDLE_BYTES = 64

# every symbol include/nkbk.h declares: (restype, argtypes)
SYMBOLS = {
    "nkbk_abi_version": (c_int, []),
    "nkbk_last_error": (c_char_p, []),
    "nkbk_launch_count": (c_int64, []),
    "nkbk_k1_overlap_previous": (c_int, [c_int]),
    "nkbk_heads_one_launch": (c_int, [c_int]),
    "nkbk_preprocess_crops": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                      POINTER(c_uint8), POINTER(c_float), POINTER(c_float), c_int, c_void_p, c_int,
                                      c_void_p, c_void_p, c_void_p]),
    "nkbk_preprocess_crops_aug": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                          c_int, POINTER(c_uint8), POINTER(c_float), POINTER(c_float), c_int,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_int, POINTER(c_uint8), c_void_p,
                                          c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "nkbk_debug_k1_timeline": (c_int64, [c_void_p, c_int64]),
    "nkbk_build_hsv_luts": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "nkbk_debug_hsv_shift": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_void_p]),
    "nkbk_debug_brightness_contrast_lut": (c_int, [c_float, c_float, c_void_p]),
    "nkbk_debug_axis_table": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "nkbk_debug_letterbox": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "nkbk_heads_reduce_buf_len": (c_int64, [c_int, c_int, c_int]),
    "nkbk_heads_workspace_bytes": (c_int64, [c_int, c_int, c_int, c_int]),
    "nkbk_heads_fwd_loss_bwd": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, POINTER(c_int32), c_int,
                                        c_void_p, c_int, c_float, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_size_t, c_void_p]),
    "nkbk_heads_step": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, POINTER(c_int32), c_int, c_void_p,
                                c_int, c_float, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_size_t, c_void_p]),
    "nkbk_heads_train_step": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, POINTER(c_int32), c_int,
                                      c_void_p, c_int, c_float, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int, c_void_p,
                                      c_size_t, c_void_p]),
    "nkbk_heads_last_path": (c_int, []),
    "nkbk_heads_weights_version": (None, [c_int64]),
    "nkbk_debug_fused_timing": (c_int, [c_void_p, c_int]),
    "nkbk_loss_workspace_bytes": (c_int64, [c_int, c_int]),
    "nkbk_loss_fwd_bwd": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_int32), c_int, c_void_p, c_int, c_float,
                                  c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "nkbk_loss_rows": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_int32), c_int, c_void_p, c_int, c_float,
                               c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "nkbk_heads_finalize": (c_int, [c_void_p, c_int, POINTER(c_int32), c_int, c_void_p, c_void_p, c_void_p, c_int64,
                                    c_void_p]),
    "nkbk_heads_demb": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(c_int32), c_int, c_int, c_int, c_void_p, c_void_p,
                                c_int, c_void_p]),
    "nkbk_argmax_confusion": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_int32), c_int, c_void_p, c_void_p,
                                      c_void_p, c_void_p]),
    "nkbk_auc_workspace_bytes": (c_int64, [c_int64, c_int]),
    "nkbk_roc_auc_counts": (c_int, [c_void_p, c_int64, c_int, POINTER(c_int32), c_int, c_void_p, c_void_p, c_void_p,
                                    c_size_t, c_void_p]),
    "nkbk_comm_unique_id": (c_int, [c_void_p]),
    "nkbk_comm_init": (c_int, [c_int, c_int, c_void_p, c_int]),
    "nkbk_comm_world": (c_int, []),
    "nkbk_allreduce_heads": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "nkbk_comm_shutdown": (c_int, []),
    "nkbk_peer_init": (c_int, [c_int, c_int, c_int, c_int64, c_int64, c_void_p]),
    "nkbk_peer_connect": (c_int, [c_void_p]),
    "nkbk_peer_world": (c_int, []),
    "nkbk_peer_allreduce_finalize": (c_int, [c_void_p, c_int, POINTER(c_int32), c_int, c_void_p, c_void_p, c_void_p,
                                             c_int64, c_void_p]),
    "nkbk_peer_status": (c_int, [POINTER(c_int32)]),
    "nkbk_peer_disconnect": (c_int, []),
    "nkbk_peer_shutdown": (c_int, []),
}


class NkbkError(RuntimeError):
    """A libnkbk call returned a negative status."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"libnkbk error {code}: {msg}")
        self.code = code


_lib = None


def lib() -> ctypes.CDLL:
    """Load libnkbk.so (once).  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m nkb_classification_b200.build` "
                "(nvcc, sm_100a).  This package has no CPU or PyTorch fallback."
            )
        h = ctypes.CDLL(str(LIB_PATH), mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(h, name)
            fn.restype = res
            fn.argtypes = args
        if h.nkbk_abi_version() != 1:
            raise RuntimeError(f"libnkbk ABI version {h.nkbk_abi_version()} != 1")
        _lib = h
    return _lib


def check(rc: int) -> None:
    """Map a status code to the exception type the reference raises at the same check."""
    if rc == NKBK_OK:
        return
    msg = lib().nkbk_last_error().decode("utf-8", "replace")
    if rc in (NKBK_E_ARG, NKBK_E_SHAPE):
        raise ValueError(f"libnkbk: {msg}")
    if rc == NKBK_E_UNSUPPORTED:
        raise NotImplementedError(f"libnkbk: {msg}")
    raise NkbkError(rc, msg)


def launch_count() -> int:
    return int(lib().nkbk_launch_count())
