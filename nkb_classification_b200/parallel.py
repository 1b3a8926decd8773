"""Batch sharding across the GPUs of one box (K4).

One process per GPU.  Frames (with all their boxes) are sharded contiguously
across ranks; the only data-path exchange is one sum-all-reduce of the packed
K2 reduce buffer per training step and one of the int64 confusion counts,
issued through libnkbk's own NCCL communicator (nkbk_allreduce_heads).
``torch.distributed`` is used only for rendezvous (broadcasting the NCCL
unique id) and for barriers / timing in bench.py.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from ._lib import UNIQUE_ID_BYTES, check, lib


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of ``n_items`` for ``rank`` (first n % world ranks get one more)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_frames(frame_idx: np.ndarray, n_frames: int, rank: int, world: int):
    """Shard by FRAME so a rank owns the source pixels of all its crops.
    Returns (frame_begin, frame_end, crop_mask) with crop_mask selecting this rank's crops."""
    fb, fe = shard_range(n_frames, rank, world)
    frame_idx = np.asarray(frame_idx)
    mask = (frame_idx >= fb) & (frame_idx < fe)
    return fb, fe, mask


class Communicator:
    """libnkbk's NCCL communicator, bootstrapped over an existing torch.distributed group."""

    def __init__(self):
        self.rank, self.world = 0, 1
        self.active = False

    def init_from_torch_distributed(self, device: torch.device) -> "Communicator":
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised first (it carries the NCCL unique id)")
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        if self.world == 1:
            return self
        idbuf = (ctypes.c_uint8 * UNIQUE_ID_BYTES)()
        if self.rank == 0:
            check(lib().nkbk_comm_unique_id(idbuf))
        payload = [bytes(idbuf) if self.rank == 0 else None]
        dist.broadcast_object_list(payload, src=0)
        raw = (ctypes.c_uint8 * UNIQUE_ID_BYTES).from_buffer_copy(payload[0])
        check(lib().nkbk_comm_init(self.rank, self.world, raw, device.index or 0))
        self.active = True
        return self

    def allreduce_heads(self, reduce_buf: Optional[torch.Tensor], cm: Optional[torch.Tensor]) -> None:
        """Sum both payloads across ranks, in place, on the current stream.  No-op for world == 1."""
        if self.world == 1:
            return
        if not self.active:
            raise RuntimeError("communicator not initialised")
        dev = (reduce_buf if reduce_buf is not None else cm).device
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        check(lib().nkbk_allreduce_heads(
            None if reduce_buf is None else ctypes.c_void_p(reduce_buf.data_ptr()),
            0 if reduce_buf is None else reduce_buf.numel(),
            None if cm is None else ctypes.c_void_p(cm.data_ptr()),
            0 if cm is None else cm.numel(), st))

    def shutdown(self) -> None:
        if self.active:
            lib().nkbk_comm_shutdown()
            self.active = False
