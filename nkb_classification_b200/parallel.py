"""Batch sharding across the GPUs of one box (K4).

One process per GPU.  Frames (with all their boxes) are sharded contiguously
across ranks; the only data-path exchange is one sum-all-reduce of the packed
K2 reduce buffer per training step and one of the int64 confusion counts.
Two transports for that step live in libnkbk:

* ``peer``  (K4', default when the GPUs can map each other's memory): ONE kernel
  pushes the payload into every peer's cudaIpc-mapped inbox over NVLink, waits on
  flags, sums in rank order and applies the K2 finalize (nkbk_peer_allreduce_finalize);
* ``nccl``  (K4): grouped ncclAllReduce + the separate finalize launch.

``torch.distributed`` is used only for rendezvous (broadcasting the NCCL unique
id, all-gathering the 64-byte IPC handles) and for barriers / timing in bench.py.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from ._lib import IPC_HANDLE_BYTES, UNIQUE_ID_BYTES, check, lib


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
        self.active = False          # NCCL communicator (K4)
        self.peer_active = False     # NVLink peer-memory transport (K4')
        self._peer_cap = (0, 0)
        self._peer_failed = False

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

    def init_peer(self, device: torch.device, max_f32: int, max_i64: int) -> bool:
        """Set up the NVLink peer-memory transport (K4') for payloads of up to ``max_f32`` fp32 + ``max_i64`` int64
        elements.  Collective: every rank must call it.  Returns False (and leaves the NCCL transport in charge) when
        the devices cannot map each other's memory; all ranks take the same decision."""
        import torch.distributed as dist
        if self._peer_failed:
            return False
        if self.peer_active:
            if max_f32 <= self._peer_cap[0] and max_i64 <= self._peer_cap[1]:
                return True
            # grow to the running maximum of both payloads, so callers alternating between shapes (per-head dropout:
            # large f32 / small cm, then the reverse) re-create the IPC mapping at most once per new maximum
            max_f32, max_i64 = max(max_f32, self._peer_cap[0]), max(max_i64, self._peer_cap[1])
            self._shutdown_peer()
        if self.world > 1 and not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised first (it carries the IPC handles)")
        hbuf = (ctypes.c_uint8 * IPC_HANDLE_BYTES)()
        rc = lib().nkbk_peer_init(self.rank, self.world, device.index or 0, int(max_f32), int(max_i64), hbuf)
        if self.world == 1:
            check(rc)
        ok = rc == 0
        if self.world > 1:
            # every rank walks through both exchanges whatever happened locally (no IPC support in this container,
            # no P2P path between the devices, ...), so that all of them reach the same decision without deadlock
            handles: List[Optional[bytes]] = [None] * self.world
            dist.all_gather_object(handles, bytes(hbuf) if ok else None)
            if all(h is not None for h in handles):
                raw = (ctypes.c_uint8 * (IPC_HANDLE_BYTES * self.world)).from_buffer_copy(b"".join(handles))
                ok = lib().nkbk_peer_connect(raw) == 0
            else:
                ok = False
            flags = [None] * self.world
            dist.all_gather_object(flags, ok)          # also the barrier nkbk_peer_connect asks for
            ok = all(flags)
        if not ok:
            lib().nkbk_peer_shutdown()
            self._peer_failed = True     # do not retry the collective set-up every step
            return False
        self.peer_active = True
        self._peer_cap = (int(max_f32), int(max_i64))
        return True

    def _shutdown_peer(self) -> None:
        if self.peer_active:
            # two-phase: nobody frees its inbox while a peer still maps it
            lib().nkbk_peer_disconnect()
            if self.world > 1:
                import torch.distributed as dist
                if dist.is_initialized():
                    dist.barrier()
            lib().nkbk_peer_shutdown()
            self.peer_active = False

    def peer_status(self) -> int:
        """0 = healthy; r + 1 = a wait on rank r timed out (host-synchronous)."""
        st = ctypes.c_int32(0)
        check(lib().nkbk_peer_status(ctypes.byref(st)))
        return int(st.value)

    def exchange_finalize(self, bufs, cm_total: Optional[torch.Tensor] = None, cm_step: Optional[torch.Tensor] = None,
                          transport: str = "peer") -> str:
        """The path's exchange step + K2 finalize for one training step: sums ``bufs.reduce_buf`` (and ``cm_step``)
        over the ranks, divides by the global denominators, folds the step confusion counts into ``cm_total``.
        ``transport``: "peer" = K4' (one fused kernel over NVLink peer memory; set up collectively on first use,
        silently replaced by "nccl" when the GPUs cannot map each other's memory), "nccl" = K4 + finalize launch.
        Returns the transport used.  world == 1: just the finalize."""
        from . import ops
        if self.world > 1 and transport == "peer":
            dev = bufs.reduce_buf.device
            n_cm = 0 if cm_step is None else cm_step.numel()
            if not self.init_peer(dev, bufs.reduce_buf.numel(), n_cm):
                transport = "nccl"
        if self.world > 1 and transport == "peer":
            ops.peer_allreduce_finalize(bufs, cm_total, cm_step)
            return "peer"
        self.allreduce_heads(bufs.reduce_buf, cm_step)
        ops.heads_finalize(bufs, cm_total, cm_step)
        return "nccl" if self.world > 1 else "local"

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
        self._shutdown_peer()
        if self.active:
            lib().nkbk_comm_shutdown()
            self.active = False
