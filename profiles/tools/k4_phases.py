#!/usr/bin/env python
"""The exchange step of the path at N ranks, measured inside the kernel (ncu cannot replay a kernel that waits for its
peers): per-CTA clock64 stamps of the fused heads step in PEER mode (NKBK_FUSED_TIMING=1) -- local reduce, push over
NVLink, flag publication, time spent waiting for the peers' flags, rank-ordered sum + finalize -- and, beside it, the
device time of the same step with the NCCL transport (grouped ncclAllReduce + finalize) and on one GPU.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/tools/k4_phases.py
"""
import ctypes
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
os.environ["NKBK_FUSED_TIMING"] = "1"


def main():
    from nkb_classification_b200 import _lib, hotpath, ops, transforms as T
    from nkb_classification_b200.parallel import Communicator
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    comm = Communicator().init_from_torch_distributed(dev)
    plan = T.compile_pipeline([T.Resize(32, 32), T.Normalize(), T.ToTensorV2()])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, B, D, classes in (("cfg5_weak_B4096", 4096, 2048, (10,)), ("cfg5_strong_B%d" % (4096 // world), 4096 // world, 2048, (10,)),
                                ("cfg4_strong_B%d" % (1024 // world), 1024 // world, 768, (2, 3, 4, 7, 14))):
        g = torch.Generator().manual_seed(rank)
        emb = torch.randn(B, D, generator=g).to(dev)
        labels = torch.stack([torch.randint(0, c, (B,), generator=g) for c in classes], 1).contiguous().to(dev)
        W = (torch.randn(sum(classes), D, generator=torch.Generator().manual_seed(1)) * 0.03).to(dev)
        b = torch.zeros(sum(classes), device=dev)
        rec = {"shape": name, "world": world, "payload_bytes": (sum(classes) * (D + 1) + 2 * len(classes)) * 4}
        for transport in ("peer", "nccl", "single"):
            hp = hotpath.HotPath(plan, classes, D, "CrossEntropyLoss", 0.0, device=dev,
                                 comm=Communicator() if transport == "single" else comm,
                                 transport="peer" if transport == "single" else transport)
            for _ in range(5):
                hp.heads_step(emb, W, b, labels)
            torch.cuda.synchronize()
            dist.barrier()
            ts, stamps = [], []
            for it in range(12):
                flush.fill_(1)
                dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                hp.heads_step(emb, W, b, labels)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
                if transport == "peer":
                    out = np.zeros((148, 16), dtype=np.uint64)
                    n = _lib.lib().nkbk_debug_fused_timing(out.ctypes.data_as(ctypes.c_void_p), 148)
                    if n and it >= 2:
                        stamps.append(out[:n].astype(np.int64))
            t = torch.tensor([float(np.median(ts))], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            rec[f"{transport}_step_us"] = float(t.item())
            if transport == "peer":
                assert comm.peer_status() == 0
                if stamps:
                    st = np.stack(stamps)
                    nspc = float(((st[:, :, 11] - st[:, :, 0]) / np.maximum(st[:, :, 10] - st[:, :, 1], 1)).mean())
                    us = lambda a, b_: float(((st[:, :, b_] - st[:, :, a]) * nspc / 1e3).mean())
                    rec["peer_phases_us"] = {"until_grid_barrier": us(1, 8), "local_reduce_and_push": us(8, 12),
                                             "publish_flags": us(12, 13), "wait_for_peers": us(13, 14),
                                             "rank_ordered_sum_finalize": us(14, 10)}
                    rec["wait_for_peers_max_us"] = float(((st[:, :, 14] - st[:, :, 13]) * nspc / 1e3).max())
                    rec["bytes_pushed_per_rank_per_step"] = rec["payload_bytes"] * (world - 1)
                rec["heads_path"] = {_lib.PATH_FUSED: "k2_fused_step"}.get(ops.heads_last_path(), "separate launches")
        if rank == 0:
            print(json.dumps(rec), flush=True)
    dist.barrier()
    comm.shutdown()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
