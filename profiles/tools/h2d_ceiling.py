#!/usr/bin/env python
"""Host->device copy ceiling of the box at N concurrent ranks: copy only, one cudaMemcpyAsync of a 398 MB pinned buffer
(= the bench's e2e frames per step) per iteration and rank, pinned pages first-touched after the process has been bound
to the CPU cores NVML reports for its GPU.  Says what the e2e leg of bench.py can reach at best.

    python profiles/tools/h2d_ceiling.py                                   (N = 1)
    python -m torch.distributed.run --nproc-per-node N ... profiles/tools/h2d_ceiling.py
Prints one JSON line on rank 0: {"n": N, "per_rank_gbs": [...], "aggregate_gbs": ..., "affinity": [...]}.
"""
import json
import os
import sys

import torch


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    aff = None
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        aff = sorted(os.sched_getaffinity(0))
    except Exception as e:  # pragma: no cover
        aff = repr(e)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    nbytes = 64 * 1080 * 1920 * 3
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.fill_(7)                       # first touch on this process's cores
    devbuf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    iters = 20
    for _ in range(3):
        devbuf.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        devbuf.copy_(host, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    gbs = nbytes * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9
    # D2H of the same size for completeness (the path's D2H is a few hundred bytes per step)
    e0.record()
    for _ in range(5):
        host.copy_(devbuf, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    d2h = nbytes * 5 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    rec = {"rank": rank, "h2d_gbs": gbs, "d2h_gbs": d2h, "cores": aff if not isinstance(aff, list) else f"{aff[0]}-{aff[-1]} ({len(aff)})"}
    if world > 1:
        out = [None] * world
        dist.all_gather_object(out, rec)
        dist.barrier()
        dist.destroy_process_group()
    else:
        out = [rec]
    if rank == 0:
        print(json.dumps({"n": world, "bytes_per_copy": nbytes, "per_rank_h2d_gbs": [r["h2d_gbs"] for r in out],
                          "aggregate_h2d_gbs": sum(r["h2d_gbs"] for r in out), "per_rank_d2h_gbs": [r["d2h_gbs"] for r in out],
                          "cores": [r["cores"] for r in out], "host_cpus": os.cpu_count()}), flush=True)


if __name__ == "__main__":
    main()
