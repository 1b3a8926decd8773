"""Run under torchrun with N >= 2 ranks: K4' (nkbk_peer_allreduce_finalize, NVLink peer memory) against the
single-GPU result and against the NCCL transport.

  * N GPUs:             one rank per GPU, torch.distributed over NCCL (gpurun --gpus N)
  * NKBK_PEER_ONE_GPU=1 every rank on cuda:0, rendezvous over gloo: the peers' inboxes are cudaIpc-mapped memory of
                        other PROCESSES on the same device, the kernels of the ranks time-slice -- slow, but it
                        exercises the same push / flag / rank-ordered-sum protocol on a 1-GPU box.

Checks, over several steps (both slot parities, growing step numbers): confusion counts bit-exact, losses / dW / db
within 1e-5 of the unsharded result, every rank bit-identical to rank 0, no wait timed out.
Prints PEER_CHECK_OK on rank 0."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    from nkb_classification_b200 import hotpath, transforms as T
    from nkb_classification_b200.parallel import Communicator, shard_frames
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    one_gpu = os.environ.get("NKBK_PEER_ONE_GPU") == "1"
    dev = torch.device("cuda", 0 if one_gpu else local)
    torch.cuda.set_device(dev)
    comm = Communicator()
    if one_gpu:
        dist.init_process_group("gloo")
        comm.rank, comm.world = rank, world          # no NCCL communicator: two ranks cannot share one device there
    else:
        dist.init_process_group("nccl", device_id=dev)
        comm.init_from_torch_distributed(dev)
    plan = T.compile_pipeline([T.Resize(32, 32), T.Normalize(), T.ToTensorV2()])

    def rel(a, e):
        return float((a - e).abs().max() / e.abs().max().clamp_min(1e-30))

    # (classes, D, frames, crops per frame, loss, gamma): an odd reduce-buffer length, a multi-CTA payload, 1 head
    # (the second case outgrows the inbox sized for the first: exercises the collective re-initialisation)
    # The first three take the fused heads step (the exchange is the epilogue of k2_fused_step), the fourth is too
    # large for it and goes through the separate launches (k4_peer_allreduce_finalize); the last one shards 7 frames
    # unevenly, so the ranks run DIFFERENT grids against the same payload slicing.
    cases = [((3, 5), 132, 6, 5, "FocalLoss", 2.0),
             ((2, 3, 4, 7, 14), 768, 16, 8, "FocalLoss", 1.0),
             ((10,), 2048, 8, 16, "CrossEntropyLoss", 0.0),
             ((3, 70, 2, 5), 1028, 8, 4, "FocalLoss", 2.0),
             ((10,), 2048, 7, 40, "CrossEntropyLoss", 0.0)]
    for ci, (classes, D, F, per, loss, gamma) in enumerate(cases):
        g = torch.Generator().manual_seed(5 + ci)
        B = F * per
        emb = torch.randn(B, D, generator=g)
        labels = torch.stack([torch.randint(0, c, (B,), generator=g) for c in classes], 1).contiguous()
        labels[3, 0] = -100
        W = torch.randn(sum(classes), D, generator=g) * 0.05
        b = torch.randn(sum(classes), generator=g) * 0.01
        fidx = np.repeat(np.arange(F), per)
        _, _, mask = shard_frames(fidx, F, rank, world)
        m = torch.from_numpy(mask)
        e_loc, l_loc = emb[m].contiguous().to(dev), labels[m].contiguous().to(dev)
        Wd, bd = W.to(dev), b.to(dev)
        peer = hotpath.HotPath(plan, classes, D, loss, gamma, device=dev, comm=comm, transport="peer")
        single = hotpath.HotPath(plan, classes, D, loss, gamma, device=dev, comm=Communicator())
        ref = single.heads_step(emb.to(dev), Wd, bd, labels.to(dev))
        torch.cuda.synchronize()
        steps = 5
        for s in range(steps):
            bufs = peer.heads_step(e_loc, Wd, bd, l_loc)
            torch.cuda.synchronize()
            from nkb_classification_b200 import _lib, ops
            assert ops.heads_last_path() == (_lib.PATH_FFMA_FWD if D == 1028 else _lib.PATH_FUSED), "unexpected heads path"
            assert peer.transport == "peer" and comm.peer_active, "peer transport was not established"
            assert comm.peer_status() == 0, f"peer wait timed out: status {comm.peer_status()}"
            assert torch.equal(peer.cm, (s + 1) * single.cm), f"case {ci} step {s}: confusion counts differ"
            assert not bool(peer.cm_step.any()), "step counts were not cleared"
            assert rel(bufs.loss, ref.loss) <= 1e-5, (ci, s, rel(bufs.loss, ref.loss))
            assert rel(bufs.dW(), ref.dW()) <= 1e-5, (ci, s, rel(bufs.dW(), ref.dW()))
            assert rel(bufs.db(), ref.db()) <= 1e-5, (ci, s)
            assert rel(bufs.denom(), ref.denom()) <= 1e-6 and rel(bufs.loss_sum(), ref.loss_sum()) <= 1e-5
            # rank-ordered sums: every rank holds the same bits
            mine = bufs.reduce_buf.detach().cpu()
            gathered = [None] * world
            dist.all_gather_object(gathered, mine)
            assert all(torch.equal(gathered[0], x) for x in gathered), f"case {ci} step {s}: ranks differ bitwise"
        if not one_gpu:   # the NCCL transport gives the same numbers (different summation order: 1e-6)
            nccl = hotpath.HotPath(plan, classes, D, loss, gamma, device=dev, comm=comm, transport="nccl")
            nb = nccl.heads_step(e_loc, Wd, bd, l_loc)
            torch.cuda.synchronize()
            assert torch.equal(nccl.cm, single.cm)
            assert rel(nb.dW(), bufs.dW()) <= 2e-6 and rel(nb.loss, bufs.loss) <= 2e-6
    dist.barrier()
    comm.shutdown()
    dist.destroy_process_group()
    if rank == 0:
        print("PEER_CHECK_OK world=%d one_gpu=%d" % (world, int(one_gpu)), flush=True)


if __name__ == "__main__":
    main()
