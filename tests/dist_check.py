"""Run under torchrun with N >= 2 GPUs: the same global batch, sharded by frame over N ranks and pushed through
K2 -> K3 -> K4 (NCCL all-reduce) -> finalize, must equal the single-GPU result: confusion counts bit-exact,
losses / gradients within 1e-5.  Prints DIST_CHECK_OK on rank 0."""
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
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    comm = Communicator().init_from_torch_distributed(dev)
    plan = T.compile_pipeline([T.Resize(32, 32), T.Normalize(), T.ToTensorV2()])
    classes, D, F, per = (2, 3, 4, 7, 14), 768, 16, 8
    g = torch.Generator().manual_seed(5)
    B = F * per
    emb = torch.randn(B, D, generator=g)
    labels = torch.stack([torch.randint(0, c, (B,), generator=g) for c in classes], 1).contiguous()
    labels[3, 1] = -100
    W = torch.randn(sum(classes), D, generator=g) * 0.05
    b = torch.zeros(sum(classes))
    fidx = np.repeat(np.arange(F), per)
    _, _, mask = shard_frames(fidx, F, rank, world)
    m = torch.from_numpy(mask)
    sharded = hotpath.HotPath(plan, classes, D, "FocalLoss", 1.0, device=dev, comm=comm)
    bufs = sharded.heads_step(emb[m].contiguous().to(dev), W.to(dev), b.to(dev), labels[m].contiguous().to(dev))
    single = hotpath.HotPath(plan, classes, D, "FocalLoss", 1.0, device=dev, comm=Communicator())
    ref = single.heads_step(emb.to(dev), W.to(dev), b.to(dev), labels.to(dev))
    torch.cuda.synchronize()

    def rel(a, e):
        return float((a - e).abs().max() / e.abs().max().clamp_min(1e-30))

    assert torch.equal(sharded.cm, single.cm), "confusion counts differ"
    assert rel(bufs.loss, ref.loss) <= 1e-5, rel(bufs.loss, ref.loss)
    assert rel(bufs.dW(), ref.dW()) <= 1e-5, rel(bufs.dW(), ref.dW())
    assert rel(bufs.db(), ref.db()) <= 1e-5
    # a second step accumulates the epoch confusion totals exactly twice
    sharded.heads_step(emb[m].contiguous().to(dev), W.to(dev), b.to(dev), labels[m].contiguous().to(dev))
    torch.cuda.synchronize()
    assert torch.equal(sharded.cm, 2 * single.cm)
    dist.barrier()
    comm.shutdown()
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_CHECK_OK world=%d" % world, flush=True)


if __name__ == "__main__":
    main()
