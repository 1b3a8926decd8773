"""Run under torchrun with N >= 2 GPUs: `engine.train_epoch` / `val_epoch` of a sharded run (cfg.communicator) against the
same epochs on one GPU over the whole data.

  * every rank feeds its contiguous shard of the frames (parallel.shard_frames) through get_dataset -> train_epoch with a
    TRAINABLE backbone (engine.sync_backbone_grads) and the fused heads step (exchange in K2's epilogue);
  * after each epoch the parameters are bit-identical on all ranks and within 1e-5 of the single-GPU run;
  * the epoch results are global and identical on every rank: gathered per-sample lists (rank order), all-reduced
    confusion matrix, ROC-AUC pair counts taken over the gathered epoch == single-GPU confusion / counts exactly.
Prints ENGINE_DIST_OK on rank 0."""
import os
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


class Backbone(torch.nn.Module):
    def __init__(self, D=64):
        super().__init__()
        self.num_features = D
        self.scale = torch.nn.Parameter(torch.ones(D))

    def forward(self, x):
        return x[:, :, ::4, ::4].reshape(x.shape[0], -1)[:, : self.num_features].float().contiguous() * self.scale


def main():
    from nkb_classification_b200 import dataset as D, engine, logging as LG, losses, model as M, transforms as T
    from nkb_classification_b200.parallel import Communicator, shard_frames
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    comm = Communicator().init_from_torch_distributed(dev)

    rng = np.random.default_rng(0)
    n_frames, per = 10, 12                       # 10 frames over N ranks: uneven shards for N = 4, 8
    frames = [rng.integers(0, 256, (120, 160, 3), dtype=np.uint8) for _ in range(n_frames)]
    boxes, fidx = [], []
    for f in range(n_frames):
        for _ in range(per):
            w, h = int(rng.integers(16, 80)), int(rng.integers(16, 80))
            x0, y0 = int(rng.integers(0, 160 - w)), int(rng.integers(0, 120 - h))
            boxes.append((x0, y0, x0 + w, y0 + h))
            fidx.append(f)
    boxes, fidx = np.array(boxes), np.array(fidx)
    labels = rng.integers(0, 4, len(fidx))
    classes = ["a", "b", "c", "d"]
    pipe = [T.Resize(32, 32), T.Normalize(), T.ToTensorV2()]

    def loader_for(mask, f0):
        idx = np.flatnonzero(mask)
        used = np.unique(fidx[idx])
        ds = D.InMemoryFrames([frames[f] for f in used], fidx[idx] - f0, labels[idx], boxes=boxes[idx], classes=classes)
        # one batch per epoch step = the rank's whole shard (global batch = all samples), in order
        return D.get_dataset({"type": "InMemoryFrames", "dataset": ds, "batch_size": int(mask.sum()), "shuffle": False,
                              "device": str(dev)}, pipe)

    def make(comm_):
        torch.manual_seed(0)
        model = M.get_model({"task": "single", "model": Backbone(), "pretrained": False, "backbone_dropout": 0.0,
                             "classifier_dropout": 0.0, "classifier_initialization": "kaiming_normal_"}, classes, dev)
        crit = {"task": "single", "type": "CrossEntropyLoss"}
        cfg = SimpleNamespace(task="single", target_column="label", enable_mixed_presicion=False, log_gradients=False,
                              disable_tqdm=True, criterion=crit, communicator=comm_)
        opt = torch.optim.SGD(model.parameters(), lr=0.05)
        return model, cfg, losses.get_loss(crit, dev), opt

    f_lo, f_hi, mask = shard_frames(fidx, n_frames, rank, world)
    sharded = loader_for(mask, f_lo)
    full = loader_for(np.ones(len(fidx), dtype=bool), 0)
    scaler = torch.amp.GradScaler("cuda", enabled=False)

    ms, cfgs, crits, opts = make(comm)
    m1, cfg1, crit1, opt1 = make(None)
    for epoch in range(3):
        res = engine.train_epoch(ms, sharded, opts, None, scaler, crits, dev, cfgs, LG.BaseLogger(cfgs, classes))
        ref = engine.train_epoch(m1, full, opt1, None, scaler, crit1, dev, cfg1, LG.BaseLogger(cfg1, classes))
        for (n_, a), (_, b) in zip(ms.named_parameters(), m1.named_parameters()):
            err = float(((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).detach())
            assert err <= 1e-5, (epoch, n_, err)
            g = [torch.empty_like(a) for _ in range(world)]
            dist.all_gather(g, a.detach().contiguous())
            assert all(torch.equal(g[0], x) for x in g), f"{n_} differs between ranks after epoch {epoch}"
        # global, rank-independent epoch results
        assert np.array_equal(res["confusion"], ref["confusion"]), "confusion counts differ from the single-GPU epoch"
        assert res["ground_truth"] == ref["ground_truth"], "gathered ground truth is not the global epoch in rank order"
        assert res["predictions"] == ref["predictions"]
        assert np.allclose(np.asarray(res["confidences"]), np.asarray(ref["confidences"]), rtol=1e-5, atol=1e-7)
        assert np.array_equal(res["roc_auc_counts"][:, 1:], ref["roc_auc_counts"][:, 1:])      # P, Q per class
        assert np.abs(res["roc_auc_counts"][:, 0] - ref["roc_auc_counts"][:, 0]).max() <= 2    # (ties of near-equal probabilities)
        assert np.allclose(res["running_loss"], ref["running_loss"], rtol=1e-5)
    vres = engine.val_epoch(ms, sharded, crits, dev, cfgs, LG.BaseLogger(cfgs, classes))
    vref = engine.val_epoch(m1, full, crit1, dev, cfg1, LG.BaseLogger(cfg1, classes))
    assert np.array_equal(vres["confusion"], vref["confusion"]) and vres["ground_truth"] == vref["ground_truth"]
    dist.barrier()
    comm.shutdown()
    dist.destroy_process_group()
    if rank == 0:
        print("ENGINE_DIST_OK world=%d" % world, flush=True)


if __name__ == "__main__":
    main()
