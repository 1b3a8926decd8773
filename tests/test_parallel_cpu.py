"""The N > 1 path on CPU: world_size-2 gloo processes exercise the frame sharding and the
sum-then-normalise contract of the K4 exchange (oracle computes the local unnormalised sums)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import heads as oh
from oracle import metrics as om


def test_shard_range_is_balanced_and_contiguous():
    from nkb_classification_b200.parallel import shard_frames, shard_range
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    fidx = np.repeat(np.arange(10), 3)
    fb, fe, mask = shard_frames(fidx, 10, 1, 4)
    assert (fb, fe) == (3, 6) and mask.sum() == 9 and set(fidx[mask]) == {3, 4, 5}
    with pytest.raises(ValueError):
        shard_range(5, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nkb_classification_b200 import ops
    from nkb_classification_b200.parallel import shard_frames
    g = torch.Generator().manual_seed(11)
    F, per, D, classes = 6, 5, 32, (3, 4)
    T, NC = len(classes), sum(classes)
    B = F * per
    emb = torch.randn(B, D, generator=g, dtype=torch.float64)
    Ws = [torch.randn(c, D, generator=g, dtype=torch.float64) for c in classes]
    bs = [torch.randn(c, generator=g, dtype=torch.float64) for c in classes]
    labels = torch.stack([torch.randint(0, c, (B,), generator=g) for c in classes], 1)
    labels[1, 0] = -100
    fidx = np.repeat(np.arange(F), per)
    _, _, mask = shard_frames(fidx, F, rank, world)
    m = torch.from_numpy(mask)
    part = oh.unnormalised_sums(emb[m], Ws, bs, labels[m], oh.LOSS_FOCAL, 1.0)
    # pack exactly as nkbk.h documents the reduce buffer: [dW NC*D | db NC | loss_sum T | denom T]
    n = int(ops.heads_reduce_buf_len(D, NC, T))
    assert n == NC * D + NC + 2 * T
    buf = torch.cat([torch.cat(part["dW_sum"]).reshape(-1), torch.cat(part["db_sum"]), torch.stack(part["loss_sum"]),
                     torch.stack(part["denom"])])
    assert buf.numel() == n
    z = [torch.nn.functional.linear(emb[m], w, b) for w, b in zip(Ws, bs)]
    cm = np.concatenate([om.confusion_matrix(labels[m][:, t].numpy(), z[t].argmax(-1).numpy(), c).reshape(-1)
                         for t, c in enumerate(classes)])
    cm_t = torch.from_numpy(cm)
    dist.all_reduce(buf)       # the two payloads of nkbk_allreduce_heads
    dist.all_reduce(cm_t)
    denom = buf[NC * D + NC + T:]
    loss = buf[NC * D + NC: NC * D + NC + T] / denom
    seg = np.concatenate([[0], np.cumsum(classes)])
    row_den = torch.repeat_interleave(denom, torch.tensor(classes))
    dW = buf[: NC * D].reshape(NC, D) / row_den[:, None]
    ref = oh.heads_loss_fwd_bwd(emb, Ws, bs, labels, oh.LOSS_FOCAL, 1.0)
    np.testing.assert_allclose(loss.numpy(), [float(x) for x in ref["loss"]], rtol=1e-12)
    np.testing.assert_allclose(dW.numpy(), torch.cat(ref["dW"]).numpy(), rtol=1e-10, atol=1e-14)
    zf = [torch.nn.functional.linear(emb, w, b) for w, b in zip(Ws, bs)]
    cm_ref = np.concatenate([om.confusion_matrix(labels[:, t].numpy(), zf[t].argmax(-1).numpy(), c).reshape(-1)
                             for t, c in enumerate(classes)])
    assert np.array_equal(cm_t.numpy(), cm_ref)
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")


def test_sharded_sums_reassemble_global_result_gloo(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def _peer_fallback_worker(rank, world, port, out_dir):
    """Without a CUDA device nkbk_peer_init fails on every rank: the collective set-up must still terminate on all
    ranks with the same answer (False -> the caller keeps the NCCL transport), and must not be retried."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nkb_classification_b200 import _lib
    from nkb_classification_b200.parallel import Communicator
    comm = Communicator()
    comm.rank, comm.world = rank, world
    ok = comm.init_peer(torch.device("cuda", 0), 1000, 10)
    assert ok is False and comm._peer_failed and not comm.peer_active
    assert comm.init_peer(torch.device("cuda", 0), 1000, 10) is False      # no second collective round
    assert _lib.lib().nkbk_peer_world() == 0
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"peer_ok{rank}"), "w").write("ok")


@pytest.mark.skipif(torch.cuda.is_available(), reason="exercises the no-CUDA fallback of the K4' set-up")
def test_peer_setup_falls_back_collectively_gloo(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_peer_fallback_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"peer_ok{r}").exists() for r in range(world))


def _backbone_sync_worker(rank, world, port, out_dir):
    """engine.sync_backbone_grads: with a trainable backbone the embedding gradient the fused heads return covers the
    LOCAL rows (already divided by the GLOBAL denominators), so the backbone gradients must be SUMMED over the ranks:
    after it every rank holds the gradient of the unsharded batch."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from types import SimpleNamespace
    from nkb_classification_b200.engine import sync_backbone_grads
    g = torch.Generator().manual_seed(3)
    x = torch.randn(12, 5, generator=g, dtype=torch.float64)
    w_head = torch.randn(5, 3, generator=g, dtype=torch.float64)

    def make():
        torch.manual_seed(0)
        m = torch.nn.Module()
        m.emb_model = torch.nn.Linear(5, 5).double()
        m.frozen = torch.nn.Linear(2, 2)
        return m

    full = make()
    (full.emb_model(x) @ w_head).sum().div(12).backward()               # the unsharded mean loss
    mine = make()
    lo, hi = rank * 6, rank * 6 + 6
    (mine.emb_model(x[lo:hi]) @ w_head).sum().div(12).backward()        # local rows, GLOBAL denominator
    comm = SimpleNamespace(world=world, rank=rank)
    assert sync_backbone_grads(mine, comm) == 2                         # weight + bias; the module without grads is skipped
    for a, b in zip(mine.emb_model.parameters(), full.emb_model.parameters()):
        assert torch.allclose(a.grad, b.grad, rtol=1e-12, atol=1e-15)
    assert sync_backbone_grads(mine, SimpleNamespace(world=1, rank=0)) == 0
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"bb_ok{rank}"), "w").write("ok")


def test_backbone_gradients_are_summed_over_ranks_gloo(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_backbone_sync_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"bb_ok{r}").exists() for r in range(world))


def _gather_worker(rank, world, port, out_dir):
    """logging.gather_rows: every rank ends up with every rank's rows in rank order, uneven shards included (the epoch
    results of a sharded run are made global with it: ADVICE round 1, scope of the logger's per-sample lists)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nkb_classification_b200.logging import gather_rows
    full = torch.arange(7 * 3, dtype=torch.float32).reshape(7, 3)
    lo, hi = (0, 4) if rank == 0 else (4, 7)
    got = gather_rows(full[lo:hi].contiguous())
    assert torch.equal(got, full)
    lab = torch.arange(7, dtype=torch.int64).reshape(7, 1)
    assert torch.equal(gather_rows(lab[lo:hi].contiguous()), lab)
    assert torch.equal(gather_rows(full[:0] if rank == 0 else full), full)        # a rank without rows
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"gather_ok{rank}"), "w").write("ok")


def test_epoch_results_gather_rows_gloo(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_gather_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"gather_ok{r}").exists() for r in range(world))
