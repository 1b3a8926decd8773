"""K2 (heads + loss + gradients) and K3 (argmax + confusion) parity on the GPU.
Tolerances (BASELINE.json): losses / gradients <= 1e-5 relative in fp32,
<= 1e-2 in bf16; predictions and confusion matrices bit-exact."""
import numpy as np
import pytest
import torch

from oracle import heads as oh
from oracle import metrics as om

pytestmark = pytest.mark.gpu

CASES = {
    "focal_g1": (oh.LOSS_FOCAL, 1.0, False),
    "focal_g2": (oh.LOSS_FOCAL, 2.0, False),
    "focal_g0p5_alpha": (oh.LOSS_FOCAL, 0.5, True),
    "focal_g2_alpha": (oh.LOSS_FOCAL, 2.0, True),
    "ce": (oh.LOSS_CE, 0.0, False),
    "ce_weight": (oh.LOSS_CE, 0.0, True),
}


def rel_err(got, exp):
    got, exp = np.asarray(got, dtype=np.float64), np.asarray(exp, dtype=np.float64)
    scale = max(np.abs(exp).max(), 1e-30)
    return np.abs(got - exp).max() / scale


def run_k2(dev, emb, Ws, bs, labels, kind, gamma, cws=None, ignore_index=-100, emb_dtype=torch.float32, demb=True):
    from nkb_classification_b200 import ops
    seg = np.concatenate([[0], np.cumsum([w.shape[0] for w in Ws])]).tolist()
    W_cat = torch.cat([w.float() for w in Ws]).contiguous().to(dev)
    b_cat = torch.cat([b.float() for b in bs]).contiguous().to(dev)
    cw = torch.cat([c.float() for c in cws]).to(dev) if cws is not None else None
    e = emb.to(dev).to(emb_dtype).contiguous()
    B, D = e.shape
    bufs = ops.HeadsBuffers(B, D, seg, dev)
    ops.heads_fwd_loss_bwd(e, W_cat, b_cat, labels.to(dev).contiguous(), bufs, kind, gamma, cw, ignore_index)
    loss = ops.heads_finalize(bufs)
    de = ops.heads_demb(bufs, W_cat, out_dtype=emb_dtype) if demb else None
    torch.cuda.synchronize()
    T = len(Ws)
    dW = [bufs.dW()[seg[t]:seg[t + 1]].cpu().numpy() for t in range(T)]
    db = [bufs.db()[seg[t]:seg[t + 1]].cpu().numpy() for t in range(T)]
    logits = [bufs.logits[:, seg[t]:seg[t + 1]].cpu().numpy() for t in range(T)]
    probs = [bufs.probs[:, seg[t]:seg[t + 1]].cpu().numpy() for t in range(T)]
    return dict(loss=loss.cpu().numpy(), dW=dW, db=db, logits=logits, probs=probs,
                demb=None if de is None else de.float().cpu().numpy(), bufs=bufs, seg=seg)


@pytest.mark.parametrize("case", sorted(CASES))
def test_k2_matches_reference_golden_fp32(cuda_device, golden_dir, case):
    g = np.load(golden_dir / "heads_golden.npz")
    T = 3
    emb = torch.from_numpy(g["emb"])
    Ws = [torch.from_numpy(g[f"W{t}"]) for t in range(T)]
    bs = [torch.from_numpy(g[f"b{t}"]) for t in range(T)]
    alphas = [torch.from_numpy(g[f"alpha{t}"]) for t in range(T)]
    labels = torch.from_numpy(g["labels"])
    kind, gamma, weighted = CASES[case]
    r = run_k2(cuda_device, emb, Ws, bs, labels, kind, gamma, alphas if weighted else None)
    tol = 1e-5
    assert rel_err(r["loss"], g[f"{case}.f64.loss"]) <= tol
    for t in range(T):
        assert rel_err(r["logits"][t], g[f"{case}.f64.logits{t}"]) <= tol
        assert rel_err(r["dW"][t], g[f"{case}.f64.dW{t}"]) <= tol, (case, t)
        assert rel_err(r["db"][t], g[f"{case}.f64.db{t}"]) <= tol
    assert rel_err(r["demb"], g[f"{case}.f64.demb"]) <= tol


@pytest.mark.parametrize("B,D,classes,kind,gamma", [
    (256, 1280, (4, 7, 2), oh.LOSS_FOCAL, 1.0),           # BASELINE config 2
    (1024, 768, (2, 3, 4, 7, 14), oh.LOSS_FOCAL, 1.0),    # BASELINE config 4 (fp32 leg)
    (1024, 768, (2, 3, 4, 7, 14), oh.LOSS_CE, 0.0),
    (32, 512, (10,), oh.LOSS_CE, 0.0),                    # BASELINE config 1
    (517, 2048, (10,), oh.LOSS_CE, 0.0),                  # config 5 width, ragged B
    (130, 260, (3, 70, 2, 5), oh.LOSS_FOCAL, 2.0),        # > 64 classes: two dW passes, ragged D
    (5, 64, (2,) * 9, oh.LOSS_FOCAL, 2.0),                # more than 8 tasks per row group
])
def test_k2_vs_oracle_fp32(cuda_device, B, D, classes, kind, gamma):
    g = torch.Generator().manual_seed(7)
    emb = torch.randn(B, D, generator=g)
    Ws = [torch.randn(c, D, generator=g) * (2.0 / D) ** 0.5 for c in classes]
    bs = [torch.randn(c, generator=g) * 0.05 for c in classes]
    labels = torch.stack([torch.randint(0, c, (B,), generator=g) for c in classes], 1)
    labels[::17, 0] = -100
    exp = oh.heads_loss_fwd_bwd(emb, Ws, bs, labels, kind, gamma, dtype=torch.float64)
    r = run_k2(cuda_device, emb, Ws, bs, labels, kind, gamma)
    tol = 1e-5
    assert rel_err(r["loss"][:-1], [float(x) for x in exp["loss"]]) <= tol
    assert rel_err(r["loss"][-1], float(exp["total"])) <= tol
    for t in range(len(classes)):
        assert rel_err(r["logits"][t], exp["logits"][t].numpy()) <= tol
        assert rel_err(r["probs"][t], exp["probs"][t].numpy()) <= tol
        assert rel_err(r["dW"][t], exp["dW"][t].numpy()) <= tol, t
        assert rel_err(r["db"][t], exp["db"][t].numpy()) <= tol, t
    assert rel_err(r["demb"], exp["demb"].numpy()) <= tol


def test_k2_bf16_embeddings(cuda_device):
    """BASELINE config 4: bf16 embeddings [1024,768], 5 heads -- tolerance 1e-2."""
    g = torch.Generator().manual_seed(7)
    B, D, classes = 1024, 768, (2, 3, 4, 7, 14)
    emb = torch.randn(B, D, generator=g).to(torch.bfloat16)
    Ws = [torch.randn(c, D, generator=g) * (2.0 / D) ** 0.5 for c in classes]
    bs = [torch.zeros(c) for c in classes]
    labels = torch.stack([torch.randint(0, c, (B,), generator=g) for c in classes], 1)
    for kind, gamma in ((oh.LOSS_FOCAL, 1.0), (oh.LOSS_CE, 0.0)):
        exp = oh.heads_loss_fwd_bwd(emb.double(), Ws, bs, labels, kind, gamma, dtype=torch.float64)
        r = run_k2(cuda_device, emb, Ws, bs, labels, kind, gamma, emb_dtype=torch.bfloat16)
        assert rel_err(r["loss"][-1], float(exp["total"])) <= 1e-2
        for t in range(len(classes)):
            assert rel_err(r["dW"][t], exp["dW"][t].numpy()) <= 1e-2
        assert rel_err(r["demb"], exp["demb"].numpy()) <= 1e-2


def test_k2_all_ignored_and_forward_only(cuda_device):
    from nkb_classification_b200 import ops
    g = torch.Generator().manual_seed(3)
    B, D, classes = 64, 128, (3, 4)
    emb = torch.randn(B, D, generator=g)
    Ws = [torch.randn(c, D, generator=g) * 0.1 for c in classes]
    bs = [torch.zeros(c) for c in classes]
    labels = torch.stack([torch.full((B,), -100), torch.randint(0, 4, (B,), generator=g)], 1)
    r = run_k2(cuda_device, emb, Ws, bs, labels, oh.LOSS_FOCAL, 1.0)
    assert r["loss"][0] == 0.0 and np.all(r["dW"][0] == 0) and np.all(r["db"][0] == 0)
    exp = oh.heads_loss_fwd_bwd(emb, Ws, bs, labels, oh.LOSS_FOCAL, 1.0)
    assert rel_err(r["loss"][1], float(exp["loss"][1])) <= 1e-5
    # forward-only buffers (val / inference): no dlogits, grads zeroed, loss still produced
    dev = cuda_device
    seg = [0, 3, 7]
    bufs = ops.HeadsBuffers(B, D, seg, dev, want_grads=False)
    W_cat = torch.cat(Ws).to(dev)
    ops.heads_fwd_loss_bwd(emb.to(dev), W_cat, torch.zeros(7, device=dev), labels.to(dev), bufs, oh.LOSS_FOCAL, 1.0)
    loss = ops.heads_finalize(bufs).cpu().numpy()
    assert rel_err(loss[1], float(exp["loss"][1])) <= 1e-5
    assert float(bufs.dW().abs().max()) == 0.0


def test_k2_is_deterministic(cuda_device):
    g = torch.Generator().manual_seed(5)
    B, D, classes = 1000, 768, (4, 7, 2)
    emb = torch.randn(B, D, generator=g)
    Ws = [torch.randn(c, D, generator=g) * 0.05 for c in classes]
    bs = [torch.zeros(c) for c in classes]
    labels = torch.stack([torch.randint(0, c, (B,), generator=g) for c in classes], 1)
    a = run_k2(cuda_device, emb, Ws, bs, labels, oh.LOSS_FOCAL, 1.0)
    b = run_k2(cuda_device, emb, Ws, bs, labels, oh.LOSS_FOCAL, 1.0)
    assert np.array_equal(a["loss"], b["loss"])
    assert all(np.array_equal(x, y) for x, y in zip(a["dW"], b["dW"]))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_k3_argmax_confusion_exact(cuda_device, dtype):
    from nkb_classification_b200 import ops
    dev = cuda_device
    rng = np.random.default_rng(21)
    classes = (2, 3, 4, 7, 14)
    seg = np.concatenate([[0], np.cumsum(classes)]).tolist()
    B = 5000
    # integer-valued logits: plenty of exact ties, representable in bf16
    z = rng.integers(-3, 4, (B, seg[-1])).astype(np.float32)
    z[5, 0:2] = np.nan
    z[6, 3] = np.nan
    z[7, 5:9] = -np.inf
    labels = np.stack([rng.integers(0, c, B) for c in classes], 1).astype(np.int64)
    labels[::13, 2] = -100
    zt = torch.from_numpy(z).to(dev).to(dtype)
    cm = torch.zeros(ops.confusion_len(seg), dtype=torch.int64, device=dev)
    lab = torch.from_numpy(labels).to(dev)
    pred, _ = ops.argmax_confusion(zt, seg, lab, cm)
    ops.argmax_confusion(zt, seg, lab, cm, want_pred=False)   # accumulates: second pass doubles the counts
    torch.cuda.synchronize()
    pred = pred.cpu().numpy()
    off = 0
    for t, C in enumerate(classes):
        zt_t = z[:, seg[t]:seg[t + 1]]
        exp_pred = torch.from_numpy(zt_t).argmax(-1).numpy()
        assert np.array_equal(exp_pred, om.argmax_first(zt_t))
        assert np.array_equal(pred[:, t], exp_pred)
        exp_cm = om.confusion_matrix(labels[:, t], exp_pred, C)
        got = cm[off: off + C * C].view(C, C).cpu().numpy()
        assert np.array_equal(got, 2 * exp_cm)
        assert om.balanced_accuracy_from_cm(got) == om.balanced_accuracy_from_cm(exp_cm)
        off += C * C


def test_k3_large_class_count_uses_global_atomics(cuda_device):
    from nkb_classification_b200 import ops
    dev = cuda_device
    rng = np.random.default_rng(22)
    C, B = 120, 3000   # 14400 bins > shared-memory histogram capacity
    z = rng.normal(size=(B, C)).astype(np.float32)
    labels = rng.integers(0, C, (B, 1)).astype(np.int64)
    cm = torch.zeros(C * C, dtype=torch.int64, device=dev)
    pred, _ = ops.argmax_confusion(torch.from_numpy(z).to(dev), [0, C], torch.from_numpy(labels).to(dev), cm)
    exp_pred = z.argmax(1)
    assert np.array_equal(pred.cpu().numpy()[:, 0], exp_pred)
    assert np.array_equal(cm.view(C, C).cpu().numpy(), om.confusion_matrix(labels[:, 0], exp_pred, C))


def test_k2_k3_end_to_end_metrics_match_reference_golden(cuda_device, golden_dir):
    """Confusion matrix from K3 -> balanced accuracy equals the reference's sklearn value bit for bit."""
    from nkb_classification_b200 import ops
    g = np.load(golden_dir / "metrics_golden.npz")
    dev = cuda_device
    names = ["a_color", "b_size", "c_kind"]
    zs = [g[f"{n}.logits"] for n in names]
    seg = np.concatenate([[0], np.cumsum([z.shape[1] for z in zs])]).tolist()
    z = torch.from_numpy(np.concatenate(zs, 1)).to(dev)
    labels = torch.from_numpy(np.stack([g[f"{n}.gt"] for n in names], 1).astype(np.int64)).to(dev)
    cm = torch.zeros(ops.confusion_len(seg), dtype=torch.int64, device=dev)
    ops.argmax_confusion(z, seg, labels, cm)
    off = 0
    accs = []
    for n, zt in zip(names, zs):
        C = zt.shape[1]
        acc = om.balanced_accuracy_from_cm(cm[off: off + C * C].view(C, C).cpu().numpy())
        assert acc == float(g[f"{n}.epoch_acc"])
        accs.append(acc)
        off += C * C
    assert float(np.mean(accs)) == float(g["epoch_acc"])


def test_sharded_equals_single_gpu_nccl(cuda_device):
    """K4: needs >= 2 visible GPUs (gpurun --gpus 2); the 2-rank gloo twin of this test runs on CPU."""
    import subprocess
    import sys
    from pathlib import Path
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    script = Path(__file__).resolve().parent / "dist_check.py"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29671", str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST_CHECK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("N,classes,quant", [
    (1000, (3, 2), None),
    (5000, (2, 3, 4, 7, 14), 32),      # config 4 heads, heavily tied scores
    (70000, (10,), None),              # > one (pos tile, neg chunk) item per column, ragged tails
    (33, (2,), 4),
    (1, (3,), None),
])
def test_k5_roc_auc_counts_exact(cuda_device, N, classes, quant):
    """K5 == the integer Mann-Whitney counts of the oracle (which equal sklearn's roc_auc_score, tested on CPU):
    bit-exact int64, including rows with ignored labels (negatives for every class of that task)."""
    from nkb_classification_b200 import ops
    from nkb_classification_b200.metrics import roc_auc_from_counts
    rng = np.random.default_rng(8)
    seg = np.concatenate([[0], np.cumsum(classes)]).tolist()
    z = torch.from_numpy(rng.normal(size=(N, seg[-1])).astype(np.float32))
    probs = torch.cat([z[:, a:b].softmax(-1) for a, b in zip(seg[:-1], seg[1:])], 1)
    if quant:
        probs = (probs * quant).round() / quant
    labels = np.stack([rng.integers(0, c, N) for c in classes], 1)
    if N > 10:
        labels[::13, 0] = -100
    got = ops.roc_auc_counts(probs.to(cuda_device), seg, torch.from_numpy(labels).to(cuda_device)).cpu().numpy()
    exp = om.roc_auc_counts(probs.numpy(), labels, seg)
    assert np.array_equal(got, exp)
    if N >= 1000:
        from sklearn.metrics import roc_auc_score
        auc = roc_auc_from_counts(got[seg[-2]:seg[-1]], classes[-1])
        t = len(classes) - 1
        if classes[-1] > 2:
            e = [roc_auc_score(labels[:, t] == c, probs[:, seg[t] + c].numpy()) for c in range(classes[-1])]
        else:
            e = roc_auc_score(labels[:, t], probs[:, seg[t] + 1].numpy())
        assert np.allclose(auc, e, rtol=0, atol=1e-12)


@pytest.mark.parametrize("emb_dtype,B,D,classes", [
    (torch.float32, 1024, 768, (2, 3, 4, 7, 14)),     # FFMA forward v3, two class passes
    (torch.float32, 4096, 2048, (10,)),                # bench shape
    (torch.float32, 37, 64, (70, 3)),                  # ragged rows, > 64 classes
    (torch.bfloat16, 1024, 768, (2, 3, 4, 7, 14)),    # tcgen05 forward, split-K
    (torch.bfloat16, 4096, 2048, (10,)),
    (torch.bfloat16, 300, 128, (40, 20)),
])
def test_fused_k3_equals_standalone_k3(cuda_device, emb_dtype, B, D, classes):
    """nkbk_heads_step (K3 in K2's epilogue, both forward kernels) gives exactly the predictions and confusion counts
    of nkbk_argmax_confusion run on the logits it emitted, and leaves every other output unchanged."""
    from nkb_classification_b200 import ops
    g = torch.Generator().manual_seed(31)
    seg = np.concatenate([[0], np.cumsum(classes)]).tolist()
    T = len(classes)
    emb = torch.randn(B, D, generator=g).to(cuda_device).to(emb_dtype)
    W = (torch.randn(seg[-1], D, generator=g) * (2.0 / D) ** 0.5).to(cuda_device)
    b = (torch.randn(seg[-1], generator=g) * 0.05).to(cuda_device)
    labels = torch.stack([torch.randint(0, c, (B,), generator=g) for c in classes], 1)
    labels[::11, 0] = -100
    labels = labels.to(cuda_device)
    ncm = ops.confusion_len(seg)
    plain = ops.HeadsBuffers(B, D, seg, cuda_device)
    ops.heads_fwd_loss_bwd(emb, W, b, labels, plain, oh.LOSS_FOCAL, 1.0)
    pred_ref, cm_ref = ops.argmax_confusion(plain.logits, seg, labels, torch.zeros(ncm, dtype=torch.int64, device=cuda_device))
    fused = ops.HeadsBuffers(B, D, seg, cuda_device)
    pred = torch.full((B, T), -7, dtype=torch.int32, device=cuda_device)
    cm = torch.zeros(ncm, dtype=torch.int64, device=cuda_device)
    ops.heads_fwd_loss_bwd(emb, W, b, labels, fused, oh.LOSS_FOCAL, 1.0, out_pred=pred, cm_step=cm)
    ops.heads_fwd_loss_bwd(emb, W, b, labels, fused, oh.LOSS_FOCAL, 1.0, out_pred=pred, cm_step=cm)   # accumulates
    torch.cuda.synchronize()
    assert torch.equal(pred, pred_ref) and torch.equal(cm, 2 * cm_ref)
    assert int(cm_ref.sum()) == B * T - len(range(0, B, 11))
    assert torch.equal(fused.logits, plain.logits) and torch.equal(fused.reduce_buf, plain.reduce_buf)
    assert torch.equal(fused.dlogits, plain.dlogits)


def _torchrun(script, world, port, env_extra=None, timeout=600):
    import os
    import subprocess
    import sys
    from pathlib import Path
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                           "--master-addr", "127.0.0.1", "--master-port", str(port),
                           str(Path(__file__).resolve().parent / script)],
                          capture_output=True, text=True, timeout=timeout, env=env)


def test_peer_allreduce_finalize_world1_equals_finalize(cuda_device):
    """K4' with a single rank is exactly nkbk_heads_finalize: same bits in dW/db/loss, same confusion fold."""
    from nkb_classification_b200 import _lib, ops
    from nkb_classification_b200.parallel import Communicator
    g = torch.Generator().manual_seed(21)
    classes, B, D = (3, 5, 2), 200, 516
    seg = np.concatenate([[0], np.cumsum(classes)]).tolist()
    emb = torch.randn(B, D, generator=g).to(cuda_device)
    W = (torch.randn(sum(classes), D, generator=g) * 0.05).to(cuda_device)
    b = torch.zeros(sum(classes), device=cuda_device)
    labels = torch.stack([torch.randint(0, c, (B,), generator=g) for c in classes], 1).to(cuda_device)
    ncm = ops.confusion_len(seg)
    comm = Communicator()
    assert comm.init_peer(cuda_device, ops.heads_reduce_buf_len(D, sum(classes), len(classes)), ncm)
    try:
        outs = []
        for fused in (False, True, True):      # two fused calls: both slot parities
            bufs = ops.HeadsBuffers(B, D, seg, cuda_device)
            cm, cm_step = torch.zeros(ncm, dtype=torch.int64, device=cuda_device), torch.zeros(ncm, dtype=torch.int64, device=cuda_device)
            ops.heads_fwd_loss_bwd(emb, W, b, labels, bufs, oh.LOSS_FOCAL, 2.0, None, -100)
            ops.argmax_confusion(bufs.logits, seg, labels, cm_step)
            (ops.peer_allreduce_finalize if fused else ops.heads_finalize)(bufs, cm, cm_step)
            torch.cuda.synchronize()
            assert int(cm.sum()) == B * len(classes) and not bool(cm_step.any())
            outs.append((bufs.reduce_buf.clone(), bufs.loss.clone(), cm))
        for o in outs[1:]:
            assert torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1]) and torch.equal(o[2], outs[0][2])
        assert comm.peer_status() == 0
        bufs_big = ops.HeadsBuffers(B, 2 * D, seg, cuda_device)      # larger than the inbox: refused, not truncated
        with pytest.raises(ValueError, match="capacity"):
            ops.peer_allreduce_finalize(bufs_big)
    finally:
        comm.shutdown()
    assert _lib.lib().nkbk_peer_world() == 0


def test_peer_allreduce_two_ranks_on_one_gpu(cuda_device):
    """K4' protocol (push into cudaIpc-mapped inboxes, flags, rank-ordered sums, finalize) with two PROCESSES sharing
    this GPU; the N-GPU NVLink run of the same script is test_peer_allreduce_multi_gpu."""
    r = _torchrun("peer_check.py", 2, 29673, {"NKBK_PEER_ONE_GPU": "1"}, timeout=900)
    assert r.returncode == 0 and "PEER_CHECK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_peer_allreduce_multi_gpu(cuda_device):
    """K4' over NVLink: needs >= 2 visible GPUs (gpurun --gpus 2)."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else (4 if n < 8 else 8)
    r = _torchrun("peer_check.py", world, 29674, timeout=900)
    assert r.returncode == 0 and "PEER_CHECK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("B,D,classes", [
    (1024, 768, (2, 3, 4, 7, 14)),     # BASELINE config 4: N padded to 32
    (4096, 2048, (10,)),               # config 5 width in bf16: N padded to 16
    (300, 128, (40, 20)),              # ragged last tile, N padded to 64
])
def test_k2_tcgen05_forward_bf16(cuda_device, B, D, classes, monkeypatch):
    """bf16 embeddings take the tcgen05 / TMEM / TMA forward.  Against an fp64 oracle that sees the SAME bf16-rounded
    operands the fp32-accumulated logits agree to ~1e-5; the FFMA path (tensor cores disabled) is the cross-check."""
    g = torch.Generator().manual_seed(9)
    emb = torch.randn(B, D, generator=g).to(torch.bfloat16)
    Ws = [(torch.randn(c, D, generator=g) * (2.0 / D) ** 0.5) for c in classes]
    bs = [torch.randn(c, generator=g) * 0.05 for c in classes]
    labels = torch.stack([torch.randint(0, c, (B,), generator=g) for c in classes], 1)
    labels[::19, 0] = -100
    Wb = [w.to(torch.bfloat16).double() for w in Ws]
    exp = oh.heads_loss_fwd_bwd(emb.double(), Wb, bs, labels, oh.LOSS_FOCAL, 1.0, dtype=torch.float64)
    # training calls of qualifying shapes take the one-launch fused kernel (exact fp32 weights, tests/test_k2_fused_gpu.py);
    # the tcgen05 forward serves forward-only calls and the shapes the fused kernel refuses -- exercised here directly
    from nkb_classification_b200 import _lib, ops
    monkeypatch.setenv("NKBK_DISABLE_FUSED_HEADS", "1")
    r = run_k2(cuda_device, emb, Ws, bs, labels, oh.LOSS_FOCAL, 1.0, emb_dtype=torch.bfloat16)
    assert ops.heads_last_path() == _lib.PATH_TC_FWD, "the tcgen05 forward did not run (silent FFMA fallback)"
    for t in range(len(classes)):
        assert rel_err(r["logits"][t], exp["logits"][t].numpy()) <= 2e-5, t
        assert rel_err(r["probs"][t], exp["probs"][t].numpy()) <= 2e-5, t
        assert rel_err(r["dW"][t], exp["dW"][t].numpy()) <= 1e-4, t
    assert rel_err(r["loss"][-1], float(exp["total"])) <= 2e-5
    monkeypatch.setenv("NKBK_DISABLE_TCGEN05", "1")
    f = run_k2(cuda_device, emb, Ws, bs, labels, oh.LOSS_FOCAL, 1.0, emb_dtype=torch.bfloat16)
    assert ops.heads_last_path() == _lib.PATH_FFMA_FWD
    monkeypatch.delenv("NKBK_DISABLE_TCGEN05")
    for t in range(len(classes)):  # FFMA path keeps W in fp32: differs by the bf16 rounding of W only
        assert rel_err(r["logits"][t], f["logits"][t]) <= 1e-2


def test_k2_tcgen05_weight_pack_is_cached_by_version(cuda_device, monkeypatch):
    """Forward-only bf16 calls (validation / inference) re-pack the bf16 copy of the weights only when the weights'
    version changes: launches drop from 3 (pack, forward, loss sums) to 2, results stay identical, and an in-place
    update of the weights (new torch version counter) is picked up."""
    from nkb_classification_b200 import _lib, ops
    dev = cuda_device
    g = torch.Generator().manual_seed(4)
    B, D, classes = 512, 768, (2, 3, 4, 7, 14)
    seg = np.concatenate([[0], np.cumsum(classes)]).tolist()
    emb = torch.randn(B, D, generator=g).to(dev).to(torch.bfloat16)
    W = (torch.randn(sum(classes), D, generator=g) * 0.05).to(dev)
    b = torch.zeros(sum(classes), device=dev)
    bufs = ops.HeadsBuffers(B, D, seg, dev, want_grads=False)
    pred = torch.empty((B, len(classes)), dtype=torch.int32, device=dev)

    def call():
        n0 = _lib.launch_count()
        ops.heads_fwd_loss_bwd(emb, W, b, None, bufs, out_pred=pred)
        assert ops.heads_last_path() == _lib.PATH_TC_FWD
        torch.cuda.synchronize()
        return _lib.launch_count() - n0, bufs.logits.clone(), pred.clone()

    n1, z1, p1 = call()
    n2, z2, p2 = call()
    assert n2 == n1 - 1 and torch.equal(z1, z2) and torch.equal(p1, p2)     # second call: no re-pack
    W.mul_(-1.0)                                                             # in-place update bumps W._version
    n3, z3, p3 = call()
    assert n3 == n1 and torch.allclose(z3, -z1, atol=1e-6)                   # re-packed: the new weights are in use
    n4, z4, _ = call()
    assert n4 == n1 - 1 and torch.equal(z3, z4)


@pytest.mark.parametrize("emb_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("classes", [[10], [4, 3, 6, 9, 8]])
def test_forward_only_k3_contention(cuda_device, emb_dtype, classes):
    """Validation / inference steps (forward-only kernels: exact-fp32 FFMA forward, tcgen05 forward for bf16): the
    confusion counts are warp-aggregated (match.any + one 64-bit atomic per warp and bin).  B = 65 536 rows that all hit
    ONE bin per task must give exact counts and cost about what uniformly spread labels cost."""
    from nkb_classification_b200 import ops
    B, D = 65536, 256
    dev = cuda_device
    seg = np.concatenate([[0], np.cumsum(classes)]).tolist()
    T, NC = len(classes), sum(classes)
    g = torch.Generator().manual_seed(3)
    W = torch.zeros(NC, D)
    for t in range(T):
        W[seg[t] + 1, :] = 1.0                       # class 1 of every task wins for positive embeddings
    emb_same = (torch.rand(B, D, generator=g) + 0.5).to(emb_dtype)
    emb_uni = torch.randn(B, D, generator=g).to(emb_dtype)
    lab_same = torch.ones(B, T, dtype=torch.int64)
    lab_uni = torch.stack([torch.randint(0, c, (B,), generator=g) for c in classes], 1).contiguous()
    bufs = ops.HeadsBuffers(B, D, seg, dev, want_grads=False)
    ncm = ops.confusion_len(seg)
    Wd, bd = W.to(dev), torch.zeros(NC, device=dev)
    pred = torch.empty((B, T), dtype=torch.int32, device=dev)

    def timed(emb, labels):
        e, l = emb.to(dev), labels.to(dev)
        cs = torch.zeros(ncm, dtype=torch.int64, device=dev)
        for _ in range(3):
            ops.heads_fwd_loss_bwd(e, Wd, bd, l, bufs, oh.LOSS_CE, 0.0, out_pred=pred, cm_step=cs)
        cs.zero_()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(10):
            ops.heads_fwd_loss_bwd(e, Wd, bd, l, bufs, oh.LOSS_CE, 0.0, out_pred=pred, cm_step=cs)
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / 10, cs.cpu().numpy(), pred.cpu().numpy().copy()

    t_same, cm_same, pred_same = timed(emb_same, lab_same)
    t_uni, cm_uni, pred_uni = timed(emb_uni, lab_uni)
    assert (pred_same == 1).all()
    off = 0
    for t, C in enumerate(classes):
        exp = np.zeros((C, C), dtype=np.int64)
        exp[1, 1] = 10 * B
        assert np.array_equal(cm_same[off: off + C * C].reshape(C, C), exp)
        ref = om.confusion_matrix(lab_uni[:, t].numpy(), pred_uni[:, t], C)       # the kernel's own predictions, counted on the host
        assert np.array_equal(cm_uni[off: off + C * C].reshape(C, C), 10 * ref)
        off += C * C
    assert t_same <= 1.25 * t_uni + 0.002, (t_same, t_uni)
