"""The fused heads step (nkbk_heads_train_step -> k2_fused_step: forward + loss + K3 + dW/db + cross-CTA sum +
finalize in ONE persistent kernel) against the CPU oracle and against the three-kernel path it replaces.
Tolerances (BASELINE.json): losses / gradients <= 1e-5 relative in fp32, <= 1e-2 in bf16; predictions and
confusion matrices bit-exact."""
import numpy as np
import pytest
import torch

from oracle import heads as oh
from oracle import metrics as om

pytestmark = pytest.mark.gpu


def rel_err(got, exp):
    got, exp = np.asarray(got, dtype=np.float64), np.asarray(exp, dtype=np.float64)
    return np.abs(got - exp).max() / max(np.abs(exp).max(), 1e-30)


def make_case(B, D, classes, seed, integer=False, ignore_every=0):
    g = torch.Generator().manual_seed(seed)
    if integer:   # integer-valued operands: fp32 sums exact -> logits (and the argmax) independent of summation order
        emb = torch.randint(-4, 5, (B, D), generator=g).float()
        Ws = [torch.randint(-2, 3, (c, D), generator=g).float() for c in classes]
        bs = [torch.randint(-1, 2, (c,), generator=g).float() for c in classes]
    else:
        emb = torch.randn(B, D, generator=g)
        Ws = [torch.randn(c, D, generator=g) * (2.0 / D) ** 0.5 for c in classes]
        bs = [torch.randn(c, generator=g) * 0.1 for c in classes]
    labels = torch.stack([torch.randint(0, c, (B,), generator=g) for c in classes], 1).contiguous()
    if ignore_every:
        labels[::ignore_every, 0] = -100
    return emb, Ws, bs, labels


def run_train_step(dev, emb, Ws, bs, labels, kind, gamma, cws=None, emb_dtype=torch.float32, steps=1):
    from nkb_classification_b200 import ops
    seg = np.concatenate([[0], np.cumsum([w.shape[0] for w in Ws])]).tolist()
    W_cat = torch.cat(Ws).contiguous().to(dev)
    b_cat = torch.cat(bs).contiguous().to(dev)
    cw = torch.cat(cws).float().to(dev) if cws is not None else None
    e = emb.to(dev).to(emb_dtype).contiguous()
    B, D = e.shape
    T = len(Ws)
    bufs = ops.HeadsBuffers(B, D, seg, dev)
    ncm = ops.confusion_len(seg)
    cm = torch.zeros(ncm, dtype=torch.int64, device=dev)
    cm_step = torch.zeros(ncm, dtype=torch.int64, device=dev)
    pred = torch.full((B, T), -7, dtype=torch.int32, device=dev)
    for _ in range(steps):
        ops.heads_train_step(e, W_cat, b_cat, labels.to(dev), bufs, kind, gamma, cw, -100, out_pred=pred, cm_total=cm,
                             cm_step=cm_step)
    path = ops.heads_last_path()
    torch.cuda.synchronize()
    return dict(bufs=bufs, seg=seg, cm=cm.cpu().numpy(), cm_step=cm_step.cpu().numpy(), pred=pred.cpu().numpy(),
                loss=bufs.loss.cpu().numpy(), path=path, W_cat=W_cat)


def check_against_oracle(r, emb, Ws, bs, labels, kind, gamma, cws=None, tol=1e-5, steps=1):
    T = len(Ws)
    seg = r["seg"]
    ref = oh.heads_loss_fwd_bwd(emb, Ws, bs, labels, kind, gamma, cws, need_demb=False)
    exp_loss = np.array([float(x) for x in ref["loss"]] + [float(ref["total"])])
    assert rel_err(r["loss"], exp_loss) <= tol
    bufs = r["bufs"]
    for t in range(T):
        a, b = seg[t], seg[t + 1]
        assert rel_err(bufs.logits[:, a:b].cpu().numpy(), ref["logits"][t].numpy()) <= tol
        assert rel_err(bufs.dW()[a:b].cpu().numpy(), ref["dW"][t].numpy()) <= tol, t
        # the bias gradient is part of the same gradient vector: its error is measured on the head's gradient scale
        # (sum_i dlogit_ij nearly cancels for a trained / symmetric head, so |db| alone is not a meaningful scale)
        scale = max(float(ref["dW"][t].abs().max()), float(ref["db"][t].abs().max()), 1e-30)
        assert np.abs(bufs.db()[a:b].cpu().numpy().astype(np.float64) - ref["db"][t].numpy()).max() <= tol * scale, t
        assert np.abs(bufs.probs[:, a:b].cpu().numpy() - ref["probs"][t].numpy()).max() <= 1e-5


@pytest.mark.parametrize("B,D,classes,kind,gamma", [
    (4096, 2048, (10,), oh.LOSS_CE, 0.0),                 # BASELINE config 5 (the bench shape): 28 rows per CTA
    (512, 2048, (10,), oh.LOSS_CE, 0.0),                  # config 5 strong-scaled to 8 GPUs: 4-row tiles
    (517, 2048, (10,), oh.LOSS_CE, 0.0),                  # ragged B
    (256, 1280, (4, 7, 2), oh.LOSS_FOCAL, 1.0),           # config 2
    (1024, 768, (2, 3, 4, 7, 14), oh.LOSS_FOCAL, 1.0),    # config 4 (fp32 leg): 30 classes = two forward passes
    (128, 768, (2, 3, 4, 7, 14), oh.LOSS_CE, 0.0),        # config 4 strong-scaled to 8 GPUs
    (32, 512, (10,), oh.LOSS_CE, 0.0),                    # config 1
    (1, 64, (3,), oh.LOSS_FOCAL, 2.0),                    # one row
    (9, 516, (3, 5, 2), oh.LOSS_FOCAL, 2.0),              # D not a multiple of 128, partial second tile
    (1187, 768, (3,), oh.LOSS_CE, 0.0),                   # config 3 width, odd B (rows per CTA 9: tiles of 8 + 1)
    (8200, 256, (5, 4), oh.LOSS_FOCAL, 0.5),              # many tiles per CTA, powf branch
])
def test_fused_step_vs_oracle_fp32(cuda_device, B, D, classes, kind, gamma):
    from nkb_classification_b200 import _lib
    emb, Ws, bs, labels = make_case(B, D, classes, seed=B + D, ignore_every=7 if B > 8 else 0)
    r = run_train_step(cuda_device, emb, Ws, bs, labels, kind, gamma)
    assert r["path"] == _lib.PATH_FUSED, "shape was expected to take the fused kernel"
    check_against_oracle(r, emb, Ws, bs, labels, kind, gamma)
    assert not r["cm_step"].any()


def test_fused_step_weighted_and_alpha(cuda_device):
    classes = (4, 7, 2)
    emb, Ws, bs, labels = make_case(300, 1280, classes, seed=3, ignore_every=5)
    g = torch.Generator().manual_seed(9)
    cws = [torch.rand(c, generator=g) + 0.5 for c in classes]
    for kind, gamma in ((oh.LOSS_CE, 0.0), (oh.LOSS_FOCAL, 2.0)):
        r = run_train_step(cuda_device, emb, Ws, bs, labels, kind, gamma, cws)
        check_against_oracle(r, emb, Ws, bs, labels, kind, gamma, cws)


@pytest.mark.parametrize("B,D,classes", [(1024, 768, (2, 3, 4, 7, 14)), (4096, 2048, (10,)), (100, 512, (10,))])
def test_fused_step_bf16_embeddings(cuda_device, B, D, classes):
    from nkb_classification_b200 import _lib
    emb, Ws, bs, labels = make_case(B, D, classes, seed=11)
    emb = emb.to(torch.bfloat16).float()   # the oracle sees the same rounded inputs
    r = run_train_step(cuda_device, emb, Ws, bs, labels, oh.LOSS_FOCAL, 1.0, emb_dtype=torch.bfloat16)
    assert r["path"] == _lib.PATH_FUSED
    check_against_oracle(r, emb, Ws, bs, labels, oh.LOSS_FOCAL, 1.0, tol=1e-5)   # exact fp32 accumulation of bf16 inputs


@pytest.mark.parametrize("B,D,classes", [(4096, 2048, (10,)), (777, 768, (2, 3, 4, 7, 14)), (37, 132, (3, 5))])
def test_fused_k3_exact_and_accumulating(cuda_device, B, D, classes):
    """Integer-valued operands: the logits are exact whatever the summation order, so predictions and confusion
    counts must equal the oracle's bit for bit; three steps accumulate three times the counts."""
    emb, Ws, bs, labels = make_case(B, D, classes, seed=5, integer=True, ignore_every=9)
    labels[1::13, -1] = 10 ** 6        # out-of-range label: ignored by the loss, not counted
    r = run_train_step(cuda_device, emb, Ws, bs, labels, oh.LOSS_CE, 0.0, steps=3)
    seg, off = r["seg"], 0
    for t, C in enumerate(classes):
        logits = (emb.double() @ Ws[t].double().T + bs[t].double()).float().numpy()
        pred = om.argmax_first(logits)
        assert np.array_equal(r["pred"][:, t], pred)
        y = labels[:, t].numpy()
        keep = (y >= 0) & (y < C)
        exp = om.confusion_matrix(y[keep], pred[keep], C)
        assert np.array_equal(r["cm"][off: off + C * C].reshape(C, C), 3 * exp)
        off += C * C
    assert not r["cm_step"].any()


def test_fused_equals_three_kernel_path(cuda_device, monkeypatch):
    """Same step through k2_fused_step and through forward_v3 + dw + finalize: integer counts identical, floating
    point within summation-order noise, and nkbk_heads_last_path reports which one ran."""
    from nkb_classification_b200 import _lib
    emb, Ws, bs, labels = make_case(1000, 1280, (4, 7, 2), seed=17, ignore_every=6)
    a = run_train_step(cuda_device, emb, Ws, bs, labels, oh.LOSS_FOCAL, 1.0)
    monkeypatch.setenv("NKBK_DISABLE_FUSED_HEADS", "1")
    b = run_train_step(cuda_device, emb, Ws, bs, labels, oh.LOSS_FOCAL, 1.0)
    assert a["path"] == _lib.PATH_FUSED and b["path"] == _lib.PATH_FFMA_FWD
    assert np.array_equal(a["cm"], b["cm"]) and np.array_equal(a["pred"], b["pred"])
    assert rel_err(a["loss"], b["loss"]) <= 2e-6
    assert rel_err(a["bufs"].reduce_buf.cpu().numpy(), b["bufs"].reduce_buf.cpu().numpy()) <= 2e-6
    assert rel_err(a["bufs"].dlogits.cpu().numpy(), b["bufs"].dlogits.cpu().numpy()) <= 2e-6


def test_fused_falls_back_for_large_heads(cuda_device):
    """NC * D beyond the register accumulators: the call still succeeds through the separate launches."""
    from nkb_classification_b200 import _lib
    classes = (3, 70, 2, 5)
    emb, Ws, bs, labels = make_case(130, 1028, classes, seed=23)
    r = run_train_step(cuda_device, emb, Ws, bs, labels, oh.LOSS_FOCAL, 2.0)
    assert r["path"] == _lib.PATH_FFMA_FWD
    check_against_oracle(r, emb, Ws, bs, labels, oh.LOSS_FOCAL, 2.0)
    assert int(r["cm"].sum()) == 130 * len(classes)
    # ... while the same heads over a narrower embedding fit
    emb, Ws, bs, labels = make_case(130, 260, classes, seed=23)
    r = run_train_step(cuda_device, emb, Ws, bs, labels, oh.LOSS_FOCAL, 2.0)
    assert r["path"] == _lib.PATH_FUSED
    check_against_oracle(r, emb, Ws, bs, labels, oh.LOSS_FOCAL, 2.0)


def test_fused_is_deterministic_and_sums_mode_matches(cuda_device):
    """Two runs give the same bits; nkbk_heads_step (sums only) followed by nkbk_heads_finalize gives the same bits
    as the one-launch nkbk_heads_train_step (both run k2_fused_step, in different modes)."""
    from nkb_classification_b200 import ops
    emb, Ws, bs, labels = make_case(2048, 2048, (10,), seed=31)
    a = run_train_step(cuda_device, emb, Ws, bs, labels, oh.LOSS_CE, 0.0)
    b = run_train_step(cuda_device, emb, Ws, bs, labels, oh.LOSS_CE, 0.0)
    assert torch.equal(a["bufs"].reduce_buf, b["bufs"].reduce_buf) and np.array_equal(a["loss"], b["loss"])
    dev = cuda_device
    seg = a["seg"]
    bufs = ops.HeadsBuffers(2048, 2048, seg, dev)
    cm_step = torch.zeros(100, dtype=torch.int64, device=dev)
    cm = torch.zeros(100, dtype=torch.int64, device=dev)
    ops.heads_fwd_loss_bwd(emb.to(dev), a["W_cat"], torch.cat(bs).to(dev), labels.to(dev), bufs, oh.LOSS_CE, 0.0,
                           cm_step=cm_step, out_pred=torch.empty((2048, 1), dtype=torch.int32, device=dev))
    ops.heads_finalize(bufs, cm, cm_step)
    torch.cuda.synchronize()
    assert torch.equal(bufs.reduce_buf, a["bufs"].reduce_buf) and torch.equal(bufs.loss, a["bufs"].loss)
    assert np.array_equal(cm.cpu().numpy(), a["cm"])


@pytest.mark.parametrize("classes", [(10,), (2, 3, 4, 7, 14)])
def test_fused_k3_contention(cuda_device, classes):
    """B = 65 536 rows that all carry the same label and all predict the same class (one confusion bin per task takes
    every count): bit-exact, and -- the counts being privatised per CTA in shared memory -- no slower than the
    uniform-label case (within 10 %)."""
    from nkb_classification_b200 import ops
    B, D = 65536, 256
    dev = cuda_device
    seg = np.concatenate([[0], np.cumsum(classes)]).tolist()
    T, NC = len(classes), sum(classes)
    g = torch.Generator().manual_seed(1)
    W = torch.zeros(NC, D)
    for t in range(T):
        W[seg[t] + 1, :] = 1.0                       # class 1 of every task wins for positive embeddings
    emb_same = torch.rand(B, D, generator=g) + 0.5
    emb_uni = torch.randn(B, D, generator=g)
    lab_same = torch.ones(B, T, dtype=torch.int64)
    lab_uni = torch.stack([torch.randint(0, c, (B,), generator=g) for c in classes], 1).contiguous()
    bufs = ops.HeadsBuffers(B, D, seg, dev)
    ncm = ops.confusion_len(seg)
    Wd, bd = W.to(dev), torch.zeros(NC, device=dev)
    pred = torch.empty((B, T), dtype=torch.int32, device=dev)

    def timed(emb, labels):
        e, l = emb.to(dev), labels.to(dev)
        cm, cs = torch.zeros(ncm, dtype=torch.int64, device=dev), torch.zeros(ncm, dtype=torch.int64, device=dev)
        for _ in range(3):
            ops.heads_train_step(e, Wd, bd, l, bufs, oh.LOSS_CE, 0.0, out_pred=pred, cm_total=cm, cm_step=cs)
        cm.zero_()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(10):
            ops.heads_train_step(e, Wd, bd, l, bufs, oh.LOSS_CE, 0.0, out_pred=pred, cm_total=cm, cm_step=cs)
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / 10, cm.cpu().numpy()

    t_same, cm_same = timed(emb_same, lab_same)
    t_uni, cm_uni = timed(emb_uni, lab_uni)
    off = 0
    for t, C in enumerate(classes):
        exp = np.zeros((C, C), dtype=np.int64)
        exp[1, 1] = 10 * B
        assert np.array_equal(cm_same[off: off + C * C].reshape(C, C), exp)
        assert int(cm_uni[off: off + C * C].sum()) == 10 * B
        off += C * C
    assert t_same <= 1.10 * t_uni + 0.002, (t_same, t_uni)


def test_k1_as_programmatic_dependent_of_the_heads_step(cuda_device):
    """nkbk_k1_overlap_previous: K1 of the next batch enqueued behind the fused heads step as a programmatic dependent
    (it may start while that kernel still runs, on the SMs it leaves free) -- in a loop of several steps both kernels
    must produce exactly what the ordinary stream order produces, and the thread's launch option is restored."""
    from nkb_classification_b200 import _lib, hotpath, transforms as T
    dev = cuda_device
    plan = T.compile_pipeline([T.Resize(224, 224), T.Normalize(), T.ToTensorV2()])
    rng = np.random.default_rng(2)
    frames = torch.from_numpy(rng.integers(0, 256, (4, 360, 640, 3), dtype=np.uint8)).to(dev)
    n = 512
    x0, y0 = rng.integers(0, 400, n), rng.integers(0, 200, n)
    boxes = torch.from_numpy(np.stack([x0, y0, x0 + rng.integers(8, 240, n), y0 + rng.integers(8, 160, n)], 1).astype(np.int32)).to(dev)
    fidx = torch.from_numpy(rng.integers(0, 4, n).astype(np.int32)).to(dev)
    g = torch.Generator().manual_seed(0)
    D, classes = 2048, (10,)
    emb = torch.randn(n, D, generator=g).to(dev)
    W, b = (torch.randn(10, D, generator=g) * 0.03).to(dev), torch.zeros(10, device=dev)
    labels = torch.randint(0, 10, (n, 1), generator=g).to(dev)

    def run(overlap):
        hp = hotpath.HotPath(plan, classes, D, "CrossEntropyLoss", 0.0, device=dev)
        outs = []
        for _ in range(4):
            bufs = hp.heads_step(emb, W, b, labels, train=True)
            if overlap:
                hp.mark_heads_done()
            img = hp.preprocess(frames, boxes, fidx, overlap_previous=overlap)
            outs.append((img.clone(), bufs.dW().clone(), bufs.loss.clone()))
        torch.cuda.synchronize()
        return outs, hp.cm.clone()

    ref, cm_ref = run(False)
    got, cm_got = run(True)
    assert _lib.lib().nkbk_k1_overlap_previous(0) == 0          # the option was restored after every launch
    for (ia, wa, la), (ib, wb, lb) in zip(ref, got):
        assert torch.equal(ia, ib) and torch.equal(wa, wb) and torch.equal(la, lb)
    assert torch.equal(cm_ref, cm_got)


def test_heads_one_launch_option_selects_the_separate_kernels(cuda_device):
    """nkbk_heads_one_launch(0) / HotPath.one_launch = False: the heads step runs as its separate kernels (forward, dW,
    finalize) instead of k2_fused_step -- same losses / gradients (1e-5), identical confusion counts and predictions,
    nkbk_heads_last_path says which ran, and the thread's option is restored after every call."""
    from nkb_classification_b200 import _lib, hotpath, ops, transforms as T
    dev = cuda_device
    plan = T.compile_pipeline([T.Resize(32, 32), T.Normalize(), T.ToTensorV2()])
    g = torch.Generator().manual_seed(4)
    B, D, classes = 700, 768, (2, 3, 4, 7, 14)
    emb = torch.randn(B, D, generator=g).to(dev)
    W, b = (torch.randn(sum(classes), D, generator=g) * 0.05).to(dev), torch.zeros(sum(classes), device=dev)
    labels = torch.stack([torch.randint(0, c, (B,), generator=g) for c in classes], 1).contiguous().to(dev)
    out = {}
    for one in (True, False):
        hp = hotpath.HotPath(plan, classes, D, "FocalLoss", 1.0, device=dev)
        hp.one_launch = one
        bufs = hp.heads_step(emb, W, b, labels, train=True)
        path = ops.heads_last_path()
        torch.cuda.synchronize()
        out[one] = (bufs.loss.clone(), bufs.dW().clone(), bufs.db().clone(), hp.cm.clone(), bufs.pred.clone(), path)
        assert _lib.lib().nkbk_heads_one_launch(1) == 1            # restored
    assert out[True][5] == _lib.PATH_FUSED and out[False][5] == _lib.PATH_FFMA_FWD
    rel = lambda a, e: float((a - e).abs().max() / e.abs().max().clamp_min(1e-30))
    assert rel(out[False][0], out[True][0]) <= 1e-5 and rel(out[False][1], out[True][1]) <= 1e-5
    assert rel(out[False][2], out[True][2]) <= 1e-5
    assert torch.equal(out[False][3], out[True][3]) and torch.equal(out[False][4], out[True][4])
