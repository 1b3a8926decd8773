"""The reference-facing API on the GPU: criterion objects, the data loader, train/val epochs and
inference(), each checked against the CPU oracle on the same inputs."""
import json
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import heads as oh
from oracle import metrics as om
from oracle import preprocess as opre

pytestmark = pytest.mark.gpu

MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def rel_err(got, exp):
    got, exp = np.asarray(got, np.float64), np.asarray(exp, np.float64)
    return np.abs(got - exp).max() / max(np.abs(exp).max(), 1e-30)


CRIT_CASES = {
    "focal_g1": {"type": "FocalLoss", "gamma": 1},
    "focal_g2": {"type": "FocalLoss"},
    "ce": {"type": "CrossEntropyLoss"},
}


@pytest.mark.parametrize("case", sorted(CRIT_CASES))
def test_multitask_criterion_matches_reference_golden(cuda_device, golden_dir, case):
    """get_loss(...)(pred, true) -> dict of losses + autograd gradients, vs the reference's own losses.py."""
    from nkb_classification_b200 import losses
    g = np.load(golden_dir / "heads_golden.npz")
    dev = cuda_device
    names = ["task0", "task1", "task2"]
    crit = losses.get_loss(dict(CRIT_CASES[case], task="multi"), dev)
    pred = {n: torch.from_numpy(g[f"{case}.f64.logits{t}"]).float().to(dev).requires_grad_(True) for t, n in enumerate(names)}
    true = {n: torch.from_numpy(g["labels"][:, t]) for t, n in enumerate(names)}   # CPU labels, like the DataLoader gives
    out = crit(pred, true)
    assert list(out.keys()) == names + ["loss"]
    out["loss"].backward()
    exp = g[f"{case}.f64.loss"]
    got = [float(out[n]) for n in names] + [float(out["loss"])]
    assert rel_err(got, exp) <= 1e-5
    # d loss / d logits from the oracle (fp64 autograd over the restated reference loss)
    for t, n in enumerate(names):
        z = torch.from_numpy(g[f"{case}.f64.logits{t}"]).requires_grad_(True)
        y = torch.from_numpy(g["labels"][:, t])
        l = oh.focal_loss(z, y, None, CRIT_CASES[case].get("gamma", 2.0)) if "Focal" in CRIT_CASES[case]["type"] \
            else oh.cross_entropy(z, y)
        l.backward()
        assert rel_err(pred[n].grad.cpu().numpy(), z.grad.numpy()) <= 1e-5, n


def test_single_task_criteria_with_weights(cuda_device, golden_dir):
    from nkb_classification_b200 import losses
    g = np.load(golden_dir / "heads_golden.npz")
    dev = cuda_device
    for t in range(3):
        alpha = g[f"alpha{t}"]
        z64 = torch.from_numpy(g[f"ce.f64.logits{t}"])
        y = torch.from_numpy(g["labels"][:, t])
        for cfg, ref in (
            ({"type": "FocalLoss", "gamma": 0.5, "alpha": alpha.tolist(), "task": "single"},
             lambda z: oh.focal_loss(z, y, torch.from_numpy(alpha).double(), 0.5)),
            ({"type": "CrossEntropyLoss", "weight": alpha.tolist(), "task": "single"},
             lambda z: oh.cross_entropy(z, y, torch.from_numpy(alpha).double())),
        ):
            crit = losses.get_loss(cfg, dev)
            z = z64.float().to(dev).requires_grad_(True)
            loss = crit(z, y.to(dev))
            loss.backward()
            zr = z64.clone().requires_grad_(True)
            lr = ref(zr)
            lr.backward()
            assert rel_err(float(loss), float(lr)) <= 1e-5
            assert rel_err(z.grad.cpu().numpy(), zr.grad.numpy()) <= 1e-5
    # bf16 logits (autocast) stay within the bf16 bar
    crit = losses.get_loss({"type": "FocalLoss", "gamma": 1, "task": "single"}, dev)
    zb = torch.from_numpy(g["ce.f64.logits1"]).to(dev).to(torch.bfloat16)
    lb = crit(zb, torch.from_numpy(g["labels"][:, 1]).to(dev))
    ref = oh.focal_loss(zb.double().cpu(), torch.from_numpy(g["labels"][:, 1]), None, 1.0)
    assert rel_err(float(lb), float(ref)) <= 1e-2


class TinyBackbone(torch.nn.Module):
    """Deterministic stand-in for timm.create_model(num_classes=0): [B,3,H,W] -> [B,num_features]."""
    num_features = 16

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(3)
        self.proj = torch.nn.Linear(3 * 4 * 4, 16)
        with torch.no_grad():
            self.proj.weight.copy_(torch.randn(16, 48, generator=g) * 0.2)
            self.proj.bias.zero_()

    def forward(self, x):
        return self.proj(torch.nn.functional.adaptive_avg_pool2d(x, 4).flatten(1))


def make_csv_dataset(root, n=22, seed=5):
    import cv2
    import pandas as pd
    rng = np.random.default_rng(seed)
    rows = []
    for i in range(n):
        h, w = int(rng.integers(30, 70)), int(rng.integers(30, 90))
        cv2.imwrite(str(root / f"i{i}.png"), rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
        rows.append({"path": f"i{i}.png", "fold": "val", "color": ["red", "green", "blue"][int(rng.integers(0, 3))],
                     "size": ["s", "l"][int(rng.integers(0, 2))]})
    pd.DataFrame(rows).to_csv(root / "ann.csv", index=False)
    return rows


def oracle_images(root, rows, plan_kw):
    import cv2
    plan = opre.Plan(mean=MEAN, std=STD, **plan_kw)
    out = []
    for r in rows:
        img = cv2.cvtColor(cv2.imread(str(root / r["path"])), cv2.COLOR_BGR2RGB)   # dataset.py:523-524
        out.append(opre.preprocess_crop(img, (0, 0, img.shape[1], img.shape[0]), plan, "cv2")[1])
    return np.stack(out)


def cfg_ns(**kw):
    base = dict(task="multi", target_names=["color", "size"], enable_mixed_presicion=False, log_gradients=False,
                show_full_current_loss_in_terminal=False, disable_tqdm=True,
                criterion={"task": "multi", "type": "FocalLoss", "gamma": 1})
    base.update(kw)
    return SimpleNamespace(**base)


def build(tmp_path, dev, pipeline_kind="letterbox"):
    from nkb_classification_b200 import dataset as D, losses, model as M, transforms as T
    rows = make_csv_dataset(tmp_path)
    if pipeline_kind == "letterbox":   # configs/multitask_config.py:130-140
        pipe = T.Compose([T.LongestMaxSize(32, always_apply=True),
                          T.PadIfNeeded(32, 32, always_apply=True, border_mode=T.BORDER_CONSTANT, value=0),
                          T.Normalize(mean=MEAN, std=STD), T.ToTensorV2()])
        plan_kw = dict(mode=opre.MODE_LETTERBOX, out_h=32, out_w=32, max_size=32)
    else:
        pipe = T.Compose([T.Resize(32, 32), T.Normalize(mean=MEAN, std=STD), T.ToTensorV2()])
        plan_kw = dict(out_h=32, out_w=32)
    data = {"type": "AnnotatedMultitaskDataset", "annotations_file": str(tmp_path / "ann.csv"),
            "target_names": ["size", "color"], "fold": "val", "image_base_dir": str(tmp_path), "batch_size": 8,
            "num_workers": 2, "shuffle": False, "device": str(dev)}
    loader = D.get_dataset(data, pipe)
    classes = loader.dataset.classes
    cfg = cfg_ns()
    model = M.get_model({"task": "multi", "model": TinyBackbone(), "pretrained": False, "backbone_dropout": 0.0,
                         "classifier_dropout": 0.0, "classifier_initialization": "kaiming_normal_"}, classes, dev)
    crit = losses.get_loss(cfg.criterion, dev)
    return rows, loader, classes, cfg, model, crit, plan_kw


def test_loader_yields_oracle_images(cuda_device, tmp_path):
    rows, loader, classes, cfg, model, crit, plan_kw = build(tmp_path, cuda_device)
    exp = oracle_images(tmp_path, rows, plan_kw)
    got, tg = [], []
    for img, target in loader:
        assert img.is_cuda and img.dtype == torch.float32 and set(target) == {"color", "size"}
        got.append(img.cpu().numpy())
        tg.append(target["color"].numpy())
    got = np.concatenate(got)
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))
    c2i = loader.dataset.class_to_idx["color"]
    assert np.concatenate(tg).tolist() == [c2i[r["color"]] for r in rows]


def test_per_head_dropout_matches_the_reference_call_sequence(cuda_device):
    """Training with classifier_dropout > 0: every head draws its own mask from the shared embedding, in ModuleDict
    order (model.py:102-116).  With the same torch seed the fused path consumes the RNG exactly like the reference
    modules do, so logits, losses and all gradients agree with `classifier[t](emb)` + the oracle loss."""
    from nkb_classification_b200 import heads, model as M
    classes = {"color": ["a", "b", "c"], "size": ["s", "l"], "kind": ["x", "y", "z", "w"]}
    net = M.get_model({"task": "multi", "model": TinyBackbone(), "pretrained": False, "backbone_dropout": 0.0,
                       "classifier_dropout": 0.3, "classifier_initialization": "kaiming_normal_"}, classes, cuda_device)
    net.train()
    fused = heads.FusedHeads(net, {"task": "multi", "type": "FocalLoss", "gamma": 1})
    names = fused.names
    g = torch.Generator().manual_seed(3)
    B = 64
    emb0 = torch.randn(B, 16, generator=g).to(cuda_device)
    target = {n: torch.randint(0, len(classes[n]), (B,), generator=g) for n in classes}
    # reference call sequence: Dropout -> Linear per head, oracle focal loss, autograd
    emb_r = emb0.clone().requires_grad_(True)
    torch.manual_seed(77)
    zs = {n: net.classifier[n](emb_r) for n in names}
    ref_losses = [oh.focal_loss(zs[n].double(), target[n].to(cuda_device), None, 1.0) for n in names]
    ref_total = sum(ref_losses)
    params = [p_ for n in names for p_ in net.classifier[n].parameters()]
    ref_grads = torch.autograd.grad(ref_total, [emb_r] + params)
    # fused path, same seed
    emb_f = emb0.clone().requires_grad_(True)
    torch.manual_seed(77)
    out = fused(emb_f, target, train=True)
    got_grads = torch.autograd.grad(out.loss[len(names)], [emb_f] + params)
    torch.cuda.synchronize()
    for t, n in enumerate(names):
        a, b = out.seg[t], out.seg[t + 1]
        e = rel_err(out.logits[:, a:b].cpu().numpy(), zs[n].detach().cpu().numpy())
        assert e <= 1e-5, f"logits of head {n}: rel err {e}"
        e = rel_err(float(out.loss[t]), float(ref_losses[t]))
        assert e <= 1e-5, f"loss of head {n}: rel err {e}"
    assert rel_err(float(out.loss[len(names)]), float(ref_total)) <= 1e-5
    for i, (gg, rg) in enumerate(zip(got_grads, ref_grads)):
        e = rel_err(gg.cpu().numpy(), rg.cpu().numpy())
        assert e <= 1e-5, f"gradient {i} (0 = emb, then weight / bias per head): rel err {e}"
    # confusion counts were accumulated once per head, eval mode takes the single segmented launch again
    assert int(fused.state["cm"].sum()) == B * len(names)
    net.eval()
    out_e = fused(emb0, target, train=False)
    z_e = torch.cat([net.classifier[n](emb0) for n in names], 1)
    assert rel_err(out_e.logits.cpu().numpy(), z_e.detach().cpu().numpy()) <= 1e-5


def test_loader_prefetch_ring_matches_inline(cuda_device, tmp_path):
    """the producer thread + copy stream + slot ring yields exactly what the inline loader yields, in
    order, over more batches than the ring has slots; abandoning an iteration and starting a new one is safe."""
    from nkb_classification_b200 import dataset as D, transforms as T
    make_csv_dataset(tmp_path, n=37)
    pipe = T.Compose([T.Resize(32, 32), T.Normalize(mean=MEAN, std=STD), T.ToTensorV2()])
    data = {"type": "AnnotatedMultitaskDataset", "annotations_file": str(tmp_path / "ann.csv"),
            "target_names": ["size", "color"], "fold": "val", "image_base_dir": str(tmp_path), "batch_size": 4,
            "num_workers": 2, "shuffle": False, "device": str(cuda_device)}
    inline = D.get_dataset(dict(data, prefetch=0), pipe)
    ring = D.get_dataset(dict(data, prefetch=2), pipe)
    assert inline.prefetch == 0 and ring.prefetch == 2 and len(ring._slots) == 4 and len(ring) == 10
    a = [(img.cpu().numpy(), t["color"].numpy()) for img, t in inline]
    for _ in zip(range(3), ring):      # abandon an iteration early: the producer must stop cleanly
        pass
    side = torch.cuda.Stream(device=cuda_device)
    with torch.cuda.stream(side):      # consumer on a non-default stream
        b = [(img.cpu().numpy(), t["color"].numpy()) for img, t in ring]
    assert len(a) == len(b) == 10
    for (ia, ta), (ib, tb) in zip(a, b):
        assert np.array_equal(ia.view(np.uint32), ib.view(np.uint32)) and np.array_equal(ta, tb)


def test_train_loader_with_fused_augmentations(cuda_device, tmp_path):
    """get_dataset(train_data, train_pipeline) with the reference's train-time ops (configs/singletask_config.py:
    162-201, all of them): the loader draws per-sample parameters from its rng and K1 applies them;
    replaying the same random stream through the oracle (cv2.flip / cv2.LUT / slice fill) gives the same bits."""
    import random
    import cv2
    from nkb_classification_b200 import dataset as D, transforms as T
    rows = make_csv_dataset(tmp_path)
    pipe = T.Compose([T.LongestMaxSize(32, always_apply=True),
                      T.PadIfNeeded(32, 32, always_apply=True, border_mode=T.BORDER_CONSTANT, value=0),
                      T.HorizontalFlip(p=0.5), T.VerticalFlip(p=0.5),
                      T.RandomBrightnessContrast(brightness_limit=(-0.2, 0.2), contrast_limit=(0.1, -0.5), p=0.5),
                      T.HueSaturationValue(hue_shift_limit=0, sat_shift_limit=10, val_shift_limit=50, p=0.5),
                      T.CoarseDropout(max_holes=4, min_holes=1, max_height=0.2, min_height=0.05, max_width=0.2,
                                      min_width=0.05, fill_value=[0, 0.5, 1], p=0.5),
                      T.Normalize(mean=MEAN, std=STD), T.ToTensorV2()])
    data = {"type": "AnnotatedMultitaskDataset", "annotations_file": str(tmp_path / "ann.csv"),
            "target_names": ["size", "color"], "fold": "val", "image_base_dir": str(tmp_path), "batch_size": 8,
            "num_workers": 2, "shuffle": False, "device": str(cuda_device)}
    loader = D.get_dataset(data, pipe)
    loader.aug_rng = random.Random(31)
    got = np.concatenate([img.cpu().numpy() for img, _ in loader])
    # oracle: same draws, batch by batch (the loader draws once per batch of 8, in sample order)
    replay = random.Random(31)
    plan = opre.Plan(mean=MEAN, std=STD, mode=opre.MODE_LETTERBOX, out_h=32, out_w=32, max_size=32)
    exp, changed = [], 0
    for i in range(0, len(rows), 8):
        chunk = rows[i:i + 8]
        b = loader.plan.draw(len(chunk), replay)
        for k, r in enumerate(chunk):
            f = int(b.flags[k])
            a = opre.AugSample(hflip=bool(f & 1), vflip=bool(f & 2), bc=bool(f & 4), alpha=float(b.alpha[k]),
                               beta=float(b.brightness[k]), holes=[tuple(int(v) for v in h) for h in b.holes[k, :f >> 8]],
                               fill=b.fill, hsv=tuple(float(v) for v in b.hsv_shift[k]) if (f & 8) else None)
            changed += int(f != 0)
            img = cv2.cvtColor(cv2.imread(str(tmp_path / r["path"])), cv2.COLOR_BGR2RGB)
            exp.append(opre.preprocess_crop(img, (0, 0, img.shape[1], img.shape[0]), plan, "cv2", aug=a)[1])
    exp = np.stack(exp)
    assert changed > len(rows) // 2
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))


def oracle_epoch(tmp_path, rows, classes, model_state, plan_kw, c2i, gamma=1.0):
    """CPU restatement of val_epoch: oracle images -> same tiny backbone (CPU, fp64) -> oracle heads / loss."""
    imgs = torch.from_numpy(oracle_images(tmp_path, rows, plan_kw)).double()
    bb = TinyBackbone().double()
    bb.load_state_dict({k[len("emb_model."):]: v.double() for k, v in model_state.items() if k.startswith("emb_model.")})
    names = list(classes.keys())
    Ws = [model_state[f"classifier.{n}.1.weight"].double() for n in names]
    bs = [model_state[f"classifier.{n}.1.bias"].double() for n in names]
    labels = torch.tensor([[c2i[n][r[n]] for n in names] for r in rows])
    return imgs, bb, names, Ws, bs, labels


def test_val_epoch_matches_oracle(cuda_device, tmp_path):
    from nkb_classification_b200 import engine, logging as L, metrics as Mx
    rows, loader, classes, cfg, model, crit, plan_kw = build(tmp_path, cuda_device)
    state = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    logger = L.BaseLogger(cfg, classes)
    res = engine.val_epoch(model, loader, crit, cuda_device, cfg, logger)
    assert set(res) >= {"running_loss", "confidences", "predictions", "ground_truth", "images", "confusion"}
    imgs, bb, names, Ws, bs, labels = oracle_epoch(tmp_path, rows, classes, state, plan_kw, loader.dataset.class_to_idx)
    with torch.no_grad():
        emb = bb(imgs)
    ref_all = oh.heads_loss_fwd_bwd(emb, Ws, bs, labels, oh.LOSS_FOCAL, 1.0)
    for t, n in enumerate(names):
        z = ref_all["logits"][t].numpy()
        assert res["ground_truth"][n] == labels[:, t].tolist()
        assert res["predictions"][n] == om.argmax_first_fast(z).tolist()
        assert rel_err(res["confidences"][n], ref_all["probs"][t].numpy()) <= 1e-5
        cm = om.confusion_matrix(labels[:, t].numpy(), om.argmax_first_fast(z), z.shape[1])
        assert np.array_equal(res["confusion"][n], cm)
    # per-batch running losses (3 batches of 8, 8, 6)
    for bi, (lo, hi) in enumerate(((0, 8), (8, 16), (16, 22))):
        r = oh.heads_loss_fwd_bwd(emb[lo:hi], Ws, bs, labels[lo:hi], oh.LOSS_FOCAL, 1.0)
        assert rel_err(res["running_loss"]["loss"][bi], float(r["total"])) <= 1e-5
    m = Mx.compute_metrics(cfg, res)
    exp_acc = np.mean([om.balanced_accuracy_from_cm(res["confusion"][n]) for n in cfg.target_names])
    assert m["epoch_acc"] == exp_acc
    # ROC-AUC came from K5's device-side pair counts; the reference's sklearn route on the same lists agrees
    assert set(res["roc_auc_counts"]) == set(names)
    sk = Mx.compute_metrics(cfg, {k: v for k, v in res.items() if k != "roc_auc_counts"})
    for n in names:
        assert np.allclose(m[n]["epoch_roc_auc"], sk[n]["epoch_roc_auc"], rtol=0, atol=1e-12, equal_nan=True)


def test_train_epoch_updates_heads_like_the_oracle(cuda_device, tmp_path):
    """One SGD epoch through train_epoch (autograd into the heads AND the backbone) == the same steps done with
    torch fp64 on the CPU from the oracle's images."""
    from nkb_classification_b200 import engine, logging as L, utils
    rows, loader, classes, cfg, model, crit, plan_kw = build(tmp_path, cuda_device, pipeline_kind="stretch")
    state = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    opt = utils.get_optimizer(model, {"type": "sgd", "lr": 0.05})
    scaler = torch.amp.GradScaler("cuda", enabled=False)
    logger = L.BaseLogger(cfg, classes)
    res = engine.train_epoch(model, loader, opt, None, scaler, crit, cuda_device, cfg, logger)
    assert len(res["running_loss"]["loss"]) == 3

    imgs, bb, names, Ws, bs, labels = oracle_epoch(tmp_path, rows, classes, state, plan_kw, loader.dataset.class_to_idx)
    params = [p for p in bb.parameters()] + Ws + bs
    for p in Ws + bs:
        p.requires_grad_(True)
    ref_opt = torch.optim.SGD(params, lr=0.05)
    ref_losses = []
    for lo, hi in ((0, 8), (8, 16), (16, 22)):
        ref_opt.zero_grad()
        emb = bb(imgs[lo:hi])
        total = sum(oh.focal_loss(torch.nn.functional.linear(emb, Ws[t], bs[t]), labels[lo:hi, t], None, 1.0)
                    for t in range(len(names)))
        total.backward()
        ref_opt.step()
        ref_losses.append(float(total))
    assert rel_err(res["running_loss"]["loss"], ref_losses) <= 2e-5
    new_state = model.state_dict()
    for t, n in enumerate(names):
        assert rel_err(new_state[f"classifier.{n}.1.weight"].cpu().numpy(), Ws[t].detach().numpy()) <= 2e-5
        assert rel_err(new_state[f"classifier.{n}.1.bias"].cpu().numpy(), bs[t].detach().numpy()) <= 2e-5
    assert rel_err(new_state["emb_model.proj.weight"].cpu().numpy(), bb.proj.weight.detach().numpy()) <= 2e-5
    torch.jit.script(model)  # still scriptable after the heads were packed (train.py:66)


def test_reference_train_config_end_to_end(cuda_device, tmp_path):
    """The whole drop-in flow on the reference's own training recipe (configs/multitask_config.py shape): letterbox
    + every train-time augmentation fused in K1, classifier_dropout 0.1 (per-head masks), focal gamma 1, bf16
    autocast, Adam + cosine schedule, weighted sampling; then validation, compute_metrics and inference().  Checks
    that it learns a separable toy problem and that every reference-facing structure has the reference's keys."""
    import random
    import cv2
    import pandas as pd
    from nkb_classification_b200 import dataset as D, engine, inference as I, logging as L, losses, metrics as Mx
    from nkb_classification_b200 import model as M, transforms as T, utils
    rng = np.random.default_rng(3)
    rows = []
    for i in range(96):   # colour decides "color", brightness decides "size": learnable from pooled pixels
        color, size = int(rng.integers(0, 3)), int(rng.integers(0, 2))
        img = rng.integers(0, 60, (int(rng.integers(40, 70)), int(rng.integers(40, 90)), 3), dtype=np.uint8)
        img[..., 2 - color] += np.uint8(90 + 80 * size)     # cv2 writes BGR
        cv2.imwrite(str(tmp_path / f"t{i}.png"), img)
        rows.append({"path": f"t{i}.png", "fold": "train" if i < 72 else "val", "color": ["red", "green", "blue"][color],
                     "size": ["s", "l"][size]})
    pd.DataFrame(rows).to_csv(tmp_path / "ann.csv", index=False)
    geo = [T.LongestMaxSize(32, always_apply=True),
           T.PadIfNeeded(32, 32, always_apply=True, border_mode=T.BORDER_CONSTANT, value=0)]
    tail = [T.Normalize(mean=MEAN, std=STD), T.ToTensorV2()]
    train_pipeline = T.Compose(geo + [
        T.HorizontalFlip(p=0.5), T.VerticalFlip(p=0.5),
        T.RandomBrightnessContrast(brightness_limit=(-0.05, 0.05), contrast_limit=(0.05, -0.05), p=0.5),
        T.HueSaturationValue(hue_shift_limit=0, sat_shift_limit=10, val_shift_limit=10, p=0.5),
        T.CoarseDropout(max_holes=2, min_holes=1, max_height=0.2, min_height=0.05, max_width=0.2, min_width=0.05,
                        fill_value=[0, 0.5, 1], p=0.5)] + tail)
    val_pipeline = T.Compose(geo + tail)
    base = {"type": "AnnotatedMultitaskDataset", "annotations_file": str(tmp_path / "ann.csv"),
            "target_names": ["size", "color"], "image_base_dir": str(tmp_path), "batch_size": 24, "num_workers": 2,
            "device": str(cuda_device)}
    random.seed(11)
    torch.manual_seed(11)
    train_loader = D.get_dataset(dict(base, fold="train", shuffle=True, weighted_sampling=True), train_pipeline)
    val_loader = D.get_dataset(dict(base, fold="val", shuffle=False), val_pipeline)
    # per-sample parameter draws in albumentations' call order from one seeded stream (the loader's default draws a whole
    # batch at once with numpy; this 72-image toy problem is sensitive to which augmentations it happens to see)
    train_loader.aug_rng = random.Random(11)
    classes = train_loader.dataset.classes
    cfg = cfg_ns(enable_mixed_presicion=True, target_names=["color", "size"])
    net = M.get_model({"task": "multi", "model": TinyBackbone(), "pretrained": False, "backbone_dropout": 0.0,
                       "classifier_dropout": 0.1, "classifier_initialization": "kaiming_normal_"}, classes, cuda_device)
    opt = utils.get_optimizer(net, {"type": "adam", "lr": 0.05, "backbone_lr": 0.02})
    sched = utils.get_scheduler(opt, {"type": "cosine", "n_epochs": 8})
    crit = losses.get_loss(cfg.criterion, cuda_device)
    scaler = torch.amp.GradScaler("cuda", enabled=False)
    first = last = None
    for epoch in range(8):
        log = L.BaseLogger(cfg, classes)
        res = engine.train_epoch(net, train_loader, opt, sched, scaler, crit, cuda_device, cfg, log)
        m = Mx.compute_metrics(cfg, res)
        loss = float(np.mean(res["running_loss"]["loss"]))
        assert np.isfinite(loss)
        first = loss if first is None else first
        last = loss
    assert last < 0.6 * first, (first, last)
    vres = engine.val_epoch(net, val_loader, crit, cuda_device, cfg, L.BaseLogger(cfg, classes))
    vm = Mx.compute_metrics(cfg, vres)
    assert set(vm) == {"color", "size", "loss", "epoch_acc"} and set(vm["color"]) == {"epoch_acc", "epoch_roc_auc", "epoch_loss"}
    assert vm["epoch_acc"] > 0.8, vm
    assert set(vres["confusion"]) == {"color", "size"} and int(vres["confusion"]["color"].sum()) == 24
    # inference() on the validation images reproduces val_epoch's predictions
    (tmp_path / "inf").mkdir()
    for r in rows[72:]:
        (tmp_path / "inf" / r["path"]).write_bytes((tmp_path / r["path"]).read_bytes())
    inf_loader = D.get_inference_dataset({"folder_path": str(tmp_path / "inf"), "batch_size": 16, "num_workers": 2,
                                          "device": str(cuda_device)}, val_pipeline)
    I.inference(net, inf_loader, classes, str(tmp_path), cuda_device, cfg)
    out = pd.read_csv(tmp_path / "inference_annotations.csv")
    assert len(out) == 24 and {"color", "size", "path"} <= set(out.columns)
    by_path = {Path(p_).name: (c, s_) for p_, c, s_ in zip(out["path"], out["color"], out["size"])}
    idx2 = {n: {v: k for k, v in train_loader.dataset.class_to_idx[n].items()} for n in ("color", "size")}
    for r, pc, ps in zip(rows[72:], vres["predictions"]["color"], vres["predictions"]["size"]):
        assert by_path[r["path"]] == (idx2["color"][pc], idx2["size"][ps])


def test_inference_writes_reference_csv(cuda_device, tmp_path):
    import pandas as pd
    from nkb_classification_b200 import dataset as D, inference as I, model as M, transforms as T
    dev = cuda_device
    img_dir = tmp_path / "imgs"
    img_dir.mkdir()
    rows = make_csv_dataset(img_dir, n=9, seed=8)
    pipe = T.Compose([T.LongestMaxSize(32), T.PadIfNeeded(32, 32, border_mode=T.BORDER_CONSTANT, value=0),
                      T.Normalize(mean=MEAN, std=STD), T.ToTensorV2()])
    (img_dir / "ann.csv").unlink()
    loader = D.get_inference_dataset({"folder_path": str(img_dir), "batch_size": 4, "num_workers": 0, "device": str(dev)}, pipe)
    classes = {"color": ["blue", "green", "red"], "size": ["l", "s"]}
    model = M.get_model({"task": "multi", "model": TinyBackbone(), "pretrained": False, "backbone_dropout": 0.0,
                         "classifier_dropout": 0.0, "classifier_initialization": "kaiming_normal_"}, classes, dev)
    cfg = SimpleNamespace(task="multi", target_names=["color", "size"], enable_mixed_presicion=False, disable_tqdm=True)
    state = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    I.inference(model, loader, classes, str(tmp_path), dev, cfg)
    table = pd.read_csv(tmp_path / "inference_annotations.csv")
    assert list(table.columns) == ["color", "size", "path"] and len(table) == 9
    order = [{"path": p.split("/")[-1]} for p in table["path"]]
    imgs = torch.from_numpy(oracle_images(img_dir, order, dict(mode=opre.MODE_LETTERBOX, out_h=32, out_w=32, max_size=32))).double()
    bb = TinyBackbone().double()
    bb.load_state_dict({k[len("emb_model."):]: v.double() for k, v in state.items() if k.startswith("emb_model.")})
    with torch.no_grad():
        emb = bb(imgs)
    for n in ("color", "size"):
        z = torch.nn.functional.linear(emb, state[f"classifier.{n}.1.weight"].double(), state[f"classifier.{n}.1.bias"].double())
        assert table[n].tolist() == [classes[n][i] for i in z.argmax(-1).tolist()]


def test_val_epoch_bf16_autocast_uses_tensor_core_heads(cuda_device, tmp_path):
    """enable_mixed_presicion=True: the backbone runs under bf16 autocast, K2 takes bf16 embeddings (tcgen05 path).
    Losses / probabilities stay within the bf16 bar (1e-2) of the fp64 oracle; the device still matches itself."""
    from nkb_classification_b200 import engine, logging as L
    rows, loader, classes, cfg, model, crit, plan_kw = build(tmp_path, cuda_device)
    cfg.enable_mixed_presicion = True
    state = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    res = engine.val_epoch(model, loader, crit, cuda_device, cfg, L.BaseLogger(cfg, classes))
    imgs, bb, names, Ws, bs, labels = oracle_epoch(tmp_path, rows, classes, state, plan_kw, loader.dataset.class_to_idx)
    with torch.no_grad():
        emb = bb(imgs)
    for bi, (lo, hi) in enumerate(((0, 8), (8, 16), (16, 22))):
        r = oh.heads_loss_fwd_bwd(emb[lo:hi], Ws, bs, labels[lo:hi], oh.LOSS_FOCAL, 1.0)
        assert rel_err(res["running_loss"]["loss"][bi], float(r["total"])) <= 1e-2
    ref_all = oh.heads_loss_fwd_bwd(emb, Ws, bs, labels, oh.LOSS_FOCAL, 1.0)
    for t, n in enumerate(names):
        # probabilities: absolute error of a bf16 backbone + bf16 operands (BASELINE's 1e-2 bar is for losses/gradients)
        assert np.abs(np.asarray(res["confidences"][n]) - ref_all["probs"][t].numpy()).max() <= 2e-2
        cm = res["confusion"][n]
        assert cm.sum() == len(rows) and np.array_equal(cm, om.confusion_matrix(res["ground_truth"][n], res["predictions"][n], cm.shape[0]))


@pytest.mark.parametrize("gamma", [0.5, 1.0, 2.0])
@pytest.mark.parametrize("with_alpha", [False, True])
def test_focal_loss_sum_and_none_match_reference_golden(cuda_device, golden_dir, gamma, with_alpha):
    """FocalLoss(reduction="sum" | "none") (losses.py:87-94 of the reference): values and logit gradients against
    fixtures produced by the reference's own FocalLoss (tests/golden/make_golden.py: make_focal_reductions)."""
    from nkb_classification_b200 import losses
    g = np.load(golden_dir / "focal_reductions_golden.npz")
    z0 = torch.from_numpy(g["z"]).float().to(cuda_device)
    y = torch.from_numpy(g["y"]).to(cuda_device)
    alpha = torch.from_numpy(g["alpha"]).float() if with_alpha else None
    gvec = torch.from_numpy(g["gvec"]).float().to(cuda_device)
    for red in ("sum", "none"):
        key = f"g{gamma}.a{int(with_alpha)}.{red}"
        crit = losses.FocalLoss(alpha=alpha, gamma=gamma, reduction=red).to(cuda_device)
        z = z0.clone().requires_grad_(True)
        val = crit(z, y)
        exp = torch.from_numpy(g[key + ".loss"])
        assert val.shape == exp.shape
        scale = float(exp.abs().max())
        assert float((val.detach().cpu().double() - exp).abs().max()) <= 1e-5 * scale
        (val if red == "sum" else (val * gvec).sum()).backward()
        dz = torch.from_numpy(g[key + ".dz"])
        assert float((z.grad.cpu().double() - dz).abs().max()) <= 1e-5 * float(dz.abs().max())
    assert "reduction='none'" in repr(losses.FocalLoss(reduction="none"))
    all_ignored = losses.FocalLoss(reduction="none").to(cuda_device)(z0, torch.full_like(y, -100))
    assert float(all_ignored) == 0.0


def test_engine_sharded_epochs_equal_single_gpu_multi_gpu(cuda_device):
    """engine.train_epoch / val_epoch with cfg.communicator over N GPUs (tests/engine_dist_check.py under torchrun):
    parameters bit-identical on all ranks and equal to the single-GPU run, epoch results global and rank independent.
    Needs >= 2 visible GPUs (gpurun --gpus 2)."""
    import subprocess
    import sys
    from pathlib import Path
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    here = Path(__file__).resolve().parent
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29679", str(here / "engine_dist_check.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "ENGINE_DIST_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
