"""The engine's training step must not synchronise the host with the device (the reference does 3T + T + 1 syncs per
step in BaseLogger.log_iter, engine.py:53 / logging.py:261-281): every step after the first runs under
torch.cuda.set_sync_debug_mode("error").  Also: TrainPbar.update_loss, and the backbone-gradient sum of sharded runs."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class Stub(torch.nn.Module):
    def __init__(self, D=64, trainable=False):
        super().__init__()
        self.num_features = D
        self.scale = torch.nn.Parameter(torch.ones(D)) if trainable else None

    def forward(self, x):
        f = x[:, :, ::4, ::4].reshape(x.shape[0], -1)[:, : self.num_features].float().contiguous()
        return f * self.scale if self.scale is not None else f


class SyncGuard:
    """Loader wrapper: from the second batch on, any host<->device synchronisation inside the loop body raises."""

    def __init__(self, inner):
        self.inner, self.dataset = inner, inner.dataset

    def __len__(self):
        return len(self.inner)

    def __iter__(self):
        try:
            for i, b in enumerate(self.inner):
                torch.cuda.set_sync_debug_mode("default")      # the loader itself may wait for its own events
                if i >= 1:
                    torch.cuda.set_sync_debug_mode("error")
                yield b
                torch.cuda.set_sync_debug_mode("default")
        finally:
            torch.cuda.set_sync_debug_mode("default")


def make_loader(dev, task, batch=12, n_frames=6, per=8):
    from nkb_classification_b200 import dataset as D, transforms as T
    rng = np.random.default_rng(0)
    frames = [rng.integers(0, 256, (120, 160, 3), dtype=np.uint8) for _ in range(n_frames)]
    boxes, fidx = [], []
    for f in range(n_frames):
        for _ in range(per):
            w, h = int(rng.integers(16, 80)), int(rng.integers(16, 80))
            x0, y0 = int(rng.integers(0, 160 - w)), int(rng.integers(0, 120 - h))
            boxes.append((x0, y0, x0 + w, y0 + h))
            fidx.append(f)
    n = len(fidx)
    if task == "single":
        labels, classes = rng.integers(0, 4, n), ["a", "b", "c", "d"]
    else:
        labels = {"color": rng.integers(0, 3, n), "size": rng.integers(0, 2, n)}
        classes = {"color": ["b", "g", "r"], "size": ["l", "s"]}
    ds = D.InMemoryFrames(frames, fidx, labels, boxes=boxes, classes=classes)
    pipe = [T.Resize(32, 32), T.Normalize(), T.ToTensorV2()]
    loader = D.get_dataset({"type": "InMemoryFrames", "dataset": ds, "batch_size": batch, "shuffle": True,
                            "device": str(dev)}, pipe)
    return loader, classes


@pytest.mark.parametrize("task", ["single", "multi"])
def test_train_and_val_steps_do_not_sync(cuda_device, task):
    from nkb_classification_b200 import engine, logging as LG, losses, model as M
    dev = cuda_device
    loader, classes = make_loader(dev, task)
    model = M.get_model({"task": task, "model": Stub(), "pretrained": False, "backbone_dropout": 0.0,
                         "classifier_dropout": 0.0, "classifier_initialization": "kaiming_normal_"}, classes, dev)
    crit = {"task": task, "type": "FocalLoss", "gamma": 1.0}
    cfg = SimpleNamespace(task=task, target_names=sorted(classes) if task == "multi" else None, target_column="label",
                          enable_mixed_presicion=False, log_gradients=False, disable_tqdm=True, criterion=crit)
    criterion = losses.get_loss(crit, dev)
    opt = torch.optim.Adam(model.classifier.parameters(), lr=1e-3)
    scaler = torch.amp.GradScaler("cuda", enabled=False)
    logger = LG.BaseLogger(cfg, classes)
    w0 = [p.detach().clone() for p in model.classifier.parameters()]
    res = engine.train_epoch(model, SyncGuard(loader), opt, None, scaler, criterion, dev, cfg, logger)
    assert any(not torch.equal(a, b) for a, b in zip(w0, model.classifier.parameters())), "the optimizer did not step"
    n = len(loader.dataset)
    if task == "single":
        assert len(res["predictions"]) == n and len(res["running_loss"]) == len(loader)
    else:
        assert len(res["predictions"]["color"]) == n and len(res["running_loss"]["loss"]) == len(loader)
    assert "confusion" in res
    # the example images of the epoch (logging.py:283-285): the first batch, copied asynchronously through pinned memory
    class First:
        def __init__(self, inner):
            self.inner, self.dataset, self.first = inner, inner.dataset, None

        def __len__(self):
            return len(self.inner)

        def __iter__(self):
            for i, (img, t) in enumerate(self.inner):
                if i == 0:
                    self.first = img.clone()
                yield img, t

    cfg.example_images = 5
    keep = First(loader)
    vres = engine.val_epoch(model, keep, criterion, dev, cfg, LG.BaseLogger(cfg, classes))
    assert "confusion" in vres
    assert vres["images"].device.type == "cpu" and torch.equal(vres["images"], keep.first[:5].cpu())
    cfg.example_images = None        # the reference's whole first batch
    vres = engine.val_epoch(model, keep, criterion, dev, cfg, LG.BaseLogger(cfg, classes))
    assert torch.equal(vres["images"], keep.first.cpu())


def test_pbar_update_loss_matches_reference_postfix(cuda_device):
    from nkb_classification_b200.engine import TrainPbar
    cfg = SimpleNamespace(task="multi", show_full_current_loss_in_terminal=True, loss_display_every=2, disable_tqdm=False)
    import io
    bar = TrainPbar(range(4), leave=False, desc="Training", cfg=cfg)
    bar.fp = io.StringIO()
    loss = {"color": torch.tensor(0.25, device=cuda_device), "loss": torch.tensor(0.75, device=cuda_device)}
    bar.update_loss(loss)
    assert bar.postfix is None                      # first step: not yet
    bar.update_loss(loss)
    assert bar.postfix == "loss color: 0.2500, loss loss: 0.7500"     # engine.py:13 of the reference
    cfg.show_full_current_loss_in_terminal = False
    bar.update_loss(loss); bar.update_loss(loss)
    assert bar.postfix == "Loss: 0.7500"            # engine.py:15
    cfg.task = "single"
    bar.update_loss(torch.tensor(1.5, device=cuda_device)); bar.update_loss(torch.tensor(1.5, device=cuda_device))
    assert bar.postfix == "Loss: 1.5000"            # engine.py:17
    bar.close()


@pytest.mark.parametrize("task", ["single", "multi"])
def test_epoch_results_as_numpy_give_the_same_metrics(cuda_device, task):
    """cfg.epoch_results_numpy: the epoch results as numpy arrays instead of the reference's Python lists -- same
    values, and compute_metrics (which wraps them in np.array / hands them to sklearn, as the reference's does) returns
    the same numbers.  Second epochs run as resident epochs (the whole epoch planned at once)."""
    from nkb_classification_b200 import engine, logging as LG, losses, metrics as MX
    dev = cuda_device
    loader, classes = make_loader(dev, task)
    from nkb_classification_b200 import model as M
    model = M.get_model({"task": task, "model": Stub(), "pretrained": False, "backbone_dropout": 0.0,
                         "classifier_dropout": 0.0, "classifier_initialization": "kaiming_normal_"}, classes, dev)
    crit = {"task": task, "type": "CrossEntropyLoss"}
    cfg = SimpleNamespace(task=task, target_names=sorted(classes) if task == "multi" else None, target_column="label",
                          enable_mixed_presicion=False, log_gradients=False, disable_tqdm=True, criterion=crit)
    criterion = losses.get_loss(crit, dev)
    out = {}
    for as_np in (False, True, False):          # (first pass fills the frame cache; the other two are resident epochs)
        cfg.epoch_results_numpy = as_np
        torch.manual_seed(3)
        res = engine.val_epoch(model, loader, criterion, dev, cfg, LG.BaseLogger(cfg, classes))
        out[as_np] = (res, MX.compute_metrics(cfg, res))
    assert loader.stats["resident_epochs"] == 2
    (rl, ml), (rn, mn) = out[False], out[True]
    pick = (lambda r, k: r[k]) if task == "single" else (lambda r, k: r[k]["color"])
    assert isinstance(pick(rl, "predictions"), list) and isinstance(pick(rn, "predictions"), np.ndarray)
    for k in ("predictions", "ground_truth", "confidences"):
        assert np.array_equal(np.asarray(pick(rl, k)), pick(rn, k))
    if task == "single":
        assert ml["epoch_acc"] == mn["epoch_acc"] and ml["epoch_loss"] == mn["epoch_loss"]
        assert np.array_equal(ml["epoch_roc_auc"], mn["epoch_roc_auc"], equal_nan=True)
    else:
        assert ml["epoch_acc"] == mn["epoch_acc"] and ml["color"]["epoch_acc"] == mn["color"]["epoch_acc"]
