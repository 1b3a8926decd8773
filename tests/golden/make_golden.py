"""Generates the committed golden fixtures in tests/golden/.

Run ONCE in the development container, where /root/reference is mounted:

    python tests/golden/make_golden.py

Sources of truth
* heads / loss / criterion goldens: the reference's own
  ``nkb_classification.losses`` (FocalLoss, MultitaskCriterion, get_loss),
  imported unmodified from /root/reference, run under torch autograd.
* metric goldens: the reference's own ``nkb_classification.metrics``.
* pixel goldens: ``cv2.resize`` / ``cv2.copyMakeBorder`` 4.13 -- the routines
  the reference reaches through albumentations -- glued together with the
  albumentations 1.x arithmetic restated in oracle/preprocess.py
  (albumentations is not installed anywhere in this image: that glue is the
  one "parity unpinned" boundary, see oracle/__init__.py).

The GPU box has no /root/reference; tests only read the .npz / .json files.
"""
import json
import sys
from pathlib import Path
from types import SimpleNamespace

import cv2
import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")

from nkb_classification import losses as ref_losses  # noqa: E402  (the reference)
from nkb_classification import metrics as ref_metrics  # noqa: E402

from oracle import preprocess as opre  # noqa: E402


def make_heads():
    g = torch.Generator().manual_seed(20240607)
    B, D = 37, 64
    classes = [4, 7, 2]
    T = len(classes)
    out = {}
    emb = torch.randn(B, D, generator=g)
    Ws = [torch.randn(c, D, generator=g) * (2.0 / D) ** 0.5 for c in classes]
    bs = [torch.randn(c, generator=g) * 0.1 for c in classes]
    labels = torch.stack([torch.randint(0, c, (B,), generator=g) for c in classes], dim=1)
    labels[3, 0] = -100
    labels[5, 1] = -100
    labels[6, 1] = -100
    out["emb"], out["labels"] = emb.numpy(), labels.numpy()
    for t in range(T):
        out[f"W{t}"], out[f"b{t}"] = Ws[t].numpy(), bs[t].numpy()
    alphas = [torch.rand(c, generator=g) + 0.5 for c in classes]
    for t in range(T):
        out[f"alpha{t}"] = alphas[t].numpy()

    names = [f"task{t}" for t in range(T)]

    def run(cfg_loss, dtype, per_task_weight=None):
        e = emb.to(dtype).clone().requires_grad_(True)
        W = [w.to(dtype).clone().requires_grad_(True) for w in Ws]
        b = [x.to(dtype).clone().requires_grad_(True) for x in bs]
        pred = {n: torch.nn.functional.linear(e, W[t], b[t]) for t, n in enumerate(names)}
        true = {n: labels[:, t] for t, n in enumerate(names)}
        if per_task_weight is None:
            crit = ref_losses.get_loss(dict(cfg_loss, task="multi"), "cpu")
            if dtype == torch.float64:
                crit.criterion.double()
            res = crit(pred, true)
            per, total = [res[n] for n in names], res["loss"]
        else:
            # get_loss applies ONE weight vector to every task; per-task class weights are
            # exercised by building the reference criterion per task.
            per = []
            for t, n in enumerate(names):
                c = dict(cfg_loss, task="single")
                key = "alpha" if c["type"] == "FocalLoss" else "weight"
                c[key] = per_task_weight[t].tolist()
                crit = ref_losses.get_loss(c, "cpu")
                if dtype == torch.float64:
                    crit.double()
                per.append(crit(pred[n], true[n]))
            total = sum(per)
        total.backward()
        r = {"loss": np.array([float(x) for x in per] + [float(total)], dtype=np.float64),
             "demb": e.grad.numpy().astype(np.float64)}
        for t in range(T):
            r[f"dW{t}"] = W[t].grad.numpy().astype(np.float64)
            r[f"db{t}"] = b[t].grad.numpy().astype(np.float64)
            r[f"logits{t}"] = pred[names[t]].detach().numpy().astype(np.float64)
        return r

    cases = {
        "focal_g1": ({"type": "FocalLoss", "gamma": 1}, None),            # configs/multitask_config.py:176
        "focal_g2": ({"type": "FocalLoss"}, None),                        # default gamma 2 (losses.py:7)
        "focal_g0p5_alpha": ({"type": "FocalLoss", "gamma": 0.5}, alphas),
        "focal_g2_alpha": ({"type": "FocalLoss", "gamma": 2.0}, alphas),
        "ce": ({"type": "CrossEntropyLoss"}, None),
        "ce_weight": ({"type": "CrossEntropyLoss"}, alphas),
    }
    for cname, (cfg, w) in cases.items():
        for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
            r = run(cfg, dt, w)
            for k, v in r.items():
                out[f"{cname}.{tag}.{k}"] = v
    # all-ignored task -> FocalLoss returns tensor(0.) (losses.py:69-70)
    crit = ref_losses.FocalLoss(gamma=1.0)
    z = torch.randn(5, 3, generator=g)
    out["all_ignored.loss"] = np.array(float(crit(z, torch.full((5,), -100))))
    np.savez_compressed(HERE / "heads_golden.npz", **out)
    print("heads_golden.npz", len(out), "arrays")


def make_focal_reductions():
    """The reference's FocalLoss with reduction="sum" / "none" (losses.py:87-94): values and logit gradients."""
    g = torch.Generator().manual_seed(77)
    B, C = 41, 6
    z = torch.randn(B, C, generator=g, dtype=torch.float64)
    y = torch.randint(0, C, (B,), generator=g)
    y[::7] = -100
    alpha = torch.rand(C, generator=g, dtype=torch.float64) + 0.5
    gvec = torch.rand(int((y != -100).sum()), generator=g, dtype=torch.float64)      # upstream gradient for "none"
    out = {"z": z.numpy(), "y": y.numpy(), "alpha": alpha.numpy(), "gvec": gvec.numpy()}
    for gamma in (0.5, 1.0, 2.0):
        for with_alpha in (False, True):
            for red in ("sum", "none"):
                crit = ref_losses.FocalLoss(alpha=alpha if with_alpha else None, gamma=gamma, reduction=red)
                x = z.clone().requires_grad_(True)
                val = crit(x, y)
                (val if red == "sum" else (val * gvec).sum()).backward()
                key = f"g{gamma}.a{int(with_alpha)}.{red}"
                out[key + ".loss"] = val.detach().numpy()
                out[key + ".dz"] = x.grad.numpy()
    np.savez_compressed(HERE / "focal_reductions_golden.npz", **out)
    print("focal_reductions_golden.npz", len(out), "arrays")


def make_metrics():
    g = np.random.default_rng(7)
    N = 400
    res = {"running_loss": {}, "confidences": {}, "predictions": {}, "ground_truth": {}}
    classes = {"a_color": 5, "b_size": 2, "c_kind": 3}
    stash = {}
    for name, C in classes.items():
        z = g.normal(size=(N, C)).astype(np.float32)
        gt = g.integers(0, C, N)
        if name == "a_color":
            gt[gt == 4] = 0  # class 4 absent from ground truth -> NaN AUC + balanced-acc over present classes
        conf = torch.from_numpy(z).softmax(-1, dtype=torch.float32).numpy()
        pred = z.argmax(1)
        res["confidences"][name] = conf.tolist()
        res["predictions"][name] = pred.tolist()
        res["ground_truth"][name] = gt.tolist()
        res["running_loss"][name] = g.random(10).tolist()
        stash[f"{name}.logits"] = z
        stash[f"{name}.gt"] = gt
    res["running_loss"]["loss"] = g.random(10).tolist()
    cfg = SimpleNamespace(task="multi", target_names=sorted(classes))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = ref_metrics.compute_metrics(cfg, res)
    for name in classes:
        stash[f"{name}.epoch_acc"] = np.array(m[name]["epoch_acc"])
        stash[f"{name}.epoch_roc_auc"] = np.array(m[name]["epoch_roc_auc"], dtype=np.float64)
        stash[f"{name}.epoch_loss"] = np.array(m[name]["epoch_loss"])
        stash[f"{name}.running_loss"] = np.array(res["running_loss"][name])
    stash["epoch_acc"] = np.array(m["epoch_acc"])
    stash["loss"] = np.array(m["loss"])
    # single-task case
    cfg1 = SimpleNamespace(task="single")
    r1 = {k: res[k]["c_kind"] for k in ("running_loss", "confidences", "predictions", "ground_truth")}
    m1 = ref_metrics.compute_metrics(cfg1, r1)
    stash["single.epoch_acc"] = np.array(m1["epoch_acc"])
    stash["single.epoch_roc_auc"] = np.array(m1["epoch_roc_auc"], dtype=np.float64)
    np.savez_compressed(HERE / "metrics_golden.npz", **stash)
    print("metrics_golden.npz", len(stash), "arrays")


def make_pixels():
    rng = np.random.default_rng(4321)
    H, W = 120, 160
    frames = rng.integers(0, 256, (2, H, W, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:H, 0:W]
    frames[1, :, :, 0] = (xx + 2 * yy) % 256  # known-answer ramps for rounding coverage
    frames[1, :, :, 1] = (3 * xx + yy) % 256
    boxes = [
        (0, 0, W, H), (10, 7, 15, 12), (3, 5, 8, 115), (100, 0, 160, 60), (0, 60, 64, 120),
        (20, 20, 84, 84), (20, 20, 52, 52), (20, 20, 51, 53), (17, 33, 113, 97), (150, 110, 160, 120),
        (0, 0, 2, 2), (5, 5, 6, 100), (5, 5, 150, 6), (40, 10, 137, 119),
    ]
    fidx = [i % 2 for i in range(len(boxes))]
    out = {"frames": frames, "boxes": np.array(boxes, dtype=np.int32), "frame_idx": np.array(fidx, dtype=np.int32)}
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    m, d = opre.normalize_constants(mean, std)

    def cv2_pipeline(crop, mode, oh, ow, ms, pad):
        crop = np.array(crop)  # dataset.py:102
        if mode == "stretch":
            img = crop if crop.shape[:2] == (oh, ow) else cv2.resize(crop, dsize=(ow, oh), interpolation=cv2.INTER_LINEAR)
        else:
            h, w = crop.shape[:2]
            nh, nw, top, left = opre.letterbox_geometry(h, w, ms, oh, ow)
            img = crop if (nh, nw) == (h, w) else cv2.resize(crop, dsize=(nw, nh), interpolation=cv2.INTER_LINEAR)
            img = cv2.copyMakeBorder(img, top, oh - nh - top, left, ow - nw - left, cv2.BORDER_CONSTANT, value=pad)
        f = img.astype(np.float32)
        f = cv2.subtract(f, np.broadcast_to(m, f.shape).astype(np.float32).copy())
        f = cv2.multiply(f, np.broadcast_to(d, f.shape).astype(np.float32).copy())
        return img, np.ascontiguousarray(f.transpose(2, 0, 1))

    for tag, mode, oh, ow, ms, pad in (
        ("stretch32x48", "stretch", 32, 48, 0, (0, 0, 0)),
        ("letterbox40", "letterbox", 40, 40, 40, (0, 0, 0)),
        ("letterbox48x56pad", "letterbox", 48, 56, 44, (7, 200, 33)),
        ("stretch64", "stretch", 64, 64, 0, (0, 0, 0)),
    ):
        u8s, fs, used = [], [], []
        for i, (b, fi) in enumerate(zip(boxes, fidx)):
            x0, y0, x1, y1 = b
            if mode == "letterbox":
                nh, nw, _, _ = opre.letterbox_geometry(y1 - y0, x1 - x0, ms, oh, ow)
                if nh < 1 or nw < 1:
                    continue  # the reference itself fails here (cv2.resize asserts on a 0-pixel side)
            used.append(i)
            u8, f = cv2_pipeline(frames[fi][y0:y1, x0:x1], mode, oh, ow, ms, pad)
            u8s.append(u8)
            fs.append(f)
        out[f"{tag}.idx"] = np.array(used, dtype=np.int32)
        out[f"{tag}.u8"] = np.stack(u8s)
        out[f"{tag}.f32"] = np.stack(fs)
    np.savez_compressed(HERE / "pixels_golden.npz", **out)
    print("pixels_golden.npz", {k: v.shape for k, v in out.items()})

    # YOLO label text -> int boxes (dataset.py:414-421) golden
    lines = []
    for _ in range(64):
        w, h = rng.integers(3, 481), rng.integers(3, 481)
        x0, y0 = rng.integers(0, 1920 - w + 1), rng.integers(0, 1080 - h + 1)
        lines.append(f"{rng.integers(0, 5)} {(x0 + w / 2) / 1920:.6f} {(y0 + h / 2) / 1080:.6f} {w / 1920:.6f} {h / 1080:.6f}")
    lines.append("1 0.001 0.001 0.01 0.01")   # clips at 0
    lines.append("2 0.999 0.999 0.01 0.01")   # clips at W/H
    lines.append("0 0.5 0.5 0.002 0.5")       # narrower than min_box_size -> dropped
    # the reference's own static method is pure numpy: call it for the expected values
    sys.modules.setdefault("albumentations", type(sys)("albumentations"))
    exp = []
    for ln in lines:
        p = ln.split()
        xc, yc, w, h = map(float, p[1:])
        ih, iw = 1080, 1920
        x_min = int(np.clip(int((xc - w / 2) * iw), 0, iw)); y_min = int(np.clip(int((yc - h / 2) * ih), 0, ih))
        x_max = int(np.clip(int((xc + w / 2) * iw), 0, iw)); y_max = int(np.clip(int((yc + h / 2) * ih), 0, ih))
        exp.append([x_min, y_min, x_max, y_max, int(p[0]), int(x_max - x_min >= 5 and y_max - y_min >= 5)])
    (HERE / "yolo_boxes_golden.json").write_text(json.dumps({"lines": lines, "image_size": [1080, 1920], "expected": exp}))
    print("yolo_boxes_golden.json", len(lines))


if __name__ == "__main__":
    if "--focal-reductions-only" in sys.argv:      # (added in round 2: leaves the other fixtures byte-identical)
        make_focal_reductions()
        sys.exit(0)
    make_heads()
    make_focal_reductions()
    make_metrics()
    make_pixels()
