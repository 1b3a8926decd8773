"""Host-side mirror of the reference API: everything that needs no GPU."""
import json
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import metrics as om
from oracle import preprocess as opre


def test_boxes_match_golden(golden_dir):
    from nkb_classification_b200 import boxes
    g = json.loads((golden_dir / "yolo_boxes_golden.json").read_text())
    size = tuple(g["image_size"])
    xywhn = np.array([[float(v) for v in ln.split()[1:]] for ln in g["lines"]])
    batch = boxes.bbox_xywhn2xyxy_batch(xywhn, size)
    for ln, exp, row in zip(g["lines"], g["expected"], batch):
        p = ln.split()
        assert list(boxes.bbox_xywhn2xyxy(*map(float, p[1:]), size)) == exp[:4]
        assert row.tolist() == exp[:4]
        assert boxes.check_boxes_sizes(*exp[:4]) == bool(exp[5])
    kept = boxes.parse_yolo_label_lines(g["lines"], size)
    assert [list(b) for b, _ in kept] == [e[:4] for e in g["expected"] if e[5]]
    # detector boxes: normalised xyxy * (W, H) then astype(int)   (metrics/det_cls_val.py:231-236)
    d = np.array([[0.1004, 0.2, 0.5, 0.9999], [0.0, 0.0, 1.0, 1.0]])
    assert np.array_equal(boxes.detector_boxes_to_int(d, 1080, 1920), opre.detector_boxes_to_int(d, 1080, 1920))
    with pytest.raises(ValueError):
        boxes.validate_boxes(np.array([[0, 0, 5, 5], [3, 3, 3, 9]]), np.array([0, 0]), [(10, 10)])


def make_yolo_dataset(root, n_img=3, seed=0):
    import cv2
    rng = np.random.default_rng(seed)
    (root / "ds" / "images").mkdir(parents=True)
    (root / "ds" / "labels").mkdir(parents=True)
    expected = []
    for i in range(n_img):
        h, w = int(rng.integers(60, 90)), int(rng.integers(80, 130))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        p = root / "ds" / "images" / f"im{i}.png"
        cv2.imwrite(str(p), img)
        lines = []
        for k in range(4):
            bw, bh = int(rng.integers(3, 40)), int(rng.integers(3, 40))
            x0, y0 = int(rng.integers(0, w - bw)), int(rng.integers(0, h - bh))
            lines.append(f"{k % 2} {(x0 + bw / 2) / w:.6f} {(y0 + bh / 2) / h:.6f} {bw / w:.6f} {bh / h:.6f}")
        (root / "ds" / "labels" / f"im{i}.txt").write_text("\n".join(lines) + "\n")
        for box, label in opre.parse_yolo_label_lines(lines, (h, w)):
            expected.append((str(p), box, label))
    (root / "ann.yaml").write_text(json.dumps({"path": "ds", "train": "images", "val": "images", "names": ["cat", "dog"]}))
    return expected


def plain_pipeline(T, size=32):
    return T.Compose([T.Resize(size, size), T.Normalize(), T.ToTensorV2()])


def test_yolo_dataset_descriptors(tmp_path):
    from nkb_classification_b200 import dataset as D, transforms as T
    expected = make_yolo_dataset(tmp_path)
    ds = D.AnnotatedYOLODataset(str(tmp_path / "ann.yaml"), fold="val", transform=D.Transforms(plain_pipeline(T)),
                                image_base_dir=str(tmp_path))
    assert ds.classes == ["cat", "dog"] and ds.class_to_idx == {"cat": 0, "dog": 1}
    assert [(p, tuple(b), l) for p, b, l in ds.list_bbox] == [(p, tuple(b), l) for p, b, l in expected]
    assert len(ds) == len(expected) and ds[0] == (expected[0][0], expected[0][1], expected[0][2])
    assert ds.get_labels().tolist() == [e[2] for e in expected]
    assert ds.transform.plan.channel_swap is True  # cv2.imread is BGR; the kernel emits RGB
    bg = D.AnnotatedYOLODataset(str(tmp_path / "ann.yaml"), fold="train", transform=None, image_base_dir=str(tmp_path),
                                generate_backgrounds=True)
    assert bg.classes[-1] == "<GENERATED>_background" and bg.class_to_idx["<GENERATED>_background"] == 2


def test_csv_datasets_and_collate(tmp_path):
    import cv2
    import pandas as pd
    from nkb_classification_b200 import dataset as D, transforms as T
    rows = []
    for i in range(6):
        p = tmp_path / f"a{i}.png"
        cv2.imwrite(str(p), np.full((20, 24, 3), i, np.uint8))
        rows.append({"path": p.name, "fold": "val" if i % 2 else "train", "color": ["red", "blue"][i % 2],
                     "size": ["s", "m", "l"][i % 3]})
    pd.DataFrame(rows).to_csv(tmp_path / "ann.csv", index=False)
    tr = D.Transforms(plain_pipeline(T))
    s = D.AnnotatedSingletaskDataset(str(tmp_path / "ann.csv"), "size", fold="train", transform=tr,
                                     image_base_dir=str(tmp_path))
    assert s.classes == ["l", "m", "s"] and len(s) == 3
    assert s[0][1] is None and s[0][2].dtype == np.int64
    m = D.AnnotatedMultitaskDataset(str(tmp_path / "ann.csv"), ["size", "color"], fold="val", transform=tr,
                                    image_base_dir=str(tmp_path))
    assert m.target_names == ["color", "size"] and list(m[0][2].keys()) == ["color", "size"]
    tgt = D.collate_targets([m[i][2] for i in range(len(m))])
    assert set(tgt) == {"color", "size"} and tgt["size"].dtype == torch.int64 and tgt["size"].shape == (3,)
    assert D.collate_targets([s[i][2] for i in range(3)]).tolist() == [s._labels[i] for i in range(3)]
    loader = D.DeviceCropLoader(m, batch_size=2, device="cuda:0")
    assert len(loader) == 2 and [len(b) for b in loader._index_batches()] == [2, 1]
    assert len(D.DeviceCropLoader(m, batch_size=2, drop_last=True)) == 1
    smp = D.ImbalancedDatasetSampler(s)
    assert len(list(iter(smp))) == 3


def test_pack_frames_aligns_rows():
    from nkb_classification_b200 import dataset as D
    rng = np.random.default_rng(0)
    frames = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in ((5, 7), (3, 16), (4, 33))]
    staging, total, desc, sizes = D.pack_frames(frames)
    buf = staging.numpy()
    for f, (off, h, w, pitch) in zip(frames, desc.tolist()):
        assert off % 256 == 0 and pitch % 16 == 0 and pitch >= w * 3
        assert np.array_equal(buf[off: off + h * pitch].reshape(h, pitch)[:, : w * 3].reshape(h, w, 3), f)
    assert sizes == [(5, 7), (3, 16), (4, 33)] and total >= sum(d[1] * d[3] for d in desc.tolist())
    assert D.pack_frames([])[2].shape == (0, 4)


def test_frame_cache_ring_allocator():
    """DeviceFrameCache: FIFO ring over one arena, pinned entries are never evicted, oversize frames are refused."""
    from nkb_classification_b200 import dataset as D
    c = D.DeviceFrameCache("cpu", 4 * 4096)              # room for four 4096-byte frames (32 x 42 px -> pitch 128)
    h, w = 32, 42
    assert D._row_pitch(w) == 128
    e = [c.alloc(f"f{i}", h, w, set()) for i in range(4)]
    assert [x[0] for x in e] == [0, 4096, 8192, 12288] and list(c.map) == ["f0", "f1", "f2", "f3"]
    assert c.get("f2") == (8192, h, w, 128, 4096)
    x = c.alloc("f4", h, w, set())                        # wraps: evicts the oldest
    assert x[0] == 0 and "f0" not in c.map and c.evictions == 1 and list(c.map) == ["f1", "f2", "f3", "f4"]
    assert c.alloc("f5", h, w, {"f1"}) is None            # the entry in the way is pinned: refused, nothing evicted
    assert list(c.map) == ["f1", "f2", "f3", "f4"]
    y = c.alloc("f5", h, w, {"f3"})
    assert y[0] == 4096 and list(c.map) == ["f2", "f3", "f4", "f5"]
    assert c.alloc("big", 1000, 1000, set()) is None      # larger than the arena
    z = c.alloc("tall", 64, 42, set())                    # 8192 bytes at head 8192: evicts f2 and f3
    assert z[0] == 8192 and list(c.map) == ["f4", "f5", "tall"]


def test_group_by_frame_keeps_the_draw_and_groups_it():
    from nkb_classification_b200 import dataset as D, transforms as T
    rng = np.random.default_rng(0)
    fidx = np.repeat(np.arange(6), 5)
    ds = D.InMemoryFrames([None] * 6, fidx, labels=rng.integers(0, 3, 30), boxes=np.tile([0, 0, 8, 8], (30, 1)),
                          transform=D.Transforms(plain_pipeline(T)))
    torch.manual_seed(0)
    plain = D.DeviceCropLoader(ds, batch_size=7, shuffle=True, device="cpu")
    grouped = D.DeviceCropLoader(ds, batch_size=7, shuffle=True, device="cpu", group_by_frame=True)
    torch.manual_seed(3)
    a = plain._order()
    torch.manual_seed(3)
    b = grouped._order()
    assert sorted(a) == sorted(b) == list(range(30))      # same samples, each once
    frames_b = [int(fidx[i]) for i in b]
    runs = [f for k, f in enumerate(frames_b) if k == 0 or frames_b[k - 1] != f]
    assert len(runs) == 6                                  # every frame is one contiguous run
    first = []
    for i in a:
        if int(fidx[i]) not in first:
            first.append(int(fidx[i]))
    assert runs == first                                   # frames in the order of their first draw
    smp = D.ImbalancedDatasetSampler(ds)
    w = D.DeviceCropLoader(ds, batch_size=7, sampler=smp, device="cpu", group_by_frame=True)
    torch.manual_seed(5)
    drawn = sorted(iter(smp))
    torch.manual_seed(5)
    assert sorted(w._order()) == drawn                     # the weighted draw (with its repeats) is kept


def test_factories_raise_like_the_reference():
    from nkb_classification_b200 import losses, utils
    with pytest.raises(NotImplementedError, match="Unknown loss type"):
        losses.get_loss({"type": "HingeLoss", "task": "single"}, "cpu")
    with pytest.raises(ValueError):
        losses.FocalLoss(reduction="median")
    lin = torch.nn.Linear(4, 2)
    model = SimpleNamespace(emb_model=torch.nn.Linear(4, 4), classifier=lin)
    with pytest.raises(NotImplementedError, match="Unknown optimizer"):
        utils.get_optimizer(model, {"type": "lion"})
    opt = utils.get_optimizer(model, {"type": "nadam", "lr": 1e-3, "classifier_lr": 1e-2, "weight_decay": 0.2})
    assert [g["lr"] for g in opt.param_groups] == [1e-3, 1e-2]
    assert utils.get_scheduler(opt, {}) is None
    with pytest.raises(NotImplementedError):
        utils.get_scheduler(opt, {"type": "warmup"})
    c2i, i2c = utils.get_classes_configs({"a": ["x", "y"], "b": ["u"]})
    assert c2i["a"]["y"] == 1 and i2c["b"][0] == "u"
    with pytest.raises(NotImplementedError):
        utils.load_classes(3)
    crit = losses.get_loss({"type": "FocalLoss", "gamma": 1, "task": "multi"}, "cpu")
    assert isinstance(crit, losses.MultitaskCriterion) and crit.criterion.gamma == 1
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        crit({"t": torch.zeros(2, 3)}, {"t": torch.zeros(2, dtype=torch.int64)})


def test_model_mirror_keeps_state_dict_keys_and_scripts():
    from nkb_classification_b200 import model as M

    class Tiny(torch.nn.Module):
        num_features = 8

        def __init__(self):
            super().__init__()
            self.conv = torch.nn.Conv2d(3, 8, 3)
            self.drop = torch.nn.Dropout(0.5)

        def forward(self, x):
            return self.drop(self.conv(x)).mean(dim=(2, 3))

    cfg = {"task": "multi", "model": Tiny(), "pretrained": False, "backbone_dropout": 0.1, "classifier_dropout": 0.2,
           "classifier_initialization": "kaiming_normal_"}
    m = M.get_model(cfg, {"color": ["r", "g", "b"], "size": ["s", "l"]})
    keys = set(m.state_dict().keys())
    assert {"classifier.color.1.weight", "classifier.color.1.bias", "classifier.size.1.weight"} <= keys
    assert m.emb_model.drop.p == 0.1 and m.classifier["size"][0].p == 0.2
    assert float(m.classifier["color"][1].bias.abs().sum()) == 0.0
    out = m(torch.zeros(2, 3, 8, 8))
    assert list(out.keys()) == ["color", "size"] and out["color"].shape == (2, 3)
    torch.jit.script(m)  # train.py:66
    m.set_backbone_state("freeze")
    assert not any(p.requires_grad for p in m.emb_model.parameters())
    s = M.get_model(dict(cfg, task="single", model=Tiny()), ["a", "b", "c", "d"])
    assert set(s.state_dict().keys()) >= {"classifier.1.weight", "classifier.1.bias"}
    with pytest.raises(RuntimeError, match="CUDA"):
        M.HeadPack(s)


def test_logger_and_metrics_reference_structure(golden_dir):
    """Reference-compatible log_iter (any device) -> get_epoch_results -> compute_metrics == reference goldens."""
    from nkb_classification_b200 import logging as L, metrics as Mx
    g = np.load(golden_dir / "metrics_golden.npz")
    names = ["a_color", "b_size", "c_kind"]
    cfg = SimpleNamespace(task="multi", target_names=names)
    classes = {n: list(range(g[f"{n}.logits"].shape[1])) for n in names}
    lg = L.BaseLogger(cfg, classes)
    assert lg.target_names == names
    N = 400
    for lo, hi, k in ((0, 150, 0), (150, 400, 1)):
        pred = {n: torch.from_numpy(g[f"{n}.logits"][lo:hi]) for n in names}
        true = {n: torch.from_numpy(g[f"{n}.gt"][lo:hi]) for n in names}
        loss = {n: torch.tensor(float(g[f"{n}.running_loss"][k])) for n in names}
        loss["loss"] = torch.tensor(float(g["loss"][k]))
        lg.log_iter(pred, true, loss)
    res = lg.get_epoch_results()
    assert set(res) >= {"running_loss", "confidences", "predictions", "ground_truth", "images"}
    assert len(res["predictions"]["a_color"]) == N and len(res["running_loss"]["loss"]) == 2
    # swap in the golden running losses (10 per task) to compare the loss-dependent entries too
    for n in names:
        res["running_loss"][n] = g[f"{n}.running_loss"].tolist()
    res["running_loss"]["loss"] = g["loss"].tolist()
    m = Mx.compute_metrics(cfg, res)
    assert m["epoch_acc"] == float(g["epoch_acc"])
    for n in names:
        assert m[n]["epoch_acc"] == float(g[f"{n}.epoch_acc"])
        np.testing.assert_array_equal(np.asarray(m[n]["epoch_roc_auc"], np.float64), g[f"{n}.epoch_roc_auc"])
        cm = om.confusion_matrix(g[f"{n}.gt"], om.argmax_first_fast(g[f"{n}.logits"]), g[f"{n}.logits"].shape[1])
        assert Mx.balanced_accuracy_from_confusion(cm) == float(g[f"{n}.epoch_acc"])
    with pytest.raises(ValueError):
        Mx.compute_metrics(SimpleNamespace(task="other"), res)


def test_loader_refuses_labels_outside_the_class_range():
    """The kernels ignore labels outside [0, C) (torch would device-assert): the loader validates them on the host."""
    import torch
    from nkb_classification_b200.dataset import validate_targets
    validate_targets(torch.tensor([0, 2, -100, 1]), ["a", "b", "c"])
    validate_targets({"x": torch.tensor([0, 1]), "y": torch.tensor([3, -100])}, {"x": [0, 1], "y": [0, 1, 2, 3]})
    validate_targets(["p0", "p1"], ["a"])                                  # inference: paths, nothing to check
    with pytest.raises(ValueError):
        validate_targets(torch.tensor([0, 3]), ["a", "b", "c"])
    with pytest.raises(ValueError):
        validate_targets(torch.tensor([0, -1]), ["a", "b", "c"])
    with pytest.raises(ValueError):
        validate_targets({"x": torch.tensor([0, 2])}, {"x": [0, 1]})
