"""Frame ingest on the GPU: whatever the sampler does, the loader must hand out the very tensors the
plain in-order path produces, while decoding each frame once and uploading no more than the crops need -- device frame
cache (epoch >= 2 uploads and decodes nothing), ROI upload, frame-grouped batches, cache eviction under prefetch.
The reference decodes a full frame per crop and ships 602,112 bytes of fp32 per 224x224 crop
(nkb_classification/dataset.py:398-409, engine.py:40)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REF_BYTES_PER_CROP = 3 * 224 * 224 * 4


def make_dataset(n_frames=12, per_frame=10, H=270, W=480, seed=0):
    from nkb_classification_b200 import dataset as D, transforms as T
    rng = np.random.default_rng(seed)
    frames = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(n_frames)]
    boxes, fidx = [], []
    for f in range(n_frames):
        for _ in range(per_frame):
            w, h = int(rng.integers(20, 200)), int(rng.integers(20, 200))
            x0, y0 = int(rng.integers(0, W - w)), int(rng.integers(0, H - h))
            boxes.append((x0, y0, x0 + w, y0 + h))
            fidx.append(f)
    labels = rng.integers(0, 5, len(fidx))
    pipe = [T.Resize(224, 224), T.Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)), T.ToTensorV2()]
    return D.InMemoryFrames(frames, fidx, labels, boxes=boxes, classes=list(range(5)), transform=D.Transforms(pipe))


def reference_outputs(ds, dev):
    """Every sample through the plain path: in order, no cache, whole frames."""
    from nkb_classification_b200 import dataset as D
    loader = D.DeviceCropLoader(ds, batch_size=16, device=dev, roi_upload=False)
    imgs, tgts = [], []
    for img, t in loader:
        imgs.append(img.clone())
        tgts.append(t)
    return torch.cat(imgs), torch.cat(tgts)


def run_epoch(loader, seed, ref_img, ref_tgt):
    torch.manual_seed(seed)
    order = loader._order()
    torch.manual_seed(seed)
    pos = 0
    for img, t in loader:
        idx = order[pos: pos + img.shape[0]]
        assert torch.equal(img, ref_img[idx]), "loader output differs from the in-order path"
        assert torch.equal(t, ref_tgt[idx])
        pos += img.shape[0]
    assert pos == len(order)
    torch.cuda.synchronize()


@pytest.mark.parametrize("prefetch", [0, 2])
def test_shuffled_loader_with_frame_cache(cuda_device, prefetch):
    from nkb_classification_b200 import dataset as D
    ds = make_dataset()
    ref_img, ref_tgt = reference_outputs(ds, cuda_device)
    n, n_frames = len(ds), 12
    ds.reads = 0
    loader = D.DeviceCropLoader(ds, batch_size=32, shuffle=True, device=cuda_device, prefetch=prefetch,
                                frame_cache_bytes=64 << 20)
    run_epoch(loader, 1, ref_img, ref_tgt)
    st = dict(loader.stats)
    assert st["decodes"] == n_frames == ds.reads          # each distinct frame decoded once, although shuffled
    assert st["crops"] == n and st["h2d_bytes"] / n <= REF_BYTES_PER_CROP
    loader.reset_stats()
    run_epoch(loader, 2, ref_img, ref_tgt)                # second epoch: everything is resident
    st2 = loader.stats
    assert st2["decodes"] == 0 and ds.reads == n_frames and st2["cache_inserts"] == 0
    meta_bytes = n * (16 + 4 + 8) + 32 * n_frames + 64    # boxes + frame indices + labels + descriptors, ONE copy
    assert st2["h2d_bytes"] <= meta_bytes and st2["resident_epochs"] == 1


def test_shuffled_loader_roi_upload_without_cache(cuda_device):
    from nkb_classification_b200 import dataset as D
    ds = make_dataset(per_frame=2, n_frames=30)           # few boxes per frame: whole-frame upload would be wasteful
    ref_img, ref_tgt = reference_outputs(ds, cuda_device)
    loader = D.DeviceCropLoader(ds, batch_size=8, shuffle=True, device=cuda_device, prefetch=2)
    run_epoch(loader, 4, ref_img, ref_tgt)
    st = loader.stats
    whole = 270 * D._row_pitch(480)
    assert st["roi_frames"] > 0 and st["h2d_bytes"] / st["crops"] <= REF_BYTES_PER_CROP
    assert st["h2d_bytes"] < 0.5 * st["decodes"] * whole  # regions of interest, not frames, crossed the bus


def test_small_cache_evicts_and_overflows_under_prefetch(cuda_device):
    """An arena that holds ~3 frames, batches that touch up to 8, two batches staged ahead: entries of staged batches
    stay pinned, the rest of the frames travel with their batch -- results unchanged."""
    from nkb_classification_b200 import dataset as D
    ds = make_dataset()
    ref_img, ref_tgt = reference_outputs(ds, cuda_device)
    frame_bytes = 270 * D._row_pitch(480)
    loader = D.DeviceCropLoader(ds, batch_size=8, shuffle=True, device=cuda_device, prefetch=2,
                                frame_cache_bytes=int(3.5 * frame_bytes))
    for ep in range(3):
        run_epoch(loader, 10 + ep, ref_img, ref_tgt)
    st = loader.stats
    assert st["cache_inserts"] > 0 and st["cache_hits"] > 0 and st["roi_frames"] + st["whole_frames"] > 0
    assert loader.cache.evictions > 0


def test_group_by_frame_and_weighted_sampling(cuda_device):
    from nkb_classification_b200 import dataset as D
    ds = make_dataset()
    ref_img, ref_tgt = reference_outputs(ds, cuda_device)
    g = D.DeviceCropLoader(ds, batch_size=20, shuffle=True, device=cuda_device, group_by_frame=True)
    run_epoch(g, 7, ref_img, ref_tgt)
    assert g.stats["decodes"] <= 12 + g.stats["batches"]   # a frame straddles at most two batches
    w = D.DeviceCropLoader(ds, batch_size=20, sampler=D.ImbalancedDatasetSampler(ds), device=cuda_device,
                           group_by_frame=True, frame_cache_bytes=64 << 20)
    run_epoch(w, 8, ref_img, ref_tgt)
    assert w.stats["decodes"] <= 12


def test_whole_image_dataset_through_the_cache(cuda_device):
    from nkb_classification_b200 import dataset as D, transforms as T
    rng = np.random.default_rng(3)
    frames = [rng.integers(0, 256, (int(rng.integers(40, 90)), int(rng.integers(40, 90)), 3), dtype=np.uint8)
              for _ in range(9)]
    pipe = [T.LongestMaxSize(64), T.PadIfNeeded(64, 64, border_mode=T.BORDER_CONSTANT, value=0),
            T.Normalize(), T.ToTensorV2()]
    ds = D.InMemoryFrames(frames, np.arange(9), rng.integers(0, 3, 9), boxes=None, classes=[0, 1, 2],
                          transform=D.Transforms(pipe))
    ref_img, ref_tgt = reference_outputs(ds, cuda_device)
    loader = D.DeviceCropLoader(ds, batch_size=4, shuffle=True, device=cuda_device, frame_cache_bytes=8 << 20)
    run_epoch(loader, 1, ref_img, ref_tgt)
    run_epoch(loader, 2, ref_img, ref_tgt)
    assert loader.stats["decodes"] == 9


def test_resident_epoch_plan_matches_streaming_path(cuda_device):
    """Epoch >= 2 of a dataset that fits the frame cache is planned at once (one metadata copy, a batch = one K1 launch
    on slices): same tensors as the per-batch path, for tensor and dict targets, host or device targets, with the
    in-batch frame sort, a ragged last batch and drop_last."""
    from nkb_classification_b200 import dataset as D
    ds = make_dataset()
    ref_img, ref_tgt = reference_outputs(ds, cuda_device)
    for kw in (dict(), dict(targets_on_device=True), dict(drop_last=True), dict(sort_within_batch=True)):
        loader = D.DeviceCropLoader(ds, batch_size=32, shuffle=True, device=cuda_device, prefetch=2,
                                    frame_cache_bytes=64 << 20, **kw)
        stream = D.DeviceCropLoader(ds, batch_size=32, shuffle=True, device=cuda_device, frame_cache_bytes=64 << 20,
                                    resident_epochs=False, **kw)
        for _ in loader:                                   # epoch 1 fills the cache
            pass
        for _ in stream:
            pass
        torch.manual_seed(5)
        a = [(img.clone(), t) for img, t in loader]
        torch.manual_seed(5)
        b = [(img.clone(), t) for img, t in stream]
        assert loader.stats["resident_epochs"] == 1 and stream.stats["resident_epochs"] == 0
        assert len(a) == len(b) == (3 if kw.get("drop_last") else 4)
        for (ia, ta), (ib, tb) in zip(a, b):
            assert torch.equal(ia, ib) and torch.equal(ta.cpu(), tb.cpu())
            assert ta.is_cuda == bool(kw.get("targets_on_device")) and ta.dtype == torch.int64
    # dict targets (multi task)
    lab = {"a": np.arange(len(ds)) % 3, "b": np.arange(len(ds)) % 2}
    dm = D.InMemoryFrames(ds.frames, ds.frame_idx, lab, boxes=ds.boxes, transform=ds.transform)
    loader = D.DeviceCropLoader(dm, batch_size=50, shuffle=True, device=cuda_device, frame_cache_bytes=64 << 20)
    for _ in loader:
        pass
    torch.manual_seed(9)
    order = loader._order()
    torch.manual_seed(9)
    pos = 0
    for img, t in loader:
        idx = order[pos: pos + img.shape[0]]
        assert torch.equal(img, ref_img[idx]) and set(t) == {"a", "b"}
        assert t["a"].tolist() == [lab["a"][i] for i in idx] and t["b"].tolist() == [lab["b"][i] for i in idx]
        pos += img.shape[0]
    assert loader.stats["resident_epochs"] == 1 and pos == len(dm)


def test_resident_epoch_with_train_pipeline_matches_streaming_path(cuda_device):
    """Train pipelines (per-sample augmentation parameters) run as resident epochs too: the parameters are drawn batch
    by batch inside the loop, exactly as the per-batch path draws them -- same `random.seed`, same shuffle, same tensors."""
    import random
    from nkb_classification_b200 import dataset as D, transforms as T
    ds = make_dataset()
    pipe = [T.Resize(224, 224), T.HorizontalFlip(p=0.5), T.VerticalFlip(p=0.5),
            T.RandomBrightnessContrast(brightness_limit=(-0.2, 0.2), contrast_limit=(0.1, -0.5), p=0.5),
            T.HueSaturationValue(hue_shift_limit=0, sat_shift_limit=10, val_shift_limit=50, p=0.5),
            T.CoarseDropout(max_holes=4, min_holes=1, max_height=0.2, min_height=0.05, max_width=0.2, min_width=0.05,
                            fill_value=[0, 0.5, 1], p=0.5),
            T.Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)), T.ToTensorV2()]
    dt = D.InMemoryFrames(ds.frames, ds.frame_idx, ds.labels, boxes=ds.boxes, classes=list(range(5)),
                          transform=D.Transforms(pipe))
    res = D.DeviceCropLoader(dt, batch_size=32, shuffle=True, device=cuda_device, frame_cache_bytes=64 << 20)
    stream = D.DeviceCropLoader(dt, batch_size=32, shuffle=True, device=cuda_device, frame_cache_bytes=64 << 20,
                                resident_epochs=False)
    for _ in res:          # first epochs fill the caches
        pass
    for _ in stream:
        pass
    outs = []
    for loader in (res, stream):
        torch.manual_seed(21)
        random.seed(21)
        outs.append([(img.clone(), t.clone()) for img, t in loader])
    assert res.stats["resident_epochs"] == 1 and stream.stats["resident_epochs"] == 0
    assert len(outs[0]) == len(outs[1]) == 4
    for (ia, ta), (ib, tb) in zip(*outs):
        assert torch.equal(ia, ib) and torch.equal(ta, tb)
    torch.manual_seed(21)
    random.seed(22)
    again = [img.clone() for img, _ in res]
    assert not all(torch.equal(a, b[0]) for a, b in zip(again, outs[0]))      # other parameters, other pixels
