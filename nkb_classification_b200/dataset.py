"""Datasets and loaders with the reference's factory interface
(nkb_classification/dataset.py:541-644), re-plumbed so that pixels are touched
once on the host (decode) and everything after that happens in K1 on the GPU.

reference (per SAMPLE, on DataLoader workers)           here (per BATCH)
  cv2.imread(frame) + cvtColor(BGR2RGB)   :398-402        each distinct frame of the batch is decoded ONCE
  img[y0:y1, x0:x1]                       :404            boxes go to the device as int32 [n,4]
  Transforms -> A.Compose (resize, pad,   :96-102         one K1 launch: crop + resize + normalize + NCHW
    normalize, ToTensorV2)                                (BGR->RGB folded into the kernel's byte selectors)
  default_collate + pin + fp32 H2D        :608-629        uint8 frames H2D from a pinned staging buffer

Datasets keep the reference's constructor arguments and attributes (``classes``,
``class_to_idx``, ``idx_to_class``, ``get_labels``); ``__getitem__`` returns a
*sample descriptor* (image path, box or None, label) instead of pixels.
The loader object exposes ``.dataset`` and ``len()`` like a DataLoader and
yields ``(img[B,3,H,W] on the device, target)`` with ``target`` a tensor
(single task) or a dict of tensors (multi task), like ``default_collate`` does.
"""
from __future__ import annotations

import glob
import os
import pickle as pkl
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Callable, Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from .boxes import bbox_xywhn2xyxy, check_boxes_sizes, validate_boxes
from .transforms import PreprocessPlan, compile_pipeline
from .utils import get_classes_configs, load_classes

IMG_EXT = [".jpg", ".jpeg", ".png"]


def _imread_bgr(path: str) -> np.ndarray:
    import cv2

    img = cv2.imread(path)  # BGR uint8 HWC; BGR->RGB happens inside K1 (channel_swap)
    if img is None:
        raise FileNotFoundError(f"cannot read image {path}")
    return img


def _image_size(path: str) -> Tuple[int, int]:
    from PIL import Image

    with Image.open(path) as im:
        w, h = im.size
    return h, w


class Transforms:
    """Reference name kept (dataset.py:89-102); here it *compiles* the pipeline for K1 instead of running it."""

    def __init__(self, transforms) -> None:
        self.transforms = transforms
        self.plan: PreprocessPlan = compile_pipeline(transforms, channel_swap=True)


class ImbalancedDatasetSampler(torch.utils.data.sampler.Sampler):
    """dataset.py:24-86: draw indices with probability 1 / count(label)."""

    def __init__(self, dataset, labels: list = None, indices: list = None, num_samples: int = None,
                 callback_get_label: Callable = None):
        self.indices = list(range(len(dataset))) if indices is None else indices
        self.num_samples = len(self.indices) if num_samples is None else num_samples
        if labels is None:
            labels = callback_get_label(dataset) if callback_get_label else dataset.get_labels()
        labels = np.asarray(labels)
        if labels.ndim > 1:
            labels = np.array(["|".join(map(str, r)) for r in labels])
        _, inv, counts = np.unique(labels, return_inverse=True, return_counts=True)
        self.weights = torch.DoubleTensor(1.0 / counts[inv])

    def __iter__(self):
        return (self.indices[i] for i in torch.multinomial(self.weights, self.num_samples, replacement=True))

    def __len__(self):
        return self.num_samples


# --------------------------------------------------------------------------------------------
# datasets: sample descriptors
# --------------------------------------------------------------------------------------------
class _DescDataset(torch.utils.data.Dataset):
    transform: Optional[Transforms] = None

    def label_at(self, idx):  # -> np.int64 scalar | dict name -> np.int64 | str path (inference)
        raise NotImplementedError

    def path_at(self, idx) -> str:
        raise NotImplementedError

    def box_at(self, idx):  # (x0,y0,x1,y1) or None for the whole image
        return None

    def __getitem__(self, idx):
        return self.path_at(idx), self.box_at(idx), self.label_at(idx)


class InferDataset(_DescDataset):
    """dataset.py:105-130: yields (tensor, path)."""

    def __init__(self, folder_path, transform=None):
        self.folder = Path(folder_path)
        self.transform = transform
        self.imgs = [str(p) for p in self.folder.iterdir() if p.suffix.lower() in IMG_EXT]

    def __len__(self):
        return len(self.imgs)

    def path_at(self, idx):
        return self.imgs[idx]

    def label_at(self, idx):
        return self.imgs[idx]


class AnnotatedSingletaskDataset(_DescDataset):
    """dataset.py:183-234."""

    def __init__(self, annotations_file, target_column, fold="test", transform=None, image_base_dir=None, classes=None,
                 **kwargs):
        import pandas as pd

        self.table = pd.read_csv(annotations_file)
        self.table = self.table[self.table["fold"] == fold]
        self.target_column = target_column
        if classes is not None:
            self.classes = load_classes(classes)
        else:
            self.classes = np.sort(np.unique(self.table[target_column].values)).tolist()
        self.class_to_idx, self.idx_to_class = get_classes_configs(self.classes)
        self.transform = transform
        if image_base_dir is not None:
            base = Path(image_base_dir)
            self.table.path = self.table.path.apply(lambda r: str(base / Path(r)))
        self._paths = self.table["path"].tolist()
        self._labels = [self.class_to_idx[v] for v in self.table[target_column].values]

    def __len__(self):
        return len(self._paths)

    def path_at(self, idx):
        return self._paths[idx]

    def label_at(self, idx):
        return np.array(self._labels[idx], dtype=np.int64)

    def get_labels(self):
        return self.table[self.target_column].values


class AnnotatedMultitaskDataset(_DescDataset):
    """dataset.py:482-538; label dict keys are the SORTED target names (:500)."""

    def __init__(self, annotations_file, target_names, fold="test", transform=None, image_base_dir=None, classes=None,
                 **kwargs):
        import pandas as pd

        self.table = pd.read_csv(annotations_file)
        self.table = self.table[self.table["fold"] == fold]
        self.target_names = [*sorted(target_names)]
        if classes is not None:
            self.classes = load_classes(classes)
        else:
            self.classes = {t: np.sort(np.unique(self.table[t].values)).tolist() for t in self.target_names}
        self.class_to_idx, self.idx_to_class = get_classes_configs(self.classes)
        self.transform = transform
        if image_base_dir is not None:
            base = Path(image_base_dir)
            self.table.path = self.table.path.apply(lambda r: str(base / Path(r)))
        self._paths = self.table["path"].tolist()
        self._labels = {t: [self.class_to_idx[t][v] for v in self.table[t].values] for t in self.target_names}

    def __len__(self):
        return len(self._paths)

    def path_at(self, idx):
        return self._paths[idx]

    def label_at(self, idx):
        return {t: np.array(self._labels[t][idx], dtype=np.int64) for t in self.target_names}

    def get_labels(self):
        return self.table[self.target_names].values


class GroupsDataset(_DescDataset):
    """dataset.py:133-180 (pickled file list + group dictionary)."""

    def __init__(self, root, ann_file, dict_path, transform=None, **kwargs):
        self.data_prefix = root
        self.transform = transform
        with Path(root, ann_file).open("rb") as f:
            data = pkl.load(f)
        with Path(dict_path).open("rb") as f:
            group_dict = pkl.load(f)
        inv_group = {v_i: k for k, v in group_dict.items() for v_i in v}
        self.class_to_idx = {k: i for i, k in enumerate(group_dict.keys())}
        self.idx_to_class = {idx: lb for lb, idx in self.class_to_idx.items()}
        self.classes = list(self.class_to_idx.keys())
        self._paths, self._labels = [], []
        for sample in data:
            sample = Path(sample)
            orig_label = sample.parent.name
            img_path = Path(root, "images_lr", orig_label, sample.name)
            assert img_path.is_file(), f"File {img_path} does not exist."
            self._paths.append(str(img_path))
            self._labels.append(self.class_to_idx[inv_group[orig_label]])

    def __len__(self):
        return len(self._paths)

    def path_at(self, idx):
        return self._paths[idx]

    def label_at(self, idx):
        return np.array(self._labels[idx], dtype=np.int64)

    def get_labels(self):
        return np.array(self._labels)


class ImageFolder(_DescDataset):
    """torchvision.datasets.ImageFolder layout (the reference's fallback, dataset.py:579-580)."""

    def __init__(self, root, transform=None):
        self.transform = transform
        self.classes = sorted(d.name for d in Path(root).iterdir() if d.is_dir())
        self.class_to_idx = {c: i for i, c in enumerate(self.classes)}
        self.idx_to_class = {i: c for c, i in self.class_to_idx.items()}
        self.imgs = []
        for c in self.classes:
            for p in sorted(Path(root, c).rglob("*")):
                if p.suffix.lower() in IMG_EXT + [".bmp", ".webp", ".ppm", ".tif", ".tiff"]:
                    self.imgs.append((str(p), self.class_to_idx[c]))

    def __len__(self):
        return len(self.imgs)

    def path_at(self, idx):
        return self.imgs[idx][0]

    def label_at(self, idx):
        return np.array(self.imgs[idx][1], dtype=np.int64)

    def get_labels(self):
        return [x[1] for x in self.imgs]


class AnnotatedYOLODataset(_DescDataset):
    """dataset.py:237-479: YOLO-format detection dataset turned into per-box classification samples."""

    def __init__(self, annotations_file, fold="train", transform=None, image_base_dir=None, min_box_size=5,
                 generate_backgrounds=False, background_generating_prob=None, background_crop_sizes=(0.1, 0.3),
                 **kwargs):
        import yaml

        self.ext = IMG_EXT
        self.min_box_size = min_box_size
        assert fold in ("train", "val", "test"), f"Got fold equals {fold}"
        self.fold, self.transform = fold, transform
        self.generate_backgrounds = generate_backgrounds
        self.background_generating_prob = background_generating_prob
        self.background_crop_sizes = background_crop_sizes
        self.attempts_to_put_bakground_crop = 1000
        assert os.path.exists(annotations_file), f"Annotations file {annotations_file} does not exist."
        with open(annotations_file, "r") as f:
            self.yaml_data = yaml.load(f, Loader=yaml.SafeLoader)
        self.idx_to_class = self.yaml_data["names"]
        if type(self.idx_to_class) is list:
            self.idx_to_class = {i: lb for i, lb in enumerate(self.idx_to_class)}
        assert set(self.idx_to_class.keys()) == set(range(len(self.idx_to_class))), \
            "Class indices should form range(0, num_classes) without skips"
        self.classes = [self.idx_to_class[i] for i in range(len(self.idx_to_class))]
        self.class_to_idx = {lb: idx for idx, lb in self.idx_to_class.items()}
        if generate_backgrounds:
            bg_idx, bg_lb = len(self.classes), "<GENERATED>_background"
            self.classes.append(bg_lb)
            self.idx_to_class[bg_idx] = bg_lb
            self.class_to_idx[bg_lb] = bg_idx
        if self.background_generating_prob is None:
            self.background_generating_prob = 1 / len(self.classes)
        if not isinstance(self.yaml_data[self.fold], list):
            self.yaml_data[self.fold] = [self.yaml_data[self.fold]]
        base = Path(image_base_dir) if image_base_dir is not None else Path("/")
        image_dirs = [base / self.yaml_data["path"] / p for p in self.yaml_data[self.fold]]
        self.list_bbox = []
        for image_filename in sorted(self.get_img_files(image_dirs)):
            image_filename = Path(image_filename)
            labels_dir = image_filename.parent.parent / "labels"
            assert labels_dir.is_dir(), f"Directory {labels_dir} does not exist"
            if image_filename.suffix.lower() not in self.ext:
                continue
            txt_file = labels_dir / (image_filename.stem + ".txt")
            if not txt_file.is_file():
                continue
            with open(txt_file, "r") as fp:
                lines = fp.readlines()
            img_height, img_width = _image_size(str(image_filename))
            image_size = (img_height, img_width)
            true_boxes = []
            for line in lines:
                if not line.split():
                    continue
                label = int(line.split()[0])
                xc, yc, w, h = tuple(map(float, line.split()[1:5]))
                box = bbox_xywhn2xyxy(xc, yc, w, h, image_size)
                true_boxes.append(box)
                if not self.check_boxes_sizes_annotation(*box):
                    continue
                self.list_bbox.append((str(image_filename), box, label))
            if self.generate_backgrounds and np.random.rand() <= self.background_generating_prob:
                self._add_background(str(image_filename), img_width, img_height, true_boxes)

    def _add_background(self, filename, img_width, img_height, true_boxes):
        """dataset.py:361-393 (random background crop kept only under the reference's intersection rule)."""
        for _ in range(self.attempts_to_put_bakground_crop):
            s = np.random.uniform(*self.background_crop_sizes)
            x0 = np.random.randint(0, int(img_width * (1 - s)))
            y0 = np.random.randint(0, int(img_height * (1 - s)))
            box = (x0, y0, x0 + int(img_width * s), y0 + int(img_height * s))
            if not self.check_boxes_sizes_annotation(*box):
                continue
            if all(self.bbox_intersect(box, tb) for tb in true_boxes):
                self.list_bbox.append((filename, box, self.class_to_idx[self.classes[-1]]))
                break

    bbox_xywhn2xyxy = staticmethod(bbox_xywhn2xyxy)

    @staticmethod
    def bbox_intersect(bbox1, bbox2):
        x1_min, y1_min, x1_max, y1_max = bbox1
        x2_min, y2_min, x2_max, y2_max = bbox2
        if x1_max < x2_min or x2_max < x1_min:
            return False
        if y1_max < y2_min or y2_max < y1_min:
            return False
        return True

    def check_boxes_sizes_annotation(self, x_min, y_min, x_max, y_max):
        return check_boxes_sizes(x_min, y_min, x_max, y_max, self.min_box_size)

    def get_img_files(self, img_path):
        f = []
        for p in img_path if isinstance(img_path, list) else [img_path]:
            p = Path(p)
            if p.is_dir():
                f += glob.glob(str(p / "**" / "*.*"), recursive=True)
            elif p.is_file():
                with open(p) as t:
                    parent = str(p.parent) + os.sep
                    f += [x.replace("./", parent) if x.startswith("./") else x for x in t.read().strip().splitlines()]
            else:
                raise FileNotFoundError(f"{p} does not exist")
        exts = tuple(e[1:] for e in self.ext)
        im_files = sorted(x for x in f if x.split(".")[-1].lower() in exts)
        assert im_files, f"No images found in {img_path}"
        return im_files

    def __len__(self):
        return len(self.list_bbox)

    def path_at(self, idx):
        return self.list_bbox[idx][0]

    def box_at(self, idx):
        return self.list_bbox[idx][1]

    def label_at(self, idx):
        return self.list_bbox[idx][2]

    def get_labels(self):
        return np.array([label for _, _, label in self.list_bbox])


# --------------------------------------------------------------------------------------------
# loader: decode once per frame, stage uint8, one K1 launch per batch
# --------------------------------------------------------------------------------------------
def pack_frames(frames: Sequence[np.ndarray], staging: Optional[torch.Tensor] = None):
    """Lay frames of arbitrary sizes into one pinned uint8 buffer with 16-byte aligned rows (so K1's TMA path
    applies).  Returns (buffer tensor view, int64 [F,4] descriptors, list of (h, w))."""
    descs, sizes, off = [], [], 0
    for f in frames:
        h, w = f.shape[:2]
        pitch = (w * 3 + 15) // 16 * 16
        descs.append((off, h, w, pitch))
        sizes.append((h, w))
        off += h * pitch
    total = max(off, 16)
    if staging is None or staging.numel() < total:
        staging = torch.empty(int(total * 1.25) + 4096, dtype=torch.uint8)
        if torch.cuda.is_available():
            staging = staging.pin_memory()
    buf = staging.numpy()
    for f, (o, h, w, pitch) in zip(frames, descs):
        buf[o: o + h * pitch].reshape(h, pitch)[:, : w * 3] = f.reshape(h, w * 3)
    return staging, total, torch.tensor(descs, dtype=torch.int64), sizes


def collate_targets(labels: list):
    """What default_collate makes of the reference's per-sample labels."""
    first = labels[0]
    if isinstance(first, dict):
        return {k: torch.from_numpy(np.stack([np.asarray(l[k], dtype=np.int64) for l in labels])) for k in first}
    if isinstance(first, str):
        return list(labels)
    return torch.from_numpy(np.asarray(labels, dtype=np.int64))


class _IngestSlot:
    """One stage of the frame-ingest ring: pinned uint8 staging, its device twin, and the two events that make reuse
    safe (`copied`: H2D finished, recorded on the copy stream; `consumed`: K1 has read the device buffer, recorded on
    the consumer's stream)."""

    def __init__(self):
        self.staging: Optional[torch.Tensor] = None
        self.dev: Optional[torch.Tensor] = None
        self.copied: Optional[torch.cuda.Event] = None
        self.consumed: Optional[torch.cuda.Event] = None


class DeviceCropLoader:
    """DataLoader stand-in: batches of sample descriptors -> K1 -> (device tensor, target).

    Frame ingest: each distinct frame of a batch is decoded ONCE (thread pool), packed as uint8
    into a pinned staging buffer with 16-byte aligned rows and copied to the device as uint8 -- the reference decodes
    per crop and ships fp32 (dataset.py:398-409, engine.py:40).  With ``prefetch`` > 0 a producer thread runs decode +
    pack + H2D (on its own copy stream) for the next batches while the consumer's stream runs K1 / the model on the
    current one; a ring of ``prefetch + 2`` slots with `copied` / `consumed` events keeps buffers from being reused
    early.  ``prefetch = 0`` does the same work inline (one slot ring, still event-guarded)."""

    def __init__(self, dataset: _DescDataset, batch_size: int, shuffle: bool = False, sampler=None,
                 num_workers: int = 0, drop_last: bool = False, device="cuda:0", out_dtype=torch.float32,
                 prefetch: Optional[int] = None):
        if dataset.transform is None:
            raise ValueError("the dataset needs a Transforms(pipeline) to compile for K1")
        self.dataset, self.batch_size, self.shuffle, self.sampler = dataset, int(batch_size), shuffle, sampler
        self.drop_last, self.device, self.out_dtype = drop_last, torch.device(device), out_dtype
        self.plan = dataset.transform.plan
        self.pool = ThreadPoolExecutor(max(1, num_workers)) if num_workers and num_workers > 0 else None
        # like torch's DataLoader: workers imply prefetching (prefetch_factor = 2)
        self.prefetch = int(prefetch) if prefetch is not None else (2 if num_workers and num_workers > 0 else 0)
        self._slots = [_IngestSlot() for _ in range(self.prefetch + 2)]
        self._copy_stream: Optional[torch.cuda.Stream] = None
        # train pipelines: where the per-sample augmentation parameters come from (None = Python's global `random`,
        # which is what albumentations draws from, so `random.seed(...)` governs it as in the reference)
        self.aug_rng = None

    def __len__(self):
        n = len(self.sampler) if self.sampler is not None else len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def _index_batches(self) -> Iterator[List[int]]:
        if self.sampler is not None:
            order = list(iter(self.sampler))
        elif self.shuffle:
            order = torch.randperm(len(self.dataset)).tolist()
        else:
            order = list(range(len(self.dataset)))
        for i in range(0, len(order), self.batch_size):
            b = order[i: i + self.batch_size]
            if len(b) < self.batch_size and self.drop_last:
                return
            yield b

    # ---- producer side: host work + H2D of one batch into ring slot `k` ----
    def _stage(self, indices: List[int], k: int) -> dict:
        slot = self._slots[k]
        if slot.consumed is not None:
            slot.consumed.synchronize()          # K1 of the batch that used this slot last has read the device buffer
        samples = [self.dataset[i] for i in indices]
        paths = [s[0] for s in samples]
        uniq: Dict[str, int] = {}
        for p in paths:
            uniq.setdefault(p, len(uniq))
        plist = list(uniq)
        frames = list(self.pool.map(_imread_bgr, plist)) if self.pool else [_imread_bgr(p) for p in plist]
        slot.staging, total, desc, sizes = pack_frames(frames, slot.staging)
        fidx = np.array([uniq[p] for p in paths], dtype=np.int32)
        boxes = np.array([s[1] if s[1] is not None else (0, 0, sizes[fi][1], sizes[fi][0])
                          for s, fi in zip(samples, fidx)], dtype=np.int32).reshape(-1, 4)
        validate_boxes(boxes, fidx, sizes)
        aug = self.plan.draw(len(samples), self.aug_rng)     # in batch order: one producer, so the stream is reproducible
        dev = self.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        if slot.dev is None or slot.dev.numel() < total:
            slot.dev = torch.empty(int(total * 1.25) + 4096, dtype=torch.uint8, device=dev)
        if slot.copied is None:
            slot.copied = torch.cuda.Event()
        with torch.cuda.stream(self._copy_stream):
            slot.dev[:total].copy_(slot.staging[:total], non_blocking=True)
            meta = dict(boxes=torch.from_numpy(boxes).to(dev, non_blocking=True),
                        fidx=torch.from_numpy(fidx).to(dev, non_blocking=True),
                        desc=desc.to(dev, non_blocking=True))
            if aug is not None:
                ops.upload_augment(aug, dev)      # train pipelines: parameters ride the copy stream too
            slot.copied.record(self._copy_stream)
        meta.update(slot=k, total=total, aug=aug, target=collate_targets([s[2] for s in samples]))
        return meta

    # ---- consumer side: K1 on the caller's current stream ----
    def _consume(self, meta: dict):
        slot = self._slots[meta["slot"]]
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(slot.copied)
        shared = [meta["boxes"], meta["fidx"], meta["desc"]]
        if meta["aug"] is not None:
            shared += [t for t in ops.upload_augment(meta["aug"], self.device) if t is not None]
        for t in shared:
            t.record_stream(cur)                 # allocated on the copy stream, read on this one
        img = ops.preprocess_crops(slot.dev[:meta["total"]], meta["boxes"], meta["fidx"], self.plan,
                                   out_dtype=self.out_dtype, frame_desc=meta["desc"], aug=meta["aug"])
        if slot.consumed is None:
            slot.consumed = torch.cuda.Event()
        slot.consumed.record(cur)
        return img, meta["target"]

    def load_batch(self, indices: List[int]):
        return self._consume(self._stage(indices, 0))

    def __iter__(self):
        if self.prefetch <= 0:
            for b in self._index_batches():
                yield self.load_batch(b)
            return
        import queue
        import threading
        q: "queue.Queue" = queue.Queue(maxsize=self.prefetch)
        stop = threading.Event()
        nslot = len(self._slots)

        def put(item) -> bool:
            while not stop.is_set():
                try:
                    q.put(item, timeout=0.1)
                    return True
                except queue.Full:
                    continue
            return False

        def producer():
            try:
                torch.cuda.set_device(self.device)
                for i, b in enumerate(self._index_batches()):
                    if stop.is_set() or not put(self._stage(b, i % nslot)):
                        return
                put(None)
            except BaseException as e:  # surfaced in the consumer
                put(e)

        th = threading.Thread(target=producer, name="nkbk-ingest", daemon=True)
        th.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                yield self._consume(item)
        finally:
            stop.set()
            th.join()


def _device_of(data: dict):
    return data.get("device", "cuda:0")


def get_dataset(data, pipeline):
    """dataset.py:541-629: same ``data`` keys (type, batch_size, num_workers, shuffle, drop_last,
    weighted_sampling, dataset kwargs); optional extra key ``device``."""
    transform = Transforms(pipeline)
    kind = data["type"]
    kwargs = {k: v for k, v in data.items() if k not in ("device", "prefetch")}
    if kind == "GroupsDataset":
        dataset = GroupsDataset(transform=transform, **kwargs)
    elif kind == "AnnotatedMultitaskDataset":
        dataset = AnnotatedMultitaskDataset(transform=transform, **kwargs)
    elif kind == "AnnotatedSingletaskDataset":
        dataset = AnnotatedSingletaskDataset(transform=transform, **kwargs)
    elif kind == "AnnotatedYOLODataset":
        dataset = AnnotatedYOLODataset(transform=transform, **kwargs)
    else:
        dataset = ImageFolder(data["root"], transform=transform)
    sampler = ImbalancedDatasetSampler(dataset) if data.get("weighted_sampling", False) else None
    return DeviceCropLoader(dataset, batch_size=data["batch_size"], shuffle=data.get("shuffle", False), sampler=sampler,
                            num_workers=data.get("num_workers", 0), drop_last=data.get("drop_last", False),
                            device=_device_of(data), prefetch=data.get("prefetch"))


def get_inference_dataset(data, pipeline):
    """dataset.py:632-644: yields (img, paths)."""
    dataset = InferDataset(folder_path=data["folder_path"], transform=Transforms(pipeline))
    return DeviceCropLoader(dataset, batch_size=data["batch_size"], num_workers=data.get("num_workers", 0),
                            device=_device_of(data), prefetch=data.get("prefetch"))
