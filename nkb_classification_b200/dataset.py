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


def _row_pitch(w: int) -> int:
    return (w * 3 + 15) // 16 * 16      # 16-byte aligned rows: K1's TMA path applies


def _unique_inverse(ids: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """np.unique(ids, return_inverse=True) for small non-negative integer ids, without the sort (O(n))."""
    ids = np.asarray(ids, dtype=np.int64)
    if ids.size == 0 or ids.min() < 0 or ids.max() > 8 * ids.size + 1024:
        return np.unique(ids, return_inverse=True)
    seen = np.zeros(int(ids.max()) + 1, dtype=bool)
    seen[ids] = True
    uniq = np.flatnonzero(seen)
    lut = np.cumsum(seen) - 1
    return uniq, lut[ids]


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

    def describe(self, indices):
        """One batch at once: (frame id per sample [n] int64, path of every id, boxes int64 [n,4] with -1 rows for
        whole-image samples, collated target).  Generic per-sample walk; datasets that keep arrays override it."""
        ids_of: Dict[str, int] = {}
        n = len(indices)
        ids = np.empty(n, dtype=np.int64)
        boxes = np.full((n, 4), -1, dtype=np.int64)
        labels = []
        for j, i in enumerate(indices):
            p, b, l = self[i]
            ids[j] = ids_of.setdefault(p, len(ids_of))
            if b is not None:
                boxes[j] = b
            labels.append(l)
        return ids, list(ids_of), boxes, collate_targets(labels)


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
        self._frame_sizes: Dict[str, Tuple[int, int]] = {}      # every frame's (h, w): sizes the device frame cache
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
            self._frame_sizes[str(image_filename)] = image_size
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

    def describe(self, indices):
        if getattr(self, "_arr_len", -1) != len(self.list_bbox):       # (re)build the sample arrays once
            paths: Dict[str, int] = {}
            self._pid = np.array([paths.setdefault(p, len(paths)) for p, _, _ in self.list_bbox], dtype=np.int64)
            self._plist = list(paths)
            self._barr = np.array([b for _, b, _ in self.list_bbox], dtype=np.int64).reshape(-1, 4)
            self._larr = np.array([l for _, _, l in self.list_bbox], dtype=np.int64)
            self._arr_len = len(self.list_bbox)
        idx = np.asarray(indices, dtype=np.int64)
        uniq, inv = _unique_inverse(self._pid[idx])
        return inv.astype(np.int64), [self._plist[u] for u in uniq], self._barr[idx], torch.from_numpy(self._larr[idx])

    def decoded_bytes(self) -> int:
        """Bytes the decoded frames that carry at least one sample take in the device frame cache."""
        used = {p for p, _, _ in self.list_bbox}
        return sum((h * _row_pitch(w) + 255) // 256 * 256 for p, (h, w) in self._frame_sizes.items() if p in used)


class InMemoryFrames(_DescDataset):
    """Already-decoded frames with per-box samples (synthetic benchmarks, tests, callers that decode elsewhere -- e.g. a
    video pipeline feeding `Evaluator.classify_crops`-style many-crops-per-frame work).  ``frames``: sequence of uint8
    HWC arrays in cv2.imread's channel order (BGR); ``boxes`` [n,4] int or None (whole frames); ``frame_idx`` [n];
    ``labels`` [n] int or dict name -> [n].  ``read_frame`` stands in for the decode, so the loader's decode count,
    frame cache and ROI upload apply unchanged."""

    def __init__(self, frames, frame_idx, labels, boxes=None, classes=None, transform=None):
        self.frames, self.frame_idx = frames, np.asarray(frame_idx, dtype=np.int64)
        self.boxes = None if boxes is None else np.asarray(boxes, dtype=np.int64).reshape(-1, 4)
        self.labels = labels
        self.transform = transform
        if classes is None:
            classes = ({k: sorted(set(np.asarray(v).tolist())) for k, v in labels.items()} if isinstance(labels, dict)
                       else sorted(set(np.asarray(labels).tolist())))
        self.classes = classes
        self.class_to_idx, self.idx_to_class = get_classes_configs(classes)
        self.reads = 0

    def __len__(self):
        return len(self.frame_idx)

    def path_at(self, idx):
        return f"mem://{int(self.frame_idx[idx])}"

    def box_at(self, idx):
        return None if self.boxes is None else tuple(int(v) for v in self.boxes[idx])

    def label_at(self, idx):
        if isinstance(self.labels, dict):
            return {k: np.array(v[idx], dtype=np.int64) for k, v in self.labels.items()}
        return np.array(self.labels[idx], dtype=np.int64)

    def get_labels(self):
        if isinstance(self.labels, dict):
            return np.stack([np.asarray(self.labels[k]) for k in sorted(self.labels)], 1)
        return np.asarray(self.labels)

    def read_frame(self, path: str) -> np.ndarray:
        self.reads += 1
        return self.frames[int(path[6:])]

    def describe(self, indices):
        idx = np.asarray(indices, dtype=np.int64)
        uniq, inv = _unique_inverse(self.frame_idx[idx])
        boxes = self.boxes[idx] if self.boxes is not None else np.full((len(idx), 4), -1, dtype=np.int64)
        if isinstance(self.labels, dict):
            target = {k: torch.from_numpy(np.asarray(v, dtype=np.int64)[idx]) for k, v in self.labels.items()}
        else:
            target = torch.from_numpy(np.asarray(self.labels, dtype=np.int64)[idx])
        return inv.astype(np.int64), [f"mem://{int(u)}" for u in uniq], boxes, target

    def decoded_bytes(self) -> int:
        return sum((f.shape[0] * _row_pitch(f.shape[1]) + 255) // 256 * 256 for f in self.frames)


# --------------------------------------------------------------------------------------------
# loader: decode once per frame, keep decoded frames in HBM, one K1 launch per batch
# --------------------------------------------------------------------------------------------
def pack_frames(frames: Sequence[np.ndarray], staging: Optional[torch.Tensor] = None):
    """Lay frames of arbitrary sizes into one pinned uint8 buffer with 16-byte aligned rows (so K1's TMA path
    applies).  Returns (buffer tensor view, total bytes, int64 [F,4] descriptors, list of (h, w))."""
    descs, sizes, off = [], [], 0
    for f in frames:
        h, w = f.shape[:2]
        pitch = _row_pitch(w)
        descs.append((off, h, w, pitch))
        sizes.append((h, w))
        off += (h * pitch + 255) // 256 * 256
    total = max(off, 16)
    if staging is None or staging.numel() < total:
        staging = torch.empty(int(total * 1.25) + 4096, dtype=torch.uint8)
        if torch.cuda.is_available():
            staging = staging.pin_memory()
    buf = staging.numpy()
    for f, (o, h, w, pitch) in zip(frames, descs):
        if pitch == w * 3 and f.flags.c_contiguous:       # rows already 16-byte multiples: one flat copy
            buf[o: o + h * pitch] = f.reshape(-1)
        else:
            buf[o: o + h * pitch].reshape(h, pitch)[:, : w * 3] = f.reshape(h, w * 3)
    return staging, total, torch.tensor(descs, dtype=torch.int64).reshape(-1, 4), sizes


def validate_targets(target, classes, ignore_index: int = -100) -> None:
    """Labels outside [0, C) that are not ``ignore_index`` would be silently IGNORED by the kernels (torch's loss
    device-asserts on them): refuse them here, on the host, where a wrong class map or label column shows up for free."""
    def check(t, n_classes, name):
        if not isinstance(t, torch.Tensor) or t.numel() == 0 or t.is_cuda or t.dtype.is_floating_point:
            return
        lo, hi = int(t.min()), int(t.max())
        if hi >= n_classes or (lo < 0 and bool(((t < 0) & (t != ignore_index)).any())):
            raise ValueError(f"label outside [0, {n_classes}) for {name or 'the target'}: min {lo}, max {hi} "
                             f"(only ignore_index = {ignore_index} may lie outside)")

    if isinstance(target, dict) and isinstance(classes, dict):
        for k, v in target.items():
            if k in classes:
                check(v, len(classes[k]), k)
    elif isinstance(target, torch.Tensor) and isinstance(classes, (list, tuple)):
        check(target, len(classes), "")


def collate_targets(labels: list):
    """What default_collate makes of the reference's per-sample labels."""
    first = labels[0]
    if isinstance(first, dict):
        return {k: torch.from_numpy(np.stack([np.asarray(l[k], dtype=np.int64) for l in labels])) for k in first}
    if isinstance(first, str):
        return list(labels)
    return torch.from_numpy(np.asarray(labels, dtype=np.int64))


class DeviceFrameCache:
    """Decoded frames resident in HBM, keyed by path: one arena, FIFO ring eviction.

    K1 addresses its sources as `base + int64 offset`, so a batch can mix frames that were uploaded epochs ago with
    frames uploaded a moment ago: from the second epoch on a dataset that fits uploads nothing and decodes nothing
    (180 GB of HBM hold ~25 k decoded 1080p frames).  Entries referenced by a batch that is staged but whose K1 has
    not run yet are pinned and never evicted; a frame that cannot be placed (arena full of pinned entries, or larger
    than the arena) simply travels with its batch like an uncached one."""

    ALIGN = 256

    def __init__(self, device, capacity_bytes: int):
        from collections import OrderedDict
        self.cap = int(capacity_bytes) // self.ALIGN * self.ALIGN
        self.buf = torch.empty(max(self.cap, self.ALIGN), dtype=torch.uint8, device=device)   # ("cpu" in host-logic tests)
        self.map: "OrderedDict[str, Tuple[int, int, int, int, int]]" = OrderedDict()   # path -> (off, h, w, pitch, nbytes)
        self.head = 0
        self.evictions = 0

    def get(self, path: str):
        return self.map.get(path)

    def _evict_range(self, lo: int, hi: int, pinned) -> bool:
        while self.map:
            path, e = next(iter(self.map.items()))       # insertion order == ring order: the oldest sits right after head
            if e[0] < hi and e[0] + e[4] > lo:
                if path in pinned:
                    return False
                self.map.popitem(last=False)
                self.evictions += 1
            else:
                break
        return True

    def alloc(self, path: str, h: int, w: int, pinned) -> Optional[Tuple[int, int, int, int, int]]:
        pitch = _row_pitch(w)
        n = (h * pitch + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        if n > self.cap:
            return None
        if self.head + n > self.cap:
            if not self._evict_range(self.head, self.cap, pinned):
                return None
            self.head = 0
        if not self._evict_range(self.head, self.head + n, pinned):
            return None
        e = (self.head, h, w, pitch, n)
        self.head += n
        self.map[path] = e
        return e


class _IngestSlot:
    """One stage of the frame-ingest ring: pinned uint8 staging, its device twin (frames that travel with the batch), the
    paths the batch pins in the frame cache, and the two events that make reuse safe (`copied`: H2D finished, recorded on
    the copy stream; `consumed`: K1 has read its sources, recorded on the consumer's stream)."""

    def __init__(self):
        self.staging: Optional[torch.Tensor] = None
        self.dev: Optional[torch.Tensor] = None
        self.copied: Optional[torch.cuda.Event] = None
        self.consumed: Optional[torch.cuda.Event] = None
        self.paths: set = set()
        self.meta_h: Optional[torch.Tensor] = None      # pinned: boxes | frame indices | descriptors of the batch
        self.meta_d: Optional[torch.Tensor] = None


class DeviceCropLoader:
    """DataLoader stand-in: batches of sample descriptors -> K1 -> (device tensor, target).

    Frame ingest, against the reference's decode-per-crop + fp32 H2D (dataset.py:398-409, engine.py:40):

    * each distinct frame of a batch is decoded ONCE (thread pool) and copied to the device as uint8;
    * ``frame_cache_bytes`` > 0: decoded frames stay in a device arena keyed by path (:class:`DeviceFrameCache`), so a
      frame is decoded and uploaded once per EPOCH at most -- whatever the sampler does -- and not at all from the second
      epoch on when the dataset fits;
    * frames that do not go through the cache travel as their region of interest only: the bounding rectangle of the
      boxes the batch takes from them (``roi_upload``), so a shuffled loader never ships more than the crops need;
    * ``group_by_frame``: batches are cut from an order that keeps the samples of a frame adjacent (frames shuffled,
      samples shuffled inside a frame; the draws of a weighted sampler are grouped the same way) -- the reference's
      i.i.d. shuffle is the default, because grouping changes which samples share a batch;
    * with ``prefetch`` > 0 a producer thread runs decode + pack + H2D (on its own copy stream) for the next batches
      while the consumer's stream runs K1 / the model; a ring of ``prefetch + 2`` slots with `copied` / `consumed`
      events keeps buffers from being reused early.  ``prefetch = 0`` does the same work inline.

    ``stats`` counts decodes, uploaded bytes, cache hits and crops since construction (or ``reset_stats()``)."""

    def __init__(self, dataset: _DescDataset, batch_size: int, shuffle: bool = False, sampler=None,
                 num_workers: int = 0, drop_last: bool = False, device="cuda:0", out_dtype=torch.float32,
                 prefetch: Optional[int] = None, frame_cache_bytes: int = 0, roi_upload: bool = True,
                 group_by_frame: bool = False, sort_within_batch: bool = False, resident_epochs: bool = True,
                 targets_on_device: bool = False):
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
        self.roi_upload, self.group_by_frame = bool(roi_upload), bool(group_by_frame)
        # reorder the samples INSIDE a batch by frame (same samples, same batch): K1 then walks the boxes of a frame
        # back to back and their overlapping source rows hit L2 instead of HBM.  Off by default: it changes the order in
        # which a batch's samples (and the logger's per-sample lists) appear.
        self.sort_within_batch = bool(sort_within_batch)
        # once every frame the epoch touches sits in the frame cache, the whole epoch is planned at once: boxes, frame
        # indices, descriptors and labels of ALL its batches go to the device in one copy and a batch is one K1 launch on
        # slices of them -- no producer thread, no per-batch host work (`_iter_resident`)
        self.resident_epochs, self.targets_on_device = bool(resident_epochs), bool(targets_on_device)
        self.cache = DeviceFrameCache(self.device, frame_cache_bytes) if frame_cache_bytes and frame_cache_bytes > 0 else None
        self._read = getattr(dataset, "read_frame", None) or _imread_bgr      # in-memory datasets supply their own
        # train pipelines: where the per-sample augmentation parameters come from (None = Python's global `random`,
        # which is what albumentations draws from, so `random.seed(...)` governs it as in the reference)
        self.aug_rng = None
        self.reset_stats()

    def reset_stats(self):
        self.stats = {"batches": 0, "crops": 0, "decodes": 0, "h2d_bytes": 0, "cache_hits": 0, "cache_inserts": 0,
                      "roi_frames": 0, "whole_frames": 0, "resident_epochs": 0}

    def __len__(self):
        n = len(self.sampler) if self.sampler is not None else len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def _order(self):
        """Sample indices of one epoch: a list, or -- on the plain shuffle / in-order paths, where nothing walks it
        element by element -- an int64 numpy array (a 100 k-element list costs milliseconds to build and to convert)."""
        if self.sampler is not None:
            order = list(iter(self.sampler))
        elif self.shuffle:
            order = torch.randperm(len(self.dataset)).numpy()
        else:
            order = np.arange(len(self.dataset), dtype=np.int64)
        if self.group_by_frame and (self.sampler is not None or self.shuffle):
            order = [int(i) for i in order]
            # keep the draw (which samples, how often) and make the samples of a frame adjacent: frames in the order of
            # their first draw -- itself random -- and samples inside a frame in draw order
            first, groups = {}, {}
            for pos, i in enumerate(order):
                p = self.dataset.path_at(i)
                first.setdefault(p, pos)
                groups.setdefault(p, []).append(i)
            order = [i for p in sorted(groups, key=first.get) for i in groups[p]]
        return order

    def _index_batches(self, order=None) -> Iterator[List[int]]:
        if order is None:
            order = self._order()
        for i in range(0, len(order), self.batch_size):
            b = order[i: i + self.batch_size]
            if len(b) < self.batch_size and self.drop_last:
                return
            yield b

    # ---- producer side: host work + H2D of one batch into ring slot `k` ----
    def _stage(self, indices: List[int], k: int) -> dict:
        slot = self._slots[k]
        if slot.consumed is not None:
            slot.consumed.synchronize()          # K1 of the batch that used this slot last has read its sources
        slot.paths = set()                       # ... so that batch no longer pins anything in the frame cache
        ids, plist, raw_boxes, target = self.dataset.describe(indices)
        validate_targets(target, getattr(self.dataset, "classes", None))
        n = len(ids)
        if self.sort_within_batch and n > 1:
            perm = np.argsort(ids, kind="stable")
            ids, raw_boxes = ids[perm], raw_boxes[perm]
            pt = torch.from_numpy(perm)
            if isinstance(target, dict):
                target = {k_: v[pt] for k_, v in target.items()}
            elif isinstance(target, torch.Tensor):
                target = target[pt]
            else:
                target = [target[i] for i in perm]
        fidx = ids.astype(np.int32)
        has_box = raw_boxes[:, 0] >= 0
        cache = self.cache
        # frames of batches that are staged but not consumed yet must survive in the cache, and so must this batch's
        pinned = set(plist)
        for sl in self._slots:
            pinned |= sl.paths
        entries: List[Optional[tuple]] = [cache.get(p) if cache is not None else None for p in plist]
        miss = [i for i, e in enumerate(entries) if e is None]
        self.stats["cache_hits"] += len(plist) - len(miss)
        mpaths = [plist[i] for i in miss]
        decoded = list(self.pool.map(self._read, mpaths)) if self.pool else [self._read(p) for p in mpaths]
        self.stats["decodes"] += len(decoded)
        # per frame: where K1 finds it.  origin = (x, y) of the uploaded region inside the frame (boxes are shifted by it)
        sizes = np.zeros((len(plist), 2), dtype=np.int64)            # (h, w) of the region K1 sees
        origin = np.zeros((len(plist), 2), dtype=np.int64)
        for i, e in enumerate(entries):
            if e is not None:
                sizes[i] = (e[1], e[2])
        if miss:                                                     # (a fully resident batch skips all of this)
            order = np.argsort(fidx, kind="stable")                  # samples grouped by frame, for the ROI rectangles
            starts = np.searchsorted(fidx[order], np.arange(len(plist) + 1))
        to_cache, to_travel = [], []            # (frame index, pixels)
        for i, fr in zip(miss, decoded):
            H, W = fr.shape[:2]
            e = cache.alloc(plist[i], H, W, pinned) if cache is not None else None
            if e is not None:
                entries[i], sizes[i] = e, (H, W)
                to_cache.append((i, fr))
                continue
            x0, y0, x1, y1 = 0, 0, W, H
            rows = order[starts[i]: starts[i + 1]]
            if self.roi_upload and len(rows) and has_box[rows].all():
                a = raw_boxes[rows]
                x0, y0 = max(0, int(a[:, 0].min())), max(0, int(a[:, 1].min()))
                x1, y1 = min(W, int(a[:, 2].max())), min(H, int(a[:, 3].max()))
                if x1 - x0 < 2 or y1 - y0 < 1:          # degenerate boxes: let validate_boxes speak on the full frame
                    x0, y0, x1, y1 = 0, 0, W, H
            self.stats["roi_frames" if (x1 - x0, y1 - y0) != (W, H) else "whole_frames"] += 1
            origin[i] = (x0, y0)
            sizes[i] = (y1 - y0, x1 - x0)
            to_travel.append((i, fr[y0:y1, x0:x1]))
        # one pinned staging buffer: [frames going into the cache | regions travelling with the batch]
        pieces = [fr for _, fr in to_cache] + [fr for _, fr in to_travel]
        slot.staging, total, pdesc, _ = pack_frames(pieces, slot.staging)
        dev = self.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        if slot.copied is None:
            slot.copied = torch.cuda.Event()
        n_cache = len(to_cache)
        travel_lo = int(pdesc[n_cache, 0]) if len(to_travel) else total
        travel_bytes = total - travel_lo if len(to_travel) else 0
        boxes = raw_boxes.copy()
        if not has_box.all():                    # whole-image samples: the box is the frame
            nb = ~has_box
            boxes[nb, 0], boxes[nb, 1] = 0, 0
            boxes[nb, 2], boxes[nb, 3] = sizes[fidx[nb], 1], sizes[fidx[nb], 0]
        if to_travel:                            # regions of interest: boxes move with their region's origin
            boxes[:, [0, 2]] -= origin[fidx, 0:1]
            boxes[:, [1, 3]] -= origin[fidx, 1:2]
        boxes = boxes.astype(np.int32)
        validate_boxes(boxes, fidx, sizes)
        aug = self.plan.draw(n, self.aug_rng)     # in batch order: one producer, so the stream is reproducible
        with torch.cuda.stream(self._copy_stream):
            # (allocated on the copy stream, so the caching allocator orders its first use after whatever freed the block;
            # the consumer records its own use in _consume)
            if travel_bytes and (slot.dev is None or slot.dev.numel() < travel_bytes):
                slot.dev = torch.empty(int(travel_bytes * 1.25) + 4096, dtype=torch.uint8, device=dev)
            base = cache.buf if cache is not None else slot.dev
            if base is None:                         # no cache and nothing travels (cannot happen: every frame is a miss)
                slot.dev = torch.empty(4096, dtype=torch.uint8, device=dev)
                base = slot.dev
            for j, (i, fr) in enumerate(to_cache):
                o, e = int(pdesc[j, 0]), entries[i]
                cache.buf[e[0]: e[0] + e[1] * e[3]].copy_(slot.staging[o: o + e[1] * e[3]], non_blocking=True)
                self.stats["h2d_bytes"] += e[1] * e[3]
            if travel_bytes:
                slot.dev[:travel_bytes].copy_(slot.staging[travel_lo:total], non_blocking=True)
                self.stats["h2d_bytes"] += travel_bytes
            desc = np.zeros((len(plist), 4), dtype=np.int64)
            for i, e in enumerate(entries):
                if e is not None:
                    desc[i] = (e[0], e[1], e[2], e[3])           # offset inside the arena (= base)
            for j, (i, fr) in enumerate(to_travel):
                o, h, w, pitch = (int(x) for x in pdesc[n_cache + j])
                # K1 addresses sources as base + int64 offset: a region in the slot's buffer is reached from the arena
                off = (slot.dev.data_ptr() - base.data_ptr()) + (o - travel_lo)
                desc[i] = (off, h, w, pitch)
            # boxes | frame indices | descriptors travel as ONE copy out of the slot's pinned staging into the slot's own
            # device buffer (recycled only after K1 of the batch has run: no allocation, no record_stream per batch)
            nb, nf, nd = boxes.nbytes, fidx.nbytes, desc.nbytes
            o_f, o_d = nb, (nb + nf + 7) // 8 * 8
            mbytes = o_d + nd
            if slot.meta_h is None or slot.meta_h.numel() < mbytes:
                slot.meta_h = torch.empty(int(mbytes * 1.5) + 256, dtype=torch.uint8).pin_memory()
                slot.meta_d = torch.empty(slot.meta_h.numel(), dtype=torch.uint8, device=dev)
            mh = slot.meta_h.numpy()
            mh[:nb] = boxes.view(np.uint8).reshape(-1)
            mh[o_f: o_f + nf] = fidx.view(np.uint8).reshape(-1)
            mh[o_d: o_d + nd] = desc.view(np.uint8).reshape(-1)
            slot.meta_d[:mbytes].copy_(slot.meta_h[:mbytes], non_blocking=True)
            md = slot.meta_d
            meta = dict(boxes=md[:nb].view(torch.int32).view(-1, 4), fidx=md[o_f: o_f + nf].view(torch.int32),
                        desc=md[o_d: o_d + nd].view(torch.int64).view(-1, 4))
            self.stats["h2d_bytes"] += nb + nf + nd
            if aug is not None:
                ops.upload_augment(aug, dev)      # train pipelines: parameters ride the copy stream too
            slot.copied.record(self._copy_stream)
        self.stats["cache_inserts"] += n_cache
        self.stats["batches"] += 1
        self.stats["crops"] += n
        slot.paths = set(plist) if cache is not None else set()
        # targets stay host tensors, as default_collate yields them, but in pinned memory (the reference's DataLoader
        # pins too): the consumer's `.to(device, non_blocking=True)` is then a true asynchronous copy, not a sync
        if torch.cuda.is_available() and self.device.type == "cuda":
            if isinstance(target, dict):
                target = {k_: self._pinned(slot, k_, v) for k_, v in target.items()}
            elif isinstance(target, torch.Tensor):
                target = self._pinned(slot, "", target)
        meta.update(slot=k, base=base, aug=aug, target=target)
        return meta

    @staticmethod
    def _pinned(slot, key, t: torch.Tensor) -> torch.Tensor:
        """A fresh pinned copy per batch (torch's caching host allocator recycles the blocks): callers may keep the
        targets of earlier batches, as they can with the reference's DataLoader, so the buffer must not be reused."""
        return t.pin_memory()

    # ---- consumer side: K1 on the caller's current stream ----
    def _consume(self, meta: dict):
        slot = self._slots[meta["slot"]]
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(slot.copied)
        shared = []          # (the slot's meta buffer lives as long as the loader: nothing to record for it)
        if slot.dev is not None:
            shared.append(slot.dev)
        if meta["aug"] is not None:
            shared += [t for t in ops.upload_augment(meta["aug"], self.device) if t is not None]
        for t in shared:
            t.record_stream(cur)                 # allocated on the copy stream, read on this one
        img = ops.preprocess_crops(meta["base"], meta["boxes"], meta["fidx"], self.plan,
                                   out_dtype=self.out_dtype, frame_desc=meta["desc"], aug=meta["aug"])
        if slot.consumed is None:
            slot.consumed = torch.cuda.Event()
        slot.consumed.record(cur)
        return img, meta["target"]

    def load_batch(self, indices: List[int]):
        return self._consume(self._stage(indices, 0))

    # ---- resident epochs: the whole epoch planned at once ----
    def _plan_resident(self, order: List[int]) -> Optional[dict]:
        """Everything K1 needs for EVERY batch of the epoch, on the device after one copy -- or None when a frame of the
        epoch is not in the cache (first epoch, dataset larger than the arena).  Train pipelines draw their per-sample
        augmentation parameters batch by batch inside the loop (`PreprocessPlan.draw`, vectorised: ~2.7 ms per 4096
        samples; the hue / sat / val tables are built on the device)."""
        cache = self.cache
        if (not self.resident_epochs or cache is None or len(order) == 0
                or self.device.type != "cuda" or type(self.dataset).describe is _DescDataset.describe):
            return None
        if self.drop_last:
            order = order[: len(order) // self.batch_size * self.batch_size]
            if len(order) == 0:
                return None
        ids, plist, raw_boxes, target = self.dataset.describe(order)
        entries = [cache.get(p) for p in plist]
        if any(e is None for e in entries) or not isinstance(target, (dict, torch.Tensor)):
            return None
        validate_targets(target, getattr(self.dataset, "classes", None))
        n, bs = len(ids), self.batch_size
        perm = None
        if self.sort_within_batch and n > 1:
            if len(plist) <= 65536 and n % bs == 0:               # 16-bit keys: numpy's stable sort is a radix sort
                perm = (np.argsort(ids.astype(np.uint16).reshape(-1, bs), axis=1, kind="stable")
                        + (np.arange(n // bs, dtype=np.int64) * bs)[:, None]).reshape(-1)
            else:
                perm = np.lexsort((ids, np.arange(n) // bs))     # stable: by frame inside each batch, draw order kept
            ids, raw_boxes = ids[perm], raw_boxes[perm]
            pt = torch.from_numpy(perm)
            target = {k_: v[pt] for k_, v in target.items()} if isinstance(target, dict) else target[pt]
        desc = np.array([(e[0], e[1], e[2], e[3]) for e in entries], dtype=np.int64).reshape(-1, 4)
        fidx = ids.astype(np.int32)
        boxes = raw_boxes.astype(np.int64, copy=True)
        nb_ = boxes[:, 0] < 0                                    # whole-image samples: the box is the frame
        if nb_.any():
            boxes[nb_, 0], boxes[nb_, 1] = 0, 0
            boxes[nb_, 2], boxes[nb_, 3] = desc[fidx[nb_], 2], desc[fidx[nb_], 1]
        boxes = boxes.astype(np.int32)
        validate_boxes(boxes, fidx, desc[:, 1:3])
        names = sorted(target) if isinstance(target, dict) else None
        lab = (torch.stack([target[k_].reshape(-1) for k_ in names], 0) if names is not None
               else target.reshape(1, -1)).to(torch.int64).contiguous()          # [T, n]: a task's batch slice is contiguous
        # one pinned buffer, one H2D: [boxes | frame indices | descriptors | labels]
        parts = [boxes.view(np.uint8).reshape(-1), fidx.view(np.uint8).reshape(-1), desc.view(np.uint8).reshape(-1),
                 lab.numpy().view(np.uint8).reshape(-1)]
        offs, total = [], 0
        for a in parts:
            offs.append(total)
            total += (a.nbytes + 15) // 16 * 16
        host = torch.empty(total, dtype=torch.uint8).pin_memory()
        hv = host.numpy()
        for o, a in zip(offs, parts):
            hv[o: o + a.nbytes] = a
        dev = host.to(self.device, non_blocking=True)
        self.stats["h2d_bytes"] += total

        def view(t, i, dt, shape):
            return t[offs[i]: offs[i] + parts[i].nbytes].view(dt).view(*shape)

        lab_src = dev if self.targets_on_device else host      # host targets: pinned views, as default_collate + pin
        lab_t = view(lab_src, 3, torch.int64, (-1, n))
        self.stats["cache_hits"] += len(plist)
        return dict(n=n, boxes=view(dev, 0, torch.int32, (n, 4)), fidx=view(dev, 1, torch.int32, (n,)),
                    desc=view(dev, 2, torch.int64, (-1, 4)), labels=lab_t, names=names, keep=(host, dev))

    def _iter_resident(self, plan: dict):
        bs, n, names = self.batch_size, plan["n"], plan["names"]
        self.stats["resident_epochs"] += 1
        for a in range(0, n, bs):
            b = min(n, a + bs)
            aug = self.plan.draw(b - a, self.aug_rng)            # None for a validation / inference pipeline
            img = ops.preprocess_crops(self.cache.buf, plan["boxes"][a:b], plan["fidx"][a:b], self.plan,
                                       out_dtype=self.out_dtype, frame_desc=plan["desc"], aug=aug)
            lab = plan["labels"][:, a:b]
            target = lab[0] if names is None else {k_: lab[t] for t, k_ in enumerate(names)}
            self.stats["batches"] += 1
            self.stats["crops"] += b - a
            yield img, target

    def __iter__(self):
        order = self._order()
        plan = self._plan_resident(order)
        if plan is not None:
            yield from self._iter_resident(plan)
            return
        if self.prefetch <= 0:
            for b in self._index_batches(order):
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
                for i, b in enumerate(self._index_batches(order)):
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


_LOADER_KEYS = ("device", "prefetch", "frame_cache_bytes", "roi_upload", "group_by_frame", "sort_within_batch",
                "resident_epochs", "targets_on_device")


def _frame_cache_bytes(data: dict, device, dataset) -> int:
    """``frame_cache_bytes``: bytes of HBM for decoded frames; 0 = off; "auto" (default for datasets whose frames are
    shared by many samples) = what the dataset's decoded frames need, at most half of the free device memory."""
    shared = data["type"] in ("AnnotatedYOLODataset", "InMemoryFrames")
    v = data.get("frame_cache_bytes", "auto" if shared else 0)
    if v == "auto":
        if not torch.cuda.is_available() or not hasattr(dataset, "decoded_bytes"):
            return 0
        free, _ = torch.cuda.mem_get_info(torch.device(device))
        return int(min(dataset.decoded_bytes() + (1 << 20), free // 2))
    return int(v or 0)


def get_dataset(data, pipeline):
    """dataset.py:541-629: same ``data`` keys (type, batch_size, num_workers, shuffle, drop_last,
    weighted_sampling, dataset kwargs); optional extra keys ``device``, ``prefetch``, ``frame_cache_bytes``,
    ``roi_upload``, ``group_by_frame`` (see :class:`DeviceCropLoader`)."""
    transform = Transforms(pipeline)
    kind = data["type"]
    kwargs = {k: v for k, v in data.items() if k not in _LOADER_KEYS}
    if kind == "GroupsDataset":
        dataset = GroupsDataset(transform=transform, **kwargs)
    elif kind == "AnnotatedMultitaskDataset":
        dataset = AnnotatedMultitaskDataset(transform=transform, **kwargs)
    elif kind == "AnnotatedSingletaskDataset":
        dataset = AnnotatedSingletaskDataset(transform=transform, **kwargs)
    elif kind == "AnnotatedYOLODataset":
        dataset = AnnotatedYOLODataset(transform=transform, **kwargs)
    elif kind == "InMemoryFrames":       # synthetic / already-decoded frames (bench.py --api engine, tests)
        dataset = data["dataset"]
        dataset.transform = transform
    else:
        dataset = ImageFolder(data["root"], transform=transform)
    sampler = ImbalancedDatasetSampler(dataset) if data.get("weighted_sampling", False) else None
    dev = _device_of(data)
    return DeviceCropLoader(dataset, batch_size=data["batch_size"], shuffle=data.get("shuffle", False), sampler=sampler,
                            num_workers=data.get("num_workers", 0), drop_last=data.get("drop_last", False),
                            device=dev, prefetch=data.get("prefetch"), frame_cache_bytes=_frame_cache_bytes(data, dev, dataset),
                            roi_upload=data.get("roi_upload", True), group_by_frame=data.get("group_by_frame", False),
                            sort_within_batch=data.get("sort_within_batch", False),
                            resident_epochs=data.get("resident_epochs", True),
                            targets_on_device=data.get("targets_on_device", False))


def get_inference_dataset(data, pipeline):
    """dataset.py:632-644: yields (img, paths)."""
    dataset = InferDataset(folder_path=data["folder_path"], transform=Transforms(pipeline))
    return DeviceCropLoader(dataset, batch_size=data["batch_size"], num_workers=data.get("num_workers", 0),
                            device=_device_of(data), prefetch=data.get("prefetch"))
