"""Host-side box arithmetic (row a1 of SURVEY.md section 8): float64, once per dataset.

Mirrors ``AnnotatedYOLODataset.bbox_xywhn2xyxy`` / ``check_boxes_sizes_annotation``
(nkb_classification/dataset.py:414-421, :433-434), the label-file parse
(:350-359) and ``Evaluator.classify_crops``' box conversion
(metrics/det_cls_val.py:231-236).  This stays on the host by design: it is
text parsing and a handful of float64 operations per box, done at dataset
construction, and its int() truncation / clip semantics must be bit-exact.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import numpy as np


def bbox_xywhn2xyxy(x_center, y_center, width, height, image_size) -> Tuple[int, int, int, int]:
    """Same signature and result as the reference's static method (dataset.py:414-421)."""
    image_height, image_width = image_size
    x_min = np.clip(int((x_center - width / 2) * image_width), 0, image_width)
    y_min = np.clip(int((y_center - height / 2) * image_height), 0, image_height)
    x_max = np.clip(int((x_center + width / 2) * image_width), 0, image_width)
    y_max = np.clip(int((y_center + height / 2) * image_height), 0, image_height)
    return int(x_min), int(y_min), int(x_max), int(y_max)


def bbox_xywhn2xyxy_batch(xywhn: np.ndarray, image_size) -> np.ndarray:
    """Vectorised form: float64 [n,4] (xc, yc, w, h) -> int32 [n,4]; np.trunc == Python int()."""
    a = np.asarray(xywhn, dtype=np.float64).reshape(-1, 4)
    ih, iw = image_size
    xc, yc, w, h = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
    out = np.empty((a.shape[0], 4), dtype=np.int64)
    out[:, 0] = np.clip(np.trunc((xc - w / 2) * iw), 0, iw)
    out[:, 1] = np.clip(np.trunc((yc - h / 2) * ih), 0, ih)
    out[:, 2] = np.clip(np.trunc((xc + w / 2) * iw), 0, iw)
    out[:, 3] = np.clip(np.trunc((yc + h / 2) * ih), 0, ih)
    return out.astype(np.int32)


def check_boxes_sizes(x_min, y_min, x_max, y_max, min_box_size: int = 5) -> bool:
    """dataset.py:433-434."""
    return bool(x_max - x_min >= min_box_size and y_max - y_min >= min_box_size)


def parse_yolo_label_lines(lines: Iterable[str], image_size, min_box_size: int = 5):
    """``cls xc yc w h`` rows -> list of ((x0,y0,x1,y1), label), dropping small boxes (dataset.py:350-359)."""
    out = []
    for line in lines:
        parts = line.split()
        if not parts:
            continue
        label = int(parts[0])
        x_center, y_center, width, height = tuple(map(float, parts[1:5]))
        box = bbox_xywhn2xyxy(x_center, y_center, width, height, image_size)
        if not check_boxes_sizes(*box, min_box_size=min_box_size):
            continue
        out.append((box, label))
    return out


def detector_boxes_to_int(boxes_xyxyn: np.ndarray, img_h: int, img_w: int) -> np.ndarray:
    """metrics/det_cls_val.py:231-236: normalised xyxy scaled by (W, H) then ``astype(int)``."""
    b = np.array(boxes_xyxyn, dtype=np.float64, copy=True).reshape(-1, 4)
    b[:, [0, 2]] *= img_w
    b[:, [1, 3]] *= img_h
    return b.astype(int).astype(np.int32)


def validate_boxes(boxes: np.ndarray, frame_idx: np.ndarray, frame_sizes: Sequence[Tuple[int, int]]) -> None:
    """Raise ValueError for a box that is empty or leaves its frame (cv2 would assert on such a crop)."""
    boxes = np.asarray(boxes).reshape(-1, 4)
    frame_idx = np.asarray(frame_idx).reshape(-1)
    hw = np.asarray(frame_sizes, dtype=np.int64).reshape(-1, 2)
    if (frame_idx < 0).any() or (frame_idx >= len(hw)).any():
        raise ValueError("frame index out of range")
    h, w = hw[frame_idx, 0], hw[frame_idx, 1]
    bad = (boxes[:, 0] < 0) | (boxes[:, 1] < 0) | (boxes[:, 2] > w) | (boxes[:, 3] > h) | \
          (boxes[:, 2] <= boxes[:, 0]) | (boxes[:, 3] <= boxes[:, 1])
    if bad.any():
        i = int(np.flatnonzero(bad)[0])
        raise ValueError(f"box {i} = {boxes[i].tolist()} is empty or outside its {int(h[i])}x{int(w[i])} frame")
