"""CPU oracle: argmax, confusion matrix, balanced accuracy, epoch metrics.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates:

* ``BaseLogger.log_iter`` argmax               nkb_classification/logging.py:272-281
  (first maximal index wins ties; a NaN counts as maximal -- torch semantics)
* ``compute_targetwise_metrics``               nkb_classification/metrics.py:7-51
* ``compute_metrics``                          nkb_classification/metrics.py:54-70
* sklearn ``balanced_accuracy_score`` from an integer confusion matrix
  (SURVEY.md section 9.5).
"""
from __future__ import annotations

import warnings
from typing import Dict, List, Sequence

import numpy as np


def argmax_first(logits: np.ndarray) -> np.ndarray:
    """torch.argmax semantics on a [B, C] array: NaN is maximal, lowest index
    wins among equals."""
    logits = np.asarray(logits)
    B, C = logits.shape
    out = np.zeros(B, dtype=np.int64)
    for i in range(B):
        best, bi = logits[i, 0], 0
        for j in range(1, C):
            v = logits[i, j]
            if np.isnan(best):
                break
            if np.isnan(v) or v > best:
                best, bi = v, j
        out[i] = bi
    return out


def argmax_first_fast(logits: np.ndarray) -> np.ndarray:
    """Vectorised equivalent of ``argmax_first`` (np.argmax already returns the
    first maximum and treats NaN as maximal)."""
    return np.argmax(np.asarray(logits), axis=1).astype(np.int64)


def confusion_matrix(gt: np.ndarray, pred: np.ndarray, n_classes: int, ignore_index: int = -100) -> np.ndarray:
    """CM[gt][pred] += 1 in int64; rows with gt outside [0, C) (ignore_index)
    are skipped."""
    cm = np.zeros((n_classes, n_classes), dtype=np.int64)
    gt = np.asarray(gt).astype(np.int64)
    pred = np.asarray(pred).astype(np.int64)
    ok = (gt >= 0) & (gt < n_classes)
    np.add.at(cm, (gt[ok], pred[ok]), 1)
    return cm


def balanced_accuracy_from_cm(cm: np.ndarray) -> float:
    """sklearn.metrics.balanced_accuracy_score restated on the matrix: mean
    over classes present in gt of diag / row-sum, float64."""
    cm = np.asarray(cm, dtype=np.int64)
    row = cm.sum(axis=1)
    present = row > 0
    if not present.any():
        return float("nan")
    per_class = np.diag(cm)[present].astype(np.float64) / row[present].astype(np.float64)
    return float(np.mean(per_class))


def compute_targetwise_metrics(epoch_results, target_name=None):
    """metrics.py:7-51, calling sklearn exactly as the reference does."""
    from sklearn.metrics import balanced_accuracy_score, roc_auc_score
    from sklearn.preprocessing import label_binarize

    if target_name is None:
        running_loss = epoch_results["running_loss"]
        confidences = epoch_results["confidences"]
        predictions = epoch_results["predictions"]
        ground_truth = epoch_results["ground_truth"]
    else:
        running_loss = epoch_results["running_loss"][target_name]
        confidences = epoch_results["confidences"][target_name]
        predictions = epoch_results["predictions"][target_name]
        ground_truth = epoch_results["ground_truth"][target_name]
    confidences = np.array(confidences)
    n_classes = confidences.shape[1]
    gt_classes = np.unique(ground_truth)
    gt_n_classes = len(gt_classes)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        epoch_acc = balanced_accuracy_score(ground_truth, predictions)
    if n_classes > 2:
        epoch_roc_auc = np.full(n_classes, np.nan)
        if gt_n_classes > 1:
            gt_bin = label_binarize(ground_truth, classes=range(n_classes))
            for c in gt_classes:
                epoch_roc_auc[c] = roc_auc_score(gt_bin[:, c], confidences[:, c])
    else:
        epoch_roc_auc = np.nan
        if gt_n_classes > 1:
            epoch_roc_auc = roc_auc_score(ground_truth, confidences[:, 1])
    return {"epoch_acc": epoch_acc, "epoch_roc_auc": epoch_roc_auc, "epoch_loss": np.mean(running_loss)}


def compute_metrics(task: str, epoch_results, target_names: Sequence[str] = ()):
    """metrics.py:54-70."""
    if task == "single":
        m = compute_targetwise_metrics(epoch_results)
        m["loss"] = epoch_results["running_loss"]
        return m
    elif task == "multi":
        m = {t: compute_targetwise_metrics(epoch_results, t) for t in target_names}
        m["loss"] = epoch_results["running_loss"]["loss"]
        m["epoch_acc"] = np.mean([m[t]["epoch_acc"] for t in target_names])
        return m
    raise ValueError(f"Unknown task type {task} for metric computation")


def roc_auc_counts(probs: np.ndarray, labels: np.ndarray, seg_offsets: Sequence[int]) -> np.ndarray:
    """int64 [NC, 3] = (num2, P, Q) per class column, the exact integers behind
    ``roc_auc_score(label_binarize(gt)[:, c], conf[:, c])`` (metrics.py:36-38): P / Q positives / negatives,
    num2 = 2 * #{pos > neg} + #{pos == neg}; AUC = num2 / (2 P Q) (Mann-Whitney U with mid-ranks == the trapezoid of
    the ROC curve with tied thresholds merged).  Plain numpy: sort the negatives, binary-search every positive."""
    probs = np.asarray(probs, dtype=np.float32)
    labels = np.asarray(labels, dtype=np.int64).reshape(probs.shape[0], -1)
    T = len(seg_offsets) - 1
    out = np.zeros((seg_offsets[-1], 3), dtype=np.int64)
    for t in range(T):
        for k in range(seg_offsets[t + 1] - seg_offsets[t]):
            c = seg_offsets[t] + k
            is_pos = labels[:, t] == k
            pos, neg = probs[is_pos, c], np.sort(probs[~is_pos, c])
            lo = np.searchsorted(neg, pos, side="left").astype(np.int64)
            hi = np.searchsorted(neg, pos, side="right").astype(np.int64)
            out[c] = (2 * lo.sum() + (hi - lo).sum(), pos.size, neg.size)
    return out
