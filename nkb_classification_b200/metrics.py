"""Epoch metrics with the reference's ``compute_metrics`` interface
(nkb_classification/metrics.py).  Balanced accuracy comes from K3's integer
confusion matrix when the epoch results carry one (bit-identical to sklearn's
``balanced_accuracy_score``).  ROC-AUC comes from K5's integer
pair counts when the epoch results carry them (``"roc_auc_counts"``: AUC =
num2 / (2 P Q), the Mann-Whitney form of sklearn's trapezoid, equal to
``roc_auc_score`` to float64 rounding); otherwise sklearn runs on the fp32
probabilities K2 emitted, exactly as the reference does."""
from __future__ import annotations

import warnings

import numpy as np
from sklearn.metrics import balanced_accuracy_score, roc_auc_score
from sklearn.preprocessing import label_binarize


def balanced_accuracy_from_confusion(cm) -> float:
    """mean over classes present in the ground truth of diag / row-sum (float64)."""
    cm = np.asarray(cm, dtype=np.int64)
    row = cm.sum(axis=1)
    present = row > 0
    if not present.any():
        return float("nan")
    return float(np.mean(np.diag(cm)[present].astype(np.float64) / row[present].astype(np.float64)))


def roc_auc_from_counts(counts, n_classes: int):
    """K5's int64 [C, 3] (num2, P, Q) -> what metrics.py:33-42 returns: for C > 2 an array with the one-vs-rest AUC of
    every class present in the ground truth (NaN elsewhere; all NaN with fewer than two classes present), for C == 2
    the scalar AUC of class 1."""
    counts = np.asarray(counts, dtype=np.int64).reshape(n_classes, 3)
    present = counts[:, 1] > 0
    with np.errstate(invalid="ignore", divide="ignore"):
        auc = counts[:, 0].astype(np.float64) / (2.0 * counts[:, 1].astype(np.float64) * counts[:, 2].astype(np.float64))
    if n_classes > 2:
        out = np.full(n_classes, np.nan)
        if present.sum() > 1:
            out[present] = auc[present]
        return out
    return float(auc[1]) if present.sum() > 1 else np.nan


def compute_targetwise_metrics(epoch_results, target_name=None):
    """metrics.py:7-51."""
    pick = (lambda k: epoch_results[k]) if target_name is None else (lambda k: epoch_results[k][target_name])
    running_loss, confidences = pick("running_loss"), np.array(pick("confidences"))
    predictions, ground_truth = pick("predictions"), pick("ground_truth")
    n_classes = confidences.shape[1]
    gt_classes = np.unique(ground_truth)
    gt_n_classes = len(gt_classes)
    if gt_n_classes < n_classes:
        warnings.warn("\nNumber of classes in ground truth is less than number of classes in predicted confidences. \n"
                      "Some of ROC AUC metric values will be NaN\n")
    cm = epoch_results.get("confusion")
    if cm is not None and target_name is not None:
        cm = cm[target_name]
    if cm is not None:
        epoch_acc = balanced_accuracy_from_confusion(cm)
    else:
        epoch_acc = balanced_accuracy_score(ground_truth, predictions)
    auc_counts = epoch_results.get("roc_auc_counts")
    if auc_counts is not None and target_name is not None:
        auc_counts = auc_counts[target_name]
    if auc_counts is not None:
        epoch_roc_auc = roc_auc_from_counts(auc_counts, n_classes)
    elif n_classes > 2:
        epoch_roc_auc = np.full(n_classes, np.nan)
        if gt_n_classes > 1:
            ground_truth_bin = label_binarize(ground_truth, classes=range(n_classes))
            for gt_class in gt_classes:
                epoch_roc_auc[gt_class] = roc_auc_score(ground_truth_bin[:, gt_class], confidences[:, gt_class])
    else:
        epoch_roc_auc = np.nan
        if gt_n_classes > 1:
            epoch_roc_auc = roc_auc_score(ground_truth, confidences[:, 1])
    return {"epoch_acc": epoch_acc, "epoch_roc_auc": epoch_roc_auc, "epoch_loss": np.mean(running_loss)}


def compute_metrics(cfg, epoch_results: dict):
    """metrics.py:54-70."""
    if cfg.task == "single":
        metrics = compute_targetwise_metrics(epoch_results)
        metrics["loss"] = epoch_results["running_loss"]
        return metrics
    elif cfg.task == "multi":
        target_names = cfg.target_names
        metrics = {t: compute_targetwise_metrics(epoch_results, t) for t in target_names}
        metrics["loss"] = epoch_results["running_loss"]["loss"]
        metrics["epoch_acc"] = np.mean([metrics[t]["epoch_acc"] for t in target_names])
        return metrics
    raise ValueError(f"Unknown task type {cfg.task} for metric computation")
