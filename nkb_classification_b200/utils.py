"""Optimizer / scheduler / class-map helpers with the reference's signatures
(nkb_classification/utils.py).  Stock torch on purpose: they are outside the hot
path (component 8) and only have to keep the config surface."""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
from torch.optim import SGD, Adam, NAdam, RAdam, SparseAdam, lr_scheduler


def get_optimizer(model, cfg_optimizer):
    """utils.py:10-42: two parameter groups, ``emb_model`` and ``classifier``."""
    base_lr = cfg_optimizer.get("lr", 0.001)
    backbone_lr = cfg_optimizer.get("backbone_lr", base_lr)
    classifier_lr = cfg_optimizer.get("classifier_lr", base_lr)
    base_wd = cfg_optimizer.get("weight_decay", 0.0)
    parameters = [
        {"params": model.emb_model.parameters(), "lr": backbone_lr,
         "weight_decay": cfg_optimizer.get("backbone_weight_decay", base_wd)},
        {"params": model.classifier.parameters(), "lr": classifier_lr,
         "weight_decay": cfg_optimizer.get("classifier_weight_decay", base_wd)},
    ]
    kind = cfg_optimizer["type"].lower()
    if kind == "adam":
        return Adam(parameters)
    if kind == "radam":
        return RAdam(parameters)
    if kind == "nadam":
        return NAdam(parameters, decoupled_weight_decay=True)
    if kind == "sparse_adam":
        return SparseAdam(parameters)
    if kind == "sgd":
        return SGD(parameters)
    raise NotImplementedError(f'Unknown optimizer in config: {cfg_optimizer["type"]}')


def get_scheduler(opt, lr_policy):
    """utils.py:45-61."""
    if len(lr_policy) == 0:
        return None
    kind = lr_policy["type"]
    if kind == "step":
        return lr_scheduler.StepLR(opt, step_size=lr_policy["step_size"], gamma=lr_policy["gamma"])
    if kind == "multistep":
        return lr_scheduler.MultiStepLR(opt, milestones=lr_policy["steps"], gamma=lr_policy["gamma"])
    if kind == "cosine":
        return lr_scheduler.CosineAnnealingLR(opt, T_max=lr_policy["n_epochs"])
    raise NotImplementedError("Learning rate policy {} not implemented.".format(kind))


def save_classes(classes, save_path):
    if isinstance(classes, (list, dict)):
        with open(save_path, "w") as f:
            json.dump(classes, f)
    else:
        raise NotImplementedError(f"unknown classes config type {type(classes)}")


def load_classes(classes):
    if isinstance(classes, (list, dict)):
        return classes
    if isinstance(classes, (str, Path)):
        with open(classes, "r") as f:
            return json.load(f)
    raise NotImplementedError(f"unknown classes config type {type(classes)}")


def get_classes_configs(classes):
    """utils.py:82-98: (class_to_idx, idx_to_class) for a list (single task) or dict of lists (multi task)."""
    if isinstance(classes, list):
        class_to_idx = {cls: idx for idx, cls in enumerate(classes)}
        return class_to_idx, {idx: cls for cls, idx in class_to_idx.items()}
    if isinstance(classes, dict):
        class_to_idx = {t: {cls: idx for idx, cls in enumerate(classes[t])} for t in classes}
        idx_to_class = {t: {idx: cls for cls, idx in class_to_idx[t].items()} for t in classes}
        return class_to_idx, idx_to_class
    raise NotImplementedError(f"unknown classes config type {type(classes)}")


def read_py_config(path):
    path = Path(path)
    sys.path.append(str(path.parent))
    return f"import {path.stem} as cfg"


def convert_dict_types_recursive(_dict):
    for key in _dict:
        if isinstance(_dict[key], dict):
            _dict[key] = convert_dict_types_recursive(_dict[key])
        elif isinstance(_dict[key], np.ndarray):
            _dict[key] = list(_dict[key])
    return _dict
