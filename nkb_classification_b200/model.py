"""Classifier modules with the reference's structure, attribute names and
``state_dict`` keys (nkb_classification/model.py): a stock backbone
(``emb_model``, out of scope for this path) and per-task ``Dropout -> Linear``
heads (``classifier``).

The modules stay plain ``nn.Module``s -- ``torch.jit.script(model)`` every
epoch (train.py:66) and checkpoints keep working -- and ``forward`` keeps the
reference semantics.  The fused path is reached from the engine:
``emb = model.emb_model(x)`` followed by :func:`fused_heads_loss`, which reads the
heads' weights through :class:`HeadPack` (one contiguous [sum C_t, D] buffer the
per-head ``nn.Linear`` parameters are views of, so no per-step concatenation).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Union

import torch
from torch import nn

try:  # the reference's backbones; absent in this image
    import timm  # type: ignore
except Exception:  # pragma: no cover
    timm = None
try:
    import unicom  # type: ignore
except Exception:  # pragma: no cover
    unicom = None


def _torchvision_backbone(name: str, pretrained: bool):
    """Stand-in honouring ``timm.create_model(name, num_classes=0)``: [B,3,H,W] -> [B,num_features]."""
    import torchvision

    fn = {
        "resnet18": "resnet18", "resnet34": "resnet34", "resnet50": "resnet50", "efficientnet_b0": "efficientnet_b0",
        "convnext_tiny": "convnext_tiny", "mobilenetv3_large_100": "mobilenet_v3_large",
        "vit_base_patch16_224": "vit_b_16",
    }.get(name)
    if fn is None:
        raise NotImplementedError(f"backbone {name!r}: timm is not installed and no torchvision stand-in is mapped")
    if pretrained:
        raise NotImplementedError("pretrained weights need timm / network access; use pretrained=False")
    m = getattr(torchvision.models, fn)(weights=None)
    if hasattr(m, "fc"):
        nf = m.fc.in_features
        m.fc = nn.Identity()
    elif hasattr(m, "heads"):
        nf = m.heads.head.in_features
        m.heads = nn.Identity()
    else:
        last = m.classifier[-1]
        nf = last.in_features
        m.classifier[-1] = nn.Identity()
    m.num_features = nf
    return m, nf


def get_emb_model(cfg_model: dict):
    """model.py:74-85.  ``cfg_model['model']`` may also be an nn.Module with ``num_features`` (tests, custom nets)."""
    name = cfg_model["model"]
    if isinstance(name, nn.Module):
        return name, int(name.num_features)
    if name.lower().startswith("unicom"):
        if unicom is None:
            raise NotImplementedError("unicom is not installed")
        emb_model, _ = unicom.load(name.split()[1])
        return emb_model, emb_model.feature[-2].out_features
    if timm is not None:
        emb_model = timm.create_model(name, pretrained=cfg_model["pretrained"], num_classes=0)
        return emb_model, emb_model.num_features
    return _torchvision_backbone(name, cfg_model.get("pretrained", False))


def _init_linear(classifier: nn.Module, strategy: str):
    """model.py:45-57."""
    for param in classifier.parameters():
        if param.ndim >= 2:
            if strategy == "kaiming_normal_":
                nn.init.kaiming_normal_(param, nonlinearity="relu")
            elif strategy == "kaiming_uniform_":
                nn.init.kaiming_uniform_(param, nonlinearity="relu")
            elif strategy == "xavier_normal_":
                nn.init.xavier_normal_(param)
            elif strategy == "xavier_uniform_":
                nn.init.xavier_uniform_(param)
        else:
            nn.init.zeros_(param)


class _ClassifierBase(nn.Module):
    def set_backbone_state(self, state: str):
        for param in self.emb_model.parameters():
            if state == "freeze":
                param.requires_grad = False
            elif state == "unfreeze":
                param.requires_grad = True

    @staticmethod
    def set_dropout(model: nn.Module, drop_rate: float = 0.2) -> None:
        for child in model.children():
            if isinstance(child, torch.nn.Dropout):
                child.p = drop_rate
            _ClassifierBase.set_dropout(child, drop_rate=drop_rate)

    get_emb_model = staticmethod(get_emb_model)


class SingletaskClassifier(_ClassifierBase):
    """model.py:17-85.  state_dict keys: emb_model.*, classifier.1.{weight,bias}."""

    def __init__(self, cfg_model: dict, classes: list):
        super().__init__()
        self.emb_model, self.emb_size = get_emb_model(cfg_model)
        self.set_dropout(self.emb_model, cfg_model["backbone_dropout"])
        self.classifier = nn.Sequential(
            nn.Dropout(cfg_model["classifier_dropout"]),
            nn.Linear(self.emb_size, len(classes)),
        )
        _init_linear(self.classifier, cfg_model["classifier_initialization"])

    def forward(self, x: torch.Tensor):
        emb = self.emb_model(x)
        return self.classifier(emb)

    def head_linears(self) -> List[nn.Linear]:
        return [self.classifier[1]]

    def head_names(self) -> Optional[List[str]]:
        return None


class MultitaskClassifier(_ClassifierBase):
    """model.py:88-159.  state_dict keys: emb_model.*, classifier.<task>.1.{weight,bias}; forward returns a dict in
    ModuleDict order (= order of the ``classes`` dict)."""

    def __init__(self, cfg_model: dict, classes: dict):
        super().__init__()
        self.emb_model, self.emb_size = get_emb_model(cfg_model)
        self.set_dropout(self.emb_model, cfg_model["backbone_dropout"])
        self.classifier = nn.ModuleDict()
        for target_name in classes:
            self.classifier[target_name] = nn.Sequential(
                nn.Dropout(cfg_model["classifier_dropout"]),
                nn.Linear(self.emb_size, len(classes[target_name])),
            )
        for classifier in self.classifier.values():
            _init_linear(classifier, cfg_model["classifier_initialization"])

    def forward(self, x: torch.Tensor):
        emb = self.emb_model(x)
        return {task_name: classifier(emb) for task_name, classifier in self.classifier.items()}

    def head_linears(self) -> List[nn.Linear]:
        return [seq[1] for seq in self.classifier.values()]

    def head_names(self) -> Optional[List[str]]:
        return list(self.classifier.keys())


def get_model(cfg_model, classes, device="cpu", compile: bool = False):
    """model.py:162-177."""
    if cfg_model.get("scripted", False):
        model = torch.jit.load(cfg_model["checkpoint"], map_location="cpu")
    else:
        if cfg_model["task"] == "single":
            model = SingletaskClassifier(cfg_model, classes)
        elif cfg_model["task"] == "multi":
            model = MultitaskClassifier(cfg_model, classes)
        else:
            raise ValueError(f"Unknown task type {cfg_model['task']}")
        chkpt = cfg_model.get("checkpoint", None)
        if chkpt is not None:
            model.load_state_dict(torch.load(chkpt, map_location="cpu"))
    model.to(device)
    if compile:
        model = torch.compile(model, dynamic=True)
    return model


class HeadPack:
    """All heads of a classifier as one contiguous W_cat [sum C_t, D] / b_cat [sum C_t] pair.

    The per-head ``nn.Linear`` parameters are re-pointed at views of the packed buffers (their names, shapes and
    values are unchanged, so ``state_dict`` / optimizers / TorchScript see nothing different); in-place optimizer
    updates of the per-head parameters therefore update the packed buffers K2 reads.
    """

    def __init__(self, model: nn.Module):
        linears = model.head_linears()
        self.names = model.head_names()
        self.linears = linears
        dev = linears[0].weight.device
        if dev.type != "cuda":
            raise RuntimeError("HeadPack needs the model on a CUDA device (no CPU path)")
        self.D = linears[0].in_features
        self.classes = [l.out_features for l in linears]
        self.seg = [0]
        for c in self.classes:
            self.seg.append(self.seg[-1] + c)
        NC = self.seg[-1]
        self.W_cat = torch.empty((NC, self.D), dtype=torch.float32, device=dev)
        self.b_cat = torch.empty((NC,), dtype=torch.float32, device=dev)
        self.dropout_p = [float(p) for p in self._dropouts(model)]
        self.repack()

    def repack(self) -> None:
        """(Re-)point every head's weight / bias at its slice of the packed buffers, keeping the values."""
        with torch.no_grad():
            for t, l in enumerate(self.linears):
                a, b = self.seg[t], self.seg[t + 1]
                if l.weight.data_ptr() != self.W_cat[a:b].data_ptr():
                    self.W_cat[a:b].copy_(l.weight)
                    l.weight.data = self.W_cat[a:b]
                if l.bias.data_ptr() != self.b_cat[a:b].data_ptr():
                    self.b_cat[a:b].copy_(l.bias)
                    l.bias.data = self.b_cat[a:b]

    @staticmethod
    def _dropouts(model):
        if isinstance(model.classifier, nn.ModuleDict):
            return [seq[0].p for seq in model.classifier.values()]
        return [model.classifier[0].p]

    def in_sync(self) -> bool:
        """False if someone replaced a head's storage (e.g. load_state_dict(assign=True)); then re-pack."""
        return all(l.weight.data_ptr() == self.W_cat[a:b].data_ptr() and l.bias.data_ptr() == self.b_cat[a:b].data_ptr()
                   for l, a, b in zip(self.linears, self.seg[:-1], self.seg[1:]))

    def params(self) -> List[torch.nn.Parameter]:
        out = []
        for l in self.linears:
            out += [l.weight, l.bias]
        return out
