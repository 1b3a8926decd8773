"""Loss objects with the reference's names, constructor arguments and call
semantics (nkb_classification/losses.py), computed by libnkbk's fused
softmax + CE/focal + logit-gradient kernel instead of ~10 ATen launches per task.

    get_loss(cfg_loss, device)   losses.py:154-176   same config keys / errors
    FocalLoss(alpha, gamma, reduction, ignore_index)  losses.py:10-94
    CrossEntropyLoss(weight)     nn.CrossEntropyLoss  losses.py:155-159
    MultitaskCriterion           losses.py:97-151     dict in -> dict out, "loss" = unweighted sum

``criterion(pred, true)`` returns tensors that carry autograd history back to
the logits (a custom Function hands out the kernel's dlogits), so
``scaler.scale(loss).backward()`` in engine.py works unchanged.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Optional

import torch
from torch import Tensor, nn

from . import ops
from ._lib import LOSS_CE, LOSS_FOCAL

DEFAULT_FOCAL_GAMMA = 2.0


class _FusedLoss(torch.autograd.Function):
    """loss vector [T+1] from logits; backward = grad_out[-1]-style contraction with the kernel's dlogits."""

    @staticmethod
    def forward(ctx, logits, labels, seg, loss_kind, gamma, class_weight, ignore_index):
        need_grad = logits.requires_grad
        z = logits.detach()
        if z.dtype not in (torch.float32, torch.bfloat16):
            z = z.float()  # fp16 autocast logits: widen (exact) -- the kernel computes in fp32 either way
        z = z.contiguous()
        loss, dl, _ = ops.loss_fwd_bwd(z, seg, labels, loss_kind, gamma, class_weight, ignore_index,
                                       want_probs=False, want_grad=need_grad)
        ctx.seg = seg
        ctx.in_dtype = logits.dtype
        ctx.save_for_backward(dl)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        (dl,) = ctx.saved_tensors
        if dl is None:
            return (None,) * 7
        seg = ctx.seg
        T = len(seg) - 1
        # d(out[t])/dz = dl restricted to task t's columns; d(out[T])/dz = dl (the total is the plain sum)
        per_col = torch.repeat_interleave(grad_loss[:T], torch.tensor([b - a for a, b in zip(seg[:-1], seg[1:])],
                                                                      device=grad_loss.device)) + grad_loss[T]
        return (dl * per_col).to(ctx.in_dtype), None, None, None, None, None, None


class _FusedRowLoss(torch.autograd.Function):
    """Per-row loss terms [B] (single task) from logits; backward scales the kernel's unnormalised dlogits row-wise."""

    @staticmethod
    def forward(ctx, logits, labels, loss_kind, gamma, class_weight, ignore_index):
        need_grad = logits.requires_grad
        z = logits.detach()
        if z.dtype not in (torch.float32, torch.bfloat16):
            z = z.float()
        z = z.contiguous()
        rows, dl = ops.loss_rows(z, [0, z.shape[1]], labels, loss_kind, gamma, class_weight, ignore_index,
                                 want_grad=need_grad)
        ctx.in_dtype = logits.dtype
        ctx.save_for_backward(dl)
        return rows[:, 0]

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        if dl is None:
            return (None,) * 6
        return (dl * g[:, None]).to(ctx.in_dtype), None, None, None, None, None


def _as_labels_2d(y: Tensor, device) -> Tensor:
    y = y.to(device=device, dtype=torch.int64)
    return y.reshape(-1, 1).contiguous()


class _SingleTaskLoss(nn.Module):
    kind = LOSS_CE

    def __init__(self, class_weight: Optional[Tensor], gamma: float, ignore_index: int):
        super().__init__()
        self.register_buffer("_class_weight", None if class_weight is None else class_weight.detach().float().clone())
        self.gamma = gamma
        self.ignore_index = ignore_index

    def forward(self, x: Tensor, y: Tensor) -> Tensor:
        if not x.is_cuda:
            raise RuntimeError("nkb_classification_b200 losses run on CUDA tensors only (no CPU fallback)")
        if x.ndim > 2:  # (N, C, d1, ..) -> (N*d1*.., C)  losses.py:60-65
            c = x.shape[1]
            x = x.permute(0, *range(2, x.ndim), 1).reshape(-1, c)
            y = y.reshape(-1)
        cw = self._class_weight
        if cw is not None and cw.device != x.device:
            cw = cw.to(x.device)
        reduction = getattr(self, "reduction", "mean")
        if reduction != "mean":          # losses.py:87-94 of the reference: "sum" / "none" over the rows that are kept
            y2 = _as_labels_2d(y, x.device)
            rows = _FusedRowLoss.apply(x, y2, self.kind, float(self.gamma), cw, int(self.ignore_index))
            if reduction == "sum":
                return rows.sum()
            keep = y2[:, 0] != self.ignore_index       # (boolean indexing synchronises, as the reference's x[mask] does)
            if not bool(keep.any()):
                return torch.zeros((), device=x.device)   # the reference returns a CPU tensor(0.) here
            return rows[keep]
        out = _FusedLoss.apply(x, _as_labels_2d(y, x.device), [0, x.shape[1]], self.kind, float(self.gamma), cw,
                               int(self.ignore_index))
        return out[0]


class FocalLoss(_SingleTaskLoss):
    """Same constructor and reductions as the reference's FocalLoss (losses.py:23-94): "mean" (what get_loss builds) is
    one fused launch; "sum" / "none" take the per-row terms of nkbk_loss_rows."""
    kind = LOSS_FOCAL

    def __init__(self, alpha: Optional[Tensor] = None, gamma: float = DEFAULT_FOCAL_GAMMA, reduction: str = "mean",
                 ignore_index: int = -100):
        if reduction not in ("mean", "sum", "none"):
            raise ValueError('Reduction must be one of: "mean", "sum", "none".')
        super().__init__(alpha, gamma, ignore_index)
        self.alpha = alpha
        self.reduction = reduction

    def __repr__(self):
        return (f"FocalLoss(alpha={self.alpha!r}, gamma={self.gamma!r}, ignore_index={self.ignore_index!r}, "
                f"reduction={self.reduction!r})")


class CrossEntropyLoss(_SingleTaskLoss):
    """nn.CrossEntropyLoss(weight) semantics: sum_i w[y_i] * nll_i / sum_i w[y_i]."""
    kind = LOSS_CE

    def __init__(self, weight: Optional[Tensor] = None, ignore_index: int = -100):
        super().__init__(weight, 0.0, ignore_index)
        self.weight = weight


class MultitaskCriterion:
    """losses.py:97-151: per-task criterion, returns {task: loss_t, ..., "loss": sum_t loss_t}.
    All tasks go through ONE kernel launch (logits concatenated along the class axis)."""

    def __init__(self, criterion: _SingleTaskLoss, device):
        self.criterion = criterion
        self.device = device
        self.criterion.to(device)

    def __call__(self, pred: dict, true: dict):
        assert pred.keys() == true.keys()
        names = list(pred.keys())
        zs = [pred[n] for n in names]
        dev = zs[0].device
        if not zs[0].is_cuda:
            raise RuntimeError("nkb_classification_b200 losses run on CUDA tensors only (no CPU fallback)")
        seg = [0]
        for z in zs:
            seg.append(seg[-1] + z.shape[1])
        z_cat = torch.cat(zs, dim=1)
        labels = torch.stack([true[n].to(device=dev, dtype=torch.int64).reshape(-1) for n in names], dim=1).contiguous()
        c = self.criterion
        cw = c._class_weight
        if cw is not None:
            # the reference applies the same weight vector to every task (losses.py:155-169)
            for z in zs:
                if z.shape[1] != cw.numel():
                    raise RuntimeError("weight tensor should be defined either for all classes or no classes")
            cw = cw.to(dev).repeat(len(zs))
        out = _FusedLoss.apply(z_cat, labels, seg, c.kind, float(c.gamma), cw, int(c.ignore_index))
        separate_loss = defaultdict()
        for t, n in enumerate(names):
            separate_loss[n] = out[t]
        separate_loss["loss"] = out[len(names)]
        return separate_loss


def get_loss(cfg_loss, device):
    """Same keys and error as the reference's get_loss (losses.py:154-176)."""
    if cfg_loss["type"] == "CrossEntropyLoss":
        weight = None
        if "weight" in cfg_loss:
            weight = torch.tensor(cfg_loss["weight"], dtype=torch.float)
        loss = CrossEntropyLoss(weight).to(device)
    elif cfg_loss["type"] == "FocalLoss":
        alpha = None
        if "alpha" in cfg_loss:
            alpha = torch.tensor(cfg_loss["alpha"], dtype=torch.float)
        gamma = DEFAULT_FOCAL_GAMMA
        if "gamma" in cfg_loss:
            gamma = cfg_loss["gamma"]
        loss = FocalLoss(alpha, gamma).to(device)
    else:
        raise NotImplementedError(f'Unknown loss type in config: {cfg_loss["type"]}')
    if cfg_loss["task"] == "multi":
        return MultitaskCriterion(loss, device)
    return loss
