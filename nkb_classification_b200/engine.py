"""train_epoch / val_epoch with the reference's signatures (nkb_classification/engine.py:20-117).

Per batch the loop does what the reference does, through the fused kernels:

    reference                                   here
    img.to(device) (fp32 NCHW, blocking)        loader already yields device tensors produced by K1
    preds = model(img)                          emb = model.emb_model(img)   (stock backbone, out of scope)
    loss = criterion(preds, target)             out = FusedHeads(emb, target)  -> K2 (+K3, +K4)
    scaler.scale(loss).backward(); step         identical (autograd reaches emb and the head parameters)
    epoch_logger.log_iter(...) (3T+T+1 syncs)   epoch_logger.log_fused(...)  (no sync)

A model that is not one of this package's classifiers (e.g. a TorchScript
module from ``get_model(scripted=True)``) takes the reference's own sequence
with this package's criterion (K2 loss kernel on the logits).
"""
from __future__ import annotations

from collections import defaultdict
from typing import Optional

import torch
from tqdm import tqdm

from .heads import FusedHeads
from .model import MultitaskClassifier, SingletaskClassifier

AUTOCAST_DTYPE = torch.bfloat16  # the reference autocasts to fp16 + GradScaler; bf16 needs no scaler on B200


class TrainPbar(tqdm):
    """engine.py:6-17 of the reference, without its per-step ``loss.item()``: the postfix is refreshed from the device
    loss every ``cfg.loss_display_every`` steps (default 50) -- one small D2H then, none otherwise."""

    def __init__(self, train_loader, leave, desc, cfg):
        super().__init__(train_loader, leave=leave, desc=desc, disable=getattr(cfg, "disable_tqdm", False))
        self.cfg = cfg
        self._every = max(1, int(getattr(cfg, "loss_display_every", 50)))
        self._seen = 0

    def update_loss(self, loss):
        """``loss``: what the reference passes (a tensor, or the criterion's dict for task='multi')."""
        self._seen += 1
        if self.disable or self._seen % self._every:
            return
        if isinstance(loss, dict):
            if getattr(self.cfg, "show_full_current_loss_in_terminal", False):
                vals = torch.stack([loss[k].detach().float().reshape(()) for k in loss]).tolist()     # one D2H
                self.set_postfix_str(", ".join(f"loss {k}: {v:.4f}" for k, v in zip(loss, vals)))
            else:
                self.set_postfix_str(f"Loss: {loss['loss'].item():.4f}")
        else:
            self.set_postfix_str(f"Loss: {loss.item():.4f}")


def sync_backbone_grads(model, comm) -> int:
    """Sharded training (cfg.communicator, world > 1): the fused heads step all-reduces the head gradients itself, but
    the embedding gradient it hands back -- and hence every backbone gradient -- covers the LOCAL rows only (already
    divided by the GLOBAL denominators).  Sum them over the ranks before the optimizer steps, or an unfrozen backbone
    silently diverges across ranks.  SUM, not mean: the division by the global batch has happened.  One flat
    all-reduce over torch.distributed (the backbone is stock PyTorch; wrap it in DDP instead and this finds nothing
    to do because DDP has averaged already -- do not combine the two).  Returns the number of tensors reduced."""
    if comm is None or comm.world <= 1:
        return 0
    grads = [p.grad for p in model.emb_model.parameters() if p.requires_grad and p.grad is not None]
    if not grads:
        return 0
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("sharded training with a trainable backbone needs torch.distributed (it carries the backbone "
                           "gradient all-reduce); freeze the backbone or initialise the process group")
    flat = torch._utils._flatten_dense_tensors(grads)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    for g, r in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
        g.copy_(r)
    return len(grads)


def _criterion_cfg(criterion, cfg):
    """The loss configuration: ``cfg.criterion`` (the reference's config attribute) or, failing that, the
    criterion object built by get_loss."""
    c = getattr(cfg, "criterion", None)
    if isinstance(c, dict):
        return c
    inner = getattr(criterion, "criterion", criterion)
    kind = "FocalLoss" if type(inner).__name__ == "FocalLoss" else "CrossEntropyLoss"
    out = {"type": kind}
    if kind == "FocalLoss":
        out["gamma"] = inner.gamma
    w = getattr(inner, "_class_weight", None)
    if w is not None:
        out["alpha" if kind == "FocalLoss" else "weight"] = w.tolist()
    return out


def _fused_for(model, criterion, cfg) -> Optional[FusedHeads]:
    base = getattr(model, "_orig_mod", model)
    if not isinstance(base, (SingletaskClassifier, MultitaskClassifier)):
        return None
    fused = getattr(base, "_nkbk_fused", None)
    if fused is None:
        fused = FusedHeads(base, _criterion_cfg(criterion, cfg), comm=getattr(cfg, "communicator", None))
        base._nkbk_fused = fused
    return fused


def _autocast(cfg):
    return torch.autocast(device_type="cuda", dtype=getattr(cfg, "autocast_dtype", AUTOCAST_DTYPE),
                          enabled=cfg.enable_mixed_presicion)


def train_epoch(model, train_loader, optimizer, scheduler, scaler, criterion, device, cfg, epoch_logger):
    model.train()
    epoch_logger.init_iter_logs()
    fused = _fused_for(model, criterion, cfg)
    base_model = getattr(model, "_orig_mod", model)
    if fused is not None:
        fused.reset_confusion()
        epoch_logger.fused = fused
    if cfg.log_gradients:
        metrics_grad_log = defaultdict(list)
    pbar = TrainPbar(train_loader, leave=False, desc="Training", cfg=cfg)

    for img, target in pbar:
        img = img.to(device, non_blocking=True)
        optimizer.zero_grad()
        if fused is not None:
            with _autocast(cfg):
                emb = model.emb_model(img)
            out = fused(emb, target, train=True)
            total = out.loss[-1] if fused.names is not None else out.loss[0]
            scaler.scale(total).backward()
            sync_backbone_grads(base_model, getattr(cfg, "communicator", None))
            scaler.step(optimizer)
            scaler.update()
            epoch_logger.log_fused(out, out.labels)
            pbar.update_loss(out.loss_like_reference())
        else:
            with _autocast(cfg):
                preds = model(img)
                if isinstance(target, torch.Tensor):
                    target = target.to(device)
                loss = criterion(preds, target)
            scaler.scale(loss["loss"] if isinstance(loss, dict) else loss).backward()
            scaler.step(optimizer)
            scaler.update()
            epoch_logger.log_iter(preds, target, loss)
            pbar.update_loss(loss)

        if cfg.log_gradients:
            total_grad = 0
            for tag, value in model.named_parameters():
                assert tag != "Total"
                if value.grad is not None:
                    grad = value.grad.norm()
                    metrics_grad_log[f"Gradients/{tag}"].append(grad)
                    total_grad += grad
            metrics_grad_log["Gradients/Total"].append(total_grad)
        epoch_logger.log_images_if_needed(img)

    if scheduler is not None:
        scheduler.step()
    results = epoch_logger.get_epoch_results()
    if cfg.log_gradients:
        results["metrics_grad_log"] = metrics_grad_log
    return results


@torch.no_grad()
def val_epoch(model, val_loader, criterion, device, cfg, epoch_logger):
    model.eval()
    epoch_logger.init_iter_logs()
    fused = _fused_for(model, criterion, cfg)
    if fused is not None:
        fused.reset_confusion()
        epoch_logger.fused = fused
    for img, target in tqdm(val_loader, leave=False, desc="Evaluating", disable=getattr(cfg, "disable_tqdm", False)):
        img = img.to(device, non_blocking=True)
        if fused is not None:
            with _autocast(cfg):
                emb = model.emb_model(img)
            out = fused(emb, target, train=False)
            epoch_logger.log_fused(out, out.labels)
        else:
            with _autocast(cfg):
                preds = model(img)
                if isinstance(target, torch.Tensor):
                    target = target.to(device)
                loss = criterion(preds, target)
            epoch_logger.log_iter(preds, target, loss)
        epoch_logger.log_images_if_needed(img)
    return epoch_logger.get_epoch_results()
