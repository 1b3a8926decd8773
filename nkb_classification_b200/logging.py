"""Per-iteration accumulation with the reference's ``BaseLogger`` interface
(nkb_classification/logging.py:218-294) and none of its per-step syncs.

The reference does, per task and per step, ``true.cpu().tolist()``,
``pred.softmax(..).cpu().tolist()``, ``pred.argmax(..).cpu().tolist()`` and
``loss.item()`` -- (3T + T + 1) device->host synchronisations.  Here every step
only appends device tensors (K2's fp32 probabilities, K3's predictions, the
labels, the loss vector); ``get_epoch_results()`` performs ONE device->host copy
per quantity per epoch and rebuilds exactly the reference's dict of lists, plus
``"confusion"`` (K3's integer matrices) and ``"roc_auc_counts"`` (K5's exact
ROC-AUC pair counts, computed on the device-resident epoch) for ``compute_metrics``.

(The reference's BaseLogger crashes for task="multi" -- it reads an attribute
that is never set, logging.py:243; target names here come from ``classes``.)
"""
from __future__ import annotations

from collections import defaultdict
from typing import List, Optional

import numpy as np
import torch


def _to_host(*tensors):
    """Device tensors -> numpy arrays through pinned memory: the copies are queued back to back and waited for once."""
    outs, on_dev = [], False
    for t in tensors:
        if t.is_cuda:
            pin = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            pin.copy_(t, non_blocking=True)
            outs.append(pin)
            on_dev = True
        else:
            outs.append(t)
    if on_dev:
        torch.cuda.current_stream().synchronize()
    return [o.numpy() for o in outs]


def gather_rows(t: torch.Tensor) -> torch.Tensor:
    """Rows of every rank, in rank order (ranks may hold different numbers of rows: uneven frame shards).  Collective over
    torch.distributed's default group; the row counts travel first, so this synchronises -- epoch end only."""
    import torch.distributed as dist
    world = dist.get_world_size()
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n)
    ns = [int(x.item()) for x in ns]
    pad = torch.zeros((max(ns),) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return torch.cat([o[:k] for o, k in zip(outs, ns)])


class BaseLogger:
    def __init__(self, cfg, classes):
        assert cfg.task in ("single", "multi")
        self.cfg = cfg
        self.task = cfg.task
        self.classes = classes
        self.target_names = None if self.task == "single" else sorted(classes)
        self.fused = None  # set by the engine: the FusedHeads whose confusion counts belong to this epoch
        self.device_roc_auc = True   # K5: ROC-AUC pair counts computed on the device at epoch end
        self.init_iter_logs()

    def init_iter_logs(self):
        self.epoch_images_example = None
        self._img_event = None
        self._gt, self._conf, self._pred, self._loss = [], [], [], []
        self._seg, self._names = None, None

    def _sharded(self) -> bool:
        """True when the epoch was computed by several ranks (the engine's fused heads carry a communicator with
        world > 1), torch.distributed can carry the gather and ``cfg.gather_epoch_results`` (default True) allows it."""
        comm = None if self.fused is None else self.fused.state.get("comm")
        if comm is None or getattr(comm, "world", 1) <= 1 or not getattr(self.cfg, "gather_epoch_results", True):
            return False
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    # ---- fused path: everything stays on the device ----
    def log_fused(self, out, labels: torch.Tensor):
        """``out``: heads.HeadsOutput of this step; ``labels``: int64 [B, T] on the device."""
        self._seg, self._names = out.seg, out.names
        self._gt.append(labels)
        self._conf.append(out.probs.clone())      # the probabilities live in the step's reusable buffers: keep a copy
        self._pred.append(out.pred)               # predictions and the loss vector are this step's own tensors
        self._loss.append(out.loss.detach())

    # ---- reference-compatible entry (pred / true / loss as the reference passes them) ----
    def log_iter(self, pred, true, loss):
        assert type(pred) == type(true)
        if isinstance(pred, dict):
            assert pred.keys() == true.keys()
            names = list(pred.keys())
            seg = [0]
            for n in names:
                seg.append(seg[-1] + pred[n].shape[1])
            dev = pred[names[0]].device
            self._seg, self._names = seg, names
            self._gt.append(torch.stack([true[n].to(dev).reshape(-1) for n in names], 1))
            self._conf.append(torch.cat([pred[n].detach().softmax(-1, dtype=torch.float32) for n in names], 1))
            self._pred.append(torch.stack([pred[n].detach().argmax(-1) for n in names], 1).to(torch.int32))
            self._loss.append(torch.stack([loss[n].detach().float() for n in names] + [loss["loss"].detach().float()]))
        else:
            self._seg, self._names = [0, pred.shape[1]], None
            self._gt.append(true.to(pred.device).reshape(-1, 1))
            self._conf.append(pred.detach().softmax(-1, dtype=torch.float32))
            self._pred.append(pred.detach().argmax(-1).reshape(-1, 1).to(torch.int32))
            l = loss.detach().float().reshape(1)
            self._loss.append(torch.cat([l, l]))

    def log_images_if_needed(self, images):
        """logging.py:283-285 of the reference keeps the WHOLE first batch on the host for the image grid of
        ``log_images`` (8 per row).  With the batches this path is built for (thousands of crops) that is a multi-GB
        pageable D2H per epoch, so only the first ``cfg.example_images`` (default 64 = 8 grid rows) are kept;
        ``cfg.example_images = None`` restores the reference's whole-batch copy."""
        if self.epoch_images_example is None:
            k = getattr(self.cfg, "example_images", 64)
            src = images if k is None else images[: int(k)]
            if src.is_cuda:
                # asynchronous copy into fresh pinned memory (a pageable `.to("cpu")` of 38 MB stalls the loop for ~17 ms
                # at the top of every epoch); get_epoch_results waits for it
                pin = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
                pin.copy_(src, non_blocking=True)
                self._img_event = torch.cuda.Event()
                self._img_event.record()
                self.epoch_images_example = pin
            else:
                self.epoch_images_example = src.to("cpu")

    def get_epoch_results(self):
        """One D2H per quantity, then the reference's structure (logging.py:287-294)."""
        if self._img_event is not None:
            self._img_event.synchronize()
        if not self._gt:
            empty = [] if self.task == "single" else defaultdict(list)
            return {"running_loss": empty, "confidences": empty, "predictions": empty, "ground_truth": empty,
                    "images": self.epoch_images_example}
        gt_d, conf_d, pred_d = torch.cat(self._gt), torch.cat(self._conf), torch.cat(self._pred)
        if self._sharded():
            # a sharded run (cfg.communicator, world > 1): the fused step's confusion counts and losses are already global,
            # so make the per-sample quantities global too -- every rank gathers every rank's rows (rank order), and the
            # ROC-AUC pair counts below are then taken over the global epoch like the confusion matrix beside them
            gt_d, conf_d, pred_d = gather_rows(gt_d), gather_rows(conf_d), gather_rows(pred_d)
        auc_counts = None
        if conf_d.is_cuda and self.device_roc_auc:
            # K5: exact ROC-AUC pair counts from the device-resident epoch, one [NC, 3] int64 D2H
            from . import ops
            auc_counts = ops.roc_auc_counts(conf_d.contiguous(), self._seg, gt_d.to(torch.int64).contiguous()).cpu().numpy()
        gt, conf, pred, loss = _to_host(gt_d, conf_d, pred_d, torch.stack(self._loss))
        seg, names = self._seg, self._names
        res = {"images": self.epoch_images_example}
        # the reference hands out Python lists (logging.py:268-281); building them costs ~0.4 us per crop -- more than
        # the whole training step of a crop on this path -- so `cfg.epoch_results_numpy = True` keeps the numpy arrays
        # (every consumer in the reference wraps the lists in np.array / passes them to sklearn anyway)
        as_np = bool(getattr(self.cfg, "epoch_results_numpy", False))
        # (floats as float64: what np.array() makes of the reference's lists of Python floats, so that downstream means /
        # sklearn calls see the same dtype either way)
        def as_array(a):
            a = np.ascontiguousarray(a)
            return a.astype(np.float64) if a.dtype.kind == "f" else a

        ls = as_array if as_np else (lambda a: a.tolist())
        if names is None:
            res["running_loss"] = ls(loss[:, 0])
            res["confidences"] = ls(conf)
            res["predictions"] = ls(pred[:, 0])
            res["ground_truth"] = ls(gt[:, 0])
        else:
            rl, cf, pr, g = defaultdict(list), defaultdict(list), defaultdict(list), defaultdict(list)
            for t, n in enumerate(names):
                rl[n] = ls(loss[:, t])
                cf[n] = ls(conf[:, seg[t]:seg[t + 1]])
                pr[n] = ls(pred[:, t])
                g[n] = ls(gt[:, t])
            rl["loss"] = ls(loss[:, len(names)])
            res.update(running_loss=rl, confidences=cf, predictions=pr, ground_truth=g)
        if auc_counts is not None:
            seg_ = self._seg
            res["roc_auc_counts"] = (auc_counts if names is None else
                                     {n: auc_counts[seg_[t]:seg_[t + 1]] for t, n in enumerate(names)})
        if self.fused is not None and self.fused.state["cm"] is not None:
            cms = self.fused.confusion_matrices()
            res["confusion"] = cms[0] if names is None else {n: cms[t] for t, n in enumerate(names)}
        return res
