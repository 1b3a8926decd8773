"""The fused heads + loss (+ metric) step as an autograd-aware call.

``fused_heads_loss`` replaces, in the engine's training / validation loop,

    preds = {t: classifier_t(emb)}            model.py:114-116
    loss  = criterion(preds, target)          losses.py:110-147, :59-94
    loss["loss"].backward() through the heads engine.py:55-58
    softmax / argmax of log_iter              logging.py:268-281

with K2 (one segmented GEMM + fused loss + gradients), K3 (argmax + confusion
counts) and K4 (all-reduce when sharded).  The returned loss vector carries
autograd history to ``emb`` and to every head parameter, so the reference's
``scaler.scale(loss).backward(); scaler.step(optimizer)`` sequence is unchanged.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import ops
from .hotpath import LOSS_KINDS
from .model import HeadPack
from .parallel import Communicator


@dataclass
class HeadsOutput:
    loss: torch.Tensor              # [T+1]: per-task mean losses, then their sum (autograd-connected)
    logits: torch.Tensor            # [B, sum C_t] fp32
    probs: Optional[torch.Tensor]   # [B, sum C_t] fp32 per-task softmax
    pred: Optional[torch.Tensor]    # [B, T] int32
    seg: List[int]
    names: Optional[List[str]]
    labels: Optional[torch.Tensor] = None   # int64 [B, T] on the device (what the kernels read; reused by the logger)

    def preds_like_reference(self):
        """Tensor (single task) or dict name -> logits view (multi task), as model.forward returns."""
        if self.names is None:
            return self.logits
        return {n: self.logits[:, a:b] for n, a, b in zip(self.names, self.seg[:-1], self.seg[1:])}

    def loss_like_reference(self):
        """Tensor (single) or dict {task: loss_t, "loss": total} (multi), as criterion returns."""
        if self.names is None:
            return self.loss[0]
        d = {n: self.loss[t] for t, n in enumerate(self.names)}
        d["loss"] = self.loss[len(self.names)]
        return d


class _FusedHeads(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, state, labels, train, *params):
        pack: HeadPack = state["pack"]
        emb_c = emb.detach()
        if emb_c.dtype == torch.float16:
            emb_c = emb_c.to(torch.bfloat16)  # reference autocasts to fp16; the kernels take bf16 | fp32
        emb_c = emb_c.contiguous()
        B = emb_c.shape[0]
        need_grads = bool(train)
        key = (B, need_grads, emb_c.dtype)
        bufs = state["bufs"].get(key)
        if bufs is None:
            if len(state["bufs"]) >= 8:          # train / val / last-batch shapes stay cached; cap the odd ones
                state["bufs"].pop(next(iter(state["bufs"])))
            bufs = ops.HeadsBuffers(B, pack.D, pack.seg, emb_c.device, want_logits=True, want_probs=True,
                                    want_grads=need_grads)
            bufs.generation = 0
            state["bufs"][key] = bufs
        bufs.generation += 1                     # backward checks that nobody has overwritten these gradients
        cm, cm_step = state.get("cm"), state.get("cm_step")
        pred = None
        if state.get("want_pred", True):
            pred = torch.empty((B, pack_T(pack)), dtype=torch.int32, device=emb_c.device)
        do_cm = pred is not None and labels is not None and cm is not None
        comm: Communicator = state["comm"]
        transport = state.get("transport", "peer")
        one_launch = need_grads and (comm.world == 1 or transport == "peer")
        peer = False
        if one_launch and comm.world > 1:
            peer = comm.init_peer(emb_c.device, bufs.reduce_buf.numel(), 0 if cm is None else cm.numel())
            one_launch = peer
        if one_launch:
            # forward + loss + K3 + dW/db + exchange (K4' as the kernel's epilogue) + finalize: ONE launch
            ops.heads_train_step(emb_c, pack.W_cat, pack.b_cat, labels, bufs, state["loss_kind"], state["gamma"],
                                 state["class_weight"], state["ignore_index"], out_pred=pred,
                                 cm_total=cm if do_cm else None, cm_step=cm_step if do_cm else None, peer=peer)
        else:
            # K3 (argmax + confusion counts) rides in K2's forward epilogue
            ops.heads_fwd_loss_bwd(emb_c, pack.W_cat, pack.b_cat, labels, bufs, state["loss_kind"], state["gamma"],
                                   state["class_weight"], state["ignore_index"], out_pred=pred,
                                   cm_step=cm_step if do_cm else None)
            comm.exchange_finalize(bufs, cm if do_cm else None, cm_step if do_cm else None,
                                   transport)
        ctx.state, ctx.bufs, ctx.emb_dtype, ctx.need_demb = state, bufs, emb.dtype, emb.requires_grad
        ctx.generation = bufs.generation
        ctx.n_params = len(params)
        state["last"] = (bufs, pred)
        return bufs.loss.clone()

    @staticmethod
    def backward(ctx, g):
        state, bufs = ctx.state, ctx.bufs
        pack: HeadPack = state["pack"]
        if bufs.dlogits is None:
            raise RuntimeError("fused heads were run with train=False; no gradients were produced")
        if bufs.generation != ctx.generation:
            raise RuntimeError("the fused heads ran again (same batch shape) before this backward: its gradient "
                               "buffers were overwritten -- call backward() before the next forward")
        T = pack_T(pack)
        # every output is a sum of per-task means; total (index T) is the usual one to differentiate
        per_task = g[:T] + g[T]
        rows = per_task if T == 1 else per_task.index_select(0, state["task_of_class"])      # [NC] (no host sync)
        demb = None
        if ctx.need_demb:
            out_dtype = torch.bfloat16 if ctx.emb_dtype in (torch.float16, torch.bfloat16) else torch.float32
            demb = ops.heads_demb(bufs, pack.W_cat, out_dtype=out_dtype,
                                  task_scale=per_task.to(torch.float32).contiguous()).to(ctx.emb_dtype)
        dW = bufs.dW() * rows[:, None]
        db = bufs.db() * rows
        grads = []
        for t in range(T):
            a, b = pack.seg[t], pack.seg[t + 1]
            grads += [dW[a:b], db[a:b]]
        return (demb, None, None, None, *grads[: ctx.n_params])


def pack_T(pack) -> int:
    return len(pack.classes)


class _TaskPack:
    """One head of a HeadPack seen as a single-task pack (row slices of the packed buffers are contiguous)."""

    def __init__(self, pack: HeadPack, t: int):
        a, b = pack.seg[t], pack.seg[t + 1]
        self.W_cat, self.b_cat = pack.W_cat[a:b], pack.b_cat[a:b]
        self.D, self.classes, self.seg = pack.D, [b - a], [0, b - a]
        self.names = None if pack.names is None else [pack.names[t]]
        self._lin = pack.linears[t]

    def params(self):
        return [self._lin.weight, self._lin.bias]


class FusedHeads:
    """Binds a classifier's heads, a loss configuration and (optionally) a communicator.

    ``cfg_loss`` is the reference's ``criterion`` config dict (``type``, optional ``gamma`` / ``alpha`` / ``weight``).
    """

    def __init__(self, model, cfg_loss: dict, comm: Optional[Communicator] = None, ignore_index: int = -100,
                 track_confusion: bool = True):
        if cfg_loss["type"] not in LOSS_KINDS:
            raise NotImplementedError(f'Unknown loss type in config: {cfg_loss["type"]}')
        self.model = model
        self.pack = HeadPack(model)
        dev = self.pack.W_cat.device
        cw = cfg_loss.get("alpha" if cfg_loss["type"] == "FocalLoss" else "weight")
        class_weight = None
        if cw is not None:
            w = torch.tensor(cw, dtype=torch.float32, device=dev)
            for c in self.pack.classes:
                if c != w.numel():
                    raise RuntimeError("weight tensor should be defined either for all classes or no classes")
            class_weight = w.repeat(len(self.pack.classes)).contiguous()
        gamma = float(cfg_loss.get("gamma", 2.0)) if cfg_loss["type"] == "FocalLoss" else 0.0
        n_cm = ops.confusion_len(self.pack.seg)
        self.state = {
            "pack": self.pack, "bufs": {}, "loss_kind": LOSS_KINDS[cfg_loss["type"]], "gamma": gamma,
            "class_weight": class_weight, "ignore_index": int(ignore_index), "comm": comm or Communicator(),
            "cm": torch.zeros(n_cm, dtype=torch.int64, device=dev) if track_confusion else None,
            "cm_step": torch.zeros(n_cm, dtype=torch.int64, device=dev) if track_confusion else None,
            # class -> task index, resident on the device: backward scales the gradient rows without a host sync
            "task_of_class": torch.repeat_interleave(torch.arange(len(self.pack.classes)),
                                                     torch.tensor(self.pack.classes)).to(dev),
        }

    @property
    def names(self):
        return self.pack.names

    def labels_tensor(self, target, device) -> torch.Tensor:
        """Reference targets (Tensor [B] or dict name -> Tensor [B]) -> contiguous int64 [B, T] on the device."""
        if isinstance(target, dict):
            cols = [target[n].to(device=device, dtype=torch.int64, non_blocking=True).reshape(-1) for n in self.pack.names]
            return torch.stack(cols, dim=1).contiguous()
        return target.to(device=device, dtype=torch.int64, non_blocking=True).reshape(-1, 1).contiguous()

    def __call__(self, emb: torch.Tensor, target, train: bool) -> HeadsOutput:
        if not self.pack.in_sync():
            self.pack.repack()
        labels = self.labels_tensor(target, emb.device)
        if train and self.model.training and max(self.pack.dropout_p) > 0.0:
            if pack_T(self.pack) > 1:
                return self._per_head_dropout(emb, labels, train)
            emb = torch.nn.functional.dropout(emb, p=self.pack.dropout_p[0], training=True)   # model.py:35
        loss = _FusedHeads.apply(emb, self.state, labels, train, *self.pack.params())
        bufs, pred = self.state["last"]
        return HeadsOutput(loss=loss, logits=bufs.logits, probs=bufs.probs, pred=pred, seg=self.pack.seg,
                           names=self.pack.names, labels=labels)

    def _per_head_dropout(self, emb: torch.Tensor, labels: torch.Tensor, train: bool) -> HeadsOutput:
        """Training with ``classifier_dropout > 0`` and several heads: every head applies its OWN ``nn.Dropout`` to
        the shared embedding (model.py:102-116), so the heads see different inputs and cannot share one GEMM.  Each
        head then runs as a single-task fused call on its own mask, drawn with ``F.dropout`` in ModuleDict order --
        the very calls (and, without autocast, the very Philox stream) the reference makes.  Eval and p = 0 keep the
        single segmented launch."""
        pack, T = self.pack, pack_T(self.pack)
        subs = self.state.setdefault("task_states", None)
        if subs is None:
            subs, off = [], 0
            for t in range(T):
                tp = _TaskPack(pack, t)
                C = tp.classes[0]
                cw = self.state["class_weight"]
                st = dict(self.state)
                st.update(pack=tp, bufs={}, last=None, task_of_class=torch.zeros(C, dtype=torch.int64, device=emb.device),
                          class_weight=None if cw is None else cw[pack.seg[t]:pack.seg[t + 1]].contiguous(),
                          cm=None if self.state["cm"] is None else self.state["cm"][off: off + C * C],
                          cm_step=None if self.state["cm_step"] is None else self.state["cm_step"][off: off + C * C])
                st.pop("task_states", None)
                subs.append(st)
                off += C * C
            self.state["task_states"] = subs
        losses, logits, probs, preds = [], [], [], []
        for t in range(T):
            st = subs[t]
            if not pack.in_sync():
                pack.repack()
            emb_t = torch.nn.functional.dropout(emb, p=pack.dropout_p[t], training=True)
            l = _FusedHeads.apply(emb_t, st, labels[:, t:t + 1].contiguous(), train, *st["pack"].params())
            bufs, pred = st["last"]
            losses.append(l[0])
            logits.append(bufs.logits)
            probs.append(bufs.probs)
            preds.append(pred)
        total = losses[0]
        for l in losses[1:]:
            total = total + l                                            # losses.py:140-147: unweighted sum
        return HeadsOutput(loss=torch.stack(losses + [total]), logits=torch.cat(logits, 1),
                           probs=None if probs[0] is None else torch.cat(probs, 1),
                           pred=None if preds[0] is None else torch.cat(preds, 1), seg=pack.seg, names=pack.names,
                           labels=labels)

    # epoch-level confusion counts (already summed over ranks)
    def reset_confusion(self):
        if self.state["cm"] is not None:
            self.state["cm"].zero_()

    def confusion_matrices(self):
        flat = self.state["cm"].cpu().numpy()
        out, off = [], 0
        for c in self.pack.classes:
            out.append(flat[off: off + c * c].reshape(c, c).copy())
            off += c * c
        return out
