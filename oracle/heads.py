"""CPU oracle: classification heads + CE / focal loss + multitask sum.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates with stock torch on
the CPU:

* head = ``nn.Sequential(nn.Dropout(p), nn.Linear(D, C_t))`` per task
  (nkb_classification/model.py:34-37, :104-108; forward :41-43, :114-116),
  evaluated at p = 0 / eval mode (per-task dropout masks are a documented
  deviation, SURVEY.md section 7 hard part 3);
* ``FocalLoss.forward``                    nkb_classification/losses.py:59-94
* ``nn.CrossEntropyLoss(weight)``          nkb_classification/losses.py:155-159
* ``MultitaskCriterion.__call__``          nkb_classification/losses.py:110-147
* ``BaseLogger.log_iter`` softmax(fp32)    nkb_classification/logging.py:268-281

Gradients come from torch autograd over this restatement (fp64 by default),
which tests/test_oracle_heads.py pins against the reference's own
``losses.py`` via the committed golden fixtures.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

LOSS_CE = 0
LOSS_FOCAL = 1
DEFAULT_FOCAL_GAMMA = 2.0  # losses.py:7
IGNORE_INDEX = -100


def focal_loss(x: torch.Tensor, y: torch.Tensor, alpha: Optional[torch.Tensor] = None,
               gamma: float = DEFAULT_FOCAL_GAMMA, ignore_index: int = IGNORE_INDEX) -> torch.Tensor:
    """losses.py:59-94 with reduction='mean': plain mean over the kept rows
    (NOT alpha-normalised); all rows ignored -> 0."""
    keep = y != ignore_index
    y = y[keep]
    if len(y) == 0:
        return torch.zeros((), dtype=x.dtype)
    x = x[keep]
    log_p = x.log_softmax(dim=-1)
    ce = F.nll_loss(log_p, y, weight=alpha, reduction="none", ignore_index=ignore_index)
    log_pt = log_p[torch.arange(len(x)), y]
    pt = log_pt.exp()
    focal_term = (1 - pt) ** gamma
    return (focal_term * ce).mean()


def cross_entropy(x: torch.Tensor, y: torch.Tensor, weight: Optional[torch.Tensor] = None,
                  ignore_index: int = IGNORE_INDEX) -> torch.Tensor:
    """``nn.CrossEntropyLoss(weight)``: sum_i w[y_i] * (-log p_i[y_i]) / sum_i w[y_i]."""
    return F.cross_entropy(x, y, weight=weight, ignore_index=ignore_index)


def heads_forward(emb: torch.Tensor, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor]):
    """model.py:114-116 in eval mode: one ``F.linear`` per task."""
    return [F.linear(emb, w, b) for w, b in zip(weights, biases)]


def heads_loss_fwd_bwd(
    emb: torch.Tensor,                      # [B, D]
    weights: Sequence[torch.Tensor],        # T x [C_t, D]
    biases: Sequence[torch.Tensor],         # T x [C_t]
    labels: torch.Tensor,                   # [B, T] int64
    loss_kind: int = LOSS_FOCAL,
    gamma: float = DEFAULT_FOCAL_GAMMA,
    class_weights: Optional[Sequence[Optional[torch.Tensor]]] = None,  # alpha / weight per task
    ignore_index: int = IGNORE_INDEX,
    dtype: torch.dtype = torch.float64,
    need_demb: bool = True,
) -> Dict[str, object]:
    """Whole K2 contract on the CPU: logits, fp32-style softmax probabilities,
    per-task losses, their unweighted sum (losses.py:134-146) and the autograd
    gradients w.r.t. every head weight / bias and the embeddings."""
    T = len(weights)
    emb_ = emb.detach().to(dtype).clone().requires_grad_(need_demb)
    Ws = [w.detach().to(dtype).clone().requires_grad_(True) for w in weights]
    bs = [b.detach().to(dtype).clone().requires_grad_(True) for b in biases]
    logits = heads_forward(emb_, Ws, bs)
    losses = []
    for t in range(T):
        cw = None
        if class_weights is not None and class_weights[t] is not None:
            cw = class_weights[t].to(dtype)
        y = labels[:, t].to(torch.int64)
        if loss_kind == LOSS_FOCAL:
            losses.append(focal_loss(logits[t], y, cw, gamma, ignore_index))
        else:
            losses.append(cross_entropy(logits[t], y, cw, ignore_index))
    total = sum(losses)
    grads = None
    if total.requires_grad:
        total.backward()
    out = {
        "logits": [z.detach() for z in logits],
        "probs": [z.detach().softmax(dim=-1) for z in logits],
        "loss": [l.detach() for l in losses],
        "total": total.detach(),
        "dW": [w.grad if w.grad is not None else torch.zeros_like(w) for w in Ws],
        "db": [b.grad if b.grad is not None else torch.zeros_like(b) for b in bs],
        "demb": (emb_.grad if (need_demb and emb_.grad is not None) else None),
    }
    return out


def unnormalised_sums(
    emb: torch.Tensor, weights, biases, labels: torch.Tensor, loss_kind: int, gamma: float,
    class_weights=None, ignore_index: int = IGNORE_INDEX, dtype=torch.float64,
):
    """The quantities a rank contributes to the K4 all-reduce: per task the SUM
    (not mean) of per-row losses, the denominator (kept rows, or sum of
    w[y_i] for weighted CE) and the gradients of the loss SUM.  The global
    mean loss / gradients are sums / denominators, which is what makes the
    N-GPU result equal the 1-GPU one."""
    T = len(weights)
    emb_ = emb.detach().to(dtype)
    Ws = [w.detach().to(dtype).clone().requires_grad_(True) for w in weights]
    bs = [b.detach().to(dtype).clone().requires_grad_(True) for b in biases]
    loss_sum, denom = [], []
    for t in range(T):
        z = F.linear(emb_, Ws[t], bs[t])
        y = labels[:, t].to(torch.int64)
        keep = y != ignore_index
        cw = None
        if class_weights is not None and class_weights[t] is not None:
            cw = class_weights[t].to(dtype)
        if keep.sum() == 0:
            loss_sum.append(torch.zeros((), dtype=dtype))
            denom.append(torch.zeros((), dtype=dtype))
            continue
        zk, yk = z[keep], y[keep]
        log_p = zk.log_softmax(-1)
        log_pt = log_p[torch.arange(len(yk)), yk]
        a = cw[yk] if cw is not None else torch.ones_like(log_pt)
        if loss_kind == LOSS_FOCAL:
            pt = log_pt.exp()
            per_row = -a * (1 - pt) ** gamma * log_pt
            denom.append(torch.tensor(float(len(yk)), dtype=dtype))
        else:
            per_row = -a * log_pt
            denom.append(a.sum().detach())
        loss_sum.append(per_row.sum())
    total = sum(loss_sum)
    if total.requires_grad:
        total.backward()
    return {
        "loss_sum": [l.detach() for l in loss_sum],
        "denom": denom,
        "dW_sum": [w.grad if w.grad is not None else torch.zeros_like(w) for w in Ws],
        "db_sum": [b.grad if b.grad is not None else torch.zeros_like(b) for b in bs],
    }
