"""inference() with the reference's signature and output file (inference.py:15-70).

Per batch the reference runs ``model(imgs)``, then per task ``argmax(-1).cpu()``
and a quadratic ``pd.concat``.  Here the heads + per-task argmax are one K2
(forward only) + one K3 launch, predictions stay on the device until the end of
the loop, and the CSV is assembled once.
"""
from __future__ import annotations

from pathlib import Path
from typing import Any, Union

import numpy as np
import pandas as pd
import torch
from tqdm import tqdm

from . import ops
from .model import HeadPack, MultitaskClassifier, SingletaskClassifier
from .utils import get_classes_configs


@torch.no_grad()
def inference(model: torch.nn.Module, loader, classes: Union[list, dict], save_path: str,
              device: Union[torch.device, str], cfg: Any) -> None:
    _, idx_to_class = get_classes_configs(classes)
    task = cfg.task
    assert task in ("single", "multi")
    if task == "single":
        columns = [cfg.target_column]
    else:
        target_names = cfg.target_names
        assert set(target_names) == set(classes.keys())
        columns = target_names.copy()
    columns.append("path")

    model.eval()
    base = getattr(model, "_orig_mod", model)
    pack = HeadPack(base) if isinstance(base, (SingletaskClassifier, MultitaskClassifier)) else None
    preds_dev, paths_all = [], []
    bufs_by_shape = {}    # one set of K2 buffers per batch shape (the last batch of a folder is usually shorter)
    for imgs, img_paths in tqdm(loader, leave=False, desc="Inference", disable=getattr(cfg, "disable_tqdm", False)):
        imgs = imgs.float().to(device)
        with torch.autocast(device_type="cuda", dtype=torch.bfloat16, enabled=cfg.enable_mixed_presicion):
            if pack is not None:
                emb = base.emb_model(imgs)
            else:
                out = model(imgs)
        if pack is not None:
            emb = emb.contiguous()
            bufs = bufs_by_shape.get(emb.shape[0])
            if bufs is None:
                bufs = ops.HeadsBuffers(emb.shape[0], pack.D, pack.seg, emb.device, want_probs=False, want_grads=False)
                bufs_by_shape[emb.shape[0]] = bufs
            pred = torch.empty((emb.shape[0], len(pack.seg) - 1), dtype=torch.int32, device=emb.device)
            ops.heads_fwd_loss_bwd(emb, pack.W_cat, pack.b_cat, None, bufs, out_pred=pred)   # K2 forward + fused argmax
            names = pack.names
        else:  # scripted / foreign model: logits from the model, argmax still K3
            if isinstance(out, dict):
                names = list(out.keys())
                z = torch.cat([out[n].float() for n in names], 1).contiguous()
                seg = np.concatenate([[0], np.cumsum([out[n].shape[1] for n in names])]).tolist()
            else:
                names, z, seg = None, out.float().contiguous(), [0, out.shape[1]]
            pred, _ = ops.argmax_confusion(z, seg)
        preds_dev.append(pred)
        paths_all += list(img_paths)
    if preds_dev:
        pred = torch.cat(preds_dev).cpu().numpy()
        def names_of(mapping, idx):      # class names by fancy indexing instead of a Python loop per prediction
            lut = np.array([str(mapping[i]) for i in range(len(mapping))], dtype=object)
            return lut[idx.astype(np.int64)]

        cols = {}
        if task == "single":
            cols[columns[0]] = names_of(idx_to_class, pred[:, 0])
        else:
            for target_name in target_names:
                cols[target_name] = names_of(idx_to_class[target_name], pred[:, names.index(target_name)])
        cols["path"] = np.array(paths_all, dtype=object)
        table = pd.DataFrame(cols, columns=columns)
    else:
        table = pd.DataFrame(columns=columns)
    table.to_csv(Path(save_path, "inference_annotations.csv"), index=False)
