#!/usr/bin/env python
"""Where the time of the fused heads step (k2_fused_step) goes: per-CTA clock64 stamps at the phase boundaries
(NKBK_FUSED_TIMING=1 -> nkbk_debug_fused_timing), averaged over the CTAs, in microseconds.

    python profiles/tools/k2_phases.py        -> one JSON line per shape
"""
import ctypes
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
os.environ["NKBK_FUSED_TIMING"] = "1"

SHAPES = [("cfg5_B4096", 4096, 2048, (10,)), ("cfg5_B512", 512, 2048, (10,)), ("cfg3_B1280", 1280, 768, (3,)),
          ("cfg4_B1024", 1024, 768, (2, 3, 4, 7, 14)), ("tiny_B64", 64, 256, (4,))]
PHASES = ["tables", "weights_wait", "pass1_forward", "epilogue", "pass2_dw", "write_partials", "grid_barrier",
          "reduce_setup", "reduce_finalize"]


def main():
    from nkb_classification_b200 import _lib, ops
    dev = torch.device("cuda:0")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    for name, B, D, classes in SHAPES:
        for dt in (torch.float32, torch.bfloat16):
            seg = np.concatenate([[0], np.cumsum(classes)]).tolist()
            T, NC = len(classes), sum(classes)
            g = torch.Generator().manual_seed(0)
            W = (torch.randn(NC, D, generator=g) * (2.0 / D) ** 0.5).to(dev)
            b = torch.zeros(NC, device=dev)
            labels = torch.stack([torch.randint(0, c, (B,), generator=g) for c in classes], 1).contiguous().to(dev)
            emb = torch.randn(B, D, generator=g).to(dev).to(dt).contiguous()
            bufs = ops.HeadsBuffers(B, D, seg, dev)
            ncm = ops.confusion_len(seg)
            cm, cs = torch.zeros(ncm, dtype=torch.int64, device=dev), torch.zeros(ncm, dtype=torch.int64, device=dev)
            pred = torch.empty((B, T), dtype=torch.int32, device=dev)
            rows = []
            for it in range(6):
                flush.fill_(1)
                ops.heads_train_step(emb, W, b, labels, bufs, 0, 0.0, out_pred=pred, cm_total=cm, cm_step=cs)
                torch.cuda.synchronize()
                out = np.zeros((148, 16), dtype=np.uint64)
                n = _lib.lib().nkbk_debug_fused_timing(out.ctypes.data_as(ctypes.c_void_p), 148)
                if n == 0:
                    break
                if it >= 2:
                    rows.append(out[:n].astype(np.int64))
            if not rows:
                print(json.dumps({"shape": name, "note": "fused kernel did not run"}))
                continue
            st = np.stack(rows)                        # [iters, ctas, 12]
            wall_ns = (st[:, :, 11].max(1) - st[:, :, 0].min(1)).mean()
            cyc = st[:, :, 1:11]
            ns_per_cycle = float(((st[:, :, 11] - st[:, :, 0]) / np.maximum(cyc[:, :, -1] - cyc[:, :, 0], 1)).mean())
            d = np.diff(cyc, axis=2) * ns_per_cycle / 1e3   # us per phase per CTA
            rec = {"shape": name, "emb": str(dt).split(".")[-1], "ctas": int(st.shape[1]),
                   "kernel_wall_us": float(wall_ns / 1e3), "ns_per_cycle": ns_per_cycle,
                   "entry_skew_us": float((st[:, :, 0].max(1) - st[:, :, 0].min(1)).mean() / 1e3)}
            for i, ph in enumerate(PHASES):
                rec[ph] = {"mean": float(d[:, :, i].mean()), "max": float(d[:, :, i].max(1).mean())}
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
