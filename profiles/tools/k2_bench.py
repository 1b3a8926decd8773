#!/usr/bin/env python
"""Times the heads step (K2 + K3 + finalize) alone, per call, with CUDA events: the fused one-launch kernel
(nkbk_heads_train_step -> k2_fused_step) against the three-kernel path it replaces (NKBK_DISABLE_FUSED_HEADS=1).
L2 is flushed between calls (a 512 MB fill), so every call reads its embeddings from HBM as it does inside the step.

    python profiles/tools/k2_bench.py [--iters 30]      -> one JSON line per (shape, dtype, path)
"""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))

SHAPES = [  # (name, B, D, classes, loss kind, gamma)
    ("cfg5_B4096", 4096, 2048, (10,), 0, 0.0),
    ("cfg5_strong8_B512", 512, 2048, (10,), 0, 0.0),
    ("cfg5_strong4_B1024", 1024, 2048, (10,), 0, 0.0),
    ("cfg4_B1024", 1024, 768, (2, 3, 4, 7, 14), 1, 1.0),
    ("cfg4_strong8_B128", 128, 768, (2, 3, 4, 7, 14), 1, 1.0),
    ("cfg2_B256", 256, 1280, (4, 7, 2), 1, 1.0),
    ("cfg3_B1280", 1280, 768, (3,), 0, 0.0),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    from nkb_classification_b200 import _lib, ops
    dev = torch.device("cuda:0")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    names = {_lib.PATH_FUSED: "fused", _lib.PATH_TC_FWD: "tcgen05_fwd+dw+finalize", _lib.PATH_FFMA_FWD: "ffma_fwd+dw+finalize"}
    for name, B, D, classes, kind, gamma in SHAPES:
        if args.only and args.only not in name:
            continue
        seg = np.concatenate([[0], np.cumsum(classes)]).tolist()
        T, NC = len(classes), sum(classes)
        g = torch.Generator().manual_seed(0)
        W = (torch.randn(NC, D, generator=g) * (2.0 / D) ** 0.5).to(dev)
        b = torch.zeros(NC, device=dev)
        labels = torch.stack([torch.randint(0, c, (B,), generator=g) for c in classes], 1).contiguous().to(dev)
        emb32 = torch.randn(B, D, generator=g).to(dev)
        ncm = ops.confusion_len(seg)
        for dt in (torch.float32, torch.bfloat16):
            emb = emb32.to(dt).contiguous()
            for disable in ("0", "1"):
                os.environ["NKBK_DISABLE_FUSED_HEADS"] = disable
                bufs = ops.HeadsBuffers(B, D, seg, dev)
                cm = torch.zeros(ncm, dtype=torch.int64, device=dev)
                cs = torch.zeros(ncm, dtype=torch.int64, device=dev)
                pred = torch.empty((B, T), dtype=torch.int32, device=dev)

                def call():
                    ops.heads_train_step(emb, W, b, labels, bufs, kind, gamma, out_pred=pred, cm_total=cm, cm_step=cs)

                for _ in range(5):
                    call()
                torch.cuda.synchronize()
                l0 = _lib.launch_count()
                call()
                launches = _lib.launch_count() - l0
                path = ops.heads_last_path()
                ts = []
                for _ in range(args.iters):
                    flush.fill_(1)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    call()
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3)
                # back to back without flush (what a graph replay / a serial step tail sees when emb is L2 resident)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.iters):
                    call()
                e1.record()
                torch.cuda.synchronize()
                hot = e0.elapsed_time(e1) * 1e3 / args.iters
                print(json.dumps({"shape": name, "B": B, "D": D, "NC": NC, "T": T, "emb": str(dt).split(".")[-1],
                                  "path": names.get(path, str(path)), "launches": int(launches),
                                  "us_cold_median": float(np.median(ts)), "us_cold_min": float(np.min(ts)),
                                  "us_back_to_back": hot,
                                  "emb_MB": B * D * emb.element_size() / 1e6}), flush=True)
    os.environ.pop("NKBK_DISABLE_FUSED_HEADS", None)


if __name__ == "__main__":
    main()
