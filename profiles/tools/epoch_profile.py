#!/usr/bin/env python
"""Where a WARM epoch of the API leg spends its host time outside the per-batch loop (planning, epoch results):
cProfile over warm train_epoch calls with numpy epoch results, sorted by own time.   python profiles/tools/epoch_profile.py"""
import cProfile
import io
import pstats
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))

import bench  # noqa: E402


def main():
    import torch
    from nkb_classification_b200.synthetic import DEFAULT_WORKLOAD, WORKLOADS
    wl = WORKLOADS[DEFAULT_WORKLOAD]
    dev = torch.device("cuda:0")
    bench.api_bench(wl, dev, 8)                       # warm every cache / allocator
    pr = cProfile.Profile()
    pr.enable()
    out = bench.api_bench(wl, dev, 24)
    pr.disable()
    print({k: v for k, v in out["train_epoch"].items() if k != "note"})
    for key in ("tottime", "cumulative"):
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats(key).print_stats(28)
        print(s.getvalue()[:6000])


if __name__ == "__main__":
    main()
