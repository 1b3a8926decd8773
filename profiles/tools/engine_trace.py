#!/usr/bin/env python
"""Device timeline of the engine's training step on the bench's API leg (torch.profiler, CUDA activities only):
which kernels run per 4096-crop batch and for how long.   python profiles/tools/engine_trace.py"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))

import bench  # noqa: E402


def main():
    import torch
    from torch.profiler import ProfilerActivity, profile
    from nkb_classification_b200.synthetic import DEFAULT_WORKLOAD, WORKLOADS
    wl = WORKLOADS[DEFAULT_WORKLOAD]
    dev = torch.device("cuda:0")
    bench.api_bench(wl, dev, 4)                       # warm every cache / allocator
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        out = bench.api_bench(wl, dev, 8)
    print(out["train_epoch"])
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))


if __name__ == "__main__":
    main()
