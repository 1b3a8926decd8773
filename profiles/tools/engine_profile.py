#!/usr/bin/env python
"""Host-side cost of one engine step (cProfile over a warm train_epoch on the bench's API leg): where the Python time
of `train_epoch` goes once the GPU work is ~0.6 ms per 4096-crop batch.   python profiles/tools/engine_profile.py"""
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
    pr = cProfile.Profile()
    orig = bench.api_bench

    # profile only the second call's warm epochs: run once for warm-up of every cache
    print(orig(wl, dev, 8)["train_epoch"])
    pr.enable()
    out = orig(wl, dev, 16)
    pr.disable()
    print(out["train_epoch"])
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
    print(s.getvalue()[:9000])


if __name__ == "__main__":
    main()
