#!/usr/bin/env python
"""Per-CTA timeline of one K1 launch (TMA kernel) from in-kernel %globaltimer stamps (NKBK_K1_TIMING): where a short launch
loses time against its per-crop rate -- ramp (first CTA entries), steady state (SMs with resident CTAs) and tail (SMs that
have run dry while others still work).

    NKBK_K1_TIMING=1 python profiles/tools/k1_timeline.py [--workload cfg5_shard8]      -> one JSON line
"""
import argparse
import ctypes
import json
import os
import sys
from pathlib import Path

os.environ.setdefault("NKBK_K1_TIMING", "1")
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg5_shard8")
    args = ap.parse_args()
    from nkb_classification_b200 import _lib
    from nkb_classification_b200.parallel import Communicator
    from nkb_classification_b200.synthetic import WORKLOADS
    wl = WORKLOADS[args.workload]
    dev = torch.device("cuda:0")
    leg = bench.Leg(wl, dev, Communicator(), torch.float32, torch.float32, "peer")
    for _ in range(5):
        leg.k1()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    leg.k1()
    e1.record()
    torch.cuda.synchronize()
    cap = 1 << 20
    buf = np.zeros((cap, 3), dtype=np.uint64)
    n = int(_lib.lib().nkbk_debug_k1_timeline(ctypes.c_void_p(buf.ctypes.data), cap))
    t = buf[:n].astype(np.int64)
    sm, t0, t1 = t[:, 0], t[:, 1], t[:, 2]
    ok = t1 > 0
    sm, t0, t1 = sm[ok], t0[ok], t1[ok]
    base = t0.min()
    t0, t1 = (t0 - base) / 1e3, (t1 - base) / 1e3          # us
    end = float(t1.max())
    dur = t1 - t0
    # per SM: when its first CTA entered, when its last CTA left
    sms = np.unique(sm)
    first = np.array([t0[sm == s].min() for s in sms])
    last = np.array([t1[sm == s].max() for s in sms])
    # number of resident CTAs over time (0.5 us bins)
    bins = np.arange(0.0, end + 0.5, 0.5)
    active = np.array([int(((t0 <= b) & (t1 > b)).sum()) for b in bins])
    full = 4 * len(sms)
    out = {
        "workload": wl.name, "crops": leg.n, "ctas": int(len(t0)), "sms": int(len(sms)),
        "kernel_span_us": end, "event_us": e0.elapsed_time(e1) * 1e3,
        "cta_duration_us": {"mean": float(dur.mean()), "p10": float(np.percentile(dur, 10)), "p50": float(np.median(dur)),
                            "p90": float(np.percentile(dur, 90)), "max": float(dur.max())},
        "first_entry_us": {"p50": float(np.median(first)), "max": float(first.max())},
        "last_entry_of_any_cta_us": float(t0.max()),
        "sm_last_exit_us": {"min": float(last.min()), "p10": float(np.percentile(last, 10)), "p50": float(np.median(last)),
                            "p90": float(np.percentile(last, 90)), "max": float(last.max())},
        "mean_idle_tail_per_sm_us": float((end - last).mean()),
        "time_until_resident_ctas_reach_90pct_us": float(bins[np.argmax(active >= 0.9 * full)]) if (active >= 0.9 * full).any() else None,
        "time_when_resident_ctas_drop_below_50pct_us": float(bins[len(active) - 1 - np.argmax(active[::-1] >= 0.5 * full)]),
        "resident_ctas_every_5us": active[::10].tolist(),
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
