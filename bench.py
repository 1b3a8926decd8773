#!/usr/bin/env python
"""bench.py -- the hot path's headline metric on B200: crops/sec through
(preprocess + heads + loss + metric), device-resident (`value`), end to end
from pinned host buffers (`e2e`), the dominant kernel against the HBM roofline
(`roofline`) and the reference's CPU path timed on the same box (`cpu_baseline`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

N > 1 is launched by torchrun (one rank per GPU); per-GPU work is fixed (weak
scaling): every rank owns its own frames, the only exchange is the K4 all-reduce.
The backbone is out of scope for this path (timed separately by the caller):
K2 consumes synthetic embeddings of the named width.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

METRIC = "crops/sec (preprocess+heads+loss)"
UNIT = "crops/s"
MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--out-dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--emb-dtype", default="f32", choices=["f32", "bf16"],
                    help="dtype of the synthetic backbone output fed to K2 (bf16 = autocast path, tcgen05 forward)")
    ap.add_argument("--allreduce", default="peer", choices=["peer", "nccl"],
                    help="exchange step at N > 1: K4' fused NVLink peer-memory all-reduce + finalize, or NCCL (K4)")
    ap.add_argument("--resize", default=None, choices=["stretch", "letterbox"],
                    help="override the workload's geometry: A.Resize(S,S) or A.LongestMaxSize(S)+A.PadIfNeeded(S,S)")
    ap.add_argument("--train-aug", action="store_true",
                    help="K1 with the reference's train-time augmentations fused (configs/singletask_config.py:162-201); "
                         "parameters are drawn once on the host before the timed region")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def measured_peak_hbm():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.004):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        self.period = period_s
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, repr(e)

    _NAMES = {
        0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
        0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting",
    }

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self._NAMES.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.nv is not None:
            self._stop.clear()
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def make_heads(wl, seed=7):
    import torch
    g = torch.Generator().manual_seed(seed)
    Ws = [torch.randn(c, wl.emb_dim, generator=g) * (2.0 / wl.emb_dim) ** 0.5 for c in wl.classes]  # kaiming_normal_
    bs = [torch.zeros(c) for c in wl.classes]
    return Ws, bs


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's path on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_path_rate(wl, boxes, fidx, sample_crops, procs, steps=1, warmup=0, seed=1234):
    """crops/s of the reference CPU path on the first `sample_crops` crops (whole frames' worth)."""
    import torch
    from oracle.cpu_baseline import CpuReferencePath

    sample_crops = min(sample_crops, len(fidx))
    n_frames = int(fidx[sample_crops - 1]) + 1
    rng = np.random.default_rng(seed)
    frames = rng.integers(0, 256, (n_frames, wl.frame_h, wl.frame_w, 3), dtype=np.uint8)
    path = CpuReferencePath(frames, (wl.out_size, wl.out_size), MEAN, STD, procs=procs, max_crops=sample_crops)
    g = torch.Generator().manual_seed(7)
    emb = torch.randn(sample_crops, wl.emb_dim, generator=g)
    labels = torch.stack([torch.randint(0, c, (sample_crops,), generator=g) for c in wl.classes], 1)
    Ws, bs = make_heads(wl)
    b, f = boxes[:sample_crops], fidx[:sample_crops]
    try:
        for _ in range(warmup):
            path.step(b, f, emb, Ws, bs, labels, wl.loss, wl.gamma)
        times = [path.step(b, f, emb, Ws, bs, labels, wl.loss, wl.gamma) for _ in range(steps)]
    finally:
        path.close()
    return sample_crops * len(times) / sum(times), times, n_frames


def run_reference(args, wl):
    """--impl reference: the reference's own CPU implementation of the path (oracle port: cv2 + torch CPU),
    all host threads, bounded sample per step.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from nkb_classification_b200.synthetic import synth_boxes

    boxes, fidx = synth_boxes(wl)
    procs = os.cpu_count() or 1
    # size the per-step sample so the whole run stays within ~2 minutes
    probe_n = min(256, len(fidx))
    rate, _, _ = cpu_path_rate(wl, boxes, fidx, probe_n, procs, steps=1, warmup=1)
    budget_s = 100.0
    total_steps = max(1, args.steps + args.warmup)
    sample = int(min(len(fidx), max(probe_n, rate * budget_s / total_steps)))
    sample = max(wl.boxes_per_frame, sample // wl.boxes_per_frame * wl.boxes_per_frame)
    rate, times, n_frames = cpu_path_rate(wl, boxes, fidx, sample, procs, steps=args.steps, warmup=args.warmup)
    ms = 1e3 * float(np.mean(times))
    desc = (f"{sample} crops ({n_frames} frames) per step of {wl.name}; cv2 {__import__('cv2').__version__} + "
            f"torch CPU fp32, {procs} worker processes")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/int32->f32", "data": "synthetic",
        "config": workload_config(wl, "f32", args.gpus),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": procs, "kind": "port", "sample": desc},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(wl, out_dtype, n_gpus, transport="peer"):
    return {
        "workload": wl.name,
        "frames_per_gpu": wl.frames, "frame": f"{wl.frame_w}x{wl.frame_h}x3 u8", "boxes_per_frame": wl.boxes_per_frame,
        "crops_per_gpu_per_step": wl.crops, "out": f"3x{wl.out_size}x{wl.out_size} {out_dtype}", "resize": wl.mode,
        "emb_dim": wl.emb_dim, "heads": list(wl.classes), "loss": wl.loss, "gamma": wl.gamma,
        "train_aug": wl.train_aug,
        "backbone": "excluded (out of scope; synthetic embeddings)",
        "l2": "inputs larger than L2 (frames + output >> 126 MB per step); no flush needed",
        "parallelism": f"dp{n_gpus} (frames sharded per rank; head grads + confusion counts all-reduced by "
                       + ("one fused push/sum/finalize kernel over NVLink peer memory, K4')" if transport == "peer"
                          else "NCCL, K4)"),
        "pipelining": "value: K1 (stream A) overlaps K2/K3/K4 of another batch (stream B, higher priority), no dependency "
                      "without the backbone; serial_value: one stream, K1 -> K2 -> K3 -> K4 back to back",
    }


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_b200(args, wl):
    import torch
    import torch.distributed as dist

    from nkb_classification_b200 import _lib, hotpath, ops, transforms as T
    from nkb_classification_b200.parallel import Communicator
    from nkb_classification_b200.synthetic import k1_algorithmic_bytes, synth_boxes

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not _lib.LIB_PATH.exists() and local_rank == 0:   # normally prebuilt in-tree by __graft_entry__.build()
        from nkb_classification_b200 import build as _build
        _build.build()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # N ranks pull 398 MB of pinned frames per step each: keep every rank (and the pinned pages it first-touches)
        # on the CPU socket its GPU hangs off, otherwise the e2e leg crosses the inter-socket link
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        except Exception:
            pass
    comm = Communicator()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        comm.init_from_torch_distributed(dev)

    out_dtype = torch.float32 if args.out_dtype == "f32" else torch.bfloat16
    geo = ([T.Resize(wl.out_size, wl.out_size)] if wl.mode == "stretch" else
           [T.LongestMaxSize(wl.out_size), T.PadIfNeeded(wl.out_size, wl.out_size, border_mode=T.BORDER_CONSTANT, value=0)])
    aug_ops = []
    if args.train_aug:
        aug_ops = [T.HorizontalFlip(p=0.5), T.VerticalFlip(p=0.5),
                   T.RandomBrightnessContrast(brightness_limit=(-0.2, 0.2), contrast_limit=(0.1, -0.5), p=0.5),
                   T.HueSaturationValue(hue_shift_limit=0, sat_shift_limit=10, val_shift_limit=50, p=0.5),
                   T.CoarseDropout(max_holes=4, min_holes=1, max_height=0.2, min_height=0.05, max_width=0.2,
                                   min_width=0.05, fill_value=[0, 0.5, 1], p=0.5)]
    plan = T.compile_pipeline(geo + aug_ops + [T.Normalize(MEAN, STD), T.ToTensorV2()])
    hp = hotpath.HotPath(plan, wl.classes, wl.emb_dim, wl.loss, wl.gamma, device=dev, comm=comm, out_dtype=out_dtype,
                         transport=args.allreduce)

    # ---- synthetic inputs, generated on the device, seeded per rank ----
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    frames = torch.randint(0, 256, (wl.frames, wl.frame_h, wl.frame_w, 3), dtype=torch.uint8, device=dev, generator=g)
    boxes_np, fidx_np = synth_boxes(wl, seed=4321 + rank)
    n = len(fidx_np)
    boxes = torch.from_numpy(boxes_np).to(dev)
    fidx = torch.from_numpy(fidx_np).to(dev)
    gc = torch.Generator().manual_seed(7 + rank)
    emb = torch.randn(n, wl.emb_dim, generator=gc).to(dev)
    if args.emb_dtype == "bf16":
        emb = emb.to(torch.bfloat16)
    labels_h = torch.stack([torch.randint(0, c, (n,), generator=gc) for c in wl.classes], 1).contiguous()
    labels = labels_h.to(dev)
    Ws, bs = make_heads(wl)
    W_cat, b_cat = torch.cat(Ws).contiguous().to(dev), torch.cat(bs).contiguous().to(dev)

    aug = None
    if args.train_aug:
        import random as _random
        aug = plan.draw(n, _random.Random(99 + rank))
        _pre = hp.preprocess
        hp.preprocess = lambda fr, bx, fi, frame_desc=None: _pre(fr, bx, fi, frame_desc, aug)   # same call sites below

    def step():
        hp.preprocess(frames, boxes, fidx)
        return hp.heads_step(emb, W_cat, b_cat, labels, train=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region 1: inputs resident in HBM ----
    sampler = ClockSampler(local_rank)
    k1_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _lib.launch_count()
    sampler.start()
    barrier()
    e0.record()
    for i in range(args.steps):
        k1_ev[i][0].record()
        hp.preprocess(frames, boxes, fidx)
        k1_ev[i][1].record()
        hp.heads_step(emb, W_cat, b_cat, labels, train=True)
    e1.record()
    barrier()
    sampler.stop()
    launches = _lib.launch_count() - launches0
    dt_ms = e0.elapsed_time(e1)
    k1_ms = float(np.mean([a.elapsed_time(b) for a, b in k1_ev]))
    tmax = torch.tensor([dt_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dt_ms_max = float(tmax.item())
    serial_value = world * n * args.steps / (dt_ms_max * 1e-3)
    serial_ms = dt_ms_max / args.steps

    # ---- timed region 1b: the same K steps software-pipelined on two streams ----
    # K1 of a batch and K2/K3/K4 of another batch have no data dependency (the backbone sits between them and is
    # out of scope here), so a real loop runs preprocessing of batch i+1 while the heads/loss/metric of batch i
    # execute.  All work of all K steps still happens inside the timed region; both streams are joined before e1.
    # The heads stream has the higher priority: its short kernels take SM slots as K1's CTAs retire instead of queueing
    # behind the whole K1 grid, so K2/K3/K4 of a batch really run under the K1 of the next one.
    prio = os.environ.get("NKBK_BENCH_HEADS_PRIORITY", "high")   # experiment knob: high | same | low
    s_pre = torch.cuda.Stream(device=dev, priority=-1 if prio == "low" else 0)
    s_heads = torch.cuda.Stream(device=dev, priority=-1 if prio == "high" else 0)
    cur = torch.cuda.current_stream(dev)

    def pipelined(k):
        s_pre.wait_stream(cur)
        s_heads.wait_stream(cur)
        for _ in range(k):
            with torch.cuda.stream(s_pre):
                hp.preprocess(frames, boxes, fidx)
            with torch.cuda.stream(s_heads):
                hp.heads_step(emb, W_cat, b_cat, labels, train=True)
        cur.wait_stream(s_pre)
        cur.wait_stream(s_heads)

    pipelined(3)
    barrier()
    launches0 = _lib.launch_count()
    sampler.start()
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    pipelined(args.steps)
    p1.record()
    barrier()
    sampler.stop()
    launches = _lib.launch_count() - launches0
    tp = torch.tensor([p0.elapsed_time(p1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
    dt_ms_max = float(tp.item())
    value = world * n * args.steps / (dt_ms_max * 1e-3)

    # ---- timed region 2: end to end from pinned host memory through the public API ----
    e2e = None
    if not args.no_e2e:
        frames_h = torch.empty(frames.shape, dtype=torch.uint8).pin_memory()
        frames_h.copy_(frames)
        boxes_h = torch.from_numpy(boxes_np).pin_memory()
        fidx_h = torch.from_numpy(fidx_np).pin_memory()
        labels_p = labels_h.pin_memory()
        loss_h = torch.empty(hp.T + 1, dtype=torch.float32).pin_memory()
        cm_h = torch.empty(hp.cm.numel(), dtype=torch.int64).pin_memory()
        # double-buffered ingest: the H2D copy of step i+1 (copy stream) overlaps K1..K4 of step i (compute stream)
        nbuf = 2
        frames_d = [torch.empty_like(frames) for _ in range(nbuf)]
        boxes_d = [torch.empty_like(boxes) for _ in range(nbuf)]
        fidx_d = [torch.empty_like(fidx) for _ in range(nbuf)]
        labels_d = [torch.empty_like(labels) for _ in range(nbuf)]
        s_copy = torch.cuda.Stream(device=dev)
        copied = [torch.cuda.Event() for _ in range(nbuf)]
        consumed = [torch.cuda.Event() for _ in range(nbuf)]

        def e2e_run(k):
            s_copy.wait_stream(cur)
            for b in range(nbuf):
                consumed[b].record(cur)
            for i in range(k):
                b = i % nbuf
                with torch.cuda.stream(s_copy):
                    s_copy.wait_event(consumed[b])          # the buffer's previous contents have been preprocessed
                    frames_d[b].copy_(frames_h, non_blocking=True)
                    boxes_d[b].copy_(boxes_h, non_blocking=True)
                    fidx_d[b].copy_(fidx_h, non_blocking=True)
                    labels_d[b].copy_(labels_p, non_blocking=True)
                    copied[b].record(s_copy)
                cur.wait_event(copied[b])
                hp.preprocess(frames_d[b], boxes_d[b], fidx_d[b])
                consumed[b].record(cur)
                bufs = hp.heads_step(emb, W_cat, b_cat, labels_d[b], train=True)
                loss_h.copy_(bufs.loss, non_blocking=True)
                cm_h.copy_(hp.cm, non_blocking=True)
            cur.wait_stream(s_copy)

        e2e_steps = max(3, min(args.steps, 30))
        e2e_run(3)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        e2e_run(e2e_steps)
        a1.record()
        barrier()
        t2 = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        h2d = frames_h.numel() + boxes_h.numel() * 4 + fidx_h.numel() * 4 + labels_p.numel() * 8
        d2h = loss_h.numel() * 4 + cm_h.numel() * 8
        e2e = {"value": world * n * e2e_steps / (float(t2.item()) * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
               "note": "uint8 frames + boxes + labels H2D from pinned memory (double-buffered, copy stream overlaps compute), "
                       "loss + confusion counts D2H, every step"}

    # ---- roofline of the dominant kernel (K1) ----
    peak, peak_src = measured_peak_hbm()
    elem = 4 if out_dtype == torch.float32 else 2
    k1_bytes = k1_algorithmic_bytes(boxes_np, wl.out_size, wl.out_size, elem, wl.mode, wl.out_size)
    achieved = k1_bytes / (k1_ms * 1e-3) / 1e9
    traffic = None
    tp = ROOT / "profiles" / "k1_traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get(f"{wl.name}.{args.out_dtype}")
        except Exception:
            traffic = None
    roofline = {"kernel": "k1_crop_resize_normalize", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": k1_bytes, "launch_ms": k1_ms,
                "share_of_step": k1_ms * args.steps / dt_ms}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        procs = os.cpu_count() or 1
        sample = min(n, 4096)
        rate, times, nfr = cpu_path_rate(wl, boxes_np, fidx_np, sample, procs, steps=1, warmup=1)
        rate1, _, _ = cpu_path_rate(wl, boxes_np, fidx_np, min(n, 256), 1, steps=1, warmup=0)
        cpu = {"value": rate, "unit": UNIT, "cores": procs, "kind": "port",
               "sample": f"{sample} crops ({nfr} frames) x 1 step of {wl.name} after 1 warm-up; oracle port = cv2 "
                         f"{__import__('cv2').__version__} resize/normalize in {procs} worker processes + torch CPU fp32 "
                         f"heads/loss/backward/logger stats; albumentations absent (cv2+numpy stand in)",
               "single_core_value": rate1}

    if world > 1:
        comm.shutdown()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": dt_ms_max / args.steps, "serial_value": serial_value, "serial_ms_per_step": serial_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/int32->" + args.out_dtype + (", heads bf16 x bf16 -> f32 (tcgen05)" if args.emb_dtype == "bf16" else ", heads f32"),
        "data": "synthetic", "config": workload_config(wl, args.out_dtype, world, hp.transport),
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": sampler.summary(),
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    from nkb_classification_b200.synthetic import DEFAULT_WORKLOAD, WORKLOADS

    wl = WORKLOADS[args.workload or DEFAULT_WORKLOAD]
    if args.resize is not None and args.resize != wl.mode:
        import dataclasses
        wl = dataclasses.replace(wl, mode=args.resize, name=f"{wl.name}.{args.resize}")
    if args.train_aug:
        import dataclasses
        wl = dataclasses.replace(wl, name=f"{wl.name}.train_aug", train_aug=True)
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
