#!/usr/bin/env python
"""bench.py -- the hot path's headline metric on B200: crops/sec through
(preprocess + heads + loss + metric), device-resident (`value`), end to end
from pinned host buffers (`e2e`), the dominant kernel against the HBM roofline
(`roofline`) and the reference's CPU path timed on the same box (`cpu_baseline`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

N > 1 is launched by torchrun (one rank per GPU); per-GPU work is fixed (weak
scaling): every rank owns its own frames, the only exchange is the sum of the head
gradients and confusion counts (K4' over NVLink peer memory, or NCCL).  The same
line also carries `strong` (N > 1: a FIXED global batch sharded over the ranks,
with the one-GPU time of that batch measured in the same run), `parity_check`
(sharded == single-GPU), and at N = 1 `variants` (the other named configurations)
and `api` (crops/s through get_dataset -> train_epoch / inference).
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
    ap.add_argument("--no-graph", action="store_true", help="skip the CUDA-graph replay of the step")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the fixed-global-batch (strong scaling) leg")
    ap.add_argument("--no-variants", action="store_true", help="N = 1: skip the other named configurations")
    ap.add_argument("--no-api", action="store_true", help="N = 1: skip the drop-in API leg (get_dataset -> train_epoch, inference)")
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
        "launches_per_step": "2: K1, then the fused heads step (forward + loss + K3 + dW/db + exchange + finalize)",
        "timing": "value = the fastest of { serial: eager launches on one stream; graph: the same step replayed from a CUDA "
                  "graph; overlap: a CUDA graph whose two branches are K1 of batch i+1 and the heads step of batch i (no "
                  "dependency without the backbone); pdl / pdl_graph: ONE stream, the heads step of batch i then K1 of "
                  "batch i+1 as its programmatic dependent (nkbk_k1_overlap_previous), eager / one graph per step; "
                  "streams / streams_multi: eager launches on two streams, K1 of batch i+1 on one and the heads step of "
                  "batch i on a higher-priority one, as one launch / as its separate kernels (nkbk_heads_one_launch(0)) } -- "
                  "`mode` says which; serial_value = eager; every figure is the "
                  "median of 3 repetitions of exactly K steps",
    }


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
REPS = 3   # every timed region = `reps` repetitions of exactly K steps, each bracketed by barrier + sync; the median counts


def _plan_for(wl, train_aug):
    from nkb_classification_b200 import transforms as T
    return T.compile_pipeline(_pipeline_ops(wl, train_aug))


def _pipeline_ops(wl, train_aug):
    from nkb_classification_b200 import transforms as T
    geo = ([T.Resize(wl.out_size, wl.out_size)] if wl.mode == "stretch" else
           [T.LongestMaxSize(wl.out_size), T.PadIfNeeded(wl.out_size, wl.out_size, border_mode=T.BORDER_CONSTANT, value=0)])
    aug_ops = []
    if train_aug:
        aug_ops = [T.HorizontalFlip(p=0.5), T.VerticalFlip(p=0.5),
                   T.RandomBrightnessContrast(brightness_limit=(-0.2, 0.2), contrast_limit=(0.1, -0.5), p=0.5),
                   T.HueSaturationValue(hue_shift_limit=0, sat_shift_limit=10, val_shift_limit=50, p=0.5),
                   T.CoarseDropout(max_holes=4, min_holes=1, max_height=0.2, min_height=0.05, max_width=0.2,
                                   min_width=0.05, fill_value=[0, 0.5, 1], p=0.5)]
    return geo + aug_ops + [T.Normalize(MEAN, STD), T.ToTensorV2()]


class Leg:
    """One workload instance on one rank: synthetic inputs resident in HBM + the HotPath that steps over them."""

    def __init__(self, wl, dev, comm, out_dtype, emb_dtype, transport, train_aug=False, seed_rank=0, frame_range=None):
        import torch
        from nkb_classification_b200 import hotpath
        from nkb_classification_b200.synthetic import synth_boxes
        self.wl, self.dev = wl, dev
        self.plan = _plan_for(wl, train_aug)
        self.hp = hotpath.HotPath(self.plan, wl.classes, wl.emb_dim, wl.loss, wl.gamma, device=dev, comm=comm,
                                  out_dtype=out_dtype, transport=transport)
        g = torch.Generator(device=dev).manual_seed(1234 + seed_rank)
        frames = torch.randint(0, 256, (wl.frames, wl.frame_h, wl.frame_w, 3), dtype=torch.uint8, device=dev, generator=g)
        boxes_np, fidx_np = synth_boxes(wl, seed=4321 + seed_rank)
        gc = torch.Generator().manual_seed(7 + seed_rank)
        n_all = len(fidx_np)
        emb = torch.randn(n_all, wl.emb_dim, generator=gc)
        labels = torch.stack([torch.randint(0, c, (n_all,), generator=gc) for c in wl.classes], 1).contiguous()
        if frame_range is not None:   # strong scaling: this rank's contiguous shard of the frames (and of their crops)
            fb, fe = frame_range
            mask = (fidx_np >= fb) & (fidx_np < fe)
            frames = frames[fb:fe].contiguous()
            boxes_np, fidx_np = boxes_np[mask], (fidx_np[mask] - fb).astype(np.int32)
            m = torch.from_numpy(mask)
            emb, labels = emb[m].contiguous(), labels[m].contiguous()
        self.frames, self.boxes_np, self.fidx_np = frames, boxes_np, fidx_np
        self.n = len(fidx_np)
        self.boxes = torch.from_numpy(np.ascontiguousarray(boxes_np)).to(dev)
        self.fidx = torch.from_numpy(np.ascontiguousarray(fidx_np)).to(dev)
        self.emb = emb.to(dev).to(emb_dtype).contiguous()
        self.labels_h = labels
        self.labels = labels.to(dev)
        Ws, bs = make_heads(wl)
        self.W_cat, self.b_cat = torch.cat(Ws).contiguous().to(dev), torch.cat(bs).contiguous().to(dev)
        self.aug = None
        if train_aug:
            import random as _random
            self.aug = self.plan.draw(self.n, _random.Random(99 + seed_rank))

    def k1(self, frames=None, boxes=None, fidx=None, overlap_previous=False):
        return self.hp.preprocess(self.frames if frames is None else frames, self.boxes if boxes is None else boxes,
                                  self.fidx if fidx is None else fidx, None, self.aug, overlap_previous=overlap_previous)

    def pdl_step(self):
        """Heads step of batch i, then K1 of batch i+1 as its programmatic dependent: ONE stream, K1 starts on the SMs
        the heads step leaves free as soon as that kernel is resident."""
        b = self.heads()
        self.hp.mark_heads_done()
        self.k1(overlap_previous=True)
        return b

    def heads(self, labels=None):
        return self.hp.heads_step(self.emb, self.W_cat, self.b_cat, self.labels if labels is None else labels, train=True)

    def step(self):
        self.k1()
        return self.heads()


def _timed(fn_k_steps, barrier, world, dev, reps=REPS):
    """`fn_k_steps()` enqueues exactly K steps.  Returns the median over `reps` of (max over ranks of the device time)."""
    import torch
    import torch.distributed as dist
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        fn_k_steps()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out.append(float(t.item()))
    return float(np.median(out)), out


def _capture(leg, dev, overlap=False, pdl=False):
    """One step as a CUDA graph; None when capture is not possible.  overlap = False: K1 -> fused heads step in
    sequence.  overlap = True: the two as parallel branches of the graph -- preprocessing of batch i+1 next to the
    heads / loss / metric / exchange of batch i, which have no dependency (the backbone that sits between them is out of
    scope here); a training loop overlaps them the same way with the loader's K1 on its own stream."""
    import torch
    try:
        s = torch.cuda.Stream(device=dev)
        s2 = torch.cuda.Stream(device=dev, priority=-1)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            leg.step()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            if pdl:
                leg.pdl_step()
            elif overlap:
                s2.wait_stream(s)
                with torch.cuda.stream(s2):
                    leg.heads()
                leg.k1()
                s.wait_stream(s2)
            else:
                leg.step()
        torch.cuda.synchronize()
        return g
    except Exception as e:   # pragma: no cover
        sys.stderr.write(f"[bench] CUDA graph capture failed ({e!r}); graph numbers omitted\n")
        return None


def _measure_leg(leg, steps, barrier, world, dev, use_graph=True):
    """serial (eager launches, one stream), graph (the same step replayed from a CUDA graph) and per-kernel K1 time."""
    import torch
    from nkb_classification_b200 import _lib, ops
    for _ in range(3):
        leg.step()
    barrier()
    k1_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]

    def eager():
        for i in range(steps):
            k1_ev[i][0].record()
            leg.k1()
            k1_ev[i][1].record()
            leg.heads()

    l0 = _lib.launch_count()
    serial_ms, serial_all = _timed(eager, barrier, world, dev)
    launches = (_lib.launch_count() - l0) // REPS
    k1_ms = float(np.mean([a.elapsed_time(b) for a, b in k1_ev]))
    path = ops.heads_last_path()
    out = {"serial_ms": serial_ms / steps, "serial_all_ms": [t / steps for t in serial_all], "k1_ms": k1_ms,
           "launches": int(launches), "heads_path": {_lib.PATH_FUSED: "k2_fused_step (1 launch)",
                                                     _lib.PATH_TC_FWD: "k2_tc_heads_forward + k2_heads_dw + finalize",
                                                     _lib.PATH_FFMA_FWD: "k2_heads_forward_v3 + k2_heads_dw + finalize"}.get(path, str(path))}
    # pdl: eager, one stream -- heads step of batch i, then K1 of batch i+1 as its programmatic dependent
    def eager_pdl():
        for _ in range(steps):
            leg.pdl_step()
    try:
        eager_pdl()
        pdl_ms, pdl_all = _timed(eager_pdl, barrier, world, dev)
        out.update({"pdl_ms": pdl_ms / steps, "pdl_all_ms": [t / steps for t in pdl_all]})
    except Exception as e:   # pragma: no cover
        sys.stderr.write(f"[bench] programmatic dependent launch failed ({e!r}); pdl numbers omitted\n")
    leg.hp._heads_done = None
    # streams: eager launches on two streams -- K1 of batch i+1 on one, the heads step of batch i on a higher-priority one
    # (round 1's pipelined form); streams_multi: the same with the heads step as its separate kernels (forward, dW,
    # exchange + finalize), whose small CTAs slot in between K1's instead of waiting for whole SMs
    cur = torch.cuda.current_stream(dev)
    s_pre, s_heads = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev, priority=-1)

    def two_streams():
        s_pre.wait_stream(cur)
        s_heads.wait_stream(cur)
        for _ in range(steps):
            with torch.cuda.stream(s_pre):
                leg.k1()
            with torch.cuda.stream(s_heads):
                leg.heads()
        cur.wait_stream(s_pre)
        cur.wait_stream(s_heads)

    for key, multi in (("streams", False), ("streams_multi", True)):
        try:
            leg.hp.one_launch = not multi           # (nkbk_heads_one_launch)
            two_streams()
            t_ms, t_all = _timed(two_streams, barrier, world, dev)
            out.update({f"{key}_ms": t_ms / steps, f"{key}_all_ms": [t / steps for t in t_all]})
        except Exception as e:   # pragma: no cover
            sys.stderr.write(f"[bench] two-stream mode failed ({e!r}); {key} numbers omitted\n")
        finally:
            leg.hp.one_launch = True
    leg.step()
    for key, overlap, pdl in (("graph", False, False), ("overlap", True, False), ("pdl_graph", False, True)):
        g = _capture(leg, dev, overlap, pdl) if use_graph else None
        if g is not None:
            def replay():
                for _ in range(steps):
                    g.replay()
            replay()
            g_ms, g_all = _timed(replay, barrier, world, dev)
            out.update({f"{key}_ms": g_ms / steps, f"{key}_all_ms": [t / steps for t in g_all]})
            del g
    leg.hp._heads_done = None
    out["best_ms"], out["best_mode"] = min((out[k], k[:-3]) for k in ("serial_ms", "graph_ms", "overlap_ms", "pdl_ms",
                                                                      "pdl_graph_ms", "streams_ms", "streams_multi_ms")
                                         if k in out)
    return out


def _unique_source_bytes(boxes_np, fidx_np, wl):
    """Bytes of the UNION of the boxes of every frame (what HBM has to deliver at least once when L2 serves the overlaps)."""
    total = 0
    for f in np.unique(fidx_np):
        m = np.zeros((wl.frame_h, wl.frame_w), dtype=bool)
        for x0, y0, x1, y1 in boxes_np[fidx_np == f]:
            m[y0:y1, x0:x1] = True
        total += int(m.sum()) * 3
    return total


def _parity_check(leg_sharded, wl, dev, rank, world, out_dtype, emb_dtype):
    """N > 1: one sharded heads step must equal rank 0's single-GPU step over the WHOLE global batch (loss and mean
    gradients to 1e-5, confusion counts exactly), and K1 of the shard must equal the same rows of rank 0's K1 output."""
    import torch
    import torch.distributed as dist
    from nkb_classification_b200.parallel import Communicator
    hp = leg_sharded.hp
    hp.reset_confusion()
    img = leg_sharded.k1().clone()
    bufs = leg_sharded.heads()
    torch.cuda.synchronize()
    mine = torch.cat([bufs.reduce_buf[: hp.NC * hp.D + hp.NC].double(), bufs.loss.double()])
    cm_mine = hp.cm.clone()
    hp.reset_confusion()
    ref = torch.zeros_like(mine)
    cm_ref = torch.zeros_like(cm_mine)
    ok_img = 1.0
    if rank == 0:
        full = Leg(wl, dev, Communicator(), out_dtype, emb_dtype, "peer")
        img_full = full.k1()
        b = full.heads()
        torch.cuda.synchronize()
        ref = torch.cat([b.reduce_buf[: hp.NC * hp.D + hp.NC].double(), b.loss.double()])
        cm_ref = full.hp.cm.clone()
        if leg_sharded.aug is None:   # (augmentation draws depend on the batch length: compared without them only)
            ok_img = float(torch.equal(img_full[: img.shape[0]], img))     # rank 0 owns the first frames
        del full, img_full
    dist.broadcast(ref, 0)
    dist.broadcast(cm_ref, 0)
    err = float(((mine - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item())
    ok = float(err <= 1e-5 and bool(torch.equal(cm_mine, cm_ref))) * ok_img
    t = torch.tensor([ok, -err], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return {"status": "ok" if t[0].item() == 1.0 else "MISMATCH", "max_rel_err": float(-t[1].item()),
            "checked": "sharded loss / dW / db vs rank 0's single-GPU step over the global batch (1e-5), confusion counts "
                       "exact, K1 output of rank 0's shard bit-exact"}


class SubsampleStub:
    """Stands in for the out-of-scope backbone in the API bench: [B,3,S,S] -> [B,D] = the first D values of every image
    (one copy kernel, 8 KB read per crop), so that its cost does not mask what is being measured -- the loader, K1, the
    fused heads step, autograd, the optimizer and the logger as `train_epoch` / `inference` drive them."""

    def __new__(cls, D):
        import torch

        class _Stub(torch.nn.Module):
            def __init__(self, D):
                super().__init__()
                self.num_features = D

            def forward(self, x):
                f = x.reshape(x.shape[0], -1)
                if f.shape[1] < self.num_features:
                    f = f.repeat(1, (self.num_features + f.shape[1] - 1) // f.shape[1])
                f = f[:, : self.num_features]
                return f.contiguous() if f.dtype == torch.float32 else f.float()     # one copy kernel

        return _Stub(D)


def api_bench(wl, dev, n_batches):
    """crops/s through the drop-in API on one GPU: get_dataset -> train_epoch (engine.py:20-85 of the reference) and
    inference() (inference.py:42-70), wall clock around whole epochs of `n_batches` batches, cache-cold then cache-warm."""
    import tempfile
    from types import SimpleNamespace

    import torch

    from nkb_classification_b200 import dataset as D, engine, inference as INF, logging as LG, losses, model as M
    from nkb_classification_b200 import transforms as T
    from nkb_classification_b200.synthetic import synth_boxes

    rng = np.random.default_rng(1234)
    frames = list(rng.integers(0, 256, (wl.frames, wl.frame_h, wl.frame_w, 3), dtype=np.uint8))
    boxes, fidx = synth_boxes(wl)
    n1 = len(fidx)
    boxes_all, fidx_all = np.tile(boxes, (n_batches, 1)), np.tile(fidx, n_batches)
    labels = rng.integers(0, wl.classes[0], len(fidx_all))
    classes = [f"c{i}" for i in range(wl.classes[0])]
    ds = D.InMemoryFrames(frames, fidx_all, labels, boxes=boxes_all, classes=classes)
    pipe = [T.Resize(wl.out_size, wl.out_size), T.Normalize(MEAN, STD), T.ToTensorV2()]
    data = {"type": "InMemoryFrames", "dataset": ds, "batch_size": n1, "shuffle": True, "num_workers": 4, "prefetch": 2,
            "device": str(dev), "sort_within_batch": True}
    loader = D.get_dataset(data, pipe)
    model = M.get_model({"task": "single", "model": SubsampleStub(wl.emb_dim), "pretrained": False, "backbone_dropout": 0.0,
                         "classifier_dropout": 0.0, "classifier_initialization": "kaiming_normal_"}, classes, dev)
    cfg = SimpleNamespace(task="single", target_column="label", enable_mixed_presicion=False, log_gradients=False,
                          disable_tqdm=True, criterion={"task": "single", "type": wl.loss})
    criterion = losses.get_loss(cfg.criterion, dev)
    opt = torch.optim.Adam(model.classifier.parameters(), lr=1e-4, fused=True)
    scaler = torch.amp.GradScaler("cuda", enabled=False)
    logger = LG.BaseLogger(cfg, classes)
    out = {"batches_per_epoch": n_batches, "crops_per_batch": n1, "backbone": "SubsampleStub (first D values of every image, one copy kernel, no parameters)",
           "loader": "get_dataset(InMemoryFrames, shuffle=True, num_workers=4, prefetch=2, frame cache auto, "
                     "sort_within_batch); warm epochs are planned at once (resident epoch: one metadata copy, one K1 "
                     "launch per batch)", "optimizer": "torch.optim.Adam(fused=True) over the head parameters"}

    def timed_epoch(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    class _Clocked:              # the loader, with a clock on the batches it hands out (loop time without the epoch end)
        def __init__(self, inner):
            self.inner, self.dataset, self.t = inner, inner.dataset, []

        def __len__(self):
            return len(self.inner)

        def __iter__(self):
            self.t, self.ev = [], []
            for b in self.inner:
                self.t.append(time.perf_counter())
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                self.ev.append(e)
                yield b
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.ev.append(e)
            torch.cuda.synchronize()
            self.t.append(time.perf_counter())

    clocked = _Clocked(loader)

    def train():
        return engine.train_epoch(model, clocked, opt, None, scaler, criterion, dev, cfg, logger)

    n = len(fidx_all)
    loader.reset_stats()
    t_cold = timed_epoch(train)
    cold = dict(loader.stats)
    loader.reset_stats()
    t_warm, loop_ms = float("inf"), float("inf")
    for _ in range(2):
        t_warm = min(t_warm, timed_epoch(train))
        loop_ms = min(loop_ms, 1e3 * (clocked.t[-1] - clocked.t[1]) / (n_batches - 1))   # batches 2..last, device drained
        dev_ms = clocked.ev[1].elapsed_time(clocked.ev[-1]) / (n_batches - 1)               # the same span on the device
    warm = dict(loader.stats)
    cfg.epoch_results_numpy = True      # epoch results as numpy arrays instead of the reference's Python lists
    t_warm_np = min(timed_epoch(train) for _ in range(2))
    cfg.epoch_results_numpy = False
    out["train_epoch"] = {"cold_crops_per_s": n / t_cold, "warm_crops_per_s": n / t_warm, "warm_ms_per_batch": 1e3 * t_warm / n_batches,
                          "warm_numpy_results_crops_per_s": n / t_warm_np,
                          "resident_epochs": warm.get("resident_epochs", 0),
                          "loop_ms_per_batch": loop_ms, "loop_crops_per_s": n1 / (loop_ms * 1e-3),
                          "loop_device_ms_per_batch": dev_ms,
                          "cold_h2d_bytes_per_crop": cold["h2d_bytes"] / n, "warm_h2d_bytes_per_crop": warm["h2d_bytes"] / (2 * n),
                          "cold_decodes": cold["decodes"], "warm_decodes": warm["decodes"],
                          "note": "wall clock around the whole epoch, epoch results (one D2H per quantity, K5 ROC-AUC "
                                  "counts) included; cold = first epoch (frames decoded + uploaded once), warm = frames "
                                  "resident in the device frame cache"}
    # the same loop with the reference's TRAIN pipeline (configs/singletask_config.py:162-201): per-sample augmentation
    # parameters drawn on the host per batch (vectorised), hue / sat / val tables built on the device, K1 with the
    # augmentations fused
    try:
        import dataclasses
        ds_a = D.InMemoryFrames(frames, fidx_all, labels, boxes=boxes_all, classes=classes)
        loader_a = D.get_dataset(dict(data, dataset=ds_a), _pipeline_ops(dataclasses.replace(wl, mode="stretch"), True))
        clocked_a = _Clocked(loader_a)

        def train_a():
            return engine.train_epoch(model, clocked_a, opt, None, scaler, criterion, dev, cfg, logger)

        cfg.epoch_results_numpy = True
        timed_epoch(train_a)                                   # cold: fills the frame cache
        t_a, loop_a = float("inf"), float("inf")
        for _ in range(2):
            t_a = min(t_a, timed_epoch(train_a))
            loop_a = min(loop_a, 1e3 * (clocked_a.t[-1] - clocked_a.t[1]) / (n_batches - 1))
        cfg.epoch_results_numpy = False
        out["train_epoch_train_pipeline"] = {
            "loop_ms_per_batch": loop_a, "loop_crops_per_s": n1 / (loop_a * 1e-3), "warm_numpy_results_crops_per_s": n / t_a,
            "resident_epochs": loader_a.stats.get("resident_epochs", 0),
            "note": "train_epoch over get_dataset(..., train pipeline: flips + brightness/contrast + HueSaturationValue + "
                    "CoarseDropout fused into K1); parameters drawn per batch on the host (numpy, all samples at once)"}
        del loader_a, ds_a, clocked_a
    except Exception as e:   # pragma: no cover
        out["train_epoch_train_pipeline"] = {"error": repr(e)}

    # inference(): same frames through the inference entry point (predictions -> CSV)
    icfg = SimpleNamespace(task="single", target_column="label", enable_mixed_presicion=False, disable_tqdm=True)

    class _PathsLoader:          # inference() expects (img, paths) batches; the labels slot carries the paths
        def __init__(self, inner):
            self.inner, self.dataset = inner, inner.dataset

        def __iter__(self):
            for img, tgt in self.inner:
                yield img, ["x"] * img.shape[0]

        def __len__(self):
            return len(self.inner)

    with tempfile.TemporaryDirectory() as td:
        t_inf = min(timed_epoch(lambda: INF.inference(model, _PathsLoader(loader), classes, td, dev, icfg)) for _ in range(2))
    out["inference"] = {"crops_per_s": n / t_inf, "ms_per_batch": 1e3 * t_inf / n_batches,
                        "note": "inference(): K1 + forward-only heads with fused argmax per batch, one CSV write at the end "
                                "(included)"}
    return out


def run_b200(args, wl):
    import dataclasses
    import torch
    import torch.distributed as dist

    from nkb_classification_b200 import _lib
    from nkb_classification_b200.parallel import Communicator, shard_range
    from nkb_classification_b200.synthetic import WORKLOADS, k1_algorithmic_bytes

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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out_dtype = torch.float32 if args.out_dtype == "f32" else torch.bfloat16
    emb_dtype = torch.float32 if args.emb_dtype == "f32" else torch.bfloat16
    steps = args.steps
    sampler = ClockSampler(local_rank)

    # ---- weak leg (the headline): every rank owns wl.frames frames; inputs resident in HBM ----
    leg = Leg(wl, dev, comm, out_dtype, emb_dtype, args.allreduce, train_aug=args.train_aug, seed_rank=rank)
    n = leg.n
    for _ in range(max(args.warmup, 3)):
        leg.step()
    barrier()
    sampler.start()
    m = _measure_leg(leg, steps, barrier, world, dev, use_graph=not args.no_graph)
    sampler.stop()
    best_ms = m["best_ms"]
    value = world * n / (best_ms * 1e-3)
    serial_value = world * n / (m["serial_ms"] * 1e-3)

    # ---- strong leg (N > 1): the SAME global batch as one GPU's step, frames sharded contiguously over the ranks ----
    strong, parity = None, None
    if world > 1 and not args.no_strong:
        fb, fe = shard_range(wl.frames, rank, world)
        sleg = Leg(wl, dev, comm, out_dtype, emb_dtype, args.allreduce, train_aug=args.train_aug, seed_rank=0,
                   frame_range=(fb, fe))
        parity = _parity_check(sleg, wl, dev, rank, world, out_dtype, emb_dtype)
        sm_ = _measure_leg(sleg, steps, barrier, world, dev, use_graph=not args.no_graph)
        # the 1-GPU time of the same global batch, measured in this very run on rank 0 while the others wait
        t1 = torch.zeros(1, dtype=torch.float64, device=dev)
        if rank == 0:
            full = Leg(wl, dev, Communicator(), out_dtype, emb_dtype, "peer", train_aug=args.train_aug, seed_rank=0)
            fm = _measure_leg(full, steps, lambda: torch.cuda.synchronize(), 1, dev, use_graph=not args.no_graph)
            t1[0] = fm["best_ms"]
            del full
        dist.broadcast(t1, 0)
        s_ms = sm_["best_ms"]
        gb = wl.crops
        strong = {"global_batch": gb, "crops_per_rank": sleg.n, "value": gb / (s_ms * 1e-3), "ms_per_step": s_ms,
                  "serial_ms_per_step": sm_["serial_ms"], "graph_ms_per_step": sm_.get("graph_ms"),
                  "overlap_ms_per_step": sm_.get("overlap_ms"), "pdl_ms_per_step": sm_.get("pdl_ms"),
                  "pdl_graph_ms_per_step": sm_.get("pdl_graph_ms"), "streams_ms_per_step": sm_.get("streams_ms"),
                  "streams_multi_ms_per_step": sm_.get("streams_multi_ms"), "mode": sm_["best_mode"],
                  "k1_ms": sm_["k1_ms"], "n1_ms_per_step": float(t1.item()),
                  "speedup_vs_n1": float(t1.item()) / s_ms, "efficiency_vs_n1": float(t1.item()) / s_ms / world,
                  "heads_path": sm_["heads_path"], "gpu_launches_per_step": sm_["launches"] // steps,
                  "note": "fixed global batch = one GPU's step; frames sharded contiguously (parallel.shard_range); "
                          "n1 = the same batch on rank 0 alone, timed in this run"}
        del sleg

    # ---- end to end from pinned host memory through the public API ----
    e2e = None
    if not args.no_e2e:
        cur = torch.cuda.current_stream(dev)
        hp = leg.hp
        frames_h = torch.empty(leg.frames.shape, dtype=torch.uint8).pin_memory()
        frames_h.copy_(leg.frames)
        boxes_h = torch.from_numpy(np.ascontiguousarray(leg.boxes_np)).pin_memory()
        fidx_h = torch.from_numpy(np.ascontiguousarray(leg.fidx_np)).pin_memory()
        labels_p = leg.labels_h.pin_memory()
        loss_h = torch.empty(hp.T + 1, dtype=torch.float32).pin_memory()
        cm_h = torch.empty(hp.cm.numel(), dtype=torch.int64).pin_memory()
        # double-buffered ingest: the H2D copy of step i+1 (copy stream) overlaps K1..K4 of step i (compute stream)
        nbuf = 2
        frames_d = [torch.empty_like(leg.frames) for _ in range(nbuf)]
        boxes_d = [torch.empty_like(leg.boxes) for _ in range(nbuf)]
        fidx_d = [torch.empty_like(leg.fidx) for _ in range(nbuf)]
        labels_d = [torch.empty_like(leg.labels) for _ in range(nbuf)]
        s_copy = torch.cuda.Stream(device=dev)
        copied = [torch.cuda.Event() for _ in range(nbuf)]
        consumed = [torch.cuda.Event() for _ in range(nbuf)]
        e2e_steps = max(3, min(steps, 30))

        def e2e_run(k=e2e_steps, resident=False):
            s_copy.wait_stream(cur)
            for b in range(nbuf):
                consumed[b].record(cur)
            for i in range(k):
                b = i % nbuf
                with torch.cuda.stream(s_copy):
                    s_copy.wait_event(consumed[b])          # the buffer's previous contents have been preprocessed
                    if not resident:                        # (resident: the decoded frames sit in the device frame cache)
                        frames_d[b].copy_(frames_h, non_blocking=True)
                    boxes_d[b].copy_(boxes_h, non_blocking=True)
                    fidx_d[b].copy_(fidx_h, non_blocking=True)
                    labels_d[b].copy_(labels_p, non_blocking=True)
                    copied[b].record(s_copy)
                cur.wait_event(copied[b])
                leg.k1(frames_d[b], boxes_d[b], fidx_d[b])
                consumed[b].record(cur)
                bufs = leg.heads(labels_d[b])
                loss_h.copy_(bufs.loss, non_blocking=True)
                cm_h.copy_(hp.cm, non_blocking=True)
            cur.wait_stream(s_copy)

        e2e_run(3)
        t_ms, _ = _timed(e2e_run, barrier, world, dev, reps=2)
        h2d = frames_h.numel() + boxes_h.numel() * 4 + fidx_h.numel() * 4 + labels_p.numel() * 8
        d2h = loss_h.numel() * 4 + cm_h.numel() * 8
        e2e = {"value": world * n * e2e_steps / (t_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
               "h2d_gbs_per_rank": h2d * e2e_steps / (t_ms * 1e-3) / 1e9,
               "note": "uint8 frames + boxes + labels H2D from pinned memory (double-buffered, copy stream overlaps compute), "
                       "loss + confusion counts D2H, every step"}
        # the same step from the second epoch on: decoded frames resident in HBM (DeviceFrameCache), so a step uploads
        # boxes + frame indices + labels only and still copies loss + confusion counts back
        e2e_run(3, True)
        r_ms, _ = _timed(lambda: e2e_run(e2e_steps, True), barrier, world, dev, reps=2)
        h2d_r = boxes_h.numel() * 4 + fidx_h.numel() * 4 + labels_p.numel() * 8
        e2e["resident"] = {"value": world * n * e2e_steps / (r_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d_r),
                           "d2h_bytes_per_step": int(d2h), "fraction_of_value": None,
                           "note": "cache-warm epoch: frames already in the device frame cache (epoch >= 2 of a dataset "
                                   "that fits HBM), per-step H2D = boxes + frame indices + labels"}
        ceil = ROOT / "profiles" / "h2d_ceiling.json"
        if ceil.exists():
            try:
                c = json.loads(ceil.read_text())
                e2e["h2d_ceiling_gbs"] = c.get(f"n{world}")
                e2e["h2d_ceiling_source"] = c.get("_source")
            except Exception:
                pass
        del frames_d, frames_h

    # ---- roofline of the dominant kernel (K1): algorithmic / DRAM-level / compulsory bytes over the same launch time ----
    peak, peak_src = measured_peak_hbm()
    elem = 4 if out_dtype == torch.float32 else 2
    k1_ms = m["k1_ms"]
    k1_bytes = k1_algorithmic_bytes(leg.boxes_np, wl.out_size, wl.out_size, elem, wl.mode, wl.out_size)
    achieved = k1_bytes / (k1_ms * 1e-3) / 1e9
    traffic = None
    tp = ROOT / "profiles" / "k1_traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get(f"{wl.name}.{args.out_dtype}")
        except Exception:
            traffic = None
    uniq = _unique_source_bytes(leg.boxes_np, leg.fidx_np, wl) + n * 3 * wl.out_size * wl.out_size * elem
    roofline = {"kernel": "k1_crop_resize_normalize", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": k1_bytes, "launch_ms": k1_ms,
                "share_of_step": k1_ms / m["serial_ms"],
                # boxes of one frame overlap and are served from L2: what HBM actually moved, and its lower bound
                "dram_frac": None if traffic is None else traffic / (k1_ms * 1e-3) / 1e9 / peak,
                "unique_bytes_per_launch": uniq, "unique_bytes_frac": uniq / (k1_ms * 1e-3) / 1e9 / peak,
                "frac_note": "frac = SURVEY 8(d) algorithmic bytes (source bytes counted per crop); dram_frac = DRAM bytes "
                             "of the committed ncu capture (profiles/k1_traffic.json); unique_bytes_frac = union of the "
                             "boxes per frame + output, the compulsory traffic"}

    # ---- the other named configurations on the same clock (N = 1): heads path, K1 fraction, crops/s ----
    variants = None
    if world == 1 and not args.no_variants:
        variants = {}
        vlist = [("cfg4_vit_5heads.bf16", WORKLOADS["cfg4_vit_5heads"], torch.bfloat16, torch.bfloat16, False),
                 ("cfg1_single_224", WORKLOADS["cfg1_single_224"], torch.float32, torch.float32, False),
                 ("cfg2_multitask_256", WORKLOADS["cfg2_multitask_256"], torch.float32, torch.float32, False),
                 ("cfg3_1080p_20", WORKLOADS["cfg3_1080p_20"], torch.float32, torch.float32, False),
                 ("cfg5_1080p_64x64.letterbox", dataclasses.replace(wl, mode="letterbox"), out_dtype, emb_dtype, False),
                 ("cfg5_1080p_64x64.train_aug", wl, out_dtype, emb_dtype, True),
                 ("cfg5_1080p_64x64.bf16", wl, torch.bfloat16, torch.bfloat16, False)]
        for name, vwl, od, ed, aug in vlist:
            try:
                vleg = Leg(vwl, dev, Communicator(), od, ed, "peer", train_aug=aug)
                vm = _measure_leg(vleg, min(steps, 20), barrier, 1, dev, use_graph=not args.no_graph)
                vms = vm["best_ms"]
                ve = 4 if od == torch.float32 else 2
                vb = k1_algorithmic_bytes(vleg.boxes_np, vwl.out_size, vwl.out_size, ve, vwl.mode, vwl.out_size)
                variants[name] = {"value": vleg.n / (vms * 1e-3), "ms_per_step": vms, "serial_ms_per_step": vm["serial_ms"],
                                  "graph_ms_per_step": vm.get("graph_ms"), "overlap_ms_per_step": vm.get("overlap_ms"),
                                  "pdl_ms_per_step": vm.get("pdl_ms"),
                                  "mode": vm["best_mode"], "k1_ms": vm["k1_ms"],
                                  "k1_frac": vb / (vm["k1_ms"] * 1e-3) / 1e9 / peak, "heads_path": vm["heads_path"],
                                  "crops_per_step": vleg.n, "out": str(od).split(".")[-1], "emb": str(ed).split(".")[-1]}
                del vleg
            except Exception as e:   # a variant must never take the headline down with it
                variants[name] = {"error": repr(e)}

    # ---- the drop-in API on the same clock (N = 1): get_dataset -> train_epoch / inference, wall clock ----
    api = None
    if world == 1 and not args.no_api and not wl.whole_image:
        try:
            api = api_bench(wl, dev, 24)
            api["train_epoch"]["loop_fraction_of_serial_value"] = api["train_epoch"]["loop_crops_per_s"] / serial_value
            api["train_epoch"]["epoch_fraction_of_serial_value"] = api["train_epoch"]["warm_crops_per_s"] / serial_value
        except Exception as e:
            api = {"error": repr(e)}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        procs = os.cpu_count() or 1
        sample = min(n, 4096)
        rate, times, nfr = cpu_path_rate(wl, leg.boxes_np, leg.fidx_np, sample, procs, steps=1, warmup=1)
        rate1, _, _ = cpu_path_rate(wl, leg.boxes_np, leg.fidx_np, min(n, 256), 1, steps=1, warmup=0)
        cpu = {"value": rate, "unit": UNIT, "cores": procs, "kind": "port",
               "sample": f"{sample} crops ({nfr} frames) x 1 step of {wl.name} after 1 warm-up; oracle port = cv2 "
                         f"{__import__('cv2').__version__} resize/normalize in {procs} worker processes + torch CPU fp32 "
                         f"heads/loss/backward/logger stats; albumentations absent (cv2+numpy stand in)",
               "single_core_value": rate1}

    transport = leg.hp.transport
    if world > 1:
        comm.shutdown()
        dist.destroy_process_group()
    if rank != 0:
        return
    if e2e is not None and e2e.get("resident"):
        e2e["resident"]["fraction_of_value"] = e2e["resident"]["value"] / value
    config = workload_config(wl, args.out_dtype, world, transport)
    if m["best_mode"] == "streams_multi":
        config["launches_per_step"] = ("4: K1 on one stream; on the other the heads step as its separate kernels (forward + loss + "
                                       "K3, dW/db, exchange + finalize) -- nkbk_heads_one_launch(0)")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
        "ms_per_step": best_ms, "serial_value": serial_value, "serial_ms_per_step": m["serial_ms"],
        "graph_ms_per_step": m.get("graph_ms"), "overlap_ms_per_step": m.get("overlap_ms"),
        "pdl_ms_per_step": m.get("pdl_ms"), "pdl_graph_ms_per_step": m.get("pdl_graph_ms"),
        "streams_ms_per_step": m.get("streams_ms"), "streams_multi_ms_per_step": m.get("streams_multi_ms"),
        "mode": m["best_mode"],
        "reps": REPS, "rep_ms_per_step": {"serial": m["serial_all_ms"], "graph": m.get("graph_all_ms"),
                                          "overlap": m.get("overlap_all_ms"), "pdl": m.get("pdl_all_ms"),
                                          "pdl_graph": m.get("pdl_graph_all_ms"), "streams": m.get("streams_all_ms"),
                                          "streams_multi": m.get("streams_multi_all_ms")},
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/int32->" + args.out_dtype + (", heads bf16 emb x f32 W -> f32" if args.emb_dtype == "bf16" else ", heads f32"),
        "data": "synthetic", "config": config,
        "heads_path": m["heads_path"], "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": int(m["launches"]), "strong": strong, "parity_check": parity, "variants": variants, "api": api,
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
