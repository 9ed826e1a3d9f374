#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native CTUNet hot path.

Workload (BASELINE.json configs[2], the largest single-GPU configuration whose every kernel is ours end to end):
sliding-window inference of one synthetic 1x1x512x512x256 volume with CTUNet(depth 101, patch_frame 8), 96^3
windows, overlap 0.5, Gaussian blend, sw_batch 4, both heads blended.  One step = one whole volume (500 windows =
125 network calls + 1000 blend launches + normalise).  With --gpus N the 500 windows are split into N contiguous
chunks (one process per GPU) and the two fp32 accumulators are summed with one NCCL all-reduce each: total work is
fixed, so scaling is "strong".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `value` is volumes/s with the volume resident in HBM; `e2e` is the same metric
through the public API with the volume in pinned host memory (H2D inside the timed region) and both blended
logit volumes read back to the host.  `--impl reference` times the reference's CPU path (the fp32 oracle
restatement, oracle/) on the host cores on a bounded sample and extrapolates.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VOLUME = (512, 512, 256)
ROI = (96, 96, 96)
OVERLAP = 0.5
SW_BATCH = 4
NUM_WINDOWS = 500
KW = dict(in_channels=1, dim_conv_stem=64, out_channels=14, model_depth=101, img_size=(96, 96), frames=96, patch_frame=8)
FWD_GFLOP_PER_WINDOW = 3423.64  # SURVEY 8d [probe]: CTUNet forward, one 96^3 patch
WORKLOAD = "sliding_window 1x1x512x512x256, CTUNet(101,pf8), roi 96^3, overlap 0.5, gaussian, sw_batch 4, 2 heads"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p.get("hbm_gbs", 6650.0), tf_burst=p.get("bf16_tflops", 1590.0),
                    tf_sustained=p.get("bf16_tflops_sustained", 1400.0), src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _loop(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


class ConvProbe:
    """CUDA events around every launch of the dominant kernel (3x3x3 conv 64->64 at 96^3) inside the timed region."""

    def __init__(self):
        self.pairs, self.flops = [], 0.0

    def install(self):
        from hybrid_ctunet_b200 import ops
        self._orig = ops.gemm
        probe = self

        def wrapped(a, w, out, *, dims, **kw):
            hit = (w.ksize == 3 and w.a_c == 64 and w.n_real == 64 and tuple(dims[:3]) == (96, 96, 96))
            if not hit:
                return probe._orig(a, w, out, dims=dims, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = probe._orig(a, w, out, dims=dims, **kw)
            e1.record()
            probe.pairs.append((e0, e1))
            probe.flops = 2.0 * dims[3] * 96 ** 3 * 64 * 27 * 64
            return r
        ops.gemm = wrapped

    def remove(self):
        from hybrid_ctunet_b200 import ops
        ops.gemm = self._orig

    def result(self, peaks):
        if not self.pairs:
            return None
        ms = [a.elapsed_time(b) for a, b in self.pairs]
        avg = sum(ms) / len(ms)
        ach = self.flops / (avg * 1e-3) / 1e12
        return {"bound": "tensor", "kernel": "umma_gemm_kernel<64,4> as conv3x3x3 64->64 @96^3 x4 windows",
                "achieved": round(ach, 1), "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                "frac": round(ach / peaks["tf_sustained"], 4), "traffic": None, "launches_timed": len(ms),
                "avg_launch_ms": round(avg, 4), "flops_per_launch": self.flops, "peak_source": peaks["src"] + ", sustained"}


def cpu_baseline_sample(threads: int):
    """Reference CPU path on a bounded sample: ONE 96^3 window through the fp32 oracle (CTUNet forward) + the oracle
    blend of that window, on the host cores; extrapolated to the 500 windows of the volume."""
    from oracle import ctunet_oracle as O
    from oracle import sliding_window_oracle as SO
    from hybrid_ctunet_b200.networks.hybrid_CTUNet import CTUNet
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    sd = {k: v.detach() for k, v in CTUNet(**KW).state_dict().items()}
    torch.manual_seed(2)
    x = torch.rand(1, 1, 96, 96, 144)
    t0 = time.perf_counter()
    with torch.no_grad():
        O.ctunet_forward(sd, x[..., :96], 101, 8)  # warm-up of the thread pool / allocator on a full window
    warm = time.perf_counter() - t0
    t0 = time.perf_counter()
    with torch.no_grad():
        SO.sliding_window_inference(x[..., :96], ROI, SW_BATCH, lambda w: O.ctunet_forward(sd, w, 101, 8), overlap=OVERLAP,
                                    mode="gaussian", two_heads=True)
    per_window = time.perf_counter() - t0
    return per_window, warm


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle restatement, kind "port"),
    each step = one window (forward + blend) on all host threads, extrapolated to volumes/s."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    from oracle import ctunet_oracle as O
    from oracle import sliding_window_oracle as SO
    from hybrid_ctunet_b200.networks.hybrid_CTUNet import CTUNet
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    sd = {k: v.detach() for k, v in CTUNet(**KW).state_dict().items()}
    torch.manual_seed(2)
    x = torch.rand(1, 1, 96, 96, 96)
    pred = lambda w: O.ctunet_forward(sd, w, 101, 8)
    step = lambda: SO.sliding_window_inference(x, ROI, SW_BATCH, pred, overlap=OVERLAP, mode="gaussian", two_heads=True)
    with torch.no_grad():
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = (time.perf_counter() - t0) / args.steps
    value = 1.0 / (dt * NUM_WINDOWS)
    sample = "1 of 500 windows per step (CTUNet fp32 forward + Gaussian blend), extrapolated x500"
    print(json.dumps({
        "impl": "reference", "metric": "sliding_window_volumes_per_s", "value": value, "unit": "volumes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 * NUM_WINDOWS,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "volumes/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from hybrid_ctunet_b200 import lib
    from hybrid_ctunet_b200.networks.hybrid_CTUNet import CTUNet
    from hybrid_ctunet_b200.trainer_CTUNet import sliding_window_inference

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    lib.require_device()
    peaks = _peaks()

    torch.manual_seed(0)
    model = CTUNet(**KW).to(dev).eval()
    torch.manual_seed(2)
    host_vol = torch.rand(1, 1, *VOLUME).pin_memory()
    vol = host_vol.to(dev, non_blocking=True)

    def step(v):
        with torch.no_grad():
            return sliding_window_inference(v, ROI, SW_BATCH, model, overlap=OVERLAP, mode="gaussian", shard_group=group)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        out = step(vol)
    del out
    sync()

    # ---------------- device-resident timing (value) with the dominant-kernel probe and clock sampling
    probe = ConvProbe()
    probe.install()
    n0 = lib.launch_count()
    with ClockSampler(local) as clocks:
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out = step(vol)
        e1.record()
        sync()
    probe.remove()
    launches = lib.launch_count() - n0
    ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    del out

    # ---------------- end to end through the public API: pinned host volume in, blended logits out to the host
    host_out = [torch.empty((1, 14) + VOLUME, dtype=torch.float32).pin_memory() for _ in range(2)] if rank == 0 else None
    sync()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        v = host_vol.to(dev, non_blocking=True)
        o = step(v)
        if rank == 0:
            host_out[0].copy_(o[0], non_blocking=True)
            host_out[1].copy_(o[1], non_blocking=True)
        del o
    f1.record()
    sync()
    t = torch.tensor([f0.elapsed_time(f1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())

    if rank == 0:
        line = {
            "metric": "sliding_window_volumes_per_s", "value": 1e3 / ms, "unit": "volumes/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "windows": NUM_WINDOWS, "l2": "activations (>=0.9 GB per layer) exceed the 126 MB L2",
                       "parallelism": f"windows sharded over {world} GPU(s), 1 all-reduce per head" if world > 1 else "single GPU",
                       "launch_mode": "eager launches (CUDA-graph replay is available via enable_cuda_graph)"},
            "tflops_per_gpu": NUM_WINDOWS * FWD_GFLOP_PER_WINDOW / ms / world,
            "e2e": {"value": 1e3 / e2e_ms, "unit": "volumes/s", "h2d_bytes_per_step": host_vol.numel() * 4,
                    "d2h_bytes_per_step": 2 * 14 * VOLUME[0] * VOLUME[1] * VOLUME[2] * 4, "ms_per_step": e2e_ms},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
            "roofline": probe.result(peaks),
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            per_window, _ = cpu_baseline_sample(threads)
            line["cpu_baseline"] = {"value": 1.0 / (per_window * NUM_WINDOWS), "unit": "volumes/s", "cores": threads,
                                    "kind": "port",
                                    "sample": "1 of 500 windows (oracle CTUNet fp32 forward + Gaussian blend) on the host cores, extrapolated x500"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
