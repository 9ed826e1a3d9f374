#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native CTUNet hot path.

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): one CTUNet(depth 101, patch_frame 8)
TRAINING STEP on synthetic 96^3 patches, batch 2 per GPU, bf16 activations / fp32 weight gradients: forward, the
reference's five-head Dice-CE loss (trainer_CTUNet.py:92-103), backward through every kernel of this repo, gradient
all-reduce over NCCL when N > 1 (data parallel, main_CTUNet.py:187-189) and the AdamW update (main_CTUNet.py:190-193).
`value` = 96^3 patches/s over all ranks (weak scaling: 2 patches per GPU per step).  The second half of the metric —
whole-volume sliding-window inference volumes/s (configs[2]: 512x512x256, 96^3 windows, overlap 0.5, Gaussian blend,
windows sharded over the ranks with one NCCL all-reduce per head) — is measured in the same run and reported under
"sliding_window".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-sliding-window]

Prints ONE JSON line (rank 0).  `e2e` is the same training step driven through the public API from pinned HOST
buffers (H2D of the patches and labels inside the timed region, loss read back to the host every step).
`--impl reference` times the reference's own CPU path (fp32 oracle restatement of the modules + loss + torch autograd)
on the host cores, one 96^3 patch per step.
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

KW = dict(in_channels=1, dim_conv_stem=64, out_channels=14, model_depth=101, img_size=(96, 96), frames=96, patch_frame=8)
BATCH_PER_GPU = 2
FWD_BWD_GFLOP_PER_PATCH = 10258.03  # SURVEY 8d [probe]
FWD_GFLOP_PER_PATCH = 3423.64
WORKLOAD = "CTUNet(101,pf8) training step (fwd + 5-head Dice-CE + bwd + AdamW), 96^3 patches, batch 2/GPU, bf16"
VOLUME, ROI, OVERLAP, SW_BATCH, NUM_WINDOWS = (512, 512, 256), (96, 96, 96), 0.5, 4, 500


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
    """CUDA events around every launch of the dominant kernel inside the timed region: the tcgen05 implicit-GEMM
    3x3x3 convolution 64 -> 64 channels at 96^3 (forward AND input-gradient launches use the same kernel; 4 forward +
    4 dgrad launches per step, 1.57 TFLOP each at batch 2)."""

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
        traffic, tsrc = None, None
        prof = os.path.join(ROOT, "profiles", "r01_ncu_full_conv3_halo64.json")
        if os.path.exists(prof):  # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu capture
            traffic = json.load(open(prof)).get("traffic_bytes_per_launch")
            tsrc = "profiles/r01_ncu_full_conv3_halo64.json (ncu --set full, same kernel / shape / batch)"
        return {"bound": "tensor", "kernel": "conv3_halo_kernel<64,1,4,3>: tcgen05 conv3x3x3 64->64 @96^3 (forward + dgrad launches)",
                "achieved": round(ach, 1), "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                "frac": round(ach / peaks["tf_sustained"], 4), "traffic": traffic, "traffic_source": tsrc,
                "algorithmic_bytes_per_launch": 2 * 2 * 96 ** 3 * 64 * 2 + 27 * 64 * 64 * 2, "launches_timed": len(ms),
                "avg_launch_ms": round(avg, 4), "flops_per_launch": self.flops, "peak_source": peaks["src"] + ", sustained"}


# --------------------------------------------------------------------------------------------- reference / CPU arm
def _oracle_train_step_seconds(threads: int, warmup: int, steps: int):
    """The reference's CPU path for one training step on ONE 96^3 patch: oracle CTUNet fp32 forward + 5-head Dice-CE
    (scipy zoom labels) + torch autograd backward, all host threads."""
    from oracle import train_oracle as T
    from hybrid_ctunet_b200.networks.hybrid_CTUNet import CTUNet
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    sd = {k: v.detach() for k, v in CTUNet(**KW).state_dict().items()}
    torch.manual_seed(1)
    x = torch.rand(1, 1, 96, 96, 96)
    y = torch.randint(0, 14, (1, 1, 96, 96, 96)).float()
    for _ in range(warmup):
        T.ctunet_train_step(sd, x, y)
    t0 = time.perf_counter()
    for _ in range(steps):
        T.ctunet_train_step(sd, x, y)
    return (time.perf_counter() - t0) / steps


_OUT = None


def _emit(text: str):
    out = _OUT if _OUT is not None else sys.stdout
    out.write(text + "\n")
    out.flush()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    wu = min(args.warmup, 1)
    dt = _oracle_train_step_seconds(threads, wu, args.steps)
    value = 1.0 / dt
    sample = f"1 patch (batch 1) per step: oracle CTUNet fp32 fwd + 5-head Dice-CE + autograd bwd on the host cores ({wu} warm-up)"
    _emit(json.dumps({
        "impl": "reference", "metric": "train_patches_per_s", "value": value, "unit": "patches/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "patches/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--optimizer", default="ours", choices=["ours", "torch"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sliding-window", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: anything a library writes to file descriptor 1 during the run (NCCL's
    # version banner, for one) is sent to stderr instead
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from hybrid_ctunet_b200 import lib
    from hybrid_ctunet_b200.dp import GradientAllReduce
    from hybrid_ctunet_b200.losses import DiceCELoss, ctunet_loss
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
    warmup = max(args.warmup, 3)
    B = BATCH_PER_GPU

    torch.manual_seed(0)
    model = CTUNet(**KW).to(dev).train()
    loss_func = DiceCELoss(to_onehot_y=True, softmax=True, squared_pred=True, smooth_nr=0.0, smooth_dr=1e-6)
    if args.optimizer == "ours":   # ctu_adamw_step: one multi-tensor launch (hybrid_ctunet_b200/optim.py), same update
        from hybrid_ctunet_b200.optim import AdamW as CtuAdamW
        opt = CtuAdamW(model.parameters(), lr=1e-4, weight_decay=1e-5)
    else:
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5, fused=True)
    reducer = GradientAllReduce(model.parameters(), group) if world > 1 else None
    torch.manual_seed(1 + rank)
    host_x = torch.rand(B, 1, 96, 96, 96).pin_memory()
    host_y = torch.randint(0, 14, (B, 1, 96, 96, 96)).float().pin_memory()
    x, y = host_x.to(dev), host_y.to(dev)

    # forward + loss + backward are captured once and replayed as ONE CUDA graph (hybrid_ctunet_b200.training): the
    # ~1,900 launches of a step otherwise leave the GPU idle ~13 % of the time; all-reduce and AdamW stay eager
    from hybrid_ctunet_b200.training import GraphedTrainStep
    n_cap = lib.launch_count()
    graphed = GraphedTrainStep(model, lambda lg, t: ctunet_loss(lg, t, loss_func), x, y, warmup=1)
    launches_per_step = None

    def train_step(xd, yd):
        loss = graphed(xd, yd)
        if reducer is not None:
            reducer.reduce()
        opt.step()
        return loss

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # kernels of this library inside one replay = launches recorded during the capture (warm-up ran once before it)
    launches_per_step = (lib.launch_count() - n_cap) / 2.0
    for _ in range(warmup):
        train_step(x, y)
    sync()

    # ---------------- device-resident timing (value) with the dominant-kernel probe and clock sampling
    with ClockSampler(local) as clocks:
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            loss = train_step(x, y)
        e1.record()
        sync()
    launches = launches_per_step
    ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    loss_value = float(loss.detach())

    # dominant kernel: CUDA events around its launches in eager steps of the SAME workload (events cannot be recorded
    # inside a graph replay); 8 launches per step (4 forward + 4 input-gradient)
    static_grads = [p.grad for p in model.parameters()]  # the graph's outputs: put back after the eager probe steps
    probe = ConvProbe()
    probe.install()
    for _ in range(2):
        for p in model.parameters():
            p.grad = None
        ctunet_loss(model(x), y, loss_func).backward()
    sync()
    probe.remove()
    for p, g in zip(model.parameters(), static_grads):
        p.grad = g

    # ---------------- end to end through the public API: pinned host patches + labels in, loss out, every step
    sync()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        _ = train_step(host_x, host_y).item()  # pinned host -> the graph's static inputs (H2D), replay, loss D2H
    f1.record()
    sync()
    e2e_ms = max_over_ranks(f0.elapsed_time(f1) / args.steps)
    peak_mem = torch.cuda.max_memory_allocated() / 2 ** 30

    # ---------------- second half of the metric: sliding-window inference of one 512x512x256 volume
    sw = None
    if not args.no_sliding_window:
        model.eval()
        for p in model.parameters():
            p.grad = None
        del opt, graphed
        torch.cuda.empty_cache()
        torch.manual_seed(2)
        host_vol = torch.rand(1, 1, *VOLUME).pin_memory()
        vol = host_vol.to(dev)

        model.enable_cuda_graph()  # one graph replay per network call of 4 windows

        def infer(v):
            with torch.no_grad():
                return sliding_window_inference(v, ROI, SW_BATCH, model, overlap=OVERLAP, mode="gaussian", shard_group=group)

        out = infer(vol)
        del out
        sync()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        out = infer(vol)
        g1.record()
        sync()
        sw_ms = max_over_ranks(g0.elapsed_time(g1))
        del out
        sw = {"metric": "sliding_window_volumes_per_s", "value": 1e3 / sw_ms, "unit": "volumes/s", "ms_per_volume": sw_ms,
              "scaling": "strong", "windows": NUM_WINDOWS, "warmup": 1, "steps": 1,
              "workload": "1x1x512x512x256, roi 96^3, overlap 0.5, gaussian, sw_batch 4, 2 heads",
              "tflops_per_gpu": NUM_WINDOWS * FWD_GFLOP_PER_PATCH / sw_ms / world,
              "parallelism": f"windows sharded over {world} GPU(s), 1 NCCL all-reduce per head" if world > 1 else "single GPU"}

    if rank == 0:
        patches = B * world
        line = {
            "metric": "train_patches_per_s", "value": patches / (ms * 1e-3), "unit": "patches/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": patches,
                       "l2": "per-layer activations (0.2-0.9 GB) and the 0.7 GB of weights exceed the 126 MB L2",
                       "parallelism": f"dp{world}: one flat fp32 gradient all-reduce (NCCL) per step" if world > 1 else "single GPU",
                       "optimizer": ("hybrid_ctunet_b200.optim.AdamW (ctu_adamw_step, one launch)" if args.optimizer == "ours"
                                     else "torch.optim.AdamW(fused=True)") + ", inside the timed step",
                       "launch_mode": "forward + loss + backward replayed as one CUDA graph; all-reduce and optimizer eager",
                       "loss": "DiceCE x5 (torch ops on device) + device-side label gather"},
            "tflops_per_gpu": B * FWD_BWD_GFLOP_PER_PATCH / ms,
            "loss": loss_value, "peak_mem_gb": round(peak_mem, 2),
            "e2e": {"value": patches / (e2e_ms * 1e-3), "unit": "patches/s",
                    "h2d_bytes_per_step": (host_x.numel() + host_y.numel()) * 4, "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms},
            "gpu_launches": int(launches * args.steps),
            "gpu_launches_per_step": launches,
            "clocks": clocks.summary(),
            "roofline": probe.result(peaks),
            "sliding_window": sw,
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            dt = _oracle_train_step_seconds(threads, 0, 1)
            line["cpu_baseline"] = {"value": 1.0 / dt, "unit": "patches/s", "cores": threads, "kind": "port",
                                    "sample": "1 patch (batch 1), one step, no warm-up: oracle CTUNet fp32 fwd + 5-head Dice-CE + autograd bwd on the host cores"}
        _emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
