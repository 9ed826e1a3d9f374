#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native CTUNet hot path.

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): one CTUNet(depth 101, patch_frame 8)
TRAINING STEP on synthetic 96^3 patches, batch 2 per GPU, bf16 activations / fp32 weight gradients: forward, the
reference's five-head Dice-CE loss (trainer_CTUNet.py:92-103), backward through every kernel of this repo, gradient
all-reduce over NCCL when N > 1 (data parallel, main_CTUNet.py:187-189) and the AdamW update (main_CTUNet.py:190-193).
`value` = 96^3 patches/s over all ranks (weak scaling: 2 patches per GPU per step).

Measured in the same run and reported as sub-objects of the ONE JSON line (rank 0):
  sliding_window  configs[2]: 512x512x256 volume, 96^3 windows, overlap 0.5, Gaussian blend, windows sharded over the
                  ranks — device-resident volumes/s and `e2e` (pinned host volume in, uint8 masks out)
  config4         configs[3]: data-parallel step at batch 4 per GPU (run when N = 8, or with --config4)
  hybrid          configs[4]: Hybrid-CTUNet mask-complementation inference (CTUNet head 0 @ overlap 0.5 + TUNet @ 0.7 +
                  ensemble), host volume in, uint8 mask out inside the timed region
  roofline        the kernel class that takes the largest share of the step, with `classes`: EVERY kernel class of the
                  step (conv fwd/dgrad, wgrad, HBM-bound GEMMs, InstanceNorm, ...) with its measured rate against
                  the measured peak and its share of the step — CUDA events around each launch of one eager step of the
                  same workload
  library_gpu     the reference's torch ops (oracle restatement = the reference modules' op sequence) on the SAME
                  B200 through cuDNN / cuBLAS: the reference's own recipe (fp16 autocast, cudnn.benchmark,
                  trainer_CTUNet.py:90, main_CTUNet.py:120) and the tuned one (bf16 autocast + channels_last_3d) — the
                  real bar; N = 1 only
  cpu_baseline    the oracle training step of one patch on the host cores (a stated baseline, not the target)

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-sliding-window] [--no-hybrid]
                    [--no-library-gpu] [--no-cpu-baseline] [--config4] [--batch B]

`e2e` is the same training step driven through the public API from pinned HOST buffers (H2D of the patches and labels
inside the timed region, loss read back to the host every step).  `--impl reference` times the reference's own CPU path
(fp32 oracle restatement of the modules + loss + torch autograd) on the host cores, one 96^3 patch per step.
"""
from __future__ import annotations

import argparse
import gc
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
TKW = {k: v for k, v in KW.items() if k != "model_depth"}
BATCH_PER_GPU = 2
FWD_BWD_GFLOP_PER_PATCH = 10258.03  # SURVEY 8d [probe]
FWD_GFLOP_PER_PATCH = 3423.64
TUNET_FWD_GFLOP_PER_PATCH = 1166.41
WORKLOAD = "CTUNet(101,pf8) training step (fwd + 5-head Dice-CE + bwd + AdamW), 96^3 patches, batch 2/GPU, bf16"
VOLUME, ROI, OVERLAP, SW_BATCH, NUM_WINDOWS = (512, 512, 256), (96, 96, 96), 0.5, 4, 500
TUNET_WINDOWS = 1792


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


# --------------------------------------------------------------------------------------------- per-class roofline
class KernelClassProbe:
    """CUDA events around EVERY launch that goes through hybrid_ctunet_b200.ops in one eager training step of the
    benchmark workload (events cannot be recorded inside a graph replay; the eager step launches the same kernels with
    the same arguments, serialised on one stream).  Each launch is booked into a kernel class with its ALGORITHMIC
    work: FLOPs with the true channel counts (2 per MAC) for the tensor-bound contractions, bytes of every distinct
    tensor argument (each read or written once) for the HBM-bound kernels.  A contraction is graded on the roofline
    its arithmetic intensity puts it under (ridge = measured sustained bf16 peak / measured HBM bandwidth)."""

    ELEMENTWISE = {"in_apply": "InstanceNorm (stats / apply+LeakyReLU(+res) / backward)", "in_backward": None,
                   "in_stats": None,
                   "layernorm": "LayerNorm fwd / bwd (+ patchify)", "layernorm_backward": None, "patchify_ln": None,
                   "patchify_ln_backward": None,
                   "gelu": "GELU / cross-weight fusion (pwa) elementwise", "gelu_backward": None, "pwa_fuse": None,
                   "pwa_fuse_backward": None,
                   "attention": "attention fwd / bwd (mma.sync flash kernels)", "attention_backward": None,
                   "colsum": "glue: column sums, accumulate, casts, layout (space_to_depth, cf_to_cl, subsample, im2col)",
                   "accumulate": None, "cast_f32_bf16": None, "space_to_depth": None, "cf_to_cl": None, "subsample": None,
                   "head_backward": None,
                   "subsample_backward": None, "im2col_cin1": None, "conv_cin1": None}

    def __init__(self, peaks):
        self.peaks = peaks
        self.ridge = peaks["tf_sustained"] * 1e12 / (peaks["hbm"] * 1e9)
        self.rec = []
        self._orig = {}
        cls = None
        self.cls_of = {}
        for k, v in self.ELEMENTWISE.items():
            cls = v or cls
            self.cls_of[k] = cls

    @staticmethod
    def _bytes(args, kwargs):
        seen, tot = set(), 0
        for t in list(args) + list(kwargs.values()):
            w = getattr(t, "w", None)
            if isinstance(w, torch.Tensor):   # PackedWeight
                t = w
            if isinstance(t, torch.Tensor) and t.data_ptr() not in seen:
                seen.add(t.data_ptr())
                tot += t.numel() * t.element_size()
        return tot

    def install(self):
        from hybrid_ctunet_b200 import ops
        probe = self

        def wrap(name, fn):
            def f(*args, **kwargs):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = fn(*args, **kwargs)
                e1.record()
                by = probe._bytes(args, kwargs)
                fl, shape = 0.0, None
                if name == "gemm":
                    a, w = args[0], args[1]
                    dims = kwargs["dims"]
                    rows = dims[0] * dims[1] * dims[2] * dims[3]
                    fl = rows * (w.alg_flops_per_row or 2.0 * w.ksize ** 3 * w.a_c * w.n_real)
                    feat = ("+stats" if kwargs.get("stats") is not None else "") + ("+res" if kwargs.get("residual") is not None else "") + \
                           ("+gelu" if kwargs.get("act") else "") + ("+gelu_bwd" if kwargs.get("gelu_bwd_of") is not None else "") + \
                           ("+bias" if w.bias is not None else "") + ("+convt" if w.convt else "") + \
                           {0: "", 1: " f32", 2: " f32cf"}[kwargs.get("out_mode", 0)]
                    shape = f"k{w.ksize} {kwargs.get('a_c') or w.a_c}->{w.n_real} rows {rows}{feat}"
                    kind = "conv3" if w.ksize == 3 else ("gemm_t" if fl / max(by, 1) >= probe.ridge else "gemm_h")
                elif name == "wgrad":
                    dims = kwargs["dims"]
                    rows = dims[0] * dims[1] * dims[2] * dims[3]
                    ks = kwargs.get("ksize", 1)
                    xc = kwargs.get("x_c") or args[0].shape[-1]
                    nn = kwargs.get("n") or args[1].shape[-1]
                    fl = rows * (kwargs.get("alg_flops_per_row") or 2.0 * ks ** 3 * xc * nn)
                    shape = f"k{ks} {xc}->{nn} rows {rows}"
                    kind = "wgrad3" if ks == 3 else ("wgrad_t" if fl / max(by, 1) >= probe.ridge else "wgrad_h")
                else:
                    kind = name
                    t0 = next((a for a in args if isinstance(a, torch.Tensor)), None)
                    shape = f"{name} {tuple(t0.shape) if t0 is not None else ''}"
                probe.rec.append((kind, shape, e0, e1, fl, by))
                return r
            return f
        for name in ["gemm", "wgrad"] + list(self.ELEMENTWISE):
            self._orig[name] = getattr(ops, name)
            setattr(ops, name, wrap(name, self._orig[name]))

    def remove(self):
        from hybrid_ctunet_b200 import ops
        for name, fn in self._orig.items():
            setattr(ops, name, fn)

    CONTRACTIONS = {
        "conv3": ("tcgen05 3x3x3 conv, forward + input-gradient launches (conv3_halo / umma_gemm kernels)", "tensor"),
        "wgrad3": ("tcgen05 3x3x3 weight gradient (wgrad_halo / umma_wgrad kernels)", "tensor"),
        "gemm_t": ("tcgen05 GEMM, tensor-bound shapes (AI >= ridge: ViT / window-attention projections, deep 1x1x1)", "tensor"),
        "wgrad_t": ("tcgen05 weight gradient of the tensor-bound GEMMs", "tensor"),
        "gemm_h": ("tcgen05 GEMM, HBM-bound shapes (AI < ridge: 1x1x1 convs + IN statistics, heads, 128/256-ch token GEMMs, "
                   "ConvT / pixel-shuffle)", "hbm"),
        "wgrad_h": ("tcgen05 weight gradient of the HBM-bound GEMMs", "hbm"),
    }

    def result(self, steps: int, graph_step_ms: float):
        agg = {}
        for kind, shape, ms, fl, by in self.rec:      # (ms already resolved from the event pairs)
            if kind in self.CONTRACTIONS:
                label, bound = self.CONTRACTIONS[kind]
            else:
                label, bound = self.cls_of[kind], "hbm"
            a = agg.setdefault(label, dict(bound=bound, ms=0.0, fl=0.0, by=0.0, n=0, shapes={}))
            a["ms"] += ms; a["fl"] += fl; a["by"] += by; a["n"] += 1
            if shape is not None:
                s = a["shapes"].setdefault(shape, [0.0, 0.0, 0.0, 0])
                s[0] += ms; s[1] += fl; s[2] += by; s[3] += 1
        total = sum(a["ms"] for a in agg.values())
        classes = []
        for label, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
            if a["bound"] == "tensor":
                ach, peak, unit = a["fl"] / (a["ms"] * 1e-3) / 1e12, self.peaks["tf_sustained"], "TFLOP/s"
            else:
                ach, peak, unit = a["by"] / (a["ms"] * 1e-3) / 1e9, self.peaks["hbm"], "GB/s"
            ent = {"class": label, "bound": a["bound"], "achieved": round(ach, 1), "peak": peak, "unit": unit,
                   "frac": round(ach / peak, 4), "ms_per_step": round(a["ms"] / steps, 3),
                   "share_of_step": round(a["ms"] / total, 4), "launches_per_step": a["n"] // steps}
            if a["shapes"] and label in [v[0] for v in self.CONTRACTIONS.values()]:
                top = sorted(a["shapes"].items(), key=lambda kv: -kv[1][0])[:3]
                ent["top_shapes"] = [
                    {"shape": k, "ms_per_step": round(v[0] / steps, 3), "launches_per_step": v[3] // steps,
                     "achieved": round((v[1] / 1e12 if a["bound"] == "tensor" else v[2] / 1e9) / (v[0] * 1e-3), 1)}
                    for k, v in top]
            classes.append(ent)
        flops = sum(a["fl"] for a in agg.values())
        if os.environ.get("CTU_BENCH_DUMP_SHAPES"):   # full per-shape table of every contraction class (profiles/)
            with open(os.environ["CTU_BENCH_DUMP_SHAPES"], "w") as fh:
                for label, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
                    fh.write(f"{a['ms'] / steps:8.3f} ms x{a['n'] // steps:4d}  {label}\n")
                    for k, v in sorted(a["shapes"].items(), key=lambda kv: -kv[1][0]):
                        fh.write(f"    {v[0] / steps:8.3f} ms x{v[3] // steps:3d} avg {1e3 * v[0] / v[3]:7.1f} us  "
                                 f"{v[1] / 1e9 / max(v[0], 1e-9):8.1f} TFLOP/s {v[2] / 1e6 / max(v[0], 1e-9):8.1f} GB/s  {k}\n")
        return classes, {"eager_kernel_ms_per_step": round(total / steps, 2), "graph_step_ms": round(graph_step_ms, 2),
                         "whole_step_tflops_vs_peak": round(flops / steps / (graph_step_ms * 1e-3) / 1e12 /
                                                            self.peaks["tf_sustained"], 4),
                         "ridge_flop_per_byte": round(self.ridge, 1)}


def _traffic_for(label: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the class's top kernel, from the committed
    `ncu --set full` capture (profiles/r02_ncu_traffic.json: {class label prefix: {...}}), else None."""
    path = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    if not os.path.exists(path):
        return None, None
    table = json.load(open(path))
    for key, ent in table.items():
        if label.startswith(key):
            return ent.get("traffic_bytes_per_launch"), ent.get("source")
    return None, None


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


def _config(world: int, batch: int, optimizer: str = "ours"):
    patches = batch * world
    return {"workload": WORKLOAD.replace("batch 2/GPU", f"batch {batch}/GPU"), "global_batch": patches,
            "l2": "per-layer activations (0.2-0.9 GB) and the 0.7 GB of weights exceed the 126 MB L2",
            "parallelism": (f"dp{world}: flat fp32 gradient all-reduce (NCCL, chunked, overlapped with the AdamW of the "
                            f"previous chunk) per step") if world > 1 else "single GPU",
            "optimizer": ("hybrid_ctunet_b200.optim.AdamW (ctu_adamw_step)" if optimizer == "ours"
                          else "torch.optim.AdamW(fused=True)") + ", inside the timed step",
            "launch_mode": "forward + loss + backward replayed as one CUDA graph; all-reduce and optimizer eager",
            "loss": "DiceCE x5 on the fused kernels (ctu_dice_ce_fwd / _bwd) + device-side label gather"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    dt = _oracle_train_step_seconds(threads, args.warmup, args.steps)
    value = 1.0 / dt
    sample = ("each step = ONE 96^3 patch (batch 1) of the workload: oracle CTUNet fp32 forward + 5-head Dice-CE + torch "
              f"autograd backward on {threads} host threads (the optimizer update, <1 % of a CPU step, is not run)")
    _emit(json.dumps({
        "impl": "reference", "metric": "train_patches_per_s", "value": value, "unit": "patches/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config(args.gpus, BATCH_PER_GPU),
        "cpu_baseline": {"value": value, "unit": "patches/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------------- library-GPU arm
def _library_gpu(dev, steps: int = 3, warmup: int = 2):
    """The reference's op sequence on THIS GPU through torch's libraries (cuDNN / cuBLAS): oracle restatement of the
    modules (bit-equal to the reference modules on CPU, tests/test_oracle_vs_reference.py) + the reference's loss with its
    two scipy zoom host round trips (trainer_CTUNet.py:93-94) + autograd + torch.optim.AdamW.  Two recipes: the
    reference's own (fp16 autocast + GradScaler, cudnn.benchmark, default memory format) and a tuned one (bf16 autocast,
    channels_last_3d).  None of this repo's kernels run here."""
    from oracle import ctunet_oracle as O
    from oracle import train_oracle as T
    from hybrid_ctunet_b200.networks.hybrid_CTUNet import CTUNet
    torch.backends.cudnn.benchmark = True          # main_CTUNet.py:120
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    B = BATCH_PER_GPU
    torch.manual_seed(0)
    base = {k: v.detach().to(dev) for k, v in CTUNet(**KW).state_dict().items()}
    torch.manual_seed(1)
    x = torch.rand(B, 1, 96, 96, 96, device=dev)
    y = torch.randint(0, 14, (B, 1, 96, 96, 96), device=dev).float()
    out = {}
    for name, dtype, cl in (("reference_recipe_fp16_autocast", torch.float16, False),
                            ("tuned_bf16_autocast_channels_last_3d", torch.bfloat16, True)):
        try:
            leaves = {}
            for k, v in base.items():
                w = v.clone()
                if cl and w.dim() == 5:
                    w = w.contiguous(memory_format=torch.channels_last_3d)
                leaves[k] = w.requires_grad_()
            xi = x.contiguous(memory_format=torch.channels_last_3d) if cl else x
            opt = torch.optim.AdamW(list(leaves.values()), lr=1e-4, weight_decay=1e-5)
            scaler = torch.amp.GradScaler("cuda", enabled=(dtype == torch.float16))

            def train_step():
                for p in leaves.values():
                    p.grad = None
                with torch.autocast("cuda", dtype=dtype):
                    loss = T.ctunet_train_loss(leaves, xi, y)
                scaler.scale(loss).backward()
                scaler.step(opt)
                scaler.update()
                return loss

            def infer_call(w4):
                with torch.no_grad(), torch.autocast("cuda", dtype=dtype):
                    return O.ctunet_forward(leaves, w4)

            for _ in range(warmup):
                train_step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                train_step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            w4 = torch.rand(SW_BATCH, 1, 96, 96, 96, device=dev)
            w4 = w4.contiguous(memory_format=torch.channels_last_3d) if cl else w4
            for _ in range(warmup):
                infer_call(w4)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(steps):
                infer_call(w4)
            e1.record()
            torch.cuda.synchronize()
            ims = e0.elapsed_time(e1) / steps
            out[name] = {"train_ms_per_step": round(ms, 2), "train_patches_per_s": round(B / (ms * 1e-3), 3),
                         "infer_ms_per_4_windows": round(ims, 2),
                         "sliding_window_s_per_volume_extrapolated": round(ims * 1e-3 * NUM_WINDOWS / SW_BATCH, 2)}
            del leaves, opt
        except Exception as exc:  # an out-of-memory or unsupported-kernel failure of the LIBRARY path is a result, not ours
            out[name] = {"error": f"{type(exc).__name__}: {str(exc)[:200]}"}
        gc.collect()
        torch.cuda.empty_cache()
    out["what"] = ("oracle restatement of the reference modules (same torch ops, cuDNN/cuBLAS) + reference loss with scipy "
                   f"zoom round trips + autograd + torch AdamW, batch {B}, cudnn.benchmark=True, {warmup} warm-up + {steps} "
                   "timed steps, CUDA events; inference = one 4-window CTUNet call, volume time = x 125 calls (blend excluded)")
    torch.backends.cudnn.benchmark = False
    return out


# --------------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--optimizer", default="ours", choices=["ours", "torch"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="patches per GPU of the headline step (configs[1]: 2)")
    ap.add_argument("--config4", action="store_true", help="also time the batch-4-per-GPU step (default: only when N = 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sliding-window", action="store_true")
    ap.add_argument("--no-hybrid", action="store_true")
    ap.add_argument("--no-library-gpu", action="store_true")
    ap.add_argument("--no-class-probe", action="store_true")
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
    from hybrid_ctunet_b200.ensemble import ensemble_masks, hybrid_ctunet_inference
    from hybrid_ctunet_b200.losses import DiceCELoss, ctunet_loss
    from hybrid_ctunet_b200.networks.hybrid_CTUNet import CTUNet, TUNet
    from hybrid_ctunet_b200.trainer_CTUNet import sliding_window_inference
    from hybrid_ctunet_b200.training import GraphedTrainStep

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

    torch.manual_seed(0)
    model = CTUNet(**KW).to(dev).train()
    loss_func = DiceCELoss(to_onehot_y=True, softmax=True, squared_pred=True, smooth_nr=0.0, smooth_dr=1e-6)

    def make_optimizer():
        if args.optimizer == "ours":   # ctu_adamw_step: one multi-tensor launch (hybrid_ctunet_b200/optim.py), same update
            from hybrid_ctunet_b200.optim import AdamW as CtuAdamW
            return CtuAdamW(model.parameters(), lr=1e-4, weight_decay=1e-5)
        return torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5, fused=True)

    dp_chunks = int(os.environ.get("CTU_DP_CHUNKS", "4"))   # 1: one all-reduce, then the optimizer (A/B comparisons)

    def timed_training(B: int, probe_classes: bool):
        """Graph capture + warm-up + K timed device-resident steps + K timed end-to-end steps at batch B per GPU."""
        opt = make_optimizer()
        reducer = GradientAllReduce(model.parameters(), group) if world > 1 else None
        torch.manual_seed(1 + rank)
        host_x = torch.rand(B, 1, 96, 96, 96).pin_memory()
        host_y = torch.randint(0, 14, (B, 1, 96, 96, 96)).float().pin_memory()
        x, y = host_x.to(dev), host_y.to(dev)
        # forward + loss + backward are captured once and replayed as ONE CUDA graph (hybrid_ctunet_b200.training): the
        # ~1,900 launches of a step otherwise leave the GPU idle ~13 % of the time; all-reduce and AdamW stay eager
        n_cap = lib.launch_count()
        graphed = GraphedTrainStep(model, lambda lg, t: ctunet_loss(lg, t, loss_func), x, y, warmup=1)
        # kernels of this library inside one replay = launches recorded during the capture (warm-up ran once before it)
        launches = (lib.launch_count() - n_cap) / 2.0

        def train_step(xd, yd):
            loss = graphed(xd, yd)
            if reducer is not None:
                reducer.reduce_and_step(opt, chunks=dp_chunks)
            else:
                opt.step()
            return loss

        for _ in range(warmup):
            train_step(x, y)
        sync()
        with ClockSampler(local) as clocks:
            sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                loss = train_step(x, y)
            e1.record()
            sync()
        ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
        loss_value = float(loss.detach())

        classes = summary = None
        if probe_classes:
            # per-class roofline: one eager step of the SAME workload with CUDA events around every launch
            static_grads = [p.grad for p in model.parameters()]  # the graph's outputs: put back after the eager steps
            probe = KernelClassProbe(peaks)
            passes = []
            for it in range(3):
                if it == 1:
                    probe.install()
                if it >= 1:
                    # park the GPU on a spin kernel (~0.15 s) while the host enqueues the step: the launches then run
                    # back to back and an event pair brackets kernel time only, not the host's launch latency
                    probe.rec = []
                    torch.cuda._sleep(int(3e8))
                for p in model.parameters():
                    p.grad = None
                ctunet_loss(model(x), y, loss_func).backward()
                if it >= 1:
                    sync()
                    passes.append([(k, sh, e0.elapsed_time(e1), fl, by) for k, sh, e0, e1, fl, by in probe.rec])
            sync()
            probe.remove()
            # two probed passes, per launch the smaller time: a host hiccup (allocator, GC) while the GPU has caught up
            # with the enqueueing thread would otherwise be booked on whatever launch it delayed
            if len(passes[0]) == len(passes[1]):
                probe.rec = [(a[0], a[1], min(a[2], b[2]), a[3], a[4]) for a, b in zip(passes[0], passes[1])]
            else:
                probe.rec = passes[1]
            classes, summary = probe.result(1, ms)
            for p, g in zip(model.parameters(), static_grads):
                p.grad = g

        # end to end through the public API: pinned host patches + labels in, loss out, every step
        sync()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            _ = train_step(host_x, host_y).item()  # pinned host -> the graph's static inputs (H2D), replay, loss D2H
        f1.record()
        sync()
        e2e_ms = max_over_ranks(f0.elapsed_time(f1) / args.steps)
        res = dict(ms=ms, e2e_ms=e2e_ms, loss=loss_value, launches=launches, clocks=clocks.summary(), classes=classes,
                   summary=summary, h2d=(host_x.numel() + host_y.numel()) * 4,
                   peak_mem=torch.cuda.max_memory_allocated() / 2 ** 30)
        for p in model.parameters():
            p.grad = None
        del graphed, opt, reducer
        gc.collect()
        torch.cuda.empty_cache()
        return res

    B = args.batch
    head = timed_training(B, probe_classes=not args.no_class_probe)
    cfg4 = None
    if args.config4 or (world == 8 and B != 4):
        r4 = timed_training(4, probe_classes=False)
        cfg4 = {"metric": "train_patches_per_s", "value": 4 * world / (r4["ms"] * 1e-3), "unit": "patches/s",
                "ms_per_step": r4["ms"], "batch_per_gpu": 4, "global_batch": 4 * world, "n_gpus": world,
                "steps": args.steps, "warmup": warmup, "tflops_per_gpu": 4 * FWD_BWD_GFLOP_PER_PATCH / r4["ms"],
                "e2e": {"value": 4 * world / (r4["e2e_ms"] * 1e-3), "unit": "patches/s", "h2d_bytes_per_step": r4["h2d"],
                        "d2h_bytes_per_step": 4, "ms_per_step": r4["e2e_ms"]},
                "peak_mem_gb": round(r4["peak_mem"], 2),
                "workload": "BASELINE configs[3]: data-parallel CTUNet training, batch 4/GPU, gradient all-reduce over NVLink"}

    # ---------------- second half of the metric: sliding-window inference of one 512x512x256 volume
    sw = hyb = None
    model.eval()
    if not args.no_sliding_window:
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
        # end to end: pinned host volume in (H2D), blended logits -> mask-complementation ensemble on the device
        # (test_CTUNet.py:236-241), three uint8 masks back to the host (D2H), all inside the timed region
        host_masks = torch.empty(3, *VOLUME, dtype=torch.uint8).pin_memory()
        sync()
        g0.record()
        v = host_vol.to(dev, non_blocking=True)
        o1, o2 = infer(v)
        mk = ensemble_masks(o1[0], o2[0])
        for i, k in enumerate(("ensemble", "head1", "head2")):
            host_masks[i].copy_(mk[k], non_blocking=True)
        g1.record()
        sync()
        sw_e2e_ms = max_over_ranks(g0.elapsed_time(g1))
        del o1, o2, mk, v
        sw = {"metric": "sliding_window_volumes_per_s", "value": 1e3 / sw_ms, "unit": "volumes/s", "ms_per_volume": sw_ms,
              "scaling": "strong", "windows": NUM_WINDOWS, "warmup": 1, "steps": 1,
              "workload": "1x1x512x512x256, roi 96^3, overlap 0.5, gaussian, sw_batch 4, 2 heads",
              "tflops_per_gpu": NUM_WINDOWS * FWD_GFLOP_PER_PATCH / sw_ms / world,
              "frac_of_bf16_peak": round(NUM_WINDOWS * FWD_GFLOP_PER_PATCH / sw_ms / world / peaks["tf_sustained"], 4),
              "e2e": {"value": 1e3 / sw_e2e_ms, "unit": "volumes/s", "ms_per_volume": sw_e2e_ms,
                      "h2d_bytes_per_step": host_vol.numel() * 4, "d2h_bytes_per_step": host_masks.numel(),
                      "what": "pinned host volume -> sliding window (2 heads) -> device ensemble -> 3 uint8 masks on the host"},
              "parallelism": (f"windows sharded over {world} GPU(s), slab-wise NCCL reduction of the overlapping accumulators"
                              if world > 1 else "single GPU")}

        # ---------------- config 5: Hybrid-CTUNet = CTUNet head 0 @0.5 + independently initialised TUNet @0.7 + ensemble
        if not args.no_hybrid:
            torch.manual_seed(7)
            tunet = TUNet(**TKW).to(dev).eval()
            tunet.enable_cuda_graph()
            small = vol[:, :, :96, :96, :192].contiguous()
            hybrid_ctunet_inference(small, model, tunet, ROI, SW_BATCH, shard_group=group)   # captures TUNet's graph
            sync()
            host_mask = torch.empty(*VOLUME, dtype=torch.uint8).pin_memory()
            g0.record()
            v = host_vol.to(dev, non_blocking=True)
            res = hybrid_ctunet_inference(v, model, tunet, ROI, SW_BATCH, shard_group=group)
            host_mask.copy_(res["ensemble"], non_blocking=True)
            g1.record()
            sync()
            hy_ms = max_over_ranks(g0.elapsed_time(g1))
            tf = (NUM_WINDOWS * FWD_GFLOP_PER_PATCH + TUNET_WINDOWS * TUNET_FWD_GFLOP_PER_PATCH) / hy_ms / world
            hyb = {"metric": "hybrid_ensemble_volumes_per_s", "value": 1e3 / hy_ms, "unit": "volumes/s",
                   "ms_per_volume": hy_ms, "scaling": "strong", "n_gpus": world,
                   "windows": {"ctunet_overlap_0.5": NUM_WINDOWS, "tunet_overlap_0.7": TUNET_WINDOWS},
                   "tflops_per_gpu": tf, "h2d_bytes_per_step": host_vol.numel() * 4, "d2h_bytes_per_step": host_mask.numel(),
                   "workload": "BASELINE configs[4] (test_CTUNet_final.py:539-552): pinned host volume 1x1x512x512x256 in, "
                               "uint8 ensemble mask on the host out; warm-up = one 96x96x192 volume"}
            del tunet, res, v, small
        del vol

    line = None
    if rank == 0:
        patches = B * world
        ms = head["ms"]
        classes = head["classes"]
        roofline = None
        if classes:
            for c in classes:   # DRAM traffic of the class's representative kernel from the committed ncu --set full capture
                tb, src = _traffic_for(c["class"])
                if tb is not None:
                    c["traffic"], c["traffic_source"] = tb, src
            top = dict(classes[0])
            traffic, tsrc = _traffic_for(top["class"])
            roofline = {"bound": top["bound"], "kernel": top["class"], "achieved": top["achieved"], "peak": top["peak"],
                        "unit": top["unit"], "frac": top["frac"], "traffic": traffic, "traffic_source": tsrc,
                        "share_of_step": top["share_of_step"], "peak_source": peaks["src"] + ", sustained",
                        "how": "CUDA events around every launch of two eager steps of the benchmark workload, per launch the smaller time (same kernels "
                               "and arguments as the graph replay, serialised on one stream); achieved = algorithmic "
                               "FLOPs (true channel counts) or bytes (each tensor argument once) of the class / its "
                               "summed launch time; shares are of the summed launch time",
                        "classes": classes, **head["summary"]}
        line = {
            "metric": "train_patches_per_s", "value": patches / (ms * 1e-3), "unit": "patches/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": _config(world, B, args.optimizer),
            "tflops_per_gpu": B * FWD_BWD_GFLOP_PER_PATCH / ms,
            "frac_of_bf16_peak": round(B * FWD_BWD_GFLOP_PER_PATCH / ms / peaks["tf_sustained"], 4),
            "loss": head["loss"], "peak_mem_gb": round(head["peak_mem"], 2),
            "e2e": {"value": patches / (head["e2e_ms"] * 1e-3), "unit": "patches/s",
                    "h2d_bytes_per_step": head["h2d"], "d2h_bytes_per_step": 4, "ms_per_step": head["e2e_ms"]},
            "gpu_launches": int(head["launches"] * args.steps),
            "gpu_launches_per_step": head["launches"],
            "clocks": head["clocks"],
            "roofline": roofline,
            "sliding_window": sw,
            "config4": cfg4,
            "hybrid": hyb,
        }
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    del model
    gc.collect()
    torch.cuda.empty_cache()
    if world == 1 and not args.no_library_gpu:
        line["library_gpu"] = _library_gpu(dev)
        best = min((v["train_ms_per_step"] for v in line["library_gpu"].values() if isinstance(v, dict) and "train_ms_per_step" in v),
                   default=None)
        if best:
            line["library_gpu"]["speedup_train_step_vs_best_library"] = round(best / line["ms_per_step"], 2)
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        dt = _oracle_train_step_seconds(threads, 0, 1)
        line["cpu_baseline"] = {"value": 1.0 / dt, "unit": "patches/s", "cores": threads, "kind": "port",
                                "sample": "1 patch (batch 1), one step, no warm-up: oracle CTUNet fp32 fwd + 5-head Dice-CE + autograd bwd on the host cores"}
    _emit(json.dumps(line))


if __name__ == "__main__":
    main()
