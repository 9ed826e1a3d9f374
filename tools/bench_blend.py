"""CUDA-event timing of the sliding-window blend kernels at the BASELINE geometry (512x512x256 volume, 96^3 windows)."""
import json
import sys
sys.path.insert(0, ".")
import torch
from hybrid_ctunet_b200 import ops
from hybrid_ctunet_b200.sliding_window import compute_importance_map

HBM = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", 6539.2) if __import__("os").path.exists("MEASURED_PEAKS.json") else 6539.2
dev = "cuda"
C, X, Y, Z, R = 14, 512, 512, 256, 96
acc0 = torch.zeros(C, X, Y, Z, device=dev)
acc1 = torch.zeros(C, X, Y, Z, device=dev)
cnt = torch.rand(X, Y, Z, device=dev) + 0.5
imp = compute_importance_map((R, R, R), mode="gaussian", device="cpu").to(dev)
l0 = torch.randn(4, C, R, R, R, device=dev)
l1 = torch.randn(4, C, R, R, R, device=dev)
starts = [(0, 0, 0), (48, 96, 48), (416, 416, 160), (96, 48, 112)]


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def acc_all():
    for j, s in enumerate(starts):
        ops.blend_accumulate(l0[j], l1[j], imp, acc0, acc1, s)


ms = timed(acc_all, 20) / len(starts)
by = 2 * (C * R ** 3 * 4 * 3) + R ** 3 * 4      # per launch: 2 heads x (logits read + accumulator read + write) + importance map
print(json.dumps({"kernel": "blend_accumulate (1 window, 2 heads)", "ms": ms, "alg_GB": by / 1e9, "GBps": by / ms / 1e6,
                  "frac_of_measured_hbm": by / ms / 1e6 / HBM}))
ms = timed(lambda: ops.blend_normalize(acc0, cnt, acc0), 5)
by = C * X * Y * Z * 4 * 2 + X * Y * Z * 4
print(json.dumps({"kernel": "blend_normalize (1 head, whole volume)", "ms": ms, "alg_GB": by / 1e9, "GBps": by / ms / 1e6,
                  "frac_of_measured_hbm": by / ms / 1e6 / HBM}))
