"""Times ctu_head_bwd on the three head shapes of the training step (batch 2): CUDA events, L2 flushed between launches."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hybrid_ctunet_b200 import ops  # noqa: E402

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {}
for C, shape in ((64, (2, 96, 96, 96)), (128, (2, 48, 48, 96)), (256, (2, 24, 24, 48))):
    for acc in (False, True):
        B = shape[0]
        g = torch.randn((B, 14) + shape[1:], device=dev)
        a = torch.randn(shape + (C,), device=dev).bfloat16()
        w = torch.randn(14, C, device=dev)
        da = torch.zeros_like(a)
        dw = torch.zeros(C, 16, device=dev)
        db = torch.zeros(16, device=dev)
        ts = []
        for i in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.head_backward(g, a, w, da, dw, db, accumulate=acc)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = min(ts[1:])
        nbytes = g.numel() * 4 + a.numel() * 2 * (3 if acc else 2)
        res[f"C{C}_{'acc' if acc else 'set'}"] = {"us": round(t * 1e3, 1), "GBs": round(nbytes / t / 1e6, 1)}
print(json.dumps(res))
