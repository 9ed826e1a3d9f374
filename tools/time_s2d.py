"""Times ctu_space_to_depth on the two big shapes of the training step (CUDA events, L2 flushed between launches)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hybrid_ctunet_b200 import ops  # noqa: E402

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {}
for shape, up in (((2, 96, 96, 96, 64), (2, 2, 2)), ((2, 48, 48, 96, 128), (2, 2, 1))):
    x = torch.randn(shape, device=dev).bfloat16()
    B, X, Y, Z, C = shape
    out = torch.empty(B, X // up[0], Y // up[1], Z // up[2], up[0] * up[1] * up[2] * C, dtype=torch.bfloat16, device=dev)
    ts = []
    for i in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.space_to_depth(x, out, up)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts[1:])
    res[str(shape)] = {"us": round(t * 1e3, 1), "GBs": round(2 * x.numel() * 2 / t / 1e6, 1)}
print(json.dumps(res))
