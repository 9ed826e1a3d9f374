"""Timing of ctu_im2col_cin1 at the benchmark shapes: ResNet stem (k7, s(2,2,1)) and vit_encoder0 (k3, s1)."""
import sys
sys.path.insert(0, ".")
import torch
from hybrid_ctunet_b200 import ops
for B in (2, 4):
    x = torch.randn(B, 1, 96, 96, 96, device="cuda")
    for name, k, s, p, oshape, kpad in (("stem k7 s(2,2,1)", (7, 7, 7), (2, 2, 1), (3, 3, 3), (48, 48, 96), 384),
                                       ("k3 s1", (3, 3, 3), (1, 1, 1), (1, 1, 1), (96, 96, 96), 64)):
        col = torch.empty(B, *oshape, kpad, device="cuda", dtype=torch.bfloat16)
        for _ in range(3):
            ops.im2col_cin1(x, col, k=k, s=s, p=p)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.im2col_cin1(x, col, k=k, s=s, p=p)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        print(f"B={B} im2col {name}: {us:.1f} us  {col.numel() * 2 / us / 1e3:.0f} GB/s written")
