"""CUDA-event timing of ctu_umma_wgrad on the shapes that dominate the CTUNet training step (batch 2)."""
import os
import sys
sys.path.insert(0, ".")
import torch
from hybrid_ctunet_b200 import ops

SMALL = [(1, 3072, 768, (864, 1, 1, 1)), (1, 768, 3072, (864, 1, 1, 1)), (1, 768, 2304, (864, 1, 1, 1)), (1, 768, 768, (864, 1, 1, 1)),
         (1, 512, 128, (3456, 1, 1, 2)), (1, 128, 512, (3456, 1, 1, 2)), (3, 128, 128, (24, 12, 12, 2)), (1, 1024, 256, (432, 1, 1, 2)),
         (1, 256, 64, (27648, 1, 1, 2)), (1, 64, 256, (27648, 1, 1, 2)), (3, 64, 64, (48, 24, 24, 2)), (1, 512, 1536, (6912, 1, 1, 1))]
SHAPES = [  # (ksize, Cin, Cout, (d1, d2, d3, d4))
    (3, 128, 128, (96, 48, 48, 2)), (3, 64, 64, (96, 96, 96, 2)), (3, 64, 64, (96, 48, 48, 2)), (3, 128, 64, (96, 96, 96, 2)),
    (3, 256, 256, (48, 24, 24, 2)), (3, 512, 512, (24, 12, 12, 2)), (3, 128, 128, (24, 12, 12, 2)),
    (1, 3072, 768, (864, 1, 1, 1)), (1, 768, 3072, (864, 1, 1, 1)), (1, 128, 384, (442368, 1, 1, 1)), (1, 128, 64, (221184, 1, 1, 2)),
    (1, 64, 128, (221184, 1, 1, 2)), (1, 512, 128, (442368, 1, 1, 1)), (1, 256, 768, (55296, 1, 1, 1)),
]
torch.manual_seed(0)
print("variant", os.environ.get("CTU_WGRAD_VARIANT", "0"), "items/slot", os.environ.get("CTU_WGRAD_ITEMS_PER_SLOT", "2"))
if len(sys.argv) > 1 and sys.argv[1] == "small":
    SHAPES = SMALL
    print("splits", os.environ.get("CTU_WGRAD_SPLITS"))
for k, ci, co, dims in SHAPES:
    d1, d2, d3, d4 = dims
    x = torch.randn(d4, d3, d2, d1, ci, device="cuda").to(torch.bfloat16)
    dy = torch.randn(d4, d3, d2, d1, co, device="cuda").to(torch.bfloat16)
    dw = torch.zeros(k ** 3 * ci, co, device="cuda")
    for _ in range(3):
        ops.wgrad(x, dy, dw, dims=dims, ksize=k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        ops.wgrad(x, dy, dw, dims=dims, ksize=k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 2.0 * d1 * d2 * d3 * d4 * k ** 3 * ci * co
    print(f"{ms:8.4f} ms {fl / ms / 1e9:7.1f} TF/s  k{k} {ci}->{co} {dims}")
    del x, dy, dw
