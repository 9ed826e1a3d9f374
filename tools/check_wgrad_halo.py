"""Correctness + speed probe of the halo-reuse wgrad kernel (CTU_WGRAD_HALO=1) against torch / the per-tap kernel."""
import os
import sys
sys.path.insert(0, ".")
import torch
import torch.nn.functional as F
from hybrid_ctunet_b200 import ops

torch.cuda.init()
print("CTU_WGRAD_HALO", os.environ.get("CTU_WGRAD_HALO"), "variant", os.environ.get("CTU_WGRAD_HALO_VARIANT"), "per_slot",
      os.environ.get("CTU_WGRAD_HALO_PER_SLOT"))
torch.backends.cudnn.allow_tf32 = False
for (B, X, Y, Z, ci, co) in [(1, 4, 16, 8, 64, 64), (2, 5, 32, 16, 64, 64), (1, 3, 16, 16, 128, 64), (1, 3, 32, 8, 128, 128), (1, 4, 16, 8, 64, 128)]:
    torch.manual_seed(X + ci + co)
    x = torch.randn(B, X, Y, Z, ci, device="cuda").to(torch.bfloat16)
    dy = torch.randn(B, X, Y, Z, co, device="cuda").to(torch.bfloat16)
    dw = torch.zeros(27 * ci, co, device="cuda")
    ops.wgrad(x, dy, dw, dims=(Z, Y, X, B), ksize=3)
    w = torch.zeros(co, ci, 3, 3, 3, device="cuda", requires_grad=True)
    F.conv3d(x.float().permute(0, 4, 1, 2, 3), w, padding=1).backward(dy.float().permute(0, 4, 1, 2, 3))
    ref = w.grad.permute(2, 3, 4, 1, 0).reshape(27 * ci, co)
    print(f"shape {(B, X, Y, Z, ci, co)} rel {((dw - ref).norm() / ref.norm()).item():.6f}")
for (B, X, Y, Z, ci, co) in [(2, 96, 96, 96, 64, 64), (2, 48, 48, 96, 128, 128), (2, 96, 96, 96, 128, 64), (2, 48, 48, 96, 64, 64)]:
    x = torch.randn(B, X, Y, Z, ci, device="cuda").to(torch.bfloat16)
    dy = torch.randn(B, X, Y, Z, co, device="cuda").to(torch.bfloat16)
    dw = torch.zeros(27 * ci, co, device="cuda")
    for _ in range(3):
        ops.wgrad(x, dy, dw, dims=(Z, Y, X, B), ksize=3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.wgrad(x, dy, dw, dims=(Z, Y, X, B), ksize=3)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{ms:8.4f} ms {2.0 * B * X * Y * Z * 27 * ci * co / ms / 1e9:7.1f} TF/s wgrad {ci}->{co} @ {(X, Y, Z)} B={B}")
    del x, dy, dw
