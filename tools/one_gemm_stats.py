"""A few launches of the HBM-bound 1x1x1 GEMM with fused InstanceNorm statistics (ResNet layer-1 conv3: 64 -> 128 channels
at 48x48x96, batch 2 — umma_gemm_kernel<128,2,1,2,8>) for `ncu --set full` / timing."""
import sys
sys.path.insert(0, ".")
import torch
from hybrid_ctunet_b200 import ops
B, S, k, n = 2, 48 * 48 * 96, 64, 128
a = torch.randn(B * S, k, device="cuda").to(torch.bfloat16)
pw = ops.pack_matrix(torch.randn(n, k, device="cuda") * 0.05)
out = torch.empty(B * S, n, device="cuda", dtype=torch.bfloat16)
st = torch.zeros(B, n, 2, device="cuda", dtype=torch.float64)
for _ in range(4):
    ops.gemm(a, pw, out, dims=(S, 1, 1, B), stats=st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.gemm(a, pw, out, dims=(S, 1, 1, B), stats=st)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 100
print(f"gemm 64->128 +stats rows {B * S}: {us:.1f} us  {B * S * (k + n) * 2 / us / 1e3:.0f} GB/s")
