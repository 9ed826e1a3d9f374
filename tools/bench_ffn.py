"""Fused FFN-128 (ctu_ffn_fused) against the LN-less two-GEMM path at the inference shape (4 windows: 884736 rows)."""
import sys
sys.path.insert(0, ".")
import torch
from hybrid_ctunet_b200 import ops
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4 * 48 * 48 * 96
C, H = 128, 512
torch.manual_seed(0)
w1 = torch.randn(H, C, device="cuda") / C ** 0.5; b1 = torch.randn(H, device="cuda")
w2 = torch.randn(C, H, device="cuda") / H ** 0.5; b2 = torch.randn(C, device="cuda")
p1, p2 = ops.pack_matrix(w1, bias=b1), ops.pack_matrix(w2, bias=b2)
a = torch.randn(M, C, device="cuda").to(torch.bfloat16); x = torch.randn(M, C, device="cuda").to(torch.bfloat16)
out = torch.empty(M, C, device="cuda", dtype=torch.bfloat16); hid = torch.empty(M, H, device="cuda", dtype=torch.bfloat16)
def fused(): ops.ffn_fused(a, p1, p2, x, out)
def two():
    ops.gemm(a, p1, hid, dims=(M, 1, 1, 1), act=ops.ACT_GELU)
    ops.gemm(hid, p2, out, dims=(M, 1, 1, 1), residual=x)
for name, fn in (("fused", fused), ("two-gemm", two)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:10s} {ms:.4f} ms  {4.0 * M * C * H / ms / 1e9:.1f} TFLOP/s  alg {M * C * 2 * 3 / ms / 1e6:.0f} GB/s", flush=True)
