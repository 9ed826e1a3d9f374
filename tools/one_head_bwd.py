"""One launch set of ctu_head_bwd at the benchmark shape (for ncu / timing): python tools/one_head_bwd.py [reps]"""
import sys
sys.path.insert(0, ".")
import torch
from hybrid_ctunet_b200 import ops
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B, C = 2, 64
a = torch.randn(B, 96, 96, 96, C, device="cuda").to(torch.bfloat16)
g = torch.randn(B, 14, 96, 96, 96, device="cuda")
w = torch.randn(14, C, device="cuda") * 0.1
da = torch.empty_like(a)
dw, db = torch.zeros(C, 16, device="cuda"), torch.zeros(16, device="cuda")
ops.head_backward(g, a, w, da, dw, db)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ops.head_backward(g, a, w, da, dw, db)
e1.record()
torch.cuda.synchronize()
print(f"head_backward 2x96^3 C=64: {e0.elapsed_time(e1) / reps * 1e3:.1f} us")
