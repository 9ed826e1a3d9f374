"""A few launches of the dominant kernel (halo-reuse tcgen05 conv 64->64 @96^3, batch 2) for `ncu --set full`."""
import sys
sys.path.insert(0, ".")
import torch
from hybrid_ctunet_b200 import ops
B, X, Y, Z, ci, co = 2, 96, 96, 96, 64, 64
a = torch.randn(B, X, Y, Z, ci, device="cuda").to(torch.bfloat16)
w = torch.randn(co, 27 * ci, device="cuda") * 0.02
pw = ops.pack_matrix(w, ksize=3, a_c=ci)
out = torch.empty(B, X, Y, Z, co, device="cuda", dtype=torch.bfloat16)
st = torch.zeros(B, co, 2, device="cuda", dtype=torch.float64)
for _ in range(4):
    ops.gemm(a, pw, out, dims=(Z, Y, X, B), stats=st)
torch.cuda.synchronize()
print("done")
