"""One launch of ctu_umma_wgrad per shape (for ncu): python tools/one_wgrad.py <k> <cin> <cout> <d1> <d2> <d3> <d4>"""
import sys
sys.path.insert(0, ".")
import torch
from hybrid_ctunet_b200 import ops
k, ci, co, d1, d2, d3, d4 = (int(v) for v in sys.argv[1:8])
torch.manual_seed(0)
x = torch.randn(d4, d3, d2, d1, ci, device="cuda").to(torch.bfloat16)
dy = torch.randn(d4, d3, d2, d1, co, device="cuda").to(torch.bfloat16)
dw = torch.zeros(k ** 3 * ci, co, device="cuda")
for _ in range(2):
    ops.wgrad(x, dy, dw, dims=(d1, d2, d3, d4), ksize=k)
torch.cuda.synchronize()
print("ok")
