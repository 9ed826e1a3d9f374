"""One launch of each InstanceNorm kernel on a 2 x 96^3 x 64 tensor (for `ncu -k regex:in_`): forward apply with a
normalised residual, backward stats + apply of the same, and the residual-free forward / backward."""
import sys
sys.path.insert(0, ".")
import torch
from hybrid_ctunet_b200 import ops
BF = torch.bfloat16
shape = (2, 96, 96, 96, 64)
x = torch.randn(shape, device="cuda").to(BF)
r = torch.randn(shape, device="cuda").to(BF)
dout = torch.randn(shape, device="cuda").to(BF)
out, dx, dres = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
st = torch.zeros(2, 64, 2, dtype=torch.float64, device="cuda"); ops.in_stats(x, st)
rst = torch.zeros(2, 64, 2, dtype=torch.float64, device="cuda"); ops.in_stats(r, rst)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for mode in (2, 0):
    flush.zero_()
    ops.in_apply(x, st, out, res=r if mode else None, rstats=rst if mode else None, act=True)
    flush.zero_()
    sums = torch.zeros(2, 64, 4, dtype=torch.float64, device="cuda")
    ops.in_backward(dout, out, x if mode else None, st, dx, res=r if mode else None, rstats=rst if mode else None,
                    dres=dres if mode else None, sums=sums)
torch.cuda.synchronize()
print("ok")
