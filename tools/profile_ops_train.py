"""Per-shape CUDA-event timing of every tensor-core launch (ops.gemm / ops.wgrad) in one CTUNet training step."""
import collections
import sys
sys.path.insert(0, ".")
import torch
from hybrid_ctunet_b200 import ops
from hybrid_ctunet_b200.losses import DiceCELoss, ctunet_loss
from hybrid_ctunet_b200.networks.hybrid_CTUNet import CTUNet
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(0)
m = CTUNet(in_channels=1, dim_conv_stem=64, out_channels=14, model_depth=101, img_size=(96, 96), frames=96, patch_frame=8).cuda().train()
lf = DiceCELoss(to_onehot_y=True, softmax=True, squared_pred=True, smooth_nr=0.0, smooth_dr=1e-6)
x = torch.rand(B, 1, 96, 96, 96, device="cuda")
y = torch.randint(0, 14, (B, 1, 96, 96, 96), device="cuda").float()


def step():
    for p in m.parameters():
        p.grad = None
    ctunet_loss(m(x), y, lf).backward()


step()
torch.cuda.synchronize()
rec = []
og, ow = ops.gemm, ops.wgrad


def gemm(a, w, out, *, dims, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = og(a, w, out, dims=dims, **kw); e1.record()
    ac = kw.get("a_c") or w.a_c
    rows = dims[0] * dims[1] * dims[2] * dims[3]
    fl = 2.0 * rows * w.ksize ** 3 * ac * w.n_real
    by = rows * (ac + w.n_real) * 2 + w.w.numel() * 2
    rec.append((("gemm", w.ksize, ac, w.n_real, tuple(dims), w.block_n, bool(w.convt), kw.get("out_mode", 0), kw.get("stats") is not None), e0, e1, fl, by))
    return r


def wgrad(x_, dy, dw, *, dims, ksize=1, x_c=None, n=None):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = ow(x_, dy, dw, dims=dims, ksize=ksize, x_c=x_c, n=n); e1.record()
    xc = x_c or x_.shape[-1]; nn = n or dy.shape[-1]
    rows = dims[0] * dims[1] * dims[2] * dims[3]
    fl = 2.0 * rows * ksize ** 3 * xc * nn
    by = rows * (xc + nn) * 2
    rec.append((("wgrad", ksize, xc, nn, tuple(dims), ops.pick_wgrad_block_n(nn)), e0, e1, fl, by))
    return r


ops.gemm, ops.wgrad = gemm, wgrad
step()
torch.cuda.synchronize()
agg = collections.OrderedDict()
for k, e0, e1, fl, by in rec:
    a = agg.setdefault(k, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += e0.elapsed_time(e1); a[2] += fl; a[3] += by
tot = sum(a[1] for a in agg.values())
print(f"total tensor-kernel time {tot:.2f} ms over {len(rec)} launches")
for k, (n, ms, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{ms:8.3f} ms x{n:3d}  {fl / ms / 1e9:7.1f} TF/s {by / ms / 1e6:7.1f} GB/s  {k}")
