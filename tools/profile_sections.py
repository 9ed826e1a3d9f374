"""Per-block CUDA-event timing of one CTUNet forward (eager), grouped by engine method and parameter prefix."""
import sys, collections, json
sys.path.insert(0, ".")
import torch
from hybrid_ctunet_b200.networks.hybrid_CTUNet import CTUNet
from hybrid_ctunet_b200 import engine as E
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
records = []
def wrap(name):
    orig = getattr(E.Engine, name)
    def f(self, *a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = orig(self, *a, **k); e1.record()
        pre = a[0] if a and isinstance(a[0], str) else ""
        if name in ("up_gemm", "head"):
            pre = f"{tuple(a[0].shape)}"
        records.append((name, pre, e0, e1))
        return r
    setattr(E.Engine, name, f)
for n in ("bottleneck", "res_block", "res_block_cin1", "pixelweight_attention", "up_gemm", "head", "vit_attention", "ffn", "window_attention"):
    wrap(n)
torch.manual_seed(0)
m = CTUNet(in_channels=1, dim_conv_stem=64, out_channels=14, model_depth=101, img_size=(96, 96), frames=96, patch_frame=8).cuda().eval()
x = torch.rand(B, 1, 96, 96, 96, device="cuda")
with torch.no_grad():
    m(x); m(x); torch.cuda.synchronize(); records.clear()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(); m(x); t1.record(); torch.cuda.synchronize()
tot = t0.elapsed_time(t1)
agg = collections.OrderedDict()
for name, pre, e0, e1 in records:
    key = pre
    if name == "bottleneck": key = pre.split(".")[1] if "." in pre else pre
    if name in ("vit_attention", "ffn") and pre.startswith("vit.transformer"): key = "vit.transformer.*"
    if "vit_encoder.layers" in pre: key = ".".join(pre.split(".")[:3])
    k = f"{name:22s} {key}"
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += e0.elapsed_time(e1)
print(f"total forward {tot:.2f} ms (B={B})")
s = 0
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    s += t; print(f"{t:8.3f} ms {100*t/tot:5.1f}%  x{n:3d}  {k}")
print(f"unattributed (stem, patch embed, glue): {tot - s:.3f} ms")
