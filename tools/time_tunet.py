"""TUNet forward (4 windows, CUDA graph) with and without the second lane."""
import sys
sys.path.insert(0, ".")
import torch
from hybrid_ctunet_b200.networks.hybrid_CTUNet import TUNet
torch.manual_seed(5)
m = TUNet(in_channels=1, dim_conv_stem=64, out_channels=14, img_size=(96, 96), frames=96, patch_frame=8).cuda().eval().enable_cuda_graph()
x = torch.rand(4, 1, 96, 96, 96, device="cuda")
with torch.no_grad():
    for _ in range(3): y = m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): y = m(x)
    e1.record(); torch.cuda.synchronize()
print("tunet 4 windows", e0.elapsed_time(e1) / 10, "ms", float(y[0].float().abs().mean()))
