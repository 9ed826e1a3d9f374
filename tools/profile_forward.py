"""One CTUNet forward between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import sys
sys.path.insert(0, ".")
import torch
from hybrid_ctunet_b200.networks.hybrid_CTUNet import CTUNet
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
torch.manual_seed(0)
m = CTUNet(in_channels=1, dim_conv_stem=64, out_channels=14, model_depth=101, img_size=(96, 96), frames=96, patch_frame=8).cuda().eval()
x = torch.rand(B, 1, 96, 96, 96, device="cuda")
with torch.no_grad():
    m(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    m(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
