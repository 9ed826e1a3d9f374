"""One CTUNet training step (fwd + loss + bwd + AdamW) between cudaProfilerStart/Stop (`ncu --profile-from-start off`)."""
import sys
sys.path.insert(0, ".")
import torch
from hybrid_ctunet_b200.losses import DiceCELoss, ctunet_loss
from hybrid_ctunet_b200.networks.hybrid_CTUNet import CTUNet
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(0)
m = CTUNet(in_channels=1, dim_conv_stem=64, out_channels=14, model_depth=101, img_size=(96, 96), frames=96, patch_frame=8).cuda().train()
lf = DiceCELoss(to_onehot_y=True, softmax=True, squared_pred=True, smooth_nr=0.0, smooth_dr=1e-6)
opt = torch.optim.AdamW(m.parameters(), lr=1e-4, weight_decay=1e-5, fused=True)
x = torch.rand(B, 1, 96, 96, 96, device="cuda")
y = torch.randint(0, 14, (B, 1, 96, 96, 96), device="cuda").float()


def step():
    for p in m.parameters():
        p.grad = None
    loss = ctunet_loss(m(x), y, lf)
    loss.backward()
    opt.step()


step()
step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
