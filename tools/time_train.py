"""Times the CTUNet training step (fwd + 5-head Dice-CE loss + bwd [+ AdamW]) on cuda:0; prints JSON lines."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from hybrid_ctunet_b200 import lib  # noqa: E402
from hybrid_ctunet_b200.losses import DiceCELoss, ctunet_loss  # noqa: E402
from hybrid_ctunet_b200.networks.hybrid_CTUNet import CTUNet  # noqa: E402

KW = dict(in_channels=1, dim_conv_stem=64, out_channels=14, model_depth=101, img_size=(96, 96), frames=96, patch_frame=8)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
use_graph = len(sys.argv) > 3 and sys.argv[3] == "graph"
torch.manual_seed(0)
model = CTUNet(**KW).cuda().train()
loss_func = DiceCELoss(to_onehot_y=True, softmax=True, squared_pred=True, smooth_nr=0.0, smooth_dr=1e-6)
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5, fused=True)
torch.manual_seed(1)
x = torch.rand(B, 1, 96, 96, 96, device="cuda")
y = torch.randint(0, 14, (B, 1, 96, 96, 96), device="cuda").float()


def step(timers=None):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    for p in model.parameters():
        p.grad = None
    ev[0].record()
    logits = model(x)
    ev[1].record()
    loss = ctunet_loss(logits, y, loss_func)
    ev[2].record()
    loss.backward()
    ev[3].record()
    opt.step()
    ev[4].record()
    if timers is not None:
        torch.cuda.synchronize()
        timers.append([ev[i].elapsed_time(ev[i + 1]) for i in range(4)])
    return loss


if use_graph:
    from hybrid_ctunet_b200.training import GraphedTrainStep
    gstep = GraphedTrainStep(model, lambda lg, t: ctunet_loss(lg, t, loss_func), x, y)

    def step(timers=None):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        loss = gstep(x, y)
        ev[1].record()
        opt.step()
        ev[2].record()
        if timers is not None:
            torch.cuda.synchronize()
            a, b = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
            timers.append([a, 0.0, 0.0, b])
        return loss

for i in range(2):
    l = step()
    torch.cuda.synchronize()
    print("warmup", i, float(l.detach()), flush=True)
n0 = lib.launch_count()
t = []
torch.cuda.synchronize()
w0 = time.perf_counter()
for _ in range(steps):
    l = step(t)
torch.cuda.synchronize()
wall = (time.perf_counter() - w0) / steps * 1e3
avg = [sum(r[i] for r in t) / len(t) for i in range(4)]
print(json.dumps({"B": B, "fwd_ms": avg[0], "loss_ms": avg[1], "bwd_ms": avg[2], "opt_ms": avg[3], "wall_ms_per_step": wall,
                  "patches_per_s": B / (wall * 1e-3), "launches_per_step": (lib.launch_count() - n0) / steps,
                  "loss": float(l.detach()), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                  "tflops_fwd_bwd": B * 10258.03 / (avg[0] + avg[2])}))
