"""BASELINE config 5: Hybrid-CTUNet mask-complementation inference of one synthetic 512x512x256 volume — CTUNet head 0
blended at overlap 0.5 (500 windows) + an independently initialised TUNet blended at overlap 0.7 (1,792 windows) +
device ensemble.  Run plainly (1 GPU) or under torchrun (windows sharded over the ranks)."""
import json
import os
import sys
sys.path.insert(0, ".")
import torch
import torch.distributed as dist
from hybrid_ctunet_b200.ensemble import hybrid_ctunet_inference
from hybrid_ctunet_b200.networks.hybrid_CTUNet import CTUNet, TUNet

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
group = None
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD
kw = dict(in_channels=1, dim_conv_stem=64, out_channels=14, img_size=(96, 96), frames=96, patch_frame=8)
torch.manual_seed(0)
ctunet = CTUNet(model_depth=101, **kw).to(dev).eval().enable_cuda_graph()
torch.manual_seed(5)
tunet = TUNet(**kw).to(dev).eval().enable_cuda_graph()
torch.manual_seed(2)
vol = torch.rand(1, 1, 512, 512, 256, device=dev)
lab = torch.randint(0, 14, (512, 512, 256), device=dev).float()
out = hybrid_ctunet_inference(vol, ctunet, tunet, labels=lab, shard_group=group)   # warm-up (graph captures)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = hybrid_ctunet_inference(vol, ctunet, tunet, labels=lab, shard_group=group)
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1)], device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(t.item())
    print(json.dumps({"metric": "hybrid_ctunet_volumes_per_s", "value": 1e3 / ms, "ms_per_volume": ms, "n_gpus": world,
                      "windows": {"ctunet@0.5": 500, "tunet@0.7": 1792}, "tflops_per_gpu": 3802.0e3 / ms / world,
                      "mean_dice_ensemble": float(out["dice"][0, 1:].mean())}))
if world > 1:
    dist.destroy_process_group()
