"""Device time of one in-place fp32 all-reduce (AVG) of the gradient arena's size; torchrun --nproc-per-node N."""
import os, sys
import torch, torch.distributed as dist
dist.init_process_group("nccl")
r = int(os.environ.get("LOCAL_RANK", 0)); torch.cuda.set_device(r)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 174_000_000
x = torch.randn(n, device="cuda")
for _ in range(3): dist.all_reduce(x, op=dist.ReduceOp.AVG)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): dist.all_reduce(x, op=dist.ReduceOp.AVG)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
if r == 0:
    w = dist.get_world_size()
    print(f"world {w}: {n * 4 / 1e6:.0f} MB all-reduce {ms:.3f} ms  algbw {n * 4 / ms / 1e6:.0f} GB/s  busbw {n * 4 / ms / 1e6 * 2 * (w - 1) / w:.0f} GB/s  "
          f"env {dict((k, v) for k, v in os.environ.items() if k.startswith('NCCL'))}", flush=True)
dist.destroy_process_group()
