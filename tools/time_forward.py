"""Times the CTUNet forward (eager launches vs CUDA-graph replay) with CUDA events; prints ms and TFLOP/s."""
import sys, time, json
sys.path.insert(0, ".")
import torch
from hybrid_ctunet_b200.networks.hybrid_CTUNet import CTUNet
from hybrid_ctunet_b200 import lib
FWD_GFLOP = 3423.64
torch.manual_seed(0)
m = CTUNet(in_channels=1, dim_conv_stem=64, out_channels=14, model_depth=101, img_size=(96, 96), frames=96, patch_frame=8).cuda().eval()
res = {}
for B in (1, 2, 4):
    x = torch.rand(B, 1, 96, 96, 96, device="cuda")
    for mode in ("eager", "graph"):
        m.enable_cuda_graph(mode == "graph")
        with torch.no_grad():
            for _ in range(3):
                m(x)
            torch.cuda.synchronize()
            n0 = lib.launch_count()
            t0 = time.time()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                m(x)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            wall = (time.time() - t0) / 5 * 1e3
        res[f"B{B}_{mode}"] = dict(ms=ms, wall_ms=wall, tflops=FWD_GFLOP * B / ms, launches=(lib.launch_count() - n0) / 5,
                                   mem_gb=torch.cuda.max_memory_allocated() / 2**30)
        print(B, mode, res[f"B{B}_{mode}"], flush=True)
json.dump(res, open("gpurun_out/time_forward.json", "w"), indent=1)
