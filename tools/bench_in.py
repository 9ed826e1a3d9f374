"""CUDA-event timing of the InstanceNorm kernels (forward apply, backward stats + apply) at the shapes of one CTUNet
training step (batch 2).  GB/s counts the bytes each PASS has to move: forward apply reads x (+res) and writes out;
backward = stats pass (dout, out | x [, res]) + apply pass (same reads + dx [+ dres]).

`python tools/bench_in.py [json_out]`; an L2 flush (256 MB write) runs between timed calls.
"""
import json
import sys
sys.path.insert(0, ".")
import torch
from hybrid_ctunet_b200 import ops

BF = torch.bfloat16
# (shape [B,X,Y,Z,C], res_mode: 0 none / 1 identity residual / 2 normalised residual, calls per training step fwd, bwd)
CASES = [
    ((2, 48, 48, 96, 128), 1, 9),
    ((2, 48, 48, 96, 64), 0, 18),
    ((2, 96, 96, 96, 64), 2, 2),
    ((2, 96, 96, 96, 64), 0, 3),
    ((2, 96, 96, 96, 64), 1, 1),
    ((2, 48, 48, 96, 128), 0, 2),
    ((2, 48, 48, 96, 128), 2, 1),
    ((2, 24, 24, 48, 256), 1, 10),
    ((2, 24, 24, 48, 64), 0, 17),
    ((2, 12, 12, 24, 512), 1, 14),
    ((2, 12, 12, 24, 128), 0, 25),
    ((2, 6, 6, 12, 1024), 1, 3),
    ((2, 6, 6, 12, 256), 0, 5),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, iters=10):
    for _ in range(2):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


res = []
tot_f = tot_b = 0.0
for shape, mode, calls in CASES:
    B, C = shape[0], shape[-1]
    x = torch.randn(shape, device="cuda").to(BF)
    r = torch.randn(shape, device="cuda").to(BF) if mode else None
    dout = torch.randn(shape, device="cuda").to(BF)
    out = torch.empty_like(x)
    dx = torch.empty_like(x)
    dres = torch.empty_like(x) if mode else None
    st = torch.zeros(B, C, 2, dtype=torch.float64, device="cuda")
    ops.in_stats(x, st)
    rst = None
    if mode == 2:
        rst = torch.zeros(B, C, 2, dtype=torch.float64, device="cuda")
        ops.in_stats(r, rst)
    sums = torch.zeros(B, C, 4, dtype=torch.float64, device="cuda")
    fwd = lambda: ops.in_apply(x, st, out, res=r, rstats=rst, act=True)

    def bwd():
        sums.zero_()
        # without a residual the engine does not keep x: xhat is recovered from `out`
        ops.in_backward(dout, out, x if mode else None, st, dx, res=r, rstats=rst, dres=dres, sums=sums)

    fwd()
    tf, tb = timed(fwd), timed(bwd)
    unit = x.numel() * 2
    f_bytes = unit * (2 + (1 if mode else 0))
    reads = 2 + (1 if mode else 0) + (1 if mode == 2 else 0)
    b_bytes = unit * (2 * reads + 1 + (1 if mode else 0))
    rec = {"shape": list(shape), "res_mode": mode, "calls_per_step": calls, "fwd_us": tf, "bwd_us": tb,
           "fwd_gbs": f_bytes / tf / 1e3, "bwd_gbs": b_bytes / tb / 1e3}
    tot_f += tf * calls
    tot_b += tb * calls
    res.append(rec)
    print(json.dumps(rec))
print(json.dumps({"per_step_fwd_ms": tot_f / 1e3, "per_step_bwd_ms": tot_b / 1e3}))
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
