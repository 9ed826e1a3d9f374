"""CUDA-event timing of the attention forward / backward kernels at the shapes of one CTUNet training step (batch 2).

`python tools/bench_attention.py [json_out]` — prints per shape: forward us, backward us (delta + bwd kernels), the
TFLOP/s of each (forward 4*n*n*D per (window, head), backward 10*n*n*D).
"""
import json
import sys
sys.path.insert(0, ".")
import torch
from hybrid_ctunet_b200 import ops

BF = torch.bfloat16
# (name, mode, D, heads, batch, grid X Y Z or n)
CASES = [
    ("vit_mhsa_432x12h64", 0, 64, 12, 2, None, 432),
    ("win_768_24h32@6.6.12", 1, 32, 24, 2, (6, 6, 12), 216),
    ("win_512_16h32@12.12.24", 1, 32, 16, 2, (12, 12, 24), 216),
    ("grid_512_16h32@12.12.24", 2, 32, 16, 2, (12, 12, 24), 216),
    ("win_256_8h32@24.24.48", 1, 32, 8, 2, (24, 24, 48), 216),
    ("grid_256_8h32@24.24.48", 2, 32, 8, 2, (24, 24, 48), 216),
]


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


res = []
for name, mode, D, heads, B, g, n in CASES:
    C = D * heads
    if mode == 0:
        rows, windows, grid, bias = B * n, B, (1, 1, 1, 1), None
    else:
        X, Y, Z = g
        rows, windows, grid = B * X * Y * Z, B * X * Y * Z // 216, (B, X, Y, Z)
        bias = torch.randn(heads, n, n, device="cuda")
    qkv = (torch.randn(rows, 3 * C, device="cuda") * 0.7).to(BF)
    dout = torch.randn(rows, C, device="cuda").to(BF)
    out = torch.empty(rows, C, device="cuda", dtype=BF)
    lse = torch.empty(rows, heads, device="cuda")
    dqkv = torch.zeros(rows, 3 * C, device="cuda", dtype=BF)
    direct = D == 32 and n <= 224
    dq = None if direct else torch.zeros(rows, C, device="cuda")
    ds = torch.zeros(windows, heads, n, n, device="cuda", dtype=BF) if bias is not None else None
    bias_t = bias.transpose(1, 2).contiguous() if bias is not None else None
    fwd = lambda: ops.attention(qkv, out, dim_head=D, n=n, windows=windows, mode=mode, bias=bias, grid=grid, lse=lse)
    bwd = lambda: ops.attention_backward(qkv, out, dout, lse, dqkv, dq, dim_head=D, n=n, windows=windows, mode=mode,
                                         bias_t=bias_t, ds_out=ds, grid=grid)
    tf, tb = timed(fwd), timed(bwd)
    fl = windows * heads * n * n * D
    r = {"case": name, "fwd_us": tf, "bwd_us": tb, "fwd_tflops": 4 * fl / tf / 1e6, "bwd_tflops": 10 * fl / tb / 1e6}
    res.append(r)
    print(json.dumps(r))
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
