"""Micro-benchmark of the tcgen05 kernel on the CTUNet shapes that dominate the forward (batch 4).
usage: bench_shapes.py [name-filter] [--once]   (--once: one launch per shape, for ncu)"""
import sys, json
sys.path.insert(0, ".")
import torch
from hybrid_ctunet_b200 import ops
filt = [a for a in sys.argv[1:] if not a.startswith("--")]
once = "--once" in sys.argv
B = 4
def conv_w(ci, co, bn=None):
    w = torch.randn(co, 27 * ci, device="cuda") * 0.02
    return ops.pack_matrix(w, ksize=3, a_c=ci, block_n=bn)
def lin_w(k, n, bias=False, bn=None):
    w = torch.randn(n, k, device="cuda") * 0.02
    return ops.pack_matrix(w, bias=torch.randn(n, device="cuda") if bias else None, block_n=bn)
shapes = []
def add(name, fn, flops, bytes_):
    if not filt or any(f in name for f in filt):
        shapes.append((name, fn, flops, bytes_))
def mk_conv(name, X, Y, Z, ci, co, bn=None):
    a = torch.randn(B, X, Y, Z, ci, device="cuda").to(torch.bfloat16)
    w = conv_w(ci, co, bn)
    out = torch.empty(B, X, Y, Z, co, device="cuda", dtype=torch.bfloat16)
    st = torch.zeros(B, co, 2, device="cuda", dtype=torch.float64)
    vox = B * X * Y * Z
    add(name, lambda: ops.gemm(a, w, out, dims=(Z, Y, X, B), stats=st), 2.0 * vox * 27 * ci * co, vox * (ci + co) * 2)
def mk_lin(name, M, k, n, act=0, bias=False, res=False, bn=None, stats=False):
    a = torch.randn(M, k, device="cuda").to(torch.bfloat16)
    w = lin_w(k, n, bias, bn)
    out = torch.empty(M, n, device="cuda", dtype=torch.bfloat16)
    r = torch.randn(M, n, device="cuda").to(torch.bfloat16) if res else None
    st = torch.zeros(B, n, 2, device="cuda", dtype=torch.float64) if stats else None
    dims = (M // B, 1, 1, B) if stats else (M, 1, 1, 1)
    add(name, lambda: ops.gemm(a, w, out, dims=dims, act=act, residual=r, stats=st), 2.0 * M * k * n, M * (k + n + (n if res else 0)) * 2)
mk_conv("conv64_64@96", 96, 96, 96, 64, 64)
mk_conv("conv128_64@96", 96, 96, 96, 128, 64)
mk_conv("conv128_128@48", 48, 48, 96, 128, 128)
mk_conv("conv256_256@24", 24, 24, 48, 256, 256)
mk_conv("conv512_512@12", 12, 12, 24, 512, 512)
mk_conv("conv64_64@48(l1)", 48, 48, 96, 64, 64)
T1 = B * 48 * 48 * 96
mk_lin("ffn128_up_gelu", T1, 128, 512, act=1, bias=True)
mk_lin("ffn128_up_gelu_bn64", T1, 128, 512, act=1, bias=True, bn=64)
mk_lin("ffn128_up_noact", T1, 128, 512, bias=True)
mk_lin("ffn128_up_noact_bn64", T1, 128, 512, bias=True, bn=64)
mk_lin("ffn128_down_res", T1, 512, 128, bias=True, res=True)
mk_lin("pwa128_qkv", T1, 128, 384)
mk_lin("pwa128_out", T1, 128, 128)
mk_lin("l1_conv1_128_64_stats", T1, 128, 64, stats=True)
mk_lin("l1_conv3_64_128_stats", T1, 64, 128, stats=True)
mk_lin("vitdec_conv3_128_64@96_stats", B * 96 ** 3, 128, 64, stats=True)
T2 = B * 24 * 24 * 48
mk_lin("pwa256_qkv", T2, 256, 768)
mk_lin("ffn256_up_gelu", T2, 256, 1024, act=1, bias=True)
mk_lin("ffn256_up_gelu_bn64", T2, 256, 1024, act=1, bias=True, bn=64)
mk_lin("vit_qkv", B * 432, 768, 2304)
mk_lin("vit_ffn_up", B * 432, 768, 3072, act=1, bias=True)
mk_lin("vit2_qkv", 864, 768, 2304)
mk_lin("vit2_ffn_up", 864, 768, 3072, bias=True)
mk_lin("vit2_ffn_down", 864, 3072, 768, bias=True)
mk_lin("vit2_out", 864, 768, 768, bias=True)
mk_lin("vit2_dgrad_qkv", 864, 2304, 768)
mk_lin("vit2_qkv_bn64", 864, 768, 2304, bn=64)
mk_lin("vit2_ffn_up_bn64", 864, 768, 3072, bias=True, bn=64)
mk_lin("vit2_ffn_down_bn64", 864, 3072, 768, bias=True, bn=64)
mk_lin("vit2_out_bn64", 864, 768, 768, bias=True, bn=64)
mk_lin("vit2_dgrad_qkv_bn64", 864, 2304, 768, bn=64)
mk_lin("l3_conv1_512_128", 6912, 512, 128)
mk_lin("l3_conv1_512_128_bn64", 6912, 512, 128, bn=64)
mk_lin("l3_conv3_128_512", 6912, 128, 512)
mk_lin("l3_conv3_128_512_bn64", 6912, 128, 512, bn=64)
res = {}
for name, fn, flops, bytes_ in shapes:
    fn(); torch.cuda.synchronize()
    if once:
        continue
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    res[name] = dict(ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1), gbs=round(bytes_ / ms / 1e6, 1))
    print(f"{name:32s} {ms:8.3f} ms {flops / ms / 1e9:8.1f} TFLOP/s {bytes_ / ms / 1e6:8.1f} GB/s(alg)", flush=True)
json.dump(res, open("gpurun_out/bench_shapes.json", "w"), indent=1)
