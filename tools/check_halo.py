"""Correctness + speed probe of the halo-reuse conv kernel (CTU_CONV_HALO=1/2) against the per-tap kernel (=0)."""
import os
import sys
sys.path.insert(0, ".")
import torch
import torch.nn.functional as F
from hybrid_ctunet_b200 import ops

torch.cuda.init()
print("CTU_CONV_HALO", os.environ.get("CTU_CONV_HALO"), "variant", os.environ.get("CTU_CONV_HALO_VARIANT"))
torch.backends.cudnn.allow_tf32 = False
for (B, X, Y, Z, ci, co) in [(1, 4, 16, 8, 64, 64), (2, 6, 32, 24, 64, 64), (1, 5, 16, 16, 128, 64), (1, 3, 16, 8, 128, 128)]:
    g = torch.Generator(device="cuda").manual_seed(X + ci)
    a = torch.randn(B, X, Y, Z, ci, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(co, ci, 3, 3, 3, device="cuda", generator=g) / (27 * ci) ** 0.5
    pw = ops.pack_matrix(w.permute(0, 2, 3, 4, 1).reshape(co, -1), ksize=3, a_c=ci)
    out = torch.full((B, X, Y, Z, co), float("nan"), device="cuda", dtype=torch.bfloat16)
    st = torch.zeros(B, co, 2, device="cuda", dtype=torch.float64)
    ops.gemm(a, pw, out, dims=(Z, Y, X, B), stats=st)
    ref = F.conv3d(a.float().permute(0, 4, 1, 2, 3), w.to(torch.bfloat16).float(), padding=1).permute(0, 2, 3, 4, 1)
    rel = ((out.float() - ref).norm() / ref.norm()).item()
    o = out.double().reshape(B, -1, co)
    ok_st = torch.allclose(st[..., 0], o.sum(1), rtol=1e-6, atol=1e-3)
    print(f"shape {(B, X, Y, Z, ci, co)} rel {rel:.5f} finite {bool(torch.isfinite(out.float()).all())} stats {ok_st}")
for (B, X, Y, Z, ci, co) in [(2, 96, 96, 96, 64, 64), (2, 48, 48, 96, 128, 128), (2, 96, 96, 96, 128, 64), (2, 48, 48, 96, 64, 64)]:
    a = torch.randn(B, X, Y, Z, ci, device="cuda").to(torch.bfloat16)
    w = torch.randn(co, 27 * ci, device="cuda") * 0.02
    pw = ops.pack_matrix(w, ksize=3, a_c=ci)
    out = torch.empty(B, X, Y, Z, co, device="cuda", dtype=torch.bfloat16)
    st = torch.zeros(B, co, 2, device="cuda", dtype=torch.float64)
    for _ in range(3):
        ops.gemm(a, pw, out, dims=(Z, Y, X, B), stats=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.gemm(a, pw, out, dims=(Z, Y, X, B), stats=st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{ms:8.4f} ms {2.0 * B * X * Y * Z * 27 * ci * co / ms / 1e9:7.1f} TF/s conv {ci}->{co} @ {(X, Y, Z)} B={B}")
