"""Parity of the tcgen05 contraction kernel (GEMM / 3x3x3 conv / ConvTranspose) against fp32 torch ops that
restate the reference's layers (nn.Linear, Conv3d(bias=False), ConvTranspose3d(k=s)) on bf16-rounded operands.
bf16 x bf16 products are exact in fp32, so the only differences are accumulation order and the bf16 rounding of
the stored output: tolerance 2^-8 relative to the output scale for bf16 outputs, 1e-4 for fp32 outputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _strict_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _ops():
    from hybrid_ctunet_b200 import ops
    return ops


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _maxrel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("M,K,N,bn", [(432, 768, 2304, 128), (128, 64, 64, 64), (1000, 3072, 768, 128),
                                      (3456, 96, 512, 128), (4096, 32, 64, 64), (864, 2048, 768, 64),
                                      (300, 256, 256, 256), (260, 128, 32, 32)])
def test_linear_bf16(M, K, N, bn):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5)
    pw = ops.pack_matrix(w, block_n=bn)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, pw, out, dims=(M, 1, 1, 1))
    ref = a.float() @ w.to(torch.bfloat16).float().t()
    assert torch.isfinite(out.float()).all()
    assert _maxrel(out.float(), ref) < 2 ** -7
    assert _rel(out.float(), ref) < 2 ** -8


def test_linear_bias_gelu_residual_f32_inplace():
    ops = _ops()
    M, K, N = 864, 768, 768
    g = torch.Generator(device="cuda").manual_seed(7)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    x = torch.randn(M, N, device="cuda", generator=g)
    pw = ops.pack_matrix(w, bias=b)
    ref = x + F.linear(a.float(), w.to(torch.bfloat16).float(), b)
    xx = x.clone()
    ops.gemm(a, pw, xx, dims=(M, 1, 1, 1), out_mode=ops.OUT_F32, residual=xx)
    assert _maxrel(xx, ref) < 1e-4
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, pw, out, dims=(M, 1, 1, 1), act=ops.ACT_GELU)
    ref2 = F.gelu(F.linear(a.float(), w.to(torch.bfloat16).float(), b))
    assert _maxrel(out.float(), ref2) < 2 ** -7


@pytest.mark.parametrize("M,K,N", [(864, 768, 3072), (4000, 512, 128), (300, 128, 512)])
def test_linear_gelu_backward_epilogue(M, K, N):
    """Input gradient of Linear -> GELU -> Linear: out = (dY W) * gelu'(x) in the GEMM epilogue (CTU_RES_GELU_BWD),
    against torch autograd of the exact-erf GELU (vit.py:37)."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(M + N)
    dy = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5      # plays W^T of the down-projection
    x = (torch.randn(M, N, device="cuda", generator=g) * 1.5).to(torch.bfloat16)
    pw = ops.pack_matrix(w)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.gemm(dy, pw, out, dims=(M, 1, 1, 1), gelu_bwd_of=x)
    xr = x.float().requires_grad_()
    F.gelu(xr).backward(dy.float() @ w.to(torch.bfloat16).float().t())
    assert torch.isfinite(out.float()).all()
    assert _rel(out.float(), xr.grad) < 2 ** -8
    assert _maxrel(out.float(), xr.grad) < 2 ** -6


def test_head_channel_first_f32():
    ops = _ops()
    B, X, Y, Z, Cin, Cout = 2, 8, 12, 16, 64, 14
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn(B, X, Y, Z, Cin, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(Cout, Cin, device="cuda", generator=g) / 8
    b = torch.randn(Cout, device="cuda", generator=g)
    pw = ops.pack_matrix(w, bias=b)
    out = torch.full((B, Cout, X, Y, Z), float("nan"), device="cuda")
    ops.gemm(a, pw, out, dims=(X * Y * Z, 1, 1, B), out_mode=ops.OUT_F32_CF)
    ref = F.conv3d(a.float().permute(0, 4, 1, 2, 3), w.to(torch.bfloat16).float()[:, :, None, None, None], b)
    assert _maxrel(out, ref) < 1e-4


def _pack_conv3(w):  # [Cout, Cin, kX, kY, kZ] -> [Cout, 27*Cin] (tap-major, channel-minor)
    return w.permute(0, 2, 3, 4, 1).reshape(w.shape[0], -1)


@pytest.mark.parametrize("B,X,Y,Z,Cin,Cout,bn", [(1, 8, 8, 32, 64, 64, 64), (2, 6, 6, 12, 128, 128, 128),
                                                  (1, 12, 12, 24, 64, 128, 64), (1, 24, 24, 48, 128, 64, 64),
                                                  (1, 5, 7, 9, 64, 64, 64), (1, 96, 96, 96, 64, 64, 64),
                                                  (2, 3, 32, 32, 128, 64, 64), (1, 4, 16, 48, 64, 64, 64),
                                                  # two-plane kernel (X even, Y % 16 == 0, Z % 8 == 0, block_n 64): several
                                                  # K blocks, two N tiles, a single plane pair, batch > 1
                                                  (2, 6, 16, 8, 128, 64, 64), (1, 2, 32, 16, 64, 128, 64),
                                                  (3, 4, 16, 16, 192, 64, 64)])
def test_conv3x3x3(B, X, Y, Z, Cin, Cout, bn):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(X * Y + Cin)
    a = torch.randn(B, X, Y, Z, Cin, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(Cout, Cin, 3, 3, 3, device="cuda", generator=g) / (27 * Cin) ** 0.5
    pw = ops.pack_matrix(_pack_conv3(w), ksize=3, a_c=Cin, block_n=bn)
    out = torch.full((B, X, Y, Z, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    stats = torch.zeros(B, Cout, 2, device="cuda", dtype=torch.float64)
    ops.gemm(a, pw, out, dims=(Z, Y, X, B), stats=stats)
    ref = F.conv3d(a.float().permute(0, 4, 1, 2, 3), w.to(torch.bfloat16).float(), padding=1).permute(0, 2, 3, 4, 1)
    assert torch.isfinite(out.float()).all()
    assert _maxrel(out.float(), ref) < 2 ** -7
    assert _rel(out.float(), ref) < 2 ** -8
    # fused InstanceNorm statistics are the sums of the stored (bf16) values
    o = out.double().reshape(B, -1, Cout)
    assert torch.allclose(stats[..., 0], o.sum(1), rtol=1e-6, atol=1e-3)
    assert torch.allclose(stats[..., 1], (o * o).sum(1), rtol=1e-6, atol=1e-3)


def test_conv3x3x3_skips_zero_padded_channels():
    """ResNet layer-1 shape class: 32 live channels in 64-channel rows (resnet.py:181-186).  With a_c_live = 32 the halo
    kernel issues half of the K steps; the result must equal the full-K result exactly (the skipped products are 0)."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(5)
    B, X, Y, Z = 1, 6, 16, 32
    a = torch.randn(B, X, Y, Z, 64, device="cuda", generator=g).to(torch.bfloat16)
    a[..., 32:] = 0
    w = torch.randn(64, 64, 3, 3, 3, device="cuda", generator=g) / (27 * 32) ** 0.5
    w[32:] = 0
    w[:, 32:] = 0
    outs = []
    for live in (0, 32):
        pw = ops.pack_matrix(_pack_conv3(w), ksize=3, a_c=64, block_n=64)
        pw.a_c_live = live
        pw.x3 = None      # both runs on the one-plane kernel (the two-plane kernel has no K-skip and another summation order)
        out = torch.full((B, X, Y, Z, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
        stats = torch.zeros(B, 64, 2, device="cuda", dtype=torch.float64)
        ops.gemm(a, pw, out, dims=(Z, Y, X, B), stats=stats)
        outs.append((out, stats))
    ref = F.conv3d(a.float().permute(0, 4, 1, 2, 3), w.to(torch.bfloat16).float(), padding=1).permute(0, 2, 3, 4, 1)
    assert _rel(outs[0][0].float(), ref) < 2 ** -8
    assert torch.equal(outs[0][0], outs[1][0]) and float(outs[1][0][..., 32:].abs().max()) == 0.0
    assert torch.allclose(outs[0][1], outs[1][1], rtol=1e-9, atol=1e-6)


@pytest.mark.parametrize("B,X,Y,Z,Cin,Cout,bn", [(1, 8, 16, 32, 64, 64, 64), (2, 6, 32, 16, 128, 128, 128),
                                                  (1, 12, 16, 24, 64, 128, 64), (1, 5, 7, 9, 64, 64, 64),
                                                  (1, 6, 6, 12, 256, 256, 128)])
def test_conv3x3x3_accumulates_in_place(B, X, Y, Z, Cin, Cout, bn):
    """out = conv(a) + out, in place (the input-gradient convolution adding to a gradient that has already arrived):
    the halo-reuse kernel where the shape allows it (Y % 16 == 0, Z % 8 == 0), the per-tap kernel otherwise."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(X * Y + Cout)
    a = torch.randn(B, X, Y, Z, Cin, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(Cout, Cin, 3, 3, 3, device="cuda", generator=g) / (27 * Cin) ** 0.5
    pw = ops.pack_matrix(_pack_conv3(w), ksize=3, a_c=Cin, block_n=bn)
    cur = torch.randn(B, X, Y, Z, Cout, device="cuda", generator=g).to(torch.bfloat16)
    ref = F.conv3d(a.float().permute(0, 4, 1, 2, 3), w.to(torch.bfloat16).float(), padding=1).permute(0, 2, 3, 4, 1) + cur.float()
    ops.gemm(a, pw, cur, dims=(Z, Y, X, B), residual=cur)
    assert torch.isfinite(cur.float()).all()
    assert _maxrel(cur.float(), ref) < 2 ** -7
    assert _rel(cur.float(), ref) < 2 ** -8


@pytest.mark.parametrize("B,X,Y,Z,Cin,Cout,u", [(1, 6, 6, 12, 128, 64, (2, 2, 2)), (2, 4, 6, 8, 128, 64, (2, 2, 1)),
                                                (1, 6, 6, 12, 1024, 512, (2, 2, 2))])
def test_conv_transpose_k_eq_s(B, X, Y, Z, Cin, Cout, u):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(Cin + Cout)
    a = torch.randn(B, X, Y, Z, Cin, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(Cin, Cout, *u, device="cuda", generator=g) / Cin ** 0.5
    ux, uy, uz = u
    w2 = w.permute(2, 3, 4, 1, 0).reshape(ux * uy * uz * Cout, Cin)
    pw = ops.pack_matrix(w2, block_n=64, convt=(Cout, uz, uy, ux))
    out = torch.full((B, X * ux, Y * uy, Z * uz, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, pw, out, dims=(Z, Y, X, B))
    ref = F.conv_transpose3d(a.float().permute(0, 4, 1, 2, 3), w.to(torch.bfloat16).float(), stride=u)
    ref = ref.permute(0, 2, 3, 4, 1)
    assert torch.isfinite(out.float()).all()
    assert _maxrel(out.float(), ref) < 2 ** -7


def test_concat_by_offset_and_strided_input():
    ops = _ops()
    M, K, N = 512, 64, 64
    g = torch.Generator(device="cuda").manual_seed(11)
    big = torch.randn(M, 128, device="cuda", generator=g).to(torch.bfloat16)
    a = big[:, 64:]  # row stride 128, channel offset 64
    w = torch.randn(N, K, device="cuda", generator=g) / 8
    pw = ops.pack_matrix(w)
    out = torch.zeros(M, 128, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, pw, out, dims=(M, 1, 1, 1), out_col0=64)
    ref = a.float() @ w.to(torch.bfloat16).float().t()
    assert _maxrel(out[:, 64:].float(), ref) < 2 ** -7
    assert (out[:, :64] == 0).all()


@pytest.mark.parametrize("M,hidden,ld_extra", [(128, 512, 0), (1000, 512, 0), (128 * 148 * 3 + 77, 512, 0), (4096, 256, 64),
                                               (20000, 384, 0)])
def test_ffn_fused_matches_the_two_gemm_path_and_fp32(M, hidden, ld_extra):
    """ctu_ffn_fused (hidden activation on chip: TMEM -> GELU -> shared-memory operand of the second GEMM) against
    (a) x + W2 gelu(W1 a + b1) + b2 in fp32 with the hidden activation rounded to bf16 where the kernel rounds it, and
    (b) the two-GEMM path of the same library — hybrid_CTUNet.py:513-526 inside Residual (:434-440).  Ragged last tile,
    several tiles per CTA (persistent loop), strided rows (views into wider buffers)."""
    ops = _ops()
    C = 128
    g = torch.Generator(device="cuda").manual_seed(M + hidden)
    w1 = torch.randn(hidden, C, device="cuda", generator=g) / C ** 0.5
    b1 = torch.randn(hidden, device="cuda", generator=g) * 0.5
    w2 = torch.randn(C, hidden, device="cuda", generator=g) / hidden ** 0.5
    b2 = torch.randn(C, device="cuda", generator=g) * 0.5
    a = torch.randn(M, C + ld_extra, device="cuda", generator=g).to(torch.bfloat16)[:, :C]
    x = torch.randn(M, C + ld_extra, device="cuda", generator=g).to(torch.bfloat16)[:, :C]
    p1, p2 = ops.pack_matrix(w1, bias=b1), ops.pack_matrix(w2, bias=b2)
    out_full = torch.full((M, C + ld_extra), float("nan"), device="cuda", dtype=torch.bfloat16)
    out = out_full[:, :C]
    ops.ffn_fused(a, p1, p2, x, out)
    torch.cuda.synchronize()
    h = F.gelu(F.linear(a.float(), w1.to(torch.bfloat16).float(), b1)).to(torch.bfloat16).float()
    ref = x.float() + F.linear(h, w2.to(torch.bfloat16).float(), b2)
    assert torch.isfinite(out.float()).all()
    assert _maxrel(out.float(), ref) < 2 ** -7 and _rel(out.float(), ref) < 2 ** -8
    if ld_extra:
        assert torch.isnan(out_full[:, C:].float()).all()          # columns outside the view are untouched
    hid = torch.empty(M, hidden, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a.contiguous(), p1, hid, dims=(M, 1, 1, 1), act=ops.ACT_GELU)
    two = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
    ops.gemm(hid, p2, two, dims=(M, 1, 1, 1), residual=x.contiguous())
    assert _rel(out.float(), two.float()) < 2 ** -8
