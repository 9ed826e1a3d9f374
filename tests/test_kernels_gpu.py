"""Parity of the HBM-bound / CUDA-core kernels against fp32 torch restatements of the reference layers.
Tolerances: outputs are stored in bf16 (rel 2^-8 per element => 2^-7 max-abs relative to the output scale);
statistics are fp32 partial sums merged in fp64 (1e-6 relative); blend is bit-exact."""
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


def _maxrel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def _gen(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


@pytest.mark.parametrize("B,S,C", [(2, 1000, 64), (1, 3456, 512), (3, 77, 1024), (1, 4096, 128)])
def test_in_stats_and_apply(B, S, C):
    ops = _ops()
    g = _gen(S + C)
    x = (torch.randn(B, S, C, device="cuda", generator=g) * 2 + 0.5).to(torch.bfloat16)
    r = torch.randn(B, S, C, device="cuda", generator=g).to(torch.bfloat16)
    st = torch.zeros(B, C, 2, device="cuda", dtype=torch.float64)
    rs = torch.zeros(B, C, 2, device="cuda", dtype=torch.float64)
    ops.in_stats(x, st)
    ops.in_stats(r, rs)
    xd = x.double()
    # per-thread partial sums are fp32 (<= ~100 values each), merged in fp64
    assert torch.allclose(st[..., 0], xd.sum(1), rtol=1e-6, atol=1e-3)
    assert torch.allclose(st[..., 1], (xd * xd).sum(1), rtol=1e-6, atol=1e-3)

    def inorm(t):  # nn.InstanceNorm3d over the voxel axis
        return F.instance_norm(t.float().permute(0, 2, 1), eps=1e-5).permute(0, 2, 1)

    out = torch.empty_like(x)
    ops.in_apply(x, st, out, act=True)
    assert _maxrel(out.float(), F.leaky_relu(inorm(x), 0.01)) < 2 ** -7
    ops.in_apply(x, st, out, res=r, act=True)
    assert _maxrel(out.float(), F.leaky_relu(inorm(x) + r.float(), 0.01)) < 2 ** -7
    ops.in_apply(x, st, out, res=r, rstats=rs, act=True)
    assert _maxrel(out.float(), F.leaky_relu(inorm(x) + inorm(r), 0.01)) < 2 ** -7
    ops.in_apply(x, st, out, act=False)
    assert _maxrel(out.float(), inorm(x)) < 2 ** -7
    # strided output (concat by offset)
    big = torch.zeros(B, S, 2 * C, device="cuda", dtype=torch.bfloat16)
    ops.in_apply(x, st, big[..., C:], act=True)
    assert torch.equal(big[..., C:], ops.in_apply(x, st, out, act=True))
    assert (big[..., :C] == 0).all()


@pytest.mark.parametrize("M,C,fin,fout", [(432, 768, True, False), (864, 2048, False, False), (3456, 512, False, False),
                                          (1000, 128, True, True), (27648, 256, False, False), (50, 96, True, False)])
def test_layernorm(M, C, fin, fout):
    ops = _ops()
    g = _gen(M + C)
    x = torch.randn(M, C, device="cuda", generator=g) * 3 + 1
    if not fin:
        x = x.to(torch.bfloat16)
    gamma = torch.randn(C, device="cuda", generator=g)
    beta = torch.randn(C, device="cuda", generator=g)
    out = torch.empty(M, C, device="cuda", dtype=torch.float32 if fout else torch.bfloat16)
    ops.layernorm(x, gamma, beta, out)
    ref = F.layer_norm(x.float(), (C,), gamma, beta)
    assert _maxrel(out.float(), ref) < (1e-5 if fout else 2 ** -7)


def test_layernorm_add_pos():
    ops = _ops()
    g = _gen(5)
    B, n, C = 2, 432, 768
    x = torch.randn(B * n, C, device="cuda", generator=g)
    gamma, beta = torch.randn(C, device="cuda", generator=g), torch.randn(C, device="cuda", generator=g)
    pos = torch.randn(1, n, C, device="cuda", generator=g)
    out = torch.empty(B * n, C, device="cuda")
    ops.layernorm(x, gamma, beta, out, add=pos)
    ref = (F.layer_norm(x, (C,), gamma, beta).view(B, n, C) + pos).view(B * n, C)
    assert _maxrel(out, ref) < 1e-5


@pytest.mark.parametrize("pf,Z", [(8, 96), (16, 96), (8, 48)])
def test_patchify_ln(pf, Z):
    ops = _ops()
    g = _gen(pf)
    B, X, Y = 2, 96, 96
    img = torch.randn(B, 1, X, Y, Z, device="cuda", generator=g)
    D = 256 * pf
    gamma, beta = torch.randn(D, device="cuda", generator=g), torch.randn(D, device="cuda", generator=g)
    tokens = (X // 16) * (Y // 16) * (Z // pf)
    out = torch.empty(B * tokens, D, device="cuda", dtype=torch.bfloat16)
    ops.patchify_ln(img, pf, gamma, beta, out)
    t = img.reshape(B, 1, X // 16, 16, Y // 16, 16, Z // pf, pf).permute(0, 2, 4, 6, 3, 5, 7, 1).reshape(B * tokens, D)
    ref = F.layer_norm(t, (D,), gamma, beta)
    assert _maxrel(out.float(), ref) < 2 ** -7


@pytest.mark.parametrize("T,C", [(3456, 512), (1001, 128), (27648, 256)])
def test_pwa_fuse(T, C):
    ops = _ops()
    g = _gen(T)
    q1 = torch.randn(T, 3 * C, device="cuda", generator=g).to(torch.bfloat16)
    q2 = torch.randn(T, 3 * C, device="cuda", generator=g).to(torch.bfloat16)
    out = torch.empty(T, C, device="cuda", dtype=torch.bfloat16)
    ops.pwa_fuse(q1, q2, out)
    h = C // 32
    a, b = q1.float().view(T, 3, h, 32), q2.float().view(T, 3, h, 32)
    d1 = (b[:, 0] * a[:, 1]).sum(-1, keepdim=True) * 32 ** -0.5
    d2 = (a[:, 0] * b[:, 1]).sum(-1, keepdim=True) * 32 ** -0.5
    w = torch.cat((d1, d2), -1).softmax(-1)
    ref = (w[..., 0:1] * a[:, 2] + w[..., 1:2] * b[:, 2]).reshape(T, C)
    assert _maxrel(out.float(), ref) < 2 ** -7


def test_subsample():
    ops = _ops()
    g = _gen(9)
    x = torch.randn(2, 9, 12, 24, 64, device="cuda", generator=g).to(torch.bfloat16)
    out = torch.empty(2, 5, 6, 24, 64, device="cuda", dtype=torch.bfloat16)
    ops.subsample(x, out, (2, 2, 1))
    assert torch.equal(out, x[:, ::2, ::2, ::1])
    out2 = torch.empty(2, 5, 6, 12, 64, device="cuda", dtype=torch.bfloat16)
    ops.subsample(x, out2, (2, 2, 2))
    assert torch.equal(out2, x[:, ::2, ::2, ::2])


@pytest.mark.parametrize("B,n,heads,dh", [(2, 432, 12, 64), (1, 100, 2, 64), (3, 216, 4, 32)])
def test_attention_linear(B, n, heads, dh):
    ops = _ops()
    g = _gen(n + heads)
    C = heads * dh
    qkv = torch.randn(B * n, 3 * C, device="cuda", generator=g).to(torch.bfloat16)
    out = torch.full((B * n, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.attention(qkv, out, dim_head=dh, n=n, windows=B, mode=0)
    q, k, v = (t.reshape(B, n, heads, dh).permute(0, 2, 1, 3) for t in qkv.float().chunk(3, -1))
    ref = ((q @ k.transpose(-1, -2)) * dh ** -0.5).softmax(-1) @ v
    ref = ref.permute(0, 2, 1, 3).reshape(B * n, C)
    assert torch.isfinite(out.float()).all()
    assert _maxrel(out.float(), ref) < 2 ** -6


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("B,X,Y,Z,C", [(1, 6, 6, 12, 768), (2, 12, 12, 24, 128)])
def test_attention_windows(mode, B, X, Y, Z, C):
    ops = _ops()
    from oracle.ctunet_oracle import rel_pos_indices
    g = _gen(mode * 100 + X)
    w, dh = 6, 32
    heads = C // dh
    qkv = torch.randn(B * X * Y * Z, 3 * C, device="cuda", generator=g).to(torch.bfloat16)
    emb = torch.randn((2 * w - 1) ** 3, heads, device="cuda", generator=g)
    bias = emb[rel_pos_indices(w).cuda()].permute(2, 0, 1).contiguous()  # [heads, 216, 216]
    out = torch.full((B * X * Y * Z, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.attention(qkv, out, dim_head=dh, n=w ** 3, mode=mode, bias=bias, grid=(B, X, Y, Z), w=w)
    t = qkv.float().view(B, X, Y, Z, 3 * C)
    if mode == 1:  # 'b (h h1) (w w1) (f f1) c -> b h w f (h1 w1 f1) c'
        t = t.view(B, X // w, w, Y // w, w, Z // w, w, 3 * C).permute(0, 1, 3, 5, 2, 4, 6, 7)
    else:          # 'b (h1 h) (w1 w) (f1 f) c -> b h w f (h1 w1 f1) c'
        t = t.view(B, w, X // w, w, Y // w, w, Z // w, 3 * C).permute(0, 2, 4, 6, 1, 3, 5, 7)
    t = t.reshape(-1, w ** 3, 3 * C)
    q, k, v = (p.reshape(t.shape[0], w ** 3, heads, dh).permute(0, 2, 1, 3) for p in t.chunk(3, -1))
    o = ((q * dh ** -0.5) @ k.transpose(-1, -2) + bias).softmax(-1) @ v
    o = o.permute(0, 2, 1, 3).reshape(B, X // w, Y // w, Z // w, w, w, w, C)
    if mode == 1:
        ref = o.permute(0, 1, 4, 2, 5, 3, 6, 7).reshape(B * X * Y * Z, C)
    else:
        ref = o.permute(0, 4, 1, 5, 2, 6, 3, 7).reshape(B * X * Y * Z, C)
    assert torch.isfinite(out.float()).all()
    assert _maxrel(out.float(), ref) < 2 ** -6


@pytest.mark.parametrize("k,s,p,shape", [((7, 7, 7), (2, 2, 1), (3, 3, 3), (1, 32, 32, 48)),
                                         ((3, 3, 3), (1, 1, 1), (1, 1, 1), (2, 16, 20, 24)),
                                         ((1, 1, 1), (1, 1, 1), (0, 0, 0), (1, 8, 8, 16))])
def test_conv_cin1(k, s, p, shape):
    ops = _ops()
    g = _gen(k[0])
    B, X, Y, Z = shape
    x = torch.randn(B, 1, X, Y, Z, device="cuda", generator=g)
    w = torch.randn(64, 1, *k, device="cuda", generator=g) / (k[0] * k[1] * k[2]) ** 0.5
    ref = F.conv3d(x, w, stride=s, padding=p).permute(0, 2, 3, 4, 1)
    out = torch.full(ref.shape, float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.conv_cin1(x, w.reshape(64, -1).t().contiguous(), out, k=k, s=s, p=p)
    assert _maxrel(out.float(), ref) < 2 ** -7


def test_blend_bit_exact():
    ops = _ops()
    g = _gen(21)
    C, r, X, Y, Z = 14, (16, 16, 32), 40, 24, 48
    imp = torch.rand(*r, device="cuda", generator=g)
    acc0 = torch.zeros(C, X, Y, Z, device="cuda")
    acc1 = torch.zeros(C, X, Y, Z, device="cuda")
    cnt = torch.zeros(X, Y, Z, device="cuda")
    ref0, ref1, refc = acc0.clone(), acc1.clone(), torch.zeros(C, X, Y, Z, device="cuda")
    starts = [(0, 0, 0), (8, 0, 16), (24, 8, 16), (3, 5, 7)]  # last one exercises the unaligned scalar path
    for st in starts:
        l0 = torch.randn(C, *r, device="cuda", generator=g)
        l1 = torch.randn(C, *r, device="cuda", generator=g)
        ops.blend_accumulate(l0, l1, imp, acc0, acc1, st)
        ops.blend_count(imp, cnt, st)
        sl = (slice(None), slice(st[0], st[0] + r[0]), slice(st[1], st[1] + r[1]), slice(st[2], st[2] + r[2]))
        ref0[sl] += imp * l0   # trainer_CTUNet.py:542
        ref1[sl] += imp * l1   # trainer_CTUNet.py:544
        refc[sl] += imp        # trainer_CTUNet.py:543
    assert torch.equal(acc0, ref0) and torch.equal(acc1, ref1)
    assert torch.equal(cnt, refc[0])
    cnt.clamp_(min=1e-3)
    refc.clamp_(min=1e-3)
    out = torch.empty_like(acc0)
    ops.blend_normalize(acc0, cnt, out)
    assert torch.equal(out, ref0 / refc)


def test_cin1_k1_stats_equal_the_sums_of_the_product():
    """ctu_cin1_k1_stats: InstanceNorm sums of r[v][c] = x[v] * w[c] (vit_encoder0 conv3 with in_channels = 1,
    hybrid_CTUNet.py:75,88-91) from the moments of x, against fp64 sums of the product itself."""
    from hybrid_ctunet_b200 import ops
    torch.manual_seed(4)
    x = torch.rand(3, 1, 20, 18, 22, device="cuda") * 2 - 0.5
    w = torch.randn(64, device="cuda")
    mom = torch.zeros(3, 1, 2, device="cuda", dtype=torch.float64)
    st = torch.full((3, 64, 2), 7.0, device="cuda", dtype=torch.float64)
    ops.cin1_k1_stats(x, w, mom, st)
    r = x.double().reshape(3, -1, 1) * w.double()
    assert torch.allclose(st[..., 0], r.sum(1), rtol=1e-10, atol=1e-8)
    assert torch.allclose(st[..., 1], (r * r).sum(1), rtol=1e-10, atol=1e-8)
