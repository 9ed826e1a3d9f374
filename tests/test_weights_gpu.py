"""Multi-tensor weight packing / gradient unpacking kernels (ctu_pack_weights, ctu_unpack_grads) against torch
permutes of the same tensors — bit-exact (pure data movement + one fp32 -> bf16 rounding).

The layouts are the ones engine.WeightCache / Engine._finalize_param_grads register: forward [N_pad][taps * C_in]
K-major, transposed / tap-flipped for the input-gradient GEMMs, and the [(tap, c_in)][c_out] fp32 accumulators of the
wgrad kernels (reference parameter layouts: nn.Linear [N, K], Conv3d [Co, Ci, 3, 3, 3] — networks/vit.py:31-78,
networks/resnet.py:17-126, networks/hybrid_CTUNet.py:29-105).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _pad64(v):
    return -(-v // 64) * 64


def _table(unpack=False):
    from hybrid_ctunet_b200.engine import ItemTable
    return ItemTable(torch.device("cuda"), unpack=unpack)


@pytest.mark.parametrize("n,k", [(768, 3072), (2304, 768), (14, 64), (100, 72), (64, 2048), (130, 200), (30, 50)])
def test_pack_linear_and_transpose(n, k):
    from hybrid_ctunet_b200 import engine as E
    torch.manual_seed(n + k)
    w = torch.randn(n, k, device="cuda")
    n_pad, k_pad = _pad64(n), -(-k // 8) * 8
    fwd = torch.full((n_pad, k_pad), 7.0, device="cuda", dtype=BF)
    kt_rows, kt_cols = k, (64 if n < 64 else -(-n // 8) * 8)
    tr = torch.full((_pad64(kt_rows), kt_cols), 7.0, device="cuda", dtype=BF)
    t = _table()
    t.add(w.data_ptr(), fwd.data_ptr(), E.PACK_LIN, n_pad, k_pad, n, k, 0)
    t.add(w.data_ptr(), tr.data_ptr(), E.PACK_LIN_T, tr.shape[0], kt_cols, n, k, 0)
    t.run("ctu_pack_weights")
    ref = torch.zeros(n_pad, k_pad, device="cuda")
    ref[:n, :k] = w
    assert torch.equal(fwd, ref.to(BF))
    ref_t = torch.zeros(tr.shape[0], kt_cols, device="cuda")
    ref_t[:k, :n] = w.t()
    assert torch.equal(tr, ref_t.to(BF))


@pytest.mark.parametrize("co,ci", [(64, 64), (128, 256), (32, 32), (64, 128), (512, 512), (70, 35)])
def test_pack_conv3_and_flipped_transpose(co, ci):
    from hybrid_ctunet_b200 import engine as E
    torch.manual_seed(co * 3 + ci)
    w = torch.randn(co, ci, 3, 3, 3, device="cuda")
    cop, cip = _pad64(co), _pad64(ci)
    fwd = torch.full((cop, 27 * cip), 7.0, device="cuda", dtype=BF)
    tr = torch.full((cip, 27 * cop), 7.0, device="cuda", dtype=BF)
    t = _table()
    t.add(w.data_ptr(), fwd.data_ptr(), E.PACK_CONV3, cop, 27 * cip, co, ci, 0)
    t.add(w.data_ptr(), tr.data_ptr(), E.PACK_CONV3_T, cip, 27 * cop, co, ci, 0)
    t.run("ctu_pack_weights")
    ref = torch.zeros(cop, 27, cip, device="cuda")
    ref[:co, :, :ci] = w.reshape(co, ci, 27).permute(0, 2, 1)
    assert torch.equal(fwd, ref.reshape(cop, -1).to(BF))
    ref_t = torch.zeros(cip, 27, cop, device="cuda")
    ref_t[:ci, :, :co] = w.reshape(co, ci, 27).flip(-1).permute(1, 2, 0)
    assert torch.equal(tr, ref_t.reshape(cip, -1).to(BF))


@pytest.mark.parametrize("n,k,ld", [(768, 3072, 768), (14, 64, 64), (100, 72, 128), (3072, 768, 3072), (130, 200, 192)])
def test_unpack_linear(n, k, ld):
    from hybrid_ctunet_b200 import engine as E
    torch.manual_seed(n + 2 * k)
    buf = torch.randn(k + 5, ld, device="cuda")          # [K'][ld]: the wgrad accumulator (rows >= K, pitch >= N)
    # the destination is a slice of the flat gradient buffer: only 4-byte aligned in general
    for off in (0, 2):
        g = torch.full((n * k + off,), 7.0, device="cuda")[off:].view(n, k)
        t = _table(unpack=True)
        t.add(buf.data_ptr(), g.data_ptr(), E.PACK_LIN, n * k, ld, n, k, 0)
        t.run("ctu_unpack_grads")
        assert torch.equal(g, buf[:k, :n].t())


@pytest.mark.parametrize("co,ci", [(64, 64), (128, 256), (32, 32), (512, 512), (70, 35)])
def test_unpack_conv3(co, ci):
    from hybrid_ctunet_b200 import engine as E
    torch.manual_seed(co + 7 * ci)
    cop, cip = _pad64(co), _pad64(ci)
    buf = torch.randn(27 * cip, cop, device="cuda")      # [(tap, ci)][co]
    g = torch.full((co, ci, 3, 3, 3), 7.0, device="cuda")
    t = _table(unpack=True)
    t.add(buf.data_ptr(), g.data_ptr(), E.PACK_CONV3, g.numel(), cop, co, ci, cip)
    t.run("ctu_unpack_grads")
    ref = buf.view(27, cip, cop)[:, :ci, :co].permute(2, 1, 0).reshape(co, ci, 3, 3, 3)
    assert torch.equal(g, ref)


def test_mixed_table_matches_single_item_runs():
    """Several kinds in one table (binary search over unit0) == each item packed alone."""
    from hybrid_ctunet_b200 import engine as E
    torch.manual_seed(3)
    w1 = torch.randn(256, 128, device="cuda")
    w2 = torch.randn(64, 64, 3, 3, 3, device="cuda")
    w3 = torch.randn(128, 64, 2, 2, 2, device="cuda")   # ConvTranspose3d [ci, co, k...]
    outs = [torch.zeros(128, 256, device="cuda", dtype=BF), torch.zeros(64, 27 * 64, device="cuda", dtype=BF),
            torch.zeros(64, 27 * 64, device="cuda", dtype=BF), torch.zeros(8 * 64, 128, device="cuda", dtype=BF)]
    specs = [(w1, outs[0], E.PACK_LIN_T, 128, 256, 256, 128, 0), (w2, outs[1], E.PACK_CONV3, 64, 27 * 64, 64, 64, 0),
             (w2, outs[2], E.PACK_CONV3_T, 64, 27 * 64, 64, 64, 0), (w3, outs[3], E.PACK_CONVT, 8 * 64, 128, 128, 64, 8)]
    t = _table()
    for w, o, kind, r, c, a, b, cc in specs:
        t.add(w.data_ptr(), o.data_ptr(), kind, r, c, a, b, cc)
    t.run("ctu_pack_weights")
    together = [o.clone() for o in outs]
    for i, o in enumerate(outs):
        o.zero_()
        t.run("ctu_pack_weights", only=i)
        assert torch.equal(o, together[i])
    assert torch.equal(together[0], w1.t().to(BF))
    assert torch.equal(together[3], w3.reshape(128, 64, 8).permute(2, 1, 0).reshape(8 * 64, 128).to(BF))


# ------------------------------------------------------------------------------------------------ paired layouts
def _paired_conv3_weight(w):
    """Dense [2 co][27 (pair taps)][2 ci] weight of the 3x3x3 convolution over PAIRED rows (two z-neighbours per row, slot =
    z & 1) built from the definition: output slot s_out at pair position p sees input slot s_in at pair position p + pz - 1
    through the real z tap dz = 2 (pz - 1) + s_in - s_out."""
    co, ci = w.shape[:2]
    big = torch.zeros(2, co, 3, 3, 3, 2, ci, device=w.device)      # [s_out][o][tx][ty][pz][s_in][i]
    for s_out in range(2):
        for s_in in range(2):
            for pz in range(3):
                dz = 2 * (pz - 1) + s_in - s_out
                if -1 <= dz <= 1:
                    big[s_out, :, :, :, pz, s_in, :] = w[:, :, :, :, dz + 1].permute(0, 2, 3, 1)
    return big.reshape(2 * co, 27, 2 * ci)


@pytest.mark.parametrize("co,ci", [(32, 32), (32, 64), (128, 32), (32, 128)])
def test_pack_paired_layouts(co, ci):
    """CTU_PACK_PAIR_LIN / _T (block-diagonal 1x1x1) and CTU_PACK_PAIR_CONV3 / _T (pair-tap 3x3x3) against their definitions
    (ResNet layer-1 bottlenecks on paired rows, resnet.py:181-186)."""
    from hybrid_ctunet_b200 import engine as E
    torch.manual_seed(co + 7 * ci)
    w1 = torch.randn(co, ci, device="cuda")
    lin = torch.full((2 * co, 2 * ci), 7.0, device="cuda", dtype=BF)
    lin_t = torch.full((2 * ci, 2 * co), 7.0, device="cuda", dtype=BF)
    t = _table()
    t.add(w1.data_ptr(), lin.data_ptr(), E.PACK_PAIR_LIN, 2 * co, 2 * ci, co, ci, 0)
    t.add(w1.data_ptr(), lin_t.data_ptr(), E.PACK_PAIR_LIN_T, 2 * ci, 2 * co, co, ci, 0)
    if co == ci == 32:
        w3 = torch.randn(co, ci, 3, 3, 3, device="cuda")
        c3 = torch.full((2 * co, 27 * 2 * ci), 7.0, device="cuda", dtype=BF)
        c3_t = torch.full((2 * ci, 27 * 2 * co), 7.0, device="cuda", dtype=BF)
        t.add(w3.data_ptr(), c3.data_ptr(), E.PACK_PAIR_CONV3, 2 * co, 27 * 2 * ci, co, ci, 0)
        t.add(w3.data_ptr(), c3_t.data_ptr(), E.PACK_PAIR_CONV3_T, 2 * ci, 27 * 2 * co, co, ci, 0)
    t.run("ctu_pack_weights")
    ref = torch.block_diag(w1, w1)
    assert torch.equal(lin, ref.to(BF)) and torch.equal(lin_t, ref.t().contiguous().to(BF))
    if co == ci == 32:
        big = _paired_conv3_weight(w3)                                   # [n][ptap][k]
        assert torch.equal(c3, big.reshape(2 * co, -1).to(BF))
        assert torch.equal(c3_t, big.flip(1).permute(2, 1, 0).reshape(2 * ci, -1).to(BF))   # tap-flipped transpose


def test_paired_conv_equals_the_plain_conv():
    """The pair-tap weight applied to paired rows IS the 3x3x3 convolution: F.conv3d on [B, 32, X, Y, Z] against F.conv3d
    with the paired weight on the [B, 64, X, Y, Z/2] view (fp32, no kernels of this library: pins the layout definition)."""
    import torch.nn.functional as F
    torch.manual_seed(0)
    w = torch.randn(32, 32, 3, 3, 3, device="cuda")
    x = torch.randn(2, 32, 5, 6, 8, device="cuda")
    ref = F.conv3d(x, w, padding=1)
    big = _paired_conv3_weight(w).reshape(64, 3, 3, 3, 64).permute(0, 4, 1, 2, 3)        # [n][k][tx][ty][pz]
    pair = lambda t: t.reshape(2, 32, 5, 6, 4, 2).permute(0, 5, 1, 2, 3, 4).reshape(2, 64, 5, 6, 4)   # channel = slot * 32 + c
    out = F.conv3d(pair(x), big, padding=1)
    assert torch.allclose(out, pair(ref), atol=2e-3, rtol=1e-4)   # |out| ~ 30: fp32 summation order only


def test_unpack_paired_gradients():
    """ctu_unpack_grads kinds PAIR_LIN / PAIR_CONV3: the real weight's gradient is the sum of the blocks of the paired
    weight's gradient that hold it."""
    from hybrid_ctunet_b200 import engine as E
    torch.manual_seed(3)
    co, ci = 32, 64
    buf = torch.randn(2 * ci, 2 * co, device="cuda")                   # dW_big^T: [(s, i)][(s, o)]
    g = torch.full((co, ci), 7.0, device="cuda")
    t = _table(unpack=True)
    t.add(buf.data_ptr(), g.data_ptr(), E.PACK_PAIR_LIN, co * ci, 2 * co, co, ci, 0)
    co3 = ci3 = 32
    buf3 = torch.randn(27 * 2 * ci3, 2 * co3, device="cuda")           # [(ptap, s_in, i)][(s_out, o)]
    g3 = torch.full((co3, ci3, 3, 3, 3), 7.0, device="cuda")
    t.add(buf3.data_ptr(), g3.data_ptr(), E.PACK_PAIR_CONV3, g3.numel(), 2 * co3, co3, ci3, 0)
    t.run("ctu_unpack_grads")
    assert torch.allclose(g, (buf[:ci, :co] + buf[ci:, co:]).t())
    b = buf3.reshape(3, 3, 3, 2, ci3, 2, co3)                          # [tx][ty][pz][s_in][i][s_out][o]
    ref = torch.zeros_like(g3)
    for s_out in range(2):
        for s_in in range(2):
            for pz in range(3):
                dz = 2 * (pz - 1) + s_in - s_out
                if -1 <= dz <= 1:
                    ref[:, :, :, :, dz + 1] += b[:, :, pz, s_in, :, s_out, :].permute(3, 2, 0, 1)
    assert torch.allclose(g3, ref, atol=1e-6)


def test_x3_relayout_and_stats_fold():
    """CTU_PACK_X3_FROM_PACKED == ops.x3_layout (the weight copy of the two-plane 3x3x3 kernel); ctu_stats_fold merges the two
    column halves of paired InstanceNorm sums."""
    from hybrid_ctunet_b200 import engine as E, ops
    torch.manual_seed(5)
    packed = torch.randn(128, 27 * 64, device="cuda").to(BF)
    dst = torch.full((9 * 2 * 192, 64), 7.0, device="cuda", dtype=BF)
    t = _table()
    t.add(packed.data_ptr(), dst.data_ptr(), E.PACK_X3_FROM_PACKED, dst.shape[0], 64, 128, 64, 0)
    t.run("ctu_pack_weights")
    assert torch.equal(dst, ops.x3_layout(packed, 64))
    st = torch.randn(3, 64, 2, device="cuda", dtype=torch.float64)
    ref = (st[:, :32] + st[:, 32:]) * 0.5
    ops.stats_fold(st, 32, 0.5)
    assert torch.equal(st[:, :32], ref) and torch.equal(st[:, 32:], ref)
    wide = torch.randn(2, 256, 4, device="cuda", dtype=torch.float64)
    ref = wide[:, :128] + wide[:, 128:]
    ops.stats_fold(wide, 128, 1.0)
    assert torch.equal(wide[:, :128], ref) and torch.equal(wide[:, 128:], ref)
