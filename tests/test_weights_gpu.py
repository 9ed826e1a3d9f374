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
