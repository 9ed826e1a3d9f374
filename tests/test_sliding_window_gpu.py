"""Sliding-window inference on the CUDA blend kernels against the reference restatement
(oracle/sliding_window_oracle.py follows trainer_CTUNet.py:417-557 / trainer_CUNet.py:268-400 literally, run here
with torch ops on the same device).  With the same predictor the blended volume must be BIT-EXACT: window order,
importance map, separate multiply / add roundings and the final divide all match."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _pred2(n_cls=14):
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(1, n_cls, 1, 1, 1, device="cuda", generator=g)
    b = torch.randn(1, n_cls, 1, 1, 1, device="cuda", generator=g)

    def predictor(w):
        h0 = torch.sin(w * 3.0) * a + b
        h1 = torch.cos(w * 2.0) * b - a
        return ((h0, h0[:, :, ::2, ::2], h0[:, :, ::4, ::4]), (h1, h1 + 1))
    return predictor


@pytest.mark.parametrize("shape,roi,overlap,mode", [((1, 1, 70, 50, 45), (32, 32, 32), 0.5, "gaussian"),
                                                    ((2, 1, 64, 64, 64), (32, 32, 32), 0.5, "gaussian"),
                                                    ((1, 1, 40, 33, 37), (16, 16, 16), 0.7, "gaussian"),
                                                    ((1, 1, 50, 50, 50), (32, 32, 32), 0.25, "constant"),
                                                    ((1, 1, 20, 40, 24), (32, 32, 32), 0.5, "gaussian")])
def test_two_head_blend_bit_exact(shape, roi, overlap, mode):
    from hybrid_ctunet_b200.trainer_CTUNet import sliding_window_inference
    from oracle import sliding_window_oracle as O
    g = torch.Generator(device="cuda").manual_seed(sum(shape))
    x = torch.rand(*shape, device="cuda", generator=g)
    pred = _pred2()
    a0, a1 = sliding_window_inference(x, roi, 4, pred, overlap=overlap, mode=mode)
    r0, r1 = O.sliding_window_inference(x, roi, 4, pred, overlap=overlap, mode=mode, two_heads=True)
    assert a0.shape == r0.shape == (shape[0], 14) + tuple(shape[2:])
    assert torch.equal(a0, r0) and torch.equal(a1, r1)


def test_cuda_blend_equals_the_reference_functions_output():
    """tests/golden/sliding_window_ref.npz holds what the reference's OWN sliding_window_inference
    (trainer_CTUNet.py:417-557, executed unmodified by tests/golden/make_golden.py through oracle/ref_exec.py) returns
    for the shared parity cases; the CUDA blend over the same logits (the stand-in predictor is evaluated on the CPU,
    as the fixture's was, so the logits are the same bits) must reproduce it bit for bit."""
    import os
    import numpy as np
    from hybrid_ctunet_b200.trainer_CTUNet import sliding_window_inference
    from oracle import ref_exec
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sliding_window_ref.npz"))
    cpu_pred = ref_exec.sw_case_predictor()

    def pred(w):
        out = cpu_pred(w.cpu())
        return tuple(tuple(t.cuda() for t in grp) for grp in out)

    for i, (shape, roi, swb, overlap, mode) in enumerate(ref_exec.SW_CASES):
        vol = torch.from_numpy(z[f"vol{i}"]).cuda()
        a0, a1 = sliding_window_inference(vol, roi, swb, pred, overlap=overlap, mode=mode)
        assert torch.equal(a0.cpu(), torch.from_numpy(z[f"head0_{i}"])), i
        assert torch.equal(a1.cpu(), torch.from_numpy(z[f"head1_{i}"])), i


def test_one_head_blend_bit_exact():
    from hybrid_ctunet_b200.trainer_CUNet import sliding_window_inference
    from oracle import sliding_window_oracle as O
    g = torch.Generator(device="cuda").manual_seed(77)
    x = torch.rand(1, 1, 48, 80, 40, device="cuda", generator=g)
    pred2 = _pred2()
    pred = lambda w: pred2(w)[0]
    a = sliding_window_inference(x, (32, 32, 32), 4, pred, overlap=0.5, mode="gaussian")
    r = O.sliding_window_inference(x, (32, 32, 32), 4, pred, overlap=0.5, mode="gaussian", two_heads=False)
    assert torch.equal(a, r)


def test_full_window_and_overlap_errors():
    from hybrid_ctunet_b200.trainer_CTUNet import sliding_window_inference
    x = torch.rand(1, 1, 32, 32, 32, device="cuda")
    pred = _pred2()
    a0, _ = sliding_window_inference(x, (32, 32, 32), 4, pred, overlap=0.5, mode="gaussian")
    assert torch.allclose(a0, pred(x)[0][0], rtol=1e-6, atol=1e-6)  # one window: (imp*p)/imp
    with pytest.raises(AssertionError):
        sliding_window_inference(x, (32, 32, 32), 4, pred, overlap=1.0)


def test_blend_properties_at_the_full_baseline_geometry():
    """Size-independent properties at BASELINE.json's full geometry (1x1x512x512x256, roi 96^3, overlap 0.5, gaussian,
    500 windows), where the oracle is too slow to be the checker:
    partition of unity — a predictor that returns a per-class constant is blended back to that constant;
    window indexing — a predictor that returns its own input window reconstructs the volume;
    linearity — blend(2a - 3b) == 2 blend(a) - 3 blend(b) for predictors a, b."""
    from hybrid_ctunet_b200.trainer_CTUNet import sliding_window_inference
    shape, roi = (1, 1, 512, 512, 256), (96, 96, 96)
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.rand(*shape, device="cuda", generator=g)
    const = torch.linspace(-3.0, 4.0, 14, device="cuda").view(1, 14, 1, 1, 1)
    calls = [0]

    def p_const(w):
        calls[0] += w.shape[0]
        c = const.expand(w.shape[0], 14, *w.shape[2:])
        return ((c, c, c), (c * 2, c))

    a0, a1 = sliding_window_inference(x, roi, 4, p_const, overlap=0.5, mode="gaussian")
    assert calls[0] == 500 and a0.shape == (1, 14, 512, 512, 256)
    assert (a0 - const).abs().max().item() < 1e-5 and (a1 - 2 * const).abs().max().item() < 1e-5
    del a0, a1

    def p_ident(w):   # channel c carries (c + 1) * the window itself
        v = w * torch.arange(1, 15, device="cuda", dtype=w.dtype).view(1, 14, 1, 1, 1)
        return ((v, v, v), (-v, v))

    b0, b1 = sliding_window_inference(x, roi, 4, p_ident, overlap=0.5, mode="gaussian")
    for c in (0, 6, 13):
        assert (b0[0, c] - (c + 1) * x[0, 0]).abs().max().item() < 1e-5 * (c + 1)
        assert (b1[0, c] + (c + 1) * x[0, 0]).abs().max().item() < 1e-5 * (c + 1)

    def p_lin(w):
        i, c = p_ident(w), p_const(w)
        return ((2 * i[0][0] - 3 * c[0][0],) * 3, (2 * i[1][0] - 3 * c[1][0], i[1][1]))

    l0, l1 = sliding_window_inference(x, roi, 4, p_lin, overlap=0.5, mode="gaussian")
    assert (l0 - (2 * b0 - 3 * const)).abs().max().item() < 2e-4
    assert (l1 - (2 * b1 - 6 * const)).abs().max().item() < 2e-4


def test_ctunet_sliding_window_three_windows():
    """The real predictor: CTUNet(101, pf8) over a 96x96x144 volume (3 windows, overlap 0.5), eager vs CUDA-graph
    replay vs the oracle blend of the same model's logits."""
    from hybrid_ctunet_b200.networks.hybrid_CTUNet import CTUNet
    from hybrid_ctunet_b200.trainer_CTUNet import sliding_window_inference
    from oracle import sliding_window_oracle as O
    torch.manual_seed(0)
    m = CTUNet(in_channels=1, dim_conv_stem=64, out_channels=14, model_depth=101, img_size=(96, 96), frames=96,
               patch_frame=8).cuda().eval()
    torch.manual_seed(2)
    x = torch.rand(1, 1, 96, 96, 144).cuda()
    with torch.no_grad():
        a0, a1 = sliding_window_inference(x, (96, 96, 96), 4, m, overlap=0.5, mode="gaussian")
        r0, r1 = O.sliding_window_inference(x, (96, 96, 96), 4, m, overlap=0.5, mode="gaussian", two_heads=True)
        m.enable_cuda_graph(True)
        g0, g1 = sliding_window_inference(x, (96, 96, 96), 4, m, overlap=0.5, mode="gaussian")
    assert a0.shape == (1, 14, 96, 96, 144) and torch.isfinite(a0).all() and torch.isfinite(a1).all()
    rel = lambda u, v: ((u - v).norm() / v.norm()).item()
    # the fp64 atomics of the InstanceNorm statistics make two runs of the model differ in the last bits only
    assert rel(a0, r0) < 1e-3 and rel(a1, r1) < 1e-4
    assert rel(g0, r0) < 1e-3 and rel(g1, r1) < 1e-4


def test_ensemble_masks_match_reference_arithmetic():
    """argmax masks bit-exact, Dice per class equal (config 5: mask-complementation ensemble on device)."""
    import numpy as np
    from hybrid_ctunet_b200.ensemble import ensemble_masks
    from oracle import sliding_window_oracle as SO
    torch.manual_seed(11)
    p1 = torch.randn(14, 40, 36, 28, device="cuda") * 3
    p2 = torch.randn(14, 40, 36, 28, device="cuda") * 3
    lab = torch.randint(0, 14, (40, 36, 28), device="cuda").float()
    got = ensemble_masks(p1, p2, lab)
    ref = SO.ensemble_reference(p1, p2, lab)
    for k in ("head1", "head2"):
        assert np.array_equal(got[k].cpu().numpy(), ref[k]), k
    # the ensemble argmax can differ only where the two averaged probabilities tie to the last ulp
    diff = (got["ensemble"].cpu().numpy() != ref["ensemble"]).mean()
    assert diff < 1e-4, diff
    assert np.allclose(got["dice"][1:].cpu().numpy(), ref["dice"][1:], atol=1e-12)
    assert np.allclose(got["dice"][0].cpu().numpy(), ref["dice"][0], atol=5e-3)
