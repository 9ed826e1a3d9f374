"""`Invertd` on the device (ctu_invert_resample / ctu_invert_ensemble_argmax through hybrid_ctunet_b200.invert) against the
stepwise MONAI restatement run on the host (oracle/invert_oracle.py: pad -> torch grid_sample in float64 -> flips).
Floating point, tolerance 2e-6 of the value range for the trilinear mode (both sides interpolate in float64 and round to
float32; the source coordinates come along different routes); nearest mode is exact away from half-way samples; the fused
inverse + ensemble equals the two-step composition bit for bit."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(kw, seed=0, channels=3):
    from test_invert_cpu import _case, _geom
    img, aff, pixdim, trace, pred = _case(kw, seed=seed, channels=channels)
    return img, trace, pred, _geom(trace)


def _cases():
    from test_invert_cpu import CASES
    return CASES


@pytest.mark.parametrize("i", range(6))
@pytest.mark.parametrize("nearest", [False, True], ids=["trilinear", "nearest"])
def test_invert_pred_matches_the_oracle(i, nearest):
    from hybrid_ctunet_b200.invert import invert_pred
    from oracle import invert_oracle as IO
    from test_invert_cpu import apply_geometry, ties
    img, trace, pred, g = _setup(_cases()[i], seed=i)
    ref, _ = IO.invertd(pred, trace, nearest_interp=nearest)
    got = invert_pred(torch.from_numpy(pred).cuda(), g, nearest_interp=nearest).cpu().numpy()
    assert got.shape == ref.shape and got.dtype == np.float32
    host = apply_geometry(pred, g, 0 if nearest else 1)
    if nearest:
        clear = ~ties(g)
        assert np.array_equal(got[:, clear], ref[:, clear])
        assert np.array_equal(got[:, clear], host[:, clear])
    else:
        assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()
        assert np.abs(got - host).max() <= 2e-6 * np.abs(ref).max()


def test_margin_rim_is_trimmed_like_cropforeground_inverse():
    from hybrid_ctunet_b200.invert import invert_pred
    from oracle import invert_oracle as IO
    from test_invert_cpu import _geom
    img, trace, _, _ = _setup(dict(axcodes="LAS"), seed=5)
    c = trace["crop"]
    t2 = {k: (dict(v) if isinstance(v, dict) else v) for k, v in trace.items()}
    t2["crop"]["box_start"], t2["crop"]["box_end"] = np.asarray(c["box_start"]) - 6, np.asarray(c["box_end"]) + 6
    size = tuple(int(e - b) for b, e in zip(t2["crop"]["box_start"], t2["crop"]["box_end"]))
    big = np.random.default_rng(9).standard_normal((2,) + size).astype(np.float32)
    ref, _ = IO.invertd(big, t2)
    got = invert_pred(torch.from_numpy(big).cuda(), _geom(t2)).cpu().numpy()
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()


@pytest.mark.parametrize("i", [0, 2, 4])
def test_fused_inverse_ensemble_equals_the_two_step_composition(i):
    from hybrid_ctunet_b200.ensemble import ensemble_masks
    from hybrid_ctunet_b200.invert import invert_ensemble_masks, invert_pred
    from oracle import invert_oracle as IO
    img, trace, p1, g = _setup(_cases()[i], seed=10 + i, channels=14)
    p2 = np.random.default_rng(77 + i).standard_normal(p1.shape).astype(np.float32) * 2.0
    labels = torch.from_numpy(np.random.default_rng(5).integers(0, 14, img.shape[1:]).astype(np.float32)).cuda()
    a, b = torch.from_numpy(p1).cuda(), torch.from_numpy(p2).cuda()
    fused = invert_ensemble_masks(a, b, g, labels)
    two = ensemble_masks(invert_pred(a, g), invert_pred(b, g), labels)
    for k in ("ensemble", "head1", "head2", "counts", "dice"):
        assert torch.equal(fused[k], two[k]), k
    # and the reference's own sequence on the host: Invertd of each, softmax, mean, argmax (test_CTUNet.py:222-233)
    r1, r2 = (torch.from_numpy(IO.invertd(p, trace)[0]) for p in (p1, p2))
    s1, s2 = torch.softmax(r1, 0), torch.softmax(r2, 0)
    ref = torch.argmax((s1 + s2) / 2.0, dim=0)
    assert (fused["ensemble"].cpu().long() != ref).float().mean().item() < 1e-4   # ties between classes at float32 precision
    assert (fused["head1"].cpu().long() != torch.argmax(s1, dim=0)).float().mean().item() < 1e-4


def test_full_size_properties():
    """A 512 x 512 x 147 scan at 0.76 x 0.76 x 3.0 mm, LAS, resampled to 1.5 x 1.5 x 2.0 mm (the loader's settings) — too big
    for the host oracle.  A prediction that holds its own padded-grid coordinates must come back as the composite map
    itself wherever all eight corners lie inside the crop (trilinear interpolation of a linear function is exact);
    linearity; the fused ensemble equals the two-step one."""
    from hybrid_ctunet_b200.ensemble import ensemble_masks
    from hybrid_ctunet_b200.invert import InvertGeometry, invert_ensemble_masks, invert_pred
    shape = (512, 512, 147)
    aff = np.diag([-0.76, 0.76, 3.0, 1.0])
    aff[:3, 3] = (190.0, -170.0, -300.0)
    probe = InvertGeometry.from_file(aff, shape, (1.5, 1.5, 2.0), (0, 0, 0), (1, 1, 1))
    ps = probe.pad_size
    box_start, box_end = (11, 17, 6), (ps[0] - 9, ps[1] - 20, ps[2] - 4)
    g = InvertGeometry.from_file(aff, shape, (1.5, 1.5, 2.0), box_start, box_end)
    assert g.out_size == shape and g.pad_size == ps and abs(ps[0] - 260) <= 1 and abs(ps[2] - 220) <= 1
    n = g.pred_size
    dev = torch.device("cuda")
    grids = torch.meshgrid(*[torch.arange(n[a], dtype=torch.float32, device=dev) + g.crop_start[a] for a in range(3)], indexing="ij")
    ramp = torch.stack(grids, 0)
    out = invert_pred(ramp, g)
    idx = torch.stack(torch.meshgrid(*[torch.arange(s, dtype=torch.float64, device=dev) for s in shape], indexing="ij"), 0)
    m = torch.from_numpy(g.m).to(dev)
    c = torch.einsum("ab,bxyz->axyz", m[:, :3], idx) + m[:, 3].view(3, 1, 1, 1)
    inside = torch.ones(shape, dtype=torch.bool, device=dev)
    for a in range(3):
        c[a].clamp_(0.0, ps[a] - 1.0)
        lo = torch.floor(c[a])
        inside &= (lo >= g.crop_start[a]) & (lo + 1 <= g.crop_start[a] + n[a] - 1)
    assert inside.float().mean().item() > 0.5
    err = (out.double() - c).abs()[:, inside].max().item()
    assert err <= 1e-4, err          # float32 rounding of coordinates up to 260
    # outside the padded crop everything is the zero pad
    far = torch.ones(shape, dtype=torch.bool, device=dev)
    for a in range(3):
        far &= (c[a] < g.crop_start[a] - 1) | (c[a] > g.crop_start[a] + n[a])
    assert out[:, far].abs().max().item() == 0.0
    del idx, c
    gen = torch.Generator(device=dev).manual_seed(1)
    p1 = torch.randn((14,) + n, generator=gen, device=dev)
    p2 = torch.randn((14,) + n, generator=gen, device=dev)
    i1, i2 = invert_pred(p1, g), invert_pred(p2, g)
    lin = invert_pred(0.5 * p1 - 2.0 * p2, g)
    assert (lin - (0.5 * i1 - 2.0 * i2)).abs().max().item() <= 2e-5
    labels = torch.randint(0, 14, shape, generator=gen, device=dev).float()
    fused = invert_ensemble_masks(p1, p2, g, labels)
    two = ensemble_masks(i1, i2, labels)
    for k in ("ensemble", "head1", "head2", "counts"):
        assert torch.equal(fused[k], two[k]), k


def test_hybrid_pipeline_with_invert_step():
    """hybrid_ctunet_inference(..., invert=geometry) = sliding windows of both models -> Invertd -> ensemble
    (test_CTUNet_final.py:539-552), here with stand-in predictors so that only the plumbing is under test."""
    from hybrid_ctunet_b200.ensemble import ensemble_masks, hybrid_ctunet_inference
    from hybrid_ctunet_b200.invert import invert_pred
    from hybrid_ctunet_b200.sliding_window import sliding_window_inference_one_head
    from test_invert_cpu import _geom
    from oracle import invert_oracle as IO
    img, aff = IO.make_case(shape=(70, 64, 30), spacing_mm=(0.8, 0.8, 3.0), axcodes="LAS", seed=4)
    trace = IO.forward_trace(img, aff, (1.5, 1.5, 2.0))
    g = _geom(trace)
    vol = torch.from_numpy(trace["image"]).cuda()[None]                        # [1, 1, x, y, z] as the loader yields it
    scale = torch.linspace(0.5, 2.0, 14, device="cuda").view(1, 14, 1, 1, 1)
    ctunet = lambda w: ((torch.sin(7.0 * w * scale),), (torch.zeros_like(w).repeat(1, 14, 1, 1, 1),))
    tunet = lambda w: (torch.cos(5.0 * w * scale),)
    roi = (32, 32, 32)
    lab = torch.randint(0, 14, img.shape[1:], device="cuda").float()
    got = hybrid_ctunet_inference(vol, ctunet, tunet, roi_size=roi, labels=lab, invert=g)
    p1 = sliding_window_inference_one_head(vol, roi, 4, lambda w: (ctunet(w)[0][0],), overlap=0.5, mode="gaussian")[0]
    p2 = sliding_window_inference_one_head(vol, roi, 4, tunet, overlap=0.7, mode="gaussian")[0]
    ref = ensemble_masks(invert_pred(p1, g), invert_pred(p2, g), lab)
    assert got["ensemble"].shape == img.shape[1:]
    for k in ("ensemble", "head1", "head2", "counts"):
        assert torch.equal(got[k], ref[k]), k


def test_evaluate_cases_runs_the_scripts_closing_loop():
    """evaluate_cases = per case: sliding windows -> Invertd + ensemble + Dice; over the cases: determine_postprocessing
    (test_CTUNet_final.py:527-655).  Stand-in predictors; checked against the pieces called by hand and against the
    reference's Dice definition evaluated on the host."""
    from hybrid_ctunet_b200.ensemble import evaluate_cases, hybrid_ctunet_inference
    from hybrid_ctunet_b200.postprocess import com_dice, determine_postprocessing
    from test_invert_cpu import _geom
    from oracle import invert_oracle as IO
    scale = torch.linspace(0.5, 2.0, 14, device="cuda").view(1, 14, 1, 1, 1)
    ctunet = lambda w: ((torch.sin(9.0 * w * scale),), (torch.zeros_like(w).repeat(1, 14, 1, 1, 1),))
    tunet = lambda w: (torch.cos(4.0 * w * scale),)
    cases = []
    for seed, ax in ((1, "LAS"), (2, "RAS")):
        img, aff = IO.make_case(shape=(60, 56, 28), spacing_mm=(0.8, 0.8, 3.0), axcodes=ax, seed=seed)
        trace = IO.forward_trace(img, aff, (1.5, 1.5, 2.0))
        lab = torch.from_numpy(np.random.default_rng(seed).integers(0, 14, img.shape[1:])).cuda().float()
        cases.append({"image": torch.from_numpy(trace["image"]).cuda()[None], "label": lab[None, None], "geometry": _geom(trace),
                      "volume_per_voxel": 0.8 * 0.8 * 3.0})
    res = evaluate_cases(cases, ctunet, tunet, roi_size=(32, 32, 32))
    assert res["dice"]["ensemble"].shape == (2, 13) and len(res["masks"]) == 2
    for i, case in enumerate(cases):
        one = hybrid_ctunet_inference(case["image"], ctunet, tunet, (32, 32, 32), labels=case["label"][0, 0], invert=case["geometry"])
        assert torch.equal(one["ensemble"], res["masks"][i])
        m, l = one["ensemble"].cpu().numpy(), case["label"][0, 0].cpu().numpy()
        host = [2 * np.sum((m == j) * (l == j)) / (np.sum(m == j) + np.sum(l == j)) if np.sum(l == j) else 0.0 for j in range(1, 14)]
        assert np.allclose(res["dice"]["ensemble"][i], host, atol=1e-12)
    labels = [c["label"][0, 0] for c in cases]
    post = determine_postprocessing(res["masks"], labels, [c["volume_per_voxel"] for c in cases], advanced_postprocessing=True)
    for a, b in zip(post, res["masks_postprocessed"]):
        assert torch.equal(torch.as_tensor(a), torch.as_tensor(b))
    assert np.allclose(res["mean_organ_dice_postprocessed"], com_dice(post, labels))
