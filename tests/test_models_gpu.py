"""End-to-end parity of the drop-in CTUNet / TUNet against the reference (SURVEY 8c protocol, level 3).

Weights: torch.manual_seed(0) default init (identical to the reference's, see tests/test_state_dict.py);
input: torch.manual_seed(1) randn(1,1,96,96,96) — BASELINE.json config 1.

 * ViT-branch heads must meet the north_star bf16 tolerance: rel-L2 <= 1e-2 against the reference's fp32 output
   (golden fixture from the unmodified reference on CPU, and the fp32 oracle run on the same GPU).
 * ResNet-branch heads are numerically chaotic at random init (the reference's own bf16 autocast differs from
   its fp32 output by 0.61 rel-L2), so they are held to the yard-stick: our error against fp32 must
   not exceed 1.25x the error of torch-autocast-bf16 running the same oracle on the same GPU.
The measured numbers are written to gpurun_out/parity_ctunet.json.
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KW = dict(in_channels=1, dim_conv_stem=64, out_channels=14, img_size=(96, 96), frames=96, patch_frame=8)
SUB = (slice(None), slice(None), slice(None, None, 8), slice(None, None, 8), slice(None, None, 8))


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _agree(a, b):
    return (a.argmax(1) == b.argmax(1)).double().mean().item()


def _margin_agree(a, ref, tol=1e-2):
    """argmax agreement on the voxels whose fp32 decision is not inside the tolerance band — top-1/top-2
    margin of the reference logits > 5 x tol x ||logit vector of that voxel||.  Returns (agreement, fraction kept)."""
    top2 = ref.topk(2, dim=1).values
    keep = (top2[:, 0] - top2[:, 1]) > 5.0 * tol * ref.norm(dim=1)
    same = a.argmax(1) == ref.argmax(1)
    return same[keep].double().mean().item(), keep.double().mean().item()


@pytest.fixture(scope="module")
def ctunet_run():
    from hybrid_ctunet_b200.networks.hybrid_CTUNet import CTUNet
    from oracle import ctunet_oracle as O
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    m = CTUNet(model_depth=101, **KW).cuda().eval()
    torch.manual_seed(1)
    x = torch.randn(1, 1, 96, 96, 96).cuda()
    with torch.no_grad():
        ours = m(x)
        sd = {k: v.detach() for k, v in m.state_dict().items()}
        ref = O.ctunet_forward(sd, x)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            amp = O.ctunet_forward(sd, x)
    names = ["res_logits", "res_48", "res_24", "vit_logits", "vit_96"]
    flat = lambda o: [o[0][0], o[0][1], o[0][2], o[1][0], o[1][1]]
    rows = {}
    for n, a, r, p in zip(names, flat(ours), flat(ref), flat(amp)):
        rows[n] = dict(ours_vs_fp32=_rel(a, r), autocast_vs_fp32=_rel(p.float(), r), ours_argmax=_agree(a, r),
                       autocast_argmax=_agree(p.float(), r), shape=list(a.shape), dtype=str(a.dtype))
        rows[n]["ours_argmax_margin"], rows[n]["margin_voxels_kept"] = _margin_agree(a, r)
        rows[n]["autocast_argmax_margin"] = _margin_agree(p.float(), r)[0]
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_ctunet.json", "w") as fh:
        json.dump(rows, fh, indent=1)
    print(json.dumps(rows, indent=1))
    return dict(ours=dict(zip(names, flat(ours))), ref=dict(zip(names, flat(ref))), rows=rows)


def test_ctunet_output_structure(ctunet_run):
    o = ctunet_run["ours"]
    assert o["res_logits"].shape == (1, 14, 96, 96, 96) and o["res_48"].shape == (1, 14, 48, 48, 96)
    assert o["res_24"].shape == (1, 14, 24, 24, 48)
    assert o["vit_logits"].shape == (1, 14, 96, 96, 96) and o["vit_96"].shape == (1, 14, 96, 96, 96)
    assert all(t.dtype == torch.float32 and torch.isfinite(t).all() for t in o.values())


def test_ctunet_vit_heads_meet_bf16_tolerance(ctunet_run):
    rows = ctunet_run["rows"]
    for n in ("vit_logits", "vit_96"):
        assert rows[n]["ours_vs_fp32"] <= 1e-2, (n, rows[n])
        assert rows[n]["ours_argmax"] >= rows[n]["autocast_argmax"] - 1e-3, (n, rows[n])


def test_ctunet_vit_heads_argmax_999_outside_the_tolerance_band(ctunet_run):
    """north_star: argmax masks >= 99.9 % voxel-identical, in the measurable form of SURVEY 8c(3): over the voxels
    whose fp32 top-1/top-2 margin exceeds 5 x (1e-2 x the voxel's logit norm).  (At random init the 14 logits of a voxel
    are nearly tied for a large share of voxels; inside the band any bf16 implementation, torch's included, flips.)"""
    rows = ctunet_run["rows"]
    for n in ("vit_logits", "vit_96"):
        assert rows[n]["margin_voxels_kept"] > 0.05, (n, rows[n])
        assert rows[n]["ours_argmax_margin"] >= 0.999, (n, rows[n])


def test_ctunet_res_heads_within_autocast_yardstick(ctunet_run):
    rows = ctunet_run["rows"]
    for n in ("res_logits", "res_48", "res_24"):
        assert rows[n]["ours_vs_fp32"] <= 1.25 * rows[n]["autocast_vs_fp32"], (n, rows[n])


def test_ctunet_matches_reference_golden(ctunet_run):
    """Same comparison against the UNMODIFIED reference run on CPU (tests/golden/make_golden.py)."""
    z = np.load(os.path.join(GOLD, "ctunet_101_pf8_seed0_x1.npz"))
    o = ctunet_run["ours"]
    for n in ("vit_logits", "vit_96"):
        g = torch.from_numpy(z[n + "_sub"]).cuda()
        assert _rel(o[n][SUB], g) <= 1.2e-2, n
    # the fp32 oracle on this GPU reproduces the reference's CPU output up to fp32 accumulation-order noise
    for n in ("vit_logits", "vit_96"):
        g = torch.from_numpy(z[n + "_sub"]).cuda()
        assert _rel(ctunet_run["ref"][n][SUB], g) <= 1e-4, n


def test_tunet_matches_reference_golden():
    from hybrid_ctunet_b200.networks.hybrid_CTUNet import TUNet
    z = np.load(os.path.join(GOLD, "tunet_pf8_seed0_x1.npz"))
    torch.manual_seed(0)
    m = TUNet(**KW).cuda().eval()
    torch.manual_seed(1)
    x = torch.randn(1, 1, 96, 96, 96).cuda()
    with torch.no_grad():
        v0, v1 = m(x)
    assert _rel(v0[SUB], torch.from_numpy(z["vit_logits_sub"]).cuda()) <= 1.2e-2
    assert _rel(v1[SUB], torch.from_numpy(z["vit_96_sub"]).cuda()) <= 1.2e-2
