"""Teacher-forced per-block parity (SURVEY 8c protocol, level 2): every drop-in block is fed the reference's
fp32 input for that block and must reproduce the reference module's fp32 output — generated from the
UNMODIFIED reference by tests/golden/make_golden.py — within rel-L2 <= 1e-2 (bf16 operands, fp32 accumulate;
the north_star tolerance for bf16)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-2


def _load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    sd = {k[2:]: torch.from_numpy(z[k].astype(np.float32)) for k in z.files if k.startswith("w:")}
    io = {k: torch.from_numpy(z[k]) for k in z.files if not k.startswith("w:")}
    return sd, io


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _seeded(seed, ctor, io):
    """Blocks whose weights are too large to ship: re-draw the reference's default init from its seed (the drop-in
    modules create the same torch layers in the same order) and check the stored probe of the last parameter."""
    torch.manual_seed(seed)
    mod = ctor()
    last = list(mod.state_dict().values())[-1].flatten()[:16]
    assert torch.equal(last, io["probe"]), "seeded init no longer reproduces the reference's weights"
    return mod, mod.state_dict()


def _run(mod, sd, *inputs):
    mod.load_state_dict(sd, strict=True)
    mod = mod.cuda().eval()
    with torch.no_grad():
        return mod(*[t.cuda() for t in inputs])


def test_bottleneck_identity():
    from hybrid_ctunet_b200.networks import resnet
    sd, io = _load("bottleneck_128_32")
    y = _run(resnet.Bottleneck(128, 32), sd, io["x"])
    assert y.shape == io["y"].shape and _rel(y, io["y"]) < TOL


def test_bottleneck_stride2_downsample():
    from hybrid_ctunet_b200.networks import resnet
    import torch.nn as nn
    sd, io = _load("bottleneck_down_128_64")
    ds = nn.Sequential(resnet.get_conv_layer(3, 128, 256, kernel_size=1, stride=(2, 2, 2)), nn.Identity())
    y = _run(resnet.Bottleneck(128, 64, stride=(2, 2, 2), downsample=ds), sd, io["x"])
    assert y.shape == io["y"].shape and _rel(y, io["y"]) < TOL


@pytest.mark.parametrize("name,cin,cout", [("resblock_64_64", 64, 64), ("resblock_128_64", 128, 64), ("resblock_1_64", 1, 64)])
def test_resblock(name, cin, cout):
    from hybrid_ctunet_b200.networks import hybrid_CTUNet as H
    sd, io = _load(name)
    y = _run(H.ResBlock(3, cin, cout, 3, 1, "instance"), sd, io["x"])
    assert y.shape == io["y"].shape and _rel(y, io["y"]) < TOL


def test_pixelweight_attention():
    from hybrid_ctunet_b200.networks import hybrid_CTUNet as H
    sd, io = _load("pwa_128")
    y = _run(H.pixelweight_attention(128), sd, io["x1"], io["x2"])
    assert _rel(y, io["y"]) < TOL


def test_up_2fusion_block():
    from hybrid_ctunet_b200.networks import hybrid_CTUNet as H
    _, io = _load("up2fusion_256_128")
    mod, sd = _seeded(16, lambda: H.Up_2Fusion_Block(3, 256, 128, 3, (2, 2, 2), "instance"), io)
    y = _run(mod, sd, io["inp"], io["skip_conv"], io["skip_vit"])
    assert y.shape == io["y"].shape and _rel(y, io["y"]) < 2 * TOL  # two ResBlocks + two fusions deep


def test_upconv_block_221():
    from hybrid_ctunet_b200.networks import hybrid_CTUNet as H
    sd, io = _load("upconv_128_64")
    y = _run(H.UpConvBlock(3, 128, 64, 3, (2, 2, 1), "instance"), sd, io["x"])
    assert y.shape == io["y"].shape and _rel(y, io["y"]) < TOL


@pytest.mark.parametrize("name,f,cin,cout", [("pixelshuffle_512_256", (2, 2, 2), 512, 256),
                                             ("pixelshuffle_221_128_64", (2, 2, 1), 128, 64)])
def test_pixel_shuffle(name, f, cin, cout):
    from hybrid_ctunet_b200.networks import hybrid_CTUNet as H
    sd, io = _load(name)
    y = _run(H.PixelShuffle(3, f, cin, cout), sd, io["x"])
    assert y.shape == io["y"].shape and _rel(y, io["y"]) < TOL


def test_vit_transformer_block():
    from hybrid_ctunet_b200.networks import vit
    _, io = _load("vit_block")
    mod, sd = _seeded(20, lambda: vit.TransformerBlock(768, 12, 64, 3072), io)
    y = _run(mod, sd, io["x"])
    assert _rel(y, io["y"]) < TOL


def test_vit_small():
    from hybrid_ctunet_b200.networks import vit
    sd, io = _load("vit_small")
    m = vit.ViT(image_size=(32, 32), image_patch_size=16, frames=48, frame_patch_size=8, dim=128, depth=2, heads=2,
                mlp_dim=256)
    y = _run(m, sd, io["x"])
    assert y.shape == io["y"].shape and _rel(y, io["y"]) < TOL


def test_out_block():
    from hybrid_ctunet_b200.networks import hybrid_CTUNet as H
    sd, io = _load("outblock_64_14")
    y = _run(H.UnetOutBlock(3, 64, 14), sd, io["x"])
    assert y.shape == io["y"].shape and _rel(y, io["y"]) < TOL


def test_up_attention_block():
    from hybrid_ctunet_b200.networks import hybrid_CTUNet as H
    _, io = _load("up_attention_666")
    mod, sd = _seeded(31, lambda: H.UpAttentionBlock(3, 768, dims=[128, 256, 512, 1024]), io)
    ys = _run(mod, sd, io["x"])
    assert len(ys) == 5
    assert _rel(ys[1], io["y1"]) < TOL
    assert _rel(ys[2][:, :, ::2, ::2, ::2], io["y2"]) < TOL
    assert _rel(ys[3][:, :, ::4, ::4, ::4], io["y3"]) < 1.5 * TOL
    assert _rel(ys[4][:, :, ::8, ::8, ::4], io["y4"]) < 1.5 * TOL
