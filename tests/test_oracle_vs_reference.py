"""Pins oracle/ctunet_oracle.py against the UNMODIFIED reference modules (build container only: imports
/root/reference through oracle/monai_stub).  Same torch ops in the same order => bit-identical on CPU."""
import pytest
import torch

from oracle import ctunet_oracle as O
from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present (GPU box)")


@pytest.fixture(scope="module")
def ref():
    return ref_import.load()


def _sd(m):
    return {k: v.detach() for k, v in m.state_dict().items()}


def test_bottleneck_and_resblock(ref):
    resnet, _, hyb = ref
    torch.manual_seed(3)
    m = resnet.Bottleneck(128, 32).eval()
    x = torch.randn(1, 128, 6, 6, 8)
    with torch.no_grad():
        sd = {"blk." + k: v for k, v in _sd(m).items()}
        assert torch.equal(m(x), O.bottleneck(sd, "blk", x, 1, False))
    torch.manual_seed(7)
    m = hyb.ResBlock(3, 128, 64, 3, 1, "instance").eval()
    x = torch.randn(1, 128, 6, 6, 8)
    with torch.no_grad():
        sd = {"blk." + k: v for k, v in _sd(m).items()}
        assert torch.equal(m(x), O.res_block(sd, "blk", x, 128, 64))


def test_small_vit(ref):
    _, vit, _ = ref
    torch.manual_seed(4)
    m = vit.ViT(image_size=(32, 32), image_patch_size=16, frames=48, frame_patch_size=8, dim=128, depth=2, heads=2,
                mlp_dim=256).eval()
    x = torch.randn(2, 1, 32, 32, 48)
    with torch.no_grad():
        a, b = m(x), O.vit_forward(_sd(m), "", x, 8, depth=2, heads=2)
    assert torch.allclose(a, b, rtol=0, atol=1e-6)


def test_up_attention_and_fusion(ref):
    _, _, hyb = ref
    torch.manual_seed(5)
    m = hyb.UpAttentionBlock(3, 768, dims=[128, 256, 512, 1024]).eval()
    x = torch.randn(1, 768, 6, 6, 6)
    with torch.no_grad():
        a, b = m(x), O.up_attention_block(_sd(m), "", x)
    for u, v in zip(a, b):
        assert torch.allclose(u, v, rtol=0, atol=1e-5)
    torch.manual_seed(6)
    m = hyb.Up_2Fusion_Block(3, 256, 128, 3, (2, 2, 2), "instance").eval()
    inp, sc, sv = torch.randn(1, 256, 3, 3, 6), torch.randn(1, 128, 6, 6, 12), torch.randn(1, 128, 6, 6, 12)
    with torch.no_grad():
        sd = {"blk." + k: v for k, v in _sd(m).items()}
        assert torch.allclose(m(inp, sc, sv), O.up_2fusion_block(sd, "blk", inp, sc, sv, 128, (2, 2, 2)), rtol=0, atol=1e-5)


def test_full_ctunet_forward_bitwise(ref):
    """BASELINE config 1: CTUNet(101, pf 8) on one 96^3 patch, fp32 CPU."""
    _, _, hyb = ref
    torch.manual_seed(0)
    m = hyb.CTUNet(in_channels=1, dim_conv_stem=64, out_channels=14, model_depth=101, img_size=(96, 96), frames=96,
                   patch_frame=8).eval()
    torch.manual_seed(1)
    x = torch.randn(1, 1, 96, 96, 96)
    with torch.no_grad():
        a = m(x)
        b = O.ctunet_forward(_sd(m), x, 101, 8)
    for u, v in zip(a[0] + a[1], b[0] + b[1]):
        assert u.shape == v.shape
        assert torch.allclose(u, v, rtol=0, atol=1e-4), (u - v).abs().max()


def test_gradients_of_reference_modules_equal_oracle_autograd(ref):
    """The training-step oracle is torch autograd over the restated forward: pin its gradients against the UNMODIFIED
    reference modules' own backward (train mode, dropout 0) on two blocks and a small training loss."""
    resnet, _, hyb = ref
    from oracle import train_oracle as T
    torch.manual_seed(8)
    m = hyb.Up_2Fusion_Block(3, 256, 128, 3, (2, 2, 2), "instance").train()
    inp, sc, sv = torch.randn(1, 256, 3, 3, 6), torch.randn(1, 128, 6, 6, 12), torch.randn(1, 128, 6, 6, 12)
    tgt = torch.randint(0, 14, (1, 1, 6, 6, 12)).float()
    head = torch.randn(14, 128, 1, 1, 1)
    loss = T.dice_ce_loss(torch.nn.functional.conv3d(m(inp, sc, sv), head), tgt)
    loss.backward()
    sd = {"blk." + k: v.detach().clone().requires_grad_() for k, v in m.state_dict().items()}
    loss_o = T.dice_ce_loss(torch.nn.functional.conv3d(O.up_2fusion_block(sd, "blk", inp, sc, sv, 128, (2, 2, 2)), head), tgt)
    loss_o.backward()
    assert abs(float(loss) - float(loss_o)) < 1e-6
    for k, p in m.named_parameters():
        g = sd["blk." + k].grad
        if p.grad is None:
            assert g is None, k
        else:
            assert torch.allclose(p.grad, g, rtol=1e-4, atol=1e-7), k

    torch.manual_seed(9)
    m = resnet.Bottleneck(128, 32).train()
    x = torch.randn(1, 128, 6, 6, 8)
    m(x).square().mean().backward()
    sd = {"blk." + k: v.detach().clone().requires_grad_() for k, v in m.state_dict().items()}
    O.bottleneck(sd, "blk", x, 1, False).square().mean().backward()
    for k, p in m.named_parameters():
        assert torch.allclose(p.grad, sd["blk." + k].grad, rtol=1e-4, atol=1e-8), k
