"""Whole-network training-step checks on the GPU: TUNet (the well-conditioned ViT branch) gradient parity against
oracle autograd with the bf16-autocast yard-stick, and the CTUNet step (5-head Dice-CE loss, trainer_CTUNet.py:
92-103): every used parameter gets a finite gradient, the seven never-used conv3 weights get None (so AdamW keeps
skipping them, SURVEY 7 hard part 6), gradients near the heads match the oracle, and one AdamW step lowers the loss."""
import pytest
import torch

pytestmark = pytest.mark.gpu
KW = dict(in_channels=1, dim_conv_stem=64, out_channels=14, model_depth=101, img_size=(96, 96), frames=96, patch_frame=8)
TKW = dict(in_channels=1, dim_conv_stem=64, out_channels=14, img_size=(96, 96), frames=96, patch_frame=8)


@pytest.fixture(autouse=True)
def _strict_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def test_tunet_gradients_match_oracle():
    from hybrid_ctunet_b200.losses import DiceCELoss
    from hybrid_ctunet_b200.networks.hybrid_CTUNet import TUNet
    from oracle import ctunet_oracle as O
    torch.manual_seed(0)
    model = TUNet(**TKW).cuda().train()
    loss_func = DiceCELoss(to_onehot_y=True, softmax=True, squared_pred=True, smooth_nr=0.0, smooth_dr=1e-6)
    torch.manual_seed(1)
    x = torch.rand(1, 1, 96, 96, 96, device="cuda")
    y = torch.randint(0, 14, (1, 1, 96, 96, 96), device="cuda").float()
    lg = model(x)
    loss = loss_func(lg[0], y) + loss_func(lg[1], y)
    loss.backward()

    def ref_grads(autocast):
        sd = {k: v.detach().clone().requires_grad_() for k, v in model.state_dict().items()}
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            r = O.tunet_forward(sd, x, 8)
        l = loss_func(r[0].float(), y) + loss_func(r[1].float(), y)
        l.backward()
        return float(l.detach()), {k: v.grad for k, v in sd.items()}

    l_ref, g_ref = ref_grads(False)
    l_yard, g_yard = ref_grads(True)
    assert abs(float(loss) - l_ref) < 1e-2 * abs(l_ref)
    worst = {}
    for name, p in model.named_parameters():
        if g_ref[name] is None:
            assert p.grad is None, name
            continue
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
        e, ey = _rel(p.grad, g_ref[name]), _rel(g_yard[name], g_ref[name])
        if not e < max(3e-2, 1.5 * ey):
            worst[name] = (e, ey)
    assert not worst, worst


def test_ctunet_training_step():
    from hybrid_ctunet_b200.losses import DiceCELoss, ctunet_loss
    from hybrid_ctunet_b200.networks.hybrid_CTUNet import CTUNet
    from oracle import ctunet_oracle as O
    torch.manual_seed(0)
    model = CTUNet(**KW).cuda().train()
    loss_func = DiceCELoss(to_onehot_y=True, softmax=True, squared_pred=True, smooth_nr=0.0, smooth_dr=1e-6)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5)
    torch.manual_seed(1)
    x = torch.rand(1, 1, 96, 96, 96, device="cuda")
    y = torch.randint(0, 14, (1, 1, 96, 96, 96), device="cuda").float()
    logits = model(x)
    loss = ctunet_loss(logits, y, loss_func)
    loss.backward()
    unused = [n for n, p in model.named_parameters() if p.grad is None]
    assert len(unused) == 7 and all(n.endswith("conv3.conv.weight") for n in unused), unused
    assert sum(p.numel() for n, p in model.named_parameters() if p.grad is None) == 692224
    for n, p in model.named_parameters():
        if p.grad is not None:
            assert torch.isfinite(p.grad).all(), n
            assert float(p.grad.abs().max()) > 0, n

    # oracle autograd on the same step: loss value, and gradients of the parameters next to the heads (the deep
    # ResNet branch is chaotic at random init — SURVEY 8c — so far-from-head parameters are judged per block instead)
    sd = {k: v.detach().clone().requires_grad_() for k, v in model.state_dict().items()}
    ref = O.ctunet_forward(sd, x, 101, 8)
    l_ref = ctunet_loss(ref, y, loss_func)
    l_ref.backward()
    assert abs(float(loss) - float(l_ref)) < 2e-2 * abs(float(l_ref)), (float(loss), float(l_ref))
    # (vit.* and vit_encoder.layers.{0,1,2} also receive gradient through the ResNet-branch decoders, so only the
    # parameters fed purely by the ViT heads are compared here; the whole ViT branch is compared in the TUNet test)
    for name in ("vit_out.conv.conv.weight", "vit_out.conv.conv.bias", "decoder_linear_96x96.head.weight",
                 "vit_decoder0.conv_block.conv2.conv.weight", "vit_decoder0.conv_block.conv1.conv.weight",
                 "vit_encoder0.layer.conv2.conv.weight", "vit_encoder.layers.3.0.4.to_out.weight",
                 "vit_encoder.layers.3.0.2.fn.net.1.weight", "vit_encoder.layers.3.0.1.fn.net.0.weight"):
        e = _rel(dict(model.named_parameters())[name].grad, sd[name].grad)
        assert e < 6e-2, (name, e)

    l0 = float(loss)
    for _ in range(3):
        opt.step()
        for p in model.parameters():
            p.grad = None
        loss = ctunet_loss(model(x), y, loss_func)
        loss.backward()
    assert float(loss) < l0, (l0, float(loss))


@pytest.mark.parametrize("B,optimizer", [(1, "torch"), (2, "ours")])
def test_graphed_train_step_trains_with_fused_adamw(B, optimizer):
    """CUDA-graph replay of forward + loss + backward (hybrid_ctunet_b200.training): same loss and gradients as the eager
    step on the same weights, and the packed weights follow a FUSED optimizer (which does not bump Tensor._version) —
    torch's fused AdamW at batch 1, this package's ctu_adamw_step at batch 2 (the bench configuration)."""
    from hybrid_ctunet_b200.losses import DiceCELoss, ctunet_loss
    from hybrid_ctunet_b200.networks.hybrid_CTUNet import CTUNet
    from hybrid_ctunet_b200.training import GraphedTrainStep
    torch.manual_seed(0)
    model = CTUNet(**KW).cuda().train()
    loss_func = DiceCELoss(to_onehot_y=True, softmax=True, squared_pred=True, smooth_nr=0.0, smooth_dr=1e-6)
    torch.manual_seed(1)
    x = torch.rand(B, 1, 96, 96, 96, device="cuda")
    y = torch.randint(0, 14, (B, 1, 96, 96, 96), device="cuda").float()
    l_eager = ctunet_loss(model(x), y, loss_func)
    l_eager.backward()
    eager = float(l_eager.detach())
    g_eager = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    del l_eager  # no autograd graph (with AccumulateGrad nodes on the default stream) may survive into the capture
    for p in model.parameters():
        p.grad = None
    step = GraphedTrainStep(model, lambda lg, t: ctunet_loss(lg, t, loss_func), x, y, warmup=1)
    if optimizer == "ours":
        from hybrid_ctunet_b200.optim import AdamW
        opt = AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5)
    else:
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5, fused=True)
    losses = []
    for it in range(4):
        losses.append(float(step(x, y).detach()))
        if it == 0:
            # the captured step (two lanes + off-path parameter-gradient stream) produces the eager step's gradients
            # (same kernels; only the order of the fp32 reductions differs — that noise is re-rounded to bf16 at every
            # layer of the backward chain and reaches ~3e-3 at the far end, the patch embedding; a race or a lost
            # contribution would be O(1))
            bad = {}
            for n, p in model.named_parameters():
                if n in g_eager:
                    e = _rel(p.grad, g_eager[n])
                    if not e < 1e-2:
                        bad[n] = e
                else:
                    assert p.grad is None, n
            assert not bad, bad
        opt.step()
    assert abs(losses[0] - eager) < 1e-3 * abs(eager), (losses[0], eager)
    assert losses[-1] < losses[0] - 1e-2, losses   # the replayed graph sees the updated weights
    # and so does an eager inference call after the fused updates (train()/eval() drop the packed copies)
    model.eval()
    with torch.no_grad():
        after = float(ctunet_loss(model(x), y, loss_func))
    assert abs(after - losses[-1]) < 0.2 and after < eager


def test_cunet_training_step():
    """CUNet (ResNet encoder + transposed-conv / concat decoder, hybrid_CTUNet.py:859-937): 3 deep-supervision heads,
    gradients finite, loss and head-side gradients match oracle autograd, an optimizer step lowers the loss."""
    from hybrid_ctunet_b200.losses import DiceCELoss, deep_supervision_targets
    from hybrid_ctunet_b200.networks.hybrid_CTUNet import CUNet
    from oracle import ctunet_oracle as O
    torch.manual_seed(0)
    model = CUNet(out_channels=14, model_depth=101).cuda().train()
    lf = DiceCELoss(to_onehot_y=True, softmax=True, squared_pred=True, smooth_nr=0.0, smooth_dr=1e-6)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5, fused=True)
    torch.manual_seed(1)
    x = torch.rand(1, 1, 96, 96, 96, device="cuda")
    y = torch.randint(0, 14, (1, 1, 96, 96, 96), device="cuda").float()
    t1, t2 = deep_supervision_targets(y)

    def loss_of(lg):
        return lf(lg[0], y) + 0.5 * (lf(lg[1], t1) + 0.5 * lf(lg[2], t2))

    loss = loss_of(model(x))
    loss.backward()
    sd = {k: v.detach().clone().requires_grad_() for k, v in model.state_dict().items()}
    ref = loss_of(O.cunet_forward(sd, x, 101))
    ref.backward()
    assert abs(float(loss.detach()) - float(ref.detach())) < 2e-2 * abs(float(ref.detach()))
    for n, p in model.named_parameters():
        if sd[n].grad is None:
            assert p.grad is None, n
        else:
            assert p.grad is not None and torch.isfinite(p.grad).all(), n
    for name in ("res_out.conv.conv.weight", "res_out.conv.conv.bias", "res_out_48x48.conv.conv.weight"):
        assert _rel(dict(model.named_parameters())[name].grad, sd[name].grad) < 6e-2, name
    l0 = float(loss.detach())
    for _ in range(3):
        opt.step()
        for p in model.parameters():
            p.grad = None
        loss = loss_of(model(x))
        loss.backward()
    assert float(loss.detach()) < l0
