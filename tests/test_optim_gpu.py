"""Fused AdamW (ctu_adamw_step, hybrid_ctunet_b200.optim.AdamW) against torch.optim.AdamW — the optimizer
main_CTUNet.py:190-193 builds — on the same parameters and gradients: same update within fp32 rounding, same skipping of
grad-less parameters, interchangeable state_dict."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _params(seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    shapes = [(768, 3072), (64, 64, 3, 3, 3), (14,), (1, 432, 768), (1331, 8), (7,), (1000003,)]
    return [torch.nn.Parameter(torch.randn(s, device="cuda", generator=g)) for s in shapes]


@pytest.mark.parametrize("wd,lr", [(1e-5, 1e-4), (1e-2, 3e-3)])
def test_adamw_matches_torch(wd, lr):
    from hybrid_ctunet_b200.optim import AdamW
    ours, ref = _params(1), _params(1)
    skip = 2                                         # a parameter that never gets a gradient (the unused conv3 weights)
    o1 = AdamW(ours, lr=lr, weight_decay=wd)
    o2 = torch.optim.AdamW(ref, lr=lr, weight_decay=wd)
    # gradients as slices of ONE flat buffer at odd offsets (only 4-byte aligned), like the engine hands them out
    total = sum(p.numel() for p in ours) + 3 * len(ours)
    gen = torch.Generator(device="cuda").manual_seed(5)
    for it in range(4):
        flat = torch.randn(total, device="cuda", generator=gen)
        off = 1
        for i, (a, b) in enumerate(zip(ours, ref)):
            if i == skip:
                continue
            g = flat[off:off + a.numel()].view(a.shape)
            off += a.numel() + 3
            a.grad, b.grad = g, g.clone()
        if it == 2:  # an LR scheduler changes the rate between steps
            for grp in o1.param_groups + o2.param_groups:
                grp["lr"] = lr * 0.5
        o1.step()
        o2.step()
        for i, (a, b) in enumerate(zip(ours, ref)):
            if i == skip:
                assert torch.equal(a, b) and len(o1.state[a]) == 0
                continue
            assert (a - b).abs().max().item() <= 1e-6 * max(1.0, b.abs().max().item()), (it, i)
            assert torch.allclose(o1.state[a]["exp_avg"], o2.state[b]["exp_avg"], rtol=1e-6, atol=1e-9)
            assert torch.allclose(o1.state[a]["exp_avg_sq"], o2.state[b]["exp_avg_sq"], rtol=1e-6, atol=1e-12)
            assert float(o1.state[a]["step"]) == float(o2.state[b]["step"]) == it + 1


def test_adamw_state_dict_round_trip_with_torch():
    from hybrid_ctunet_b200.optim import AdamW
    ours, ref = _params(3), _params(3)
    o1 = AdamW(ours, lr=1e-3, weight_decay=1e-2)
    for p in ours:
        p.grad = torch.randn_like(p)
    o1.step()
    o2 = torch.optim.AdamW(ref, lr=1e-3, weight_decay=1e-2)
    # (deepcopy = what torch.save / torch.load does: load_state_dict itself keeps references to the given tensors)
    o2.load_state_dict(copy.deepcopy(o1.state_dict()))   # our checkpoint loads into torch's optimizer ...
    for a, b in zip(ours, ref):
        b.data.copy_(a.data)
        g = torch.randn_like(a)
        a.grad, b.grad = g, g.clone()
    o1.step()
    o2.step()
    for a, b in zip(ours, ref):
        assert (a - b).abs().max().item() <= 1e-6 * max(1.0, b.abs().max().item())
    o3 = AdamW(_params(3), lr=1e-3, weight_decay=1e-2)
    o3.load_state_dict(copy.deepcopy(o2.state_dict()))   # ... and torch's into ours
    assert float(o3.state[o3.param_groups[0]["params"][0]]["step"]) == 2.0


def test_adamw_rejects_what_it_does_not_implement():
    from hybrid_ctunet_b200.optim import AdamW
    with pytest.raises(NotImplementedError):
        AdamW(_params(0), amsgrad=True)
    with pytest.raises(ValueError):
        AdamW(_params(0), lr=-1.0)
    p = torch.nn.Parameter(torch.randn(8))
    p.grad = torch.randn(8)
    with pytest.raises(RuntimeError):
        AdamW([p]).step()
