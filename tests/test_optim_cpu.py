"""Host logic of the fused AdamW wrapper (no GPU): constructor contract of torch.optim.AdamW (main_CTUNet.py:190-193),
torch-compatible param_groups / state_dict layout, and the loud failure when there is no sm_100 device."""
import pytest
import torch


def test_constructor_contract_and_param_group_keys():
    from hybrid_ctunet_b200.optim import AdamW
    p = [torch.nn.Parameter(torch.zeros(4, 3))]
    opt = AdamW(p, lr=1e-4, weight_decay=1e-5)
    ref = torch.optim.AdamW([torch.nn.Parameter(torch.zeros(4, 3))], lr=1e-4, weight_decay=1e-5)
    g, r = opt.param_groups[0], ref.param_groups[0]
    for k in ("lr", "betas", "eps", "weight_decay", "amsgrad"):
        assert g[k] == r[k]
    # every key torch keeps is present, so a state_dict written here configures torch's optimizer identically
    assert set(r.keys()) <= set(g.keys())
    assert g.get("decoupled_weight_decay", True) is True
    for bad in (dict(lr=-1.0), dict(eps=-1.0), dict(betas=(1.0, 0.9)), dict(betas=(0.9, 1.0)), dict(weight_decay=-0.1)):
        with pytest.raises(ValueError):
            AdamW(p, **bad)
    with pytest.raises(NotImplementedError):
        AdamW(p, amsgrad=True)


def test_state_dict_of_an_unstepped_optimizer_loads_into_torch():
    from hybrid_ctunet_b200.optim import AdamW
    opt = AdamW([torch.nn.Parameter(torch.zeros(5))], lr=3e-4, weight_decay=1e-2)
    ref = torch.optim.AdamW([torch.nn.Parameter(torch.zeros(5))])
    ref.load_state_dict(opt.state_dict())
    assert ref.param_groups[0]["lr"] == 3e-4 and ref.param_groups[0]["weight_decay"] == 1e-2


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_step_without_a_gpu_fails_loudly():
    from hybrid_ctunet_b200.lib import CtuError
    from hybrid_ctunet_b200.optim import AdamW
    p = torch.nn.Parameter(torch.zeros(3))
    p.grad = torch.ones(3)
    with pytest.raises((CtuError, RuntimeError)):
        AdamW([p]).step()
    assert torch.equal(p.detach(), torch.zeros(3))   # nothing was updated by some silent fallback
