"""Pins the sliding-window oracle, the drop-in's host logic and the five-head loss weighting against the REFERENCE'S OWN
functions, executed unmodified in this container (oracle/ref_exec.py: AST-extracted from /root/reference/trainer_*.py,
MONAI 0.7.0 helper names bound to their restatements).  Skipped where /root/reference does not exist (the GPU box);
tests/golden/sliding_window_ref_*.npz carries the same reference outputs there (tests/test_sliding_window_gpu.py).
"""
import os
import types

import numpy as np
import pytest
import torch

from oracle import ref_exec
from oracle import sliding_window_oracle as SO

pytestmark = pytest.mark.skipif(not ref_exec.available(), reason="/root/reference is not present on this box")

GEOMETRIES = ref_exec.SW_CASES
_predictor = ref_exec.sw_case_predictor


@pytest.mark.parametrize("shape,roi,swb,overlap,mode", GEOMETRIES)
def test_oracle_equals_reference_two_heads(shape, roi, swb, overlap, mode):
    """trainer_CTUNet.py:417-557 executed as it is == oracle/sliding_window_oracle.py, bit for bit (loop order,
    14-channel count maps, importance weighting, divide, crop)."""
    ref_fn = ref_exec.sliding_window_two_heads()
    torch.manual_seed(3)
    vol = torch.rand(shape)
    r0, r1 = ref_fn(vol, roi, swb, _predictor(), overlap=overlap, mode=mode)
    o0, o1 = SO.sliding_window_inference(vol, roi, swb, _predictor(), overlap=overlap, mode=mode, two_heads=True)
    assert r0.shape == o0.shape == (shape[0], 5) + tuple(shape[2:])
    assert torch.equal(r0, o0) and torch.equal(r1, o1)


@pytest.mark.parametrize("shape,roi,swb,overlap,mode", GEOMETRIES)
def test_oracle_equals_reference_one_head(shape, roi, swb, overlap, mode):
    """trainer_CUNet.py:268-400 executed as it is == the oracle's one-head mode."""
    ref_fn = ref_exec.sliding_window_one_head()
    torch.manual_seed(4)
    vol = torch.rand(shape)
    pred = lambda w: _predictor()(w)[0]   # CUNet returns a 3-tuple; the function blends element [0]
    r = ref_fn(vol, roi, swb, pred, overlap=overlap, mode=mode)
    o = SO.sliding_window_inference(vol, roi, swb, pred, overlap=overlap, mode=mode, two_heads=False)
    assert torch.equal(r, o)


def test_reference_error_behaviour_matches_dropin_host_logic():
    from hybrid_ctunet_b200 import sliding_window as S
    ref_fn = ref_exec.sliding_window_two_heads()
    vol = torch.rand(1, 1, 20, 20, 20)
    for bad in (-0.1, 1.0):
        with pytest.raises(AssertionError):
            ref_fn(vol, (16, 16, 16), 1, _predictor(), overlap=bad)
    fns = ref_exec.extract("trainer_CTUNet.py", ["_get_scan_interval"])
    for img, roi, ov in [((512, 512, 256), (96, 96, 96), 0.5), ((512, 512, 256), (96, 96, 96), 0.7),
                         ((96, 96, 96), (96, 96, 96), 0.5), ((100, 100, 100), (2, 2, 2), 0.9)]:
        assert fns["_get_scan_interval"](img, roi, 3, ov) == S.get_scan_interval(img, roi, 3, ov)
    with pytest.raises(ValueError):
        fns["_get_scan_interval"]((10, 10), (4, 4, 4), 3, 0.5)
    with pytest.raises(ValueError):
        S.get_scan_interval((10, 10), (4, 4, 4), 3, 0.5)


def test_committed_sliding_window_golden_is_the_reference_output():
    """The fixtures the GPU box checks the CUDA blend against are what the reference function returns."""
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sliding_window_ref.npz")
    z = np.load(gold)
    ref_fn = ref_exec.sliding_window_two_heads()
    for i, (shape, roi, swb, overlap, mode) in enumerate(GEOMETRIES):
        vol = torch.from_numpy(z[f"vol{i}"])
        assert tuple(vol.shape) == shape
        r0, r1 = ref_fn(vol, roi, swb, _predictor(), overlap=overlap, mode=mode)
        assert torch.equal(r0, torch.from_numpy(z[f"head0_{i}"])) and torch.equal(r1, torch.from_numpy(z[f"head1_{i}"]))


class _TinyNet(torch.nn.Module):
    """Stand-in with CTUNet's return structure (hybrid_CTUNet.py:857) for running the reference's train_epoch on CPU."""

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.linspace(-1.0, 1.0, 5 * 4).reshape(5, 4))

    def forward(self, x):
        full = torch.einsum("bcxyz,kc->bkxyz", torch.cat([x, x * x, torch.sin(x), torch.cos(x)], 1), self.w)
        return ((full, full[:, :, ::2, ::2, :] * 1.5, full[:, :, ::4, ::4, ::2] - 0.5), (full * 0.7, full + 0.3))


def test_reference_train_epoch_step_equals_dropin_loss(monkeypatch):
    """trainer_CTUNet.py:76-109 executed as it is (amp off, one batch): the parameter after its optimizer.step() equals
    the one obtained from hybrid_ctunet_b200.losses.ctunet_loss (device-side label gather instead of the two
    scipy.ndimage.zoom round trips) — pins the five-head weighting and the label down-sampling."""
    from torch.amp import GradScaler, autocast
    from hybrid_ctunet_b200.losses import DiceCELoss, ctunet_loss
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)   # the function moves tensors with .cuda()
    fns = ref_exec.extract("trainer_CTUNet.py", ["train_epoch", "AverageMeter"],
                           dict(time=__import__("time"), autocast=lambda enabled=True: autocast("cpu", enabled=enabled),
                                GradScaler=GradScaler, distributed_all_gather=None))
    loss_func = DiceCELoss(to_onehot_y=True, softmax=True, squared_pred=True, smooth_nr=0.0, smooth_dr=1e-6)
    torch.manual_seed(5)
    x = torch.rand(2, 1, 16, 16, 8)
    y = torch.randint(0, 5, (2, 1, 16, 16, 8)).float()
    args = types.SimpleNamespace(rank=1, amp=False, distributed=False, batch_size=2, max_epochs=1, world_size=1)

    m_ref = _TinyNet()
    opt = torch.optim.SGD(m_ref.parameters(), lr=0.5)
    fns["train_epoch"](m_ref, [{"image": x, "label": y}], opt, None, 0, loss_func, args)

    m_our = _TinyNet()
    loss = ctunet_loss(m_our(x), y, loss_func)
    loss.backward()
    with torch.no_grad():
        m_our.w -= 0.5 * m_our.w.grad
    assert torch.allclose(m_ref.w, m_our.w, rtol=0, atol=1e-7), (m_ref.w - m_our.w).abs().max()
