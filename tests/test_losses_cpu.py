"""CPU checks of the loss restatement: the device-side label gather selects exactly the voxels
scipy.ndimage.zoom(order=0, prefilter=False) selects (trainer_CTUNet.py:93-94; SURVEY 8c known answers), and
DiceCELoss reproduces the MONAI 0.7 formula on a hand-computed case."""
import numpy as np
import torch
from scipy import ndimage

from hybrid_ctunet_b200.losses import DiceCELoss, ctunet_loss, deep_supervision_targets, zoom_indices


def test_zoom_indices_known_answers():
    i48 = zoom_indices(96, 48)
    assert i48[:3] == (0, 2, 4) and i48[23] == 46 and i48[24] == 49 and i48[-1] == 95
    i24 = zoom_indices(96, 24)
    assert i24[:6] == (0, 4, 8, 12, 17, 21) and i24[-5:] == (78, 83, 87, 91, 95)


def test_deep_supervision_targets_equal_scipy_zoom():
    rng = np.random.default_rng(0)
    t = rng.integers(0, 14, size=(2, 1, 96, 96, 96)).astype(np.float32)
    t1, t2 = deep_supervision_targets(torch.from_numpy(t))
    r1 = ndimage.zoom(t, (1, 1, 0.5, 0.5, 1), order=0, prefilter=False)
    r2 = ndimage.zoom(t, (1, 1, 0.25, 0.25, 0.5), order=0, prefilter=False)
    assert np.array_equal(t1.numpy(), r1) and np.array_equal(t2.numpy(), r2)


def test_zoom_nearest_equals_scipy_on_ragged_sizes_including_out_of_volume_samples():
    """scipy places out_idx * ((n-1)/(m-1)) a rounding error above n-1 for some sizes (32 -> 16, 48 -> 24) and then writes
    cval = 0 there; the restatement reproduces it (the reference's 96 -> 48 / 96 -> 24 volumes have no such sample)."""
    from hybrid_ctunet_b200.losses import zoom_nearest
    rng = np.random.default_rng(1)
    assert zoom_indices(32, 16)[-1] == -1 and min(zoom_indices(96, 48)) == 0 and min(zoom_indices(96, 24)) == 0
    for shape in ((32, 32, 16), (48, 20, 24), (64, 36, 8), (10, 12, 14)):
        t = rng.integers(1, 14, size=(2, 1) + shape).astype(np.float32)
        for zoom in ((1, 1, 0.5, 0.5, 1), (1, 1, 0.25, 0.25, 0.5)):
            ours = zoom_nearest(torch.from_numpy(t), zoom).numpy()
            assert np.array_equal(ours, ndimage.zoom(t, zoom, order=0, prefilter=False)), (shape, zoom)


def test_dice_ce_formula():
    torch.manual_seed(0)
    logits = torch.randn(2, 3, 4, 4, 4, dtype=torch.float64)
    target = torch.randint(0, 3, (2, 1, 4, 4, 4)).double()
    loss = DiceCELoss(to_onehot_y=True, softmax=True, squared_pred=True, smooth_nr=0.0, smooth_dr=1e-6)(logits, target)
    p = torch.softmax(logits, 1)
    y = torch.nn.functional.one_hot(target.squeeze(1).long(), 3).permute(0, 4, 1, 2, 3).double()
    f = 1 - 2 * (p * y).sum((2, 3, 4)) / ((p * p).sum((2, 3, 4)) + (y * y).sum((2, 3, 4)) + 1e-6)
    ce = torch.nn.functional.cross_entropy(logits, target.squeeze(1).long())
    assert abs(float(loss) - float(f.mean() + ce)) < 1e-5


def test_ctunet_loss_weights():
    lf = lambda a, b: a.mean() + 0 * b.mean()
    ones = lambda *s: torch.ones(1, 14, *s)
    t = torch.zeros(1, 1, 96, 96, 96)
    logits = ((ones(96, 96, 96), 2 * ones(48, 48, 96), 4 * ones(24, 24, 48)), (8 * ones(96, 96, 96), 16 * ones(96, 96, 96)))
    assert abs(float(ctunet_loss(logits, t, lf)) - (1 + 0.5 * (2 + 0.5 * 4) + 0.5 * (8 + 16))) < 1e-6


def test_product_loss_equals_oracle_loss():
    """hybrid_ctunet_b200.losses (device-side label gather, product) vs oracle/train_oracle.py (scipy zoom on the host
    + the MONAI formula as the reference runs it) on the same logits."""
    from oracle import train_oracle as T
    torch.manual_seed(3)
    target = torch.randint(0, 14, (1, 1, 96, 96, 96)).float()
    lf = DiceCELoss(to_onehot_y=True, softmax=True, squared_pred=True, smooth_nr=0.0, smooth_dr=1e-6)
    for shape, zoom in (((96, 96, 96), None), ((48, 48, 96), (1, 1, 0.5, 0.5, 1)), ((24, 24, 48), (1, 1, 0.25, 0.25, 0.5))):
        logits = torch.randn(1, 14, *shape)
        t_ref = target if zoom is None else T.zoom_labels(target, zoom)
        t_ours = target if zoom is None else deep_supervision_targets(target)[0 if shape[0] == 48 else 1]
        assert torch.equal(t_ref, t_ours)
        assert abs(float(lf(logits, t_ours)) - float(T.dice_ce_loss(logits, t_ref))) < 1e-5
