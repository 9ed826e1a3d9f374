"""CPU/fp32 restatement of the reference's training-step arithmetic (TEST INFRASTRUCTURE ONLY).

Follows trainer_CTUNet.py:87-109 (train_epoch): forward of CTUNet (oracle/ctunet_oracle.py), five Dice-CE terms with
the deep-supervision weights of :92-103, labels down-sampled with scipy.ndimage.zoom(order=0, prefilter=False) on
the host exactly as :93-94 do, then loss.backward() through torch autograd.  DiceCELoss is MONAI 0.7.0's
(third-party, absent from /root/reference; constructed at main_CTUNet.py:156-158 with to_onehot_y=True,
softmax=True, squared_pred=True, smooth_nr=0.0, smooth_dr=1e-6): DiceLoss(reduction="mean") over (batch, class) of
1 - (2*sum(p*y) + nr) / (sum(p^2) + sum(y^2) + dr), plus nn.CrossEntropyLoss on the squeezed integer labels.
Parity unpinned by the reference (it ships no tests); pinned here against scipy's zoom and a float64 evaluation of
the published formula (tests/test_losses_cpu.py).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F
from scipy import ndimage

from . import ctunet_oracle as O


def dice_ce_loss(logits: torch.Tensor, target: torch.Tensor, smooth_nr: float = 0.0, smooth_dr: float = 1e-6):
    """monai.losses.DiceCELoss(to_onehot_y=True, softmax=True, squared_pred=True, smooth_nr, smooth_dr) (0.7.0)."""
    n_cls = logits.shape[1]
    prob = torch.softmax(logits, 1)
    labels = torch.squeeze(target, dim=1).long()
    onehot = F.one_hot(labels, n_cls).permute(0, 4, 1, 2, 3).to(prob.dtype)
    axes = (2, 3, 4)
    inter = torch.sum(onehot * prob, dim=axes)
    ground = torch.sum(onehot ** 2, dim=axes)
    pred = torch.sum(prob ** 2, dim=axes)
    dice = torch.mean(1.0 - (2.0 * inter + smooth_nr) / (ground + pred + smooth_dr))
    return dice + F.cross_entropy(logits, labels)


def zoom_labels(target: torch.Tensor, zoom):
    """trainer_CTUNet.py:93-94: torch.from_numpy(ndimage.zoom(target.cpu().numpy(), zoom, order=0, prefilter=False))."""
    return torch.from_numpy(ndimage.zoom(target.detach().cpu().numpy(), zoom, order=0, prefilter=False)).to(target.device)


def ctunet_train_loss(sd, x, target, model_depth: int = 101, patch_frame: int = 8):
    """trainer_CTUNet.py:91-103."""
    logits = O.ctunet_forward(sd, x, model_depth, patch_frame)
    t1 = zoom_labels(target, (1, 1, 0.5, 0.5, 1))
    t2 = zoom_labels(target, (1, 1, 0.25, 0.25, 0.5))
    loss1 = dice_ce_loss(logits[0][0], target) + 0.5 * (dice_ce_loss(logits[0][1], t1) + 0.5 * dice_ce_loss(logits[0][2], t2))
    loss2 = dice_ce_loss(logits[1][0], target) + dice_ce_loss(logits[1][1], target)
    return loss1 + 0.5 * loss2


def ctunet_train_step(sd, x, target, model_depth: int = 101, patch_frame: int = 8):
    """One fwd + loss + bwd on leaf copies of `sd`; returns (loss value, {name: grad or None})."""
    leaves = {k: v.detach().clone().requires_grad_() for k, v in sd.items()}
    loss = ctunet_train_loss(leaves, x, target, model_depth, patch_frame)
    loss.backward()
    return float(loss.detach()), {k: v.grad for k, v in leaves.items()}
