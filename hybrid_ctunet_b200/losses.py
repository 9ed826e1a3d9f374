"""Loss of the CTUNet training step (trainer_CTUNet.py:90-103) with the label handling kept on the device.

`DiceCELoss` restates the subset of monai.losses.DiceCELoss (MONAI 0.7.0) the reference constructs
(main_CTUNet.py:156-158: to_onehot_y=True, softmax=True, squared_pred=True, smooth_nr=0.0, smooth_dr=1e-6); on CUDA
tensors it runs on the fused kernels ctu_dice_ce_fwd / ctu_dice_ce_bwd (csrc/loss.cu), the torch expression below
is the definition used for CPU tensors (tests) and for configurations the reference never builds:
    dice = mean over (batch, class) of 1 - (2*sum(p*y) + smooth_nr) / (sum(p^2) + sum(y^2) + smooth_dr)
    ce   = nn.CrossEntropyLoss()(logits, labels)            total = dice + ce
`deep_supervision_targets` replaces the two scipy.ndimage.zoom(order=0) host round trips per step
(trainer_CTUNet.py:93-94) with an index gather on the device that selects exactly the voxels zoom selects.
`ctunet_loss` is the reference's weighting of the five heads (trainer_CTUNet.py:92-103).
"""
from __future__ import annotations

from functools import lru_cache
from typing import Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


_FUSED_CLASSES = (2, 3, 4, 14)


class _DiceCEFused(torch.autograd.Function):
    """Dice-CE on the fused CUDA kernels (ctu_dice_ce_fwd / ctu_dice_ce_bwd): two passes over the logits in total."""

    @staticmethod
    def forward(ctx, logits, target, smooth_nr, smooth_dr, lambda_dice, lambda_ce):
        from . import lib as _lib
        lib = _lib.require_device()
        logits = logits.float().contiguous()
        target = target.float().contiguous()
        B, C = logits.shape[:2]
        S = logits.numel() // (B * C)
        if target.numel() != B * S:
            raise ValueError(f"ground truth has different shape ({tuple(target.shape)}) from input ({tuple(logits.shape)})")
        sums = torch.zeros(B * C * 3 + 1, dtype=torch.float64, device=logits.device)
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.ctu_dice_ce_fwd(logits.data_ptr(), target.data_ptr(), B, C, S, sums.data_ptr(), stream), "ctu_dice_ce_fwd")
        s = sums[:B * C * 3].view(B, C, 3)
        inter, den = s[..., 0], s[..., 1] + s[..., 2] + smooth_dr
        dice = (1.0 - (2.0 * inter + smooth_nr) / den).mean()
        ce = sums[-1] / float(B * S)
        ctx.save_for_backward(logits, target, inter, den)
        ctx.cfg = (B, C, S, smooth_nr, lambda_dice, lambda_ce)
        return (lambda_dice * dice + lambda_ce * ce).float()

    @staticmethod
    def backward(ctx, g):
        from . import lib as _lib
        lib = _lib.require_device()
        logits, target, inter, den = ctx.saved_tensors
        B, C, S, smooth_nr, lambda_dice, lambda_ce = ctx.cfg
        scale = g.double() * (lambda_dice / float(B * C))
        coef = torch.stack((scale * (-2.0 / den), scale * (2.0 * (2.0 * inter + smooth_nr) / (den * den))), dim=-1)
        coef = coef.float().contiguous()
        ce_scale = (g.double() * (lambda_ce / float(B * S))).float().reshape(1).contiguous()
        dlogits = torch.empty_like(logits)
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.ctu_dice_ce_bwd(logits.data_ptr(), target.data_ptr(), B, C, S, coef.data_ptr(), ce_scale.data_ptr(),
                                       dlogits.data_ptr(), stream), "ctu_dice_ce_bwd")
        return dlogits, None, None, None, None, None


class DiceCELoss(nn.Module):
    def __init__(self, include_background: bool = True, to_onehot_y: bool = False, sigmoid: bool = False,
                 softmax: bool = False, squared_pred: bool = False, jaccard: bool = False, reduction: str = "mean",
                 smooth_nr: float = 1e-5, smooth_dr: float = 1e-5, batch: bool = False, lambda_dice: float = 1.0,
                 lambda_ce: float = 1.0):
        super().__init__()
        if not include_background or sigmoid or jaccard or batch or reduction != "mean" or not softmax:
            raise NotImplementedError("only the configuration the reference builds is restated (main_CTUNet.py:156-158)")
        self.to_onehot_y, self.squared_pred = to_onehot_y, squared_pred
        self.smooth_nr, self.smooth_dr = float(smooth_nr), float(smooth_dr)
        self.lambda_dice, self.lambda_ce = lambda_dice, lambda_ce

    def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        n_cls = input.shape[1]
        if (input.is_cuda and self.squared_pred and target.shape[1] == 1 and n_cls in _FUSED_CLASSES and input.dim() == 5):
            # the configuration the reference trains with: fused kernels (no eager fallback on the GPU)
            return _DiceCEFused.apply(input, target, self.smooth_nr, self.smooth_dr, self.lambda_dice, self.lambda_ce)
        logits = input.float()
        labels = target.squeeze(1).long() if target.shape[1] == 1 else target.argmax(1)
        prob = torch.softmax(logits, 1)
        onehot = F.one_hot(labels, n_cls).movedim(-1, 1).to(prob.dtype) if self.to_onehot_y or target.shape[1] == 1 \
            else target.float()
        axes = tuple(range(2, input.dim()))
        inter = (prob * onehot).sum(axes)
        if self.squared_pred:
            den = (prob * prob).sum(axes) + (onehot * onehot).sum(axes)
        else:
            den = prob.sum(axes) + onehot.sum(axes)
        dice = (1.0 - (2.0 * inter + self.smooth_nr) / (den + self.smooth_dr)).mean()
        ce = F.cross_entropy(logits, labels)
        return self.lambda_dice * dice + self.lambda_ce * ce


@lru_cache(maxsize=None)
def zoom_indices(n_in: int, n_out: int) -> Tuple[int, ...]:
    """Source index per output index of scipy.ndimage.zoom(order=0, prefilter=False, grid_mode=False, mode='constant'):
    nearest neighbour of out_idx * ((n_in - 1) / (n_out - 1)), the scale rounded to double FIRST as scipy does.  Where that
    product lands above n_in - 1 by a rounding error (32 -> 16: 15 * (31/15) = 31.000000000000004) scipy treats the sample
    as outside the volume and writes cval = 0; such positions are returned as -1 (none occur for the reference's
    96 -> 48 / 96 -> 24 label volumes)."""
    if n_out == 1:
        return (0,)
    pos = np.arange(n_out, dtype=np.float64) * (float(n_in - 1) / float(n_out - 1))
    idx = np.floor(pos + 0.5).astype(np.int64)
    idx[pos > float(n_in - 1)] = -1
    return tuple(int(v) for v in idx)


_ZOOM_DEVICE_CACHE = {}


def _zoom_index_tensor(n_in: int, n_out: int, device) -> torch.Tensor:
    key = (n_in, n_out, str(device))
    t = _ZOOM_DEVICE_CACHE.get(key)
    if t is None:  # cached on the device: no host-to-device copy inside a CUDA-graph capture
        t = torch.as_tensor(zoom_indices(n_in, n_out), device=device)
        _ZOOM_DEVICE_CACHE[key] = t
    return t


def zoom_nearest(target: torch.Tensor, zoom: Sequence[float]) -> torch.Tensor:
    """ndimage.zoom(target, zoom, order=0, prefilter=False) on the device (axes are independent)."""
    out = target
    for ax, z in enumerate(zoom):
        n_in = target.shape[ax]
        n_out = int(round(n_in * z))
        if n_out != n_in:
            idx = _zoom_index_tensor(n_in, n_out, target.device)
            out = out.index_select(ax, idx.clamp(min=0))
            if min(zoom_indices(n_in, n_out)) < 0:      # scipy's out-of-volume samples (cval = 0)
                shape = [1] * out.dim()
                shape[ax] = n_out
                out = out * (idx >= 0).to(out.dtype).view(shape)
    return out


def deep_supervision_targets(target: torch.Tensor):
    """trainer_CTUNet.py:93-94: labels at 1/2,1/2,1 and 1/4,1/4,1/2 resolution."""
    return zoom_nearest(target, (1, 1, 0.5, 0.5, 1)), zoom_nearest(target, (1, 1, 0.25, 0.25, 0.5))


class _MultiHeadDiceCEFused(torch.autograd.Function):
    """sum_h weight_h * DiceCE(logits_h, target_h) for the heads of one training step: one reduction pass per head
    (ctu_dice_ce_fwd), ONE finalize launch for the scalar and every backward coefficient (ctu_dice_ce_finalize), one
    gradient pass per head (ctu_dice_ce_bwd) — no tiny torch kernels in between (they are the serial section between the
    forward and the backward of the step)."""

    @staticmethod
    def forward(ctx, cfg, weights, targets, *logits):
        from . import lib as _lib
        lib = _lib.require_device()
        smooth_nr, smooth_dr, lambda_dice, lambda_ce = cfg
        dev = logits[0].device
        stream = torch.cuda.current_stream().cuda_stream
        heads = _lib.LossHeads()
        heads.n_heads = len(logits)
        heads.lambda_dice, heads.lambda_ce, heads.smooth_nr, heads.smooth_dr = lambda_dice, lambda_ce, smooth_nr, smooth_dr
        lg, tg, so, co = [], [], 0, 0
        for h, (l, t, w) in enumerate(zip(logits, targets, weights)):
            l = l.float().contiguous()
            t = t.float().contiguous()
            B, C = l.shape[:2]
            S = l.numel() // (B * C)
            if t.numel() != B * S:
                raise ValueError(f"ground truth has different shape ({tuple(t.shape)}) from input ({tuple(l.shape)})")
            heads.B[h], heads.C[h], heads.S[h], heads.weight[h] = B, C, S, float(w)
            heads.sums_off[h], heads.coef_off[h] = so, co
            so += B * C * 3 + 1
            co += B * C * 2
            lg.append(l)
            tg.append(t)
        sums = torch.zeros(so, dtype=torch.float64, device=dev)
        coef = torch.empty(co, dtype=torch.float32, device=dev)
        ce_scale = torch.empty(len(logits), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        for h, (l, t) in enumerate(zip(lg, tg)):
            _lib.check(lib.ctu_dice_ce_fwd(l.data_ptr(), t.data_ptr(), heads.B[h], heads.C[h], heads.S[h],
                                           sums.data_ptr() + 8 * heads.sums_off[h], stream), "ctu_dice_ce_fwd")
        import ctypes as C_
        _lib.check(lib.ctu_dice_ce_finalize(C_.byref(heads), sums.data_ptr(), loss.data_ptr(), coef.data_ptr(),
                                            ce_scale.data_ptr(), stream), "ctu_dice_ce_finalize")
        ctx.save_for_backward(coef, ce_scale, *lg, *tg)
        ctx.meta = [(heads.B[h], heads.C[h], heads.S[h], heads.coef_off[h]) for h in range(len(logits))]
        return loss

    @staticmethod
    def backward(ctx, g):
        from . import lib as _lib
        lib = _lib.require_device()
        saved = ctx.saved_tensors
        n = len(ctx.meta)
        coef, ce_scale, lg, tg = saved[0], saved[1], saved[2:2 + n], saved[2 + n:2 + 2 * n]
        gf = g.float()
        coef_g, ce_g = coef * gf, ce_scale * gf          # the upstream gradient (1, or GradScaler's scale)
        stream = torch.cuda.current_stream().cuda_stream
        outs = []
        for h, (B, C, S, off) in enumerate(ctx.meta):
            d = torch.empty_like(lg[h])
            _lib.check(lib.ctu_dice_ce_bwd(lg[h].data_ptr(), tg[h].data_ptr(), B, C, S, coef_g.data_ptr() + 4 * off,
                                           ce_g.data_ptr() + 4 * h, d.data_ptr(), stream), "ctu_dice_ce_bwd")
            outs.append(d)
        return (None, None, None, *outs)


_ZOOM_I32_CACHE = {}


def _zoom_index_i32(n_in: int, n_out: int, device) -> torch.Tensor:
    key = (n_in, n_out, str(device))
    t = _ZOOM_I32_CACHE.get(key)
    if t is None:
        t = torch.as_tensor(zoom_indices(n_in, n_out), dtype=torch.int32, device=device)
        _ZOOM_I32_CACHE[key] = t
    return t


def _gather_targets(target: torch.Tensor, zoom: Sequence[float]) -> torch.Tensor:
    """zoom_nearest for a CUDA [B, 1, X, Y, Z] label volume in one launch (ctu_gather3d)."""
    from . import lib as _lib
    lib = _lib.require_device()
    t = target.float().contiguous()
    B, c, X, Y, Z = t.shape
    Xo, Yo, Zo = int(round(X * zoom[2])), int(round(Y * zoom[3])), int(round(Z * zoom[4]))
    out = torch.empty(B, c, Xo, Yo, Zo, dtype=torch.float32, device=t.device)
    ix, iy, iz = (_zoom_index_i32(a, b, t.device) for a, b in ((X, Xo), (Y, Yo), (Z, Zo)))
    _lib.check(lib.ctu_gather3d(t.data_ptr(), out.data_ptr(), B * c, X, Y, Z, Xo, Yo, Zo, ix.data_ptr(), iy.data_ptr(),
                                iz.data_ptr(), torch.cuda.current_stream().cuda_stream), "ctu_gather3d")
    return out


def ctunet_loss(logits, target: torch.Tensor, loss_func) -> torch.Tensor:
    """trainer_CTUNet.py:92-103: loss1 = l(full) + 0.5*(l(1/2) + 0.5*l(1/4)); loss = loss1 + 0.5*(l(vit) + l(vit_96)).
    With the reference's DiceCELoss configuration on CUDA tensors the five heads go through ONE fused autograd node
    (_MultiHeadDiceCEFused) and the label volumes are gathered by ctu_gather3d; anything else takes the composition."""
    heads = (logits[0][0], logits[0][1], logits[0][2], logits[1][0], logits[1][1])
    import os
    fused = (os.environ.get("CTU_FUSED_LOSS", "1") != "0"   # (0: the per-head composition, for A/B comparisons)
             and isinstance(loss_func, DiceCELoss) and loss_func.squared_pred and target.is_cuda and target.dim() == 5
             and target.shape[1] == 1 and all(h.is_cuda and h.dim() == 5 and h.shape[1] in _FUSED_CLASSES for h in heads))
    if fused:
        t1 = _gather_targets(target, (1, 1, 0.5, 0.5, 1))
        t2 = _gather_targets(target, (1, 1, 0.25, 0.25, 0.5))
        cfg = (loss_func.smooth_nr, loss_func.smooth_dr, loss_func.lambda_dice, loss_func.lambda_ce)
        return _MultiHeadDiceCEFused.apply(cfg, (1.0, 0.5, 0.25, 0.5, 0.5), (target, t1, t2, target, target), *heads)
    t1, t2 = deep_supervision_targets(target)
    loss1 = loss_func(logits[0][0], target) + 0.5 * (loss_func(logits[0][1], t1) + 0.5 * loss_func(logits[0][2], t2))
    loss2 = loss_func(logits[1][0], target) + loss_func(logits[1][1], target)
    return loss1 + 0.5 * loss2
