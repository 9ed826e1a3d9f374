"""CPU restatement of the reference's sliding-window inference (TEST INFRASTRUCTURE ONLY).

Follows trainer_CTUNet.py:417-581 (two heads) / trainer_CUNet.py:268-424 (one head) statement by statement,
including the 14-channel count maps, with MONAI 0.7.0's helpers restated from their published source
(monai/data/utils.py: dense_patch_slices, get_valid_patch_size, compute_importance_map;
monai/networks/layers/simplelayers.py: GaussianFilter / gaussian_1d / separable_filtering;
monai/utils/misc.py: fall_back_tuple).  MONAI is a third-party dependency absent from /root/reference
(requirements.txt:1 pins monai==0.7.0); the reference has no tests for this path, so it is pinned by the
known-answer values of SURVEY 8c (tests/test_sliding_window_cpu.py) — "parity unpinned" by reference tests.
"""
from __future__ import annotations

import math
from typing import Callable, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F


def fall_back_tuple(user_provided, default):
    nd = len(default)
    user = (user_provided,) * nd if not isinstance(user_provided, (tuple, list)) else tuple(user_provided)
    return tuple(u if (u and u > 0) else d for u, d in zip(user, default))


def get_valid_patch_size(image_size, patch_size):
    patch = fall_back_tuple(patch_size, image_size)
    return tuple(min(ms, ps or ms) for ms, ps in zip(image_size, patch))


def dense_patch_slices(image_size, patch_size, scan_interval):
    nd = len(image_size)
    patch_size = get_valid_patch_size(image_size, patch_size)
    scan_num = []
    for i in range(nd):
        if scan_interval[i] == 0:
            scan_num.append(1)
        else:
            num = int(math.ceil(float(image_size[i]) / scan_interval[i]))
            scan_dim = next((d for d in range(num) if d * scan_interval[i] + patch_size[i] >= image_size[i]), None)
            scan_num.append(scan_dim + 1 if scan_dim is not None else 1)
    starts = []
    for dim in range(nd):
        dim_starts = []
        for idx in range(scan_num[dim]):
            start_idx = idx * scan_interval[dim]
            start_idx -= max(start_idx + patch_size[dim] - image_size[dim], 0)
            dim_starts.append(start_idx)
        starts.append(dim_starts)
    out = np.asarray([x.flatten() for x in np.meshgrid(*starts, indexing="ij")]).T
    return [tuple(slice(int(s), int(s) + patch_size[d]) for d, s in enumerate(x)) for x in out]


def gaussian_1d(sigma, truncated: float = 4.0):
    """approx='erf' branch of monai.networks.layers.gaussian_1d."""
    sigma = torch.as_tensor(sigma, dtype=torch.float)
    tail = int(max(float(sigma) * truncated, 0.5) + 0.5)
    x = torch.arange(-tail, tail + 1, dtype=torch.float)
    t = 0.70710678 / torch.abs(sigma)
    out = 0.5 * ((t * (x + 0.5)).erf() - (t * (x - 0.5)).erf())
    return out.clamp(min=0)


def gaussian_filter_3d(x, sigmas):
    """GaussianFilter(3, sigmas, truncated=4.0, approx='erf') forward = separable zero-padded conv, axis 0 first."""
    for d, s in enumerate(sigmas):
        k = gaussian_1d(s)
        shape = [1, 1, 1, 1, 1]
        shape[d + 2] = -1
        pad = [0, 0, 0]
        pad[d] = (k.numel() - 1) // 2
        x = F.conv3d(x, k.reshape(shape), padding=pad)
    return x


def compute_importance_map(patch_size, mode="constant", sigma_scale=0.125):
    if mode == "constant":
        return torch.ones(patch_size).float()
    center = [i // 2 for i in patch_size]
    scales = (sigma_scale,) * len(patch_size) if isinstance(sigma_scale, (int, float)) else tuple(sigma_scale)
    sigmas = [i * s for i, s in zip(patch_size, scales)]
    imp = torch.zeros(patch_size)
    imp[tuple(center)] = 1
    imp = gaussian_filter_3d(imp[None, None], sigmas)[0, 0]
    imp = imp / torch.max(imp)
    imp = imp.float()
    min_non_zero = imp[imp != 0].min().item()
    return torch.clamp(imp, min=min_non_zero)


def get_scan_interval(image_size, roi_size, num_spatial_dims, overlap):
    """trainer_CTUNet.py:560-581."""
    if len(image_size) != num_spatial_dims:
        raise ValueError("image coord different from spatial dims.")
    if len(roi_size) != num_spatial_dims:
        raise ValueError("roi coord different from spatial dims.")
    scan_interval = []
    for i in range(num_spatial_dims):
        if roi_size[i] == image_size[i]:
            scan_interval.append(int(roi_size[i]))
        else:
            interval = int(roi_size[i] * (1 - overlap))
            scan_interval.append(interval if interval > 0 else 1)
    return tuple(scan_interval)


def sliding_window_inference(inputs, roi_size, sw_batch_size, predictor: Callable, overlap=0.25, mode="constant",
                             sigma_scale=0.125, padding_mode="constant", cval=0.0, two_heads=True):
    """trainer_CTUNet.py:478-557 (two_heads) / trainer_CUNet.py:329-400 (one head)."""
    num_spatial_dims = len(inputs.shape) - 2
    if overlap < 0 or overlap >= 1:
        raise AssertionError("overlap must be >= 0 and < 1.")
    image_size_ = list(inputs.shape[2:])
    batch_size = inputs.shape[0]
    roi_size = fall_back_tuple(roi_size, image_size_)
    image_size = tuple(max(image_size_[i], roi_size[i]) for i in range(num_spatial_dims))
    pad_size = []
    for k in range(len(inputs.shape) - 1, 1, -1):
        diff = max(roi_size[k - 2] - inputs.shape[k], 0)
        half = diff // 2
        pad_size.extend([half, diff - half])
    inputs = F.pad(inputs, pad=pad_size, mode=padding_mode, value=cval)
    scan_interval = get_scan_interval(image_size, roi_size, num_spatial_dims, overlap)
    slices = dense_patch_slices(image_size, roi_size, scan_interval)
    num_win = len(slices)
    total_slices = num_win * batch_size
    importance_map = compute_importance_map(get_valid_patch_size(image_size, roi_size), mode=mode,
                                            sigma_scale=sigma_scale).to(inputs.device)
    outs: List[torch.Tensor] = []
    cnts: List[torch.Tensor] = []
    for slice_g in range(0, total_slices, sw_batch_size):
        slice_range = range(slice_g, min(slice_g + sw_batch_size, total_slices))
        unravel_slice = [[slice(int(idx / num_win), int(idx / num_win) + 1), slice(None)] + list(slices[idx % num_win])
                         for idx in slice_range]
        window_data = torch.cat([inputs[tuple(w)] for w in unravel_slice])
        seg_prob = predictor(window_data)
        probs = [seg_prob[0][0], seg_prob[1][0]] if two_heads else [seg_prob[0]]
        if not outs:
            shape = [batch_size, probs[0].shape[1]] + list(image_size)
            outs = [torch.zeros(shape, dtype=torch.float32, device=inputs.device) for _ in probs]
            cnts = [torch.zeros(shape, dtype=torch.float32, device=inputs.device) for _ in probs]
        for idx, original_idx in zip(slice_range, unravel_slice):
            for o, c, p in zip(outs, cnts, probs):
                o[tuple(original_idx)] += importance_map * p[idx - slice_g]
                c[tuple(original_idx)] += importance_map
    outs = [o / c for o, c in zip(outs, cnts)]
    final_slicing = []
    for sp in range(num_spatial_dims):
        final_slicing.insert(0, slice(pad_size[sp * 2], image_size_[num_spatial_dims - sp - 1] + pad_size[sp * 2]))
    while len(final_slicing) < len(outs[0].shape):
        final_slicing.insert(0, slice(None))
    outs = [o[tuple(final_slicing)] for o in outs]
    return tuple(outs) if two_heads else outs[0]


def ensemble_reference(pred1, pred2, labels=None, n_classes: int = 14):
    """test_CTUNet.py:236-251 restated (torch softmax / argmax as the reference runs them; numpy `dice` of
    utils/utils.py:16-22): returns the three masks and, with labels, dice[3][n_classes]."""
    import numpy as np
    import torch
    i1, i2 = torch.softmax(pred1, 0), torch.softmax(pred2, 0)
    ens = (i1 + i2) / 2.0
    masks = [torch.argmax(ens, 0).cpu().numpy(), torch.argmax(i1, 0).cpu().numpy(), torch.argmax(i2, 0).cpu().numpy()]
    out = {"ensemble": masks[0], "head1": masks[1], "head2": masks[2]}
    if labels is not None:
        lab = labels.cpu().numpy()

        def dice(x, y):
            inter = np.sum(x * y)
            ys = np.sum(y)
            if ys == 0:
                return 0.0
            return 2 * inter / (np.sum(x) + ys)
        out["dice"] = np.array([[dice(m == i, lab == i) for i in range(n_classes)] for m in masks])
    return out
