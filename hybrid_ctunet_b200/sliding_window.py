"""Sliding-window inference with Gaussian blending on the sm_100a blend kernels.

Drop-in for the reference's forked MONAI-0.7 functions
    trainer_CTUNet.py:417-581  sliding_window_inference  (two blended heads: seg_prob[0][0], seg_prob[1][0])
    trainer_CUNet.py:268-424   sliding_window_inference  (one head: predictor(...)[0])
with the same signature, window order, importance map and error behaviour.  Host-side integer logic (scan
interval, window starts, importance map) restates MONAI 0.7.0's `dense_patch_slices`, `get_valid_patch_size`,
`fall_back_tuple` and `compute_importance_map`; the per-window accumulation and the final divide run in
ctu_blend_accumulate / ctu_blend_count / ctu_blend_normalize.

`device` (the reference's output/accumulator device, trainer_CTUNet.py:534) selects where the RESULT is returned; the
fp32 accumulators themselves live on the input's CUDA device, where the blend kernels run (3.76 GB per head at
512x512x256 against 180 GB of HBM), so `device="cpu"` does not bound GPU memory as it does in the reference.

Differences that do not change results: the count map is kept as ONE channel (the reference accumulates 14
identical channels, trainer_CTUNet.py:535,543) and is cached per geometry; windows may be sharded over ranks
(`shard_group`), which only changes the fp32 summation order of overlapping windows.
"""
from __future__ import annotations

import math
from typing import Any, Callable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn.functional as F

from . import ops

__all__ = ["sliding_window_inference", "sliding_window_inference_one_head", "dense_patch_starts",
           "compute_importance_map", "get_scan_interval"]

_PAD_MODES = {"constant", "reflect", "replicate", "circular"}
_BLEND_MODES = {"constant", "gaussian"}


def fall_back_tuple(user_provided, default) -> Tuple[int, ...]:
    """monai.utils.fall_back_tuple: keep positive user values, else the default (trainer_CTUNet.py:492)."""
    ndim = len(default)
    if isinstance(user_provided, (int, float)) or user_provided is None:
        user = (user_provided,) * ndim
    else:
        user = tuple(user_provided)
        if len(user) != ndim:
            raise ValueError(f"Sequence must have length {ndim}, got {len(user)}.")
    return tuple(int(u) if (u and u > 0) else int(d) for u, d in zip(user, default))


def get_valid_patch_size(image_size: Sequence[int], patch_size) -> Tuple[int, ...]:
    """monai.data.utils.get_valid_patch_size: min(image dim, patch dim or image dim)."""
    patch = fall_back_tuple(patch_size, image_size)
    return tuple(min(int(i), int(p)) for i, p in zip(image_size, patch))


def get_scan_interval(image_size: Sequence[int], roi_size: Sequence[int], num_spatial_dims: int,
                      overlap: float) -> Tuple[int, ...]:
    """trainer_CTUNet.py:560-581."""
    if len(image_size) != num_spatial_dims:
        raise ValueError("image coord different from spatial dims.")
    if len(roi_size) != num_spatial_dims:
        raise ValueError("roi coord different from spatial dims.")
    out = []
    for i in range(num_spatial_dims):
        if roi_size[i] == image_size[i]:
            out.append(int(roi_size[i]))
        else:
            interval = int(roi_size[i] * (1 - overlap))
            out.append(interval if interval > 0 else 1)
    return tuple(out)


def dense_patch_starts(image_size: Sequence[int], patch_size: Sequence[int], scan_interval: Sequence[int]) -> np.ndarray:
    """Window start coordinates [num_windows, ndim] in MONAI 0.7.0 `dense_patch_slices` order (row-major meshgrid)."""
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
        ds = []
        for idx in range(scan_num[dim]):
            s = idx * scan_interval[dim]
            s -= max(s + patch_size[dim] - image_size[dim], 0)
            ds.append(s)
        starts.append(ds)
    return np.asarray([x.flatten() for x in np.meshgrid(*starts, indexing="ij")]).T.astype(np.int64)


def _gaussian_1d_erf(sigma: float, truncated: float = 4.0) -> torch.Tensor:
    """monai.networks.layers.gaussian_1d(approx="erf"), evaluated with torch CPU ops as the reference does."""
    sig = torch.as_tensor(sigma, dtype=torch.float)
    tail = int(max(float(sig) * truncated, 0.5) + 0.5)
    x = torch.arange(-tail, tail + 1, dtype=torch.float)
    t = 0.70710678 / torch.abs(sig)
    out = 0.5 * ((t * (x + 0.5)).erf() - (t * (x - 0.5)).erf())
    return out.clamp(min=0)


def compute_importance_map(patch_size: Sequence[int], mode: str = "constant", sigma_scale=0.125,
                           device="cpu") -> torch.Tensor:
    """monai.data.utils.compute_importance_map (0.7.0).  The Gaussian map is a separable zero-padded filtering of
    a unit impulse at patch_size//2, i.e. the outer product of shifted 1-D erf kernels; built on the CPU in fp32
    (bit-equal to the reference's conv3d chain: every output voxel has a single non-zero term) then moved."""
    mode = str(getattr(mode, "value", mode)).lower()
    if mode not in _BLEND_MODES:
        raise ValueError(f"Unsupported mode: {mode}, available options are {sorted(_BLEND_MODES)}.")
    patch_size = tuple(int(p) for p in patch_size)
    if mode == "constant":
        return torch.ones(patch_size, dtype=torch.float32, device=device)
    nd = len(patch_size)
    scales = (sigma_scale,) * nd if isinstance(sigma_scale, (int, float)) else tuple(sigma_scale)
    imp = None
    for d, (p, ss) in enumerate(zip(patch_size, scales)):
        ker = _gaussian_1d_erf(p * ss)
        tail = (ker.numel() - 1) // 2
        idx = (p // 2) + tail - torch.arange(p)          # out[i] = ker[center + tail - i], zero outside the kernel
        line = torch.where((idx >= 0) & (idx < ker.numel()), ker[idx.clamp(0, ker.numel() - 1)], torch.zeros(()))
        shape = [1] * nd
        shape[d] = p
        imp = line.reshape(shape) if imp is None else imp * line.reshape(shape)
    imp = imp.expand(patch_size).contiguous() if imp.shape != patch_size else imp
    imp = imp / torch.max(imp)
    imp = imp.float()
    min_non_zero = imp[imp != 0].min().item()
    imp = torch.clamp(imp, min=min_non_zero)
    return imp.to(device)


_COUNT_CACHE = {}


def _count_map(image_size, roi, starts: np.ndarray, imp: torch.Tensor, key_extra) -> torch.Tensor:
    """Sum of importance maps over all windows (geometry only); windows added in reference order."""
    key = (tuple(image_size), tuple(roi), starts.tobytes(), str(imp.device), key_extra)
    hit = _COUNT_CACHE.get(key)
    if hit is not None:
        return hit
    cnt = torch.zeros(tuple(image_size), dtype=torch.float32, device=imp.device)
    for s in starts:
        ops.blend_count(imp, cnt, tuple(int(v) for v in s))
    if len(_COUNT_CACHE) > 8:
        _COUNT_CACHE.clear()
    _COUNT_CACHE[key] = cnt
    return cnt


def shard_window_range(total: int, sw_batch_size: int, world: int, rank: int):
    """Windows [lo, hi) of the C-ordered window list that `rank` runs: contiguous chunks (one x-slab of the accumulator per
    rank) whose length is a whole number of `sw_batch_size` calls, so that only the last non-empty rank ever issues a
    ragged predictor call — a 3-window call next to 4-window ones is another problem shape for the network (its own CUDA
    graph / eager path) and costs a full call.  500 windows, batch 4, 8 ranks: 64 x 7 + 52 (was 63 x 7 + 59: one 3-window
    call on every rank)."""
    calls = -(-total // max(int(sw_batch_size), 1))
    per = -(-calls // world) * max(int(sw_batch_size), 1)
    return min(rank * per, total), min((rank + 1) * per, total)


def _sliding_window(inputs: torch.Tensor, roi_size, sw_batch_size: int, predictor: Callable, overlap: float, mode,
                    sigma_scale, padding_mode, cval: float, sw_device, device, two_heads: bool, shard_group,
                    args, kwargs):
    num_spatial_dims = len(inputs.shape) - 2
    if overlap < 0 or overlap >= 1:
        raise AssertionError("overlap must be >= 0 and < 1.")
    if num_spatial_dims != 3:
        raise NotImplementedError("the CUDA blend path is 3-D (the reference only runs 96^3 windows)")
    if not inputs.is_cuda:
        raise RuntimeError("sliding_window_inference runs on the CUDA blend kernels: inputs must be a CUDA tensor")
    if sw_device is not None and torch.device(sw_device) != inputs.device and \
            not (torch.device(sw_device).type == "cuda" and torch.device(sw_device).index is None):
        # the reference moves each window batch to sw_device (trainer_CTUNet.py:526); here windows are views of
        # `inputs` consumed in place, so a different window device cannot be honoured — say so instead of ignoring it
        raise ValueError(f"sw_device={sw_device} differs from inputs.device={inputs.device}: move `inputs` there instead "
                         "(windows are sliced and blended on the input's CUDA device)")
    image_size_ = list(inputs.shape[2:])
    batch_size = inputs.shape[0]
    out_device = inputs.device if device is None else torch.device(device)
    roi_size = fall_back_tuple(roi_size, image_size_)
    image_size = tuple(max(image_size_[i], roi_size[i]) for i in range(num_spatial_dims))
    pad_size = []
    for k in range(len(inputs.shape) - 1, 1, -1):
        diff = max(roi_size[k - 2] - inputs.shape[k], 0)
        half = diff // 2
        pad_size.extend([half, diff - half])
    pmode = str(getattr(padding_mode, "value", padding_mode)).lower()
    if pmode not in _PAD_MODES:
        raise ValueError(f"Unsupported padding_mode: {padding_mode}, available options are {sorted(_PAD_MODES)}.")
    if any(pad_size):
        inputs = F.pad(inputs, pad=pad_size, mode=pmode, value=cval)
    scan_interval = get_scan_interval(image_size, roi_size, num_spatial_dims, overlap)
    starts = dense_patch_starts(image_size, roi_size, scan_interval)
    num_win = len(starts)
    total = num_win * batch_size
    roi = get_valid_patch_size(image_size, roi_size)
    imp = compute_importance_map(roi, mode=mode, sigma_scale=sigma_scale, device="cpu").to(inputs.device)

    # window range of this rank (contiguous chunk of the C-ordered window list; whole range without sharding)
    lo, hi = 0, total
    if shard_group is not None:
        import torch.distributed as dist
        lo, hi = shard_window_range(total, sw_batch_size, dist.get_world_size(shard_group), dist.get_rank(shard_group))

    acc1 = acc2 = None
    for g0 in range(lo, hi, sw_batch_size):
        idxs = range(g0, min(g0 + sw_batch_size, hi))
        wins = []
        for idx in idxs:
            b, s = idx // num_win, starts[idx % num_win]
            wins.append(inputs[b:b + 1, :, s[0]:s[0] + roi[0], s[1]:s[1] + roi[1], s[2]:s[2] + roi[2]])
        window_data = torch.cat(wins)
        seg = predictor(window_data, *args, **kwargs)
        p1 = seg[0][0] if two_heads else seg[0]
        p2 = seg[1][0] if two_heads else None
        if acc1 is None:
            classes = p1.shape[1]
            acc1 = torch.zeros((batch_size, classes) + image_size, dtype=torch.float32, device=inputs.device)
            if two_heads:
                acc2 = torch.zeros_like(acc1)
        p1 = p1.float().contiguous()
        p2 = p2.float().contiguous() if two_heads else None
        for j, idx in enumerate(idxs):
            b, s = idx // num_win, tuple(int(v) for v in starts[idx % num_win])
            ops.blend_accumulate(p1[j], p2[j] if two_heads else None, imp, acc1[b], acc2[b] if two_heads else None, s)

    if shard_group is not None:
        import torch.distributed as dist
        # a rank that owned no window (fewer windows than ranks) still joins the collectives, with zero accumulators;
        # it learns the class count from the ranks that ran the predictor
        ncls = torch.tensor([0 if acc1 is None else acc1.shape[1]], dtype=torch.int64, device=inputs.device)
        dist.all_reduce(ncls, op=dist.ReduceOp.MAX, group=shard_group)
        if acc1 is None:
            acc1 = torch.zeros((batch_size, int(ncls.item())) + image_size, dtype=torch.float32, device=inputs.device)
            if two_heads:
                acc2 = torch.zeros_like(acc1)
        dist.all_reduce(acc1, group=shard_group)
        if two_heads:
            dist.all_reduce(acc2, group=shard_group)

    cnt = _count_map(image_size, roi, starts, imp, (str(mode), str(sigma_scale)))
    for b in range(batch_size):
        ops.blend_normalize(acc1[b], cnt, acc1[b])
        if two_heads:
            ops.blend_normalize(acc2[b], cnt, acc2[b])

    final_slicing: List[slice] = []
    for sp in range(num_spatial_dims):
        final_slicing.insert(0, slice(pad_size[sp * 2], image_size_[num_spatial_dims - sp - 1] + pad_size[sp * 2]))
    while len(final_slicing) < len(acc1.shape):
        final_slicing.insert(0, slice(None))
    final_slicing = tuple(final_slicing)
    if two_heads:
        return (acc1[final_slicing].to(out_device), acc2[final_slicing].to(out_device))
    return acc1[final_slicing].to(out_device)


def sliding_window_inference(inputs: torch.Tensor, roi_size: Union[Sequence[int], int], sw_batch_size: int,
                             predictor: Callable[..., Any], overlap: float = 0.25, mode: str = "constant",
                             sigma_scale: Union[Sequence[float], float] = 0.125, padding_mode: str = "constant",
                             cval: float = 0.0, sw_device=None, device=None, *args: Any, shard_group=None,
                             **kwargs: Any):
    """Two-head variant (trainer_CTUNet.py:417-557): returns (blend(seg[0][0]), blend(seg[1][0]))."""
    return _sliding_window(inputs, roi_size, sw_batch_size, predictor, overlap, mode, sigma_scale, padding_mode, cval,
                           sw_device, device, True, shard_group, args, kwargs)


def sliding_window_inference_one_head(inputs: torch.Tensor, roi_size: Union[Sequence[int], int], sw_batch_size: int,
                                      predictor: Callable[..., Any], overlap: float = 0.25, mode: str = "constant",
                                      sigma_scale: Union[Sequence[float], float] = 0.125,
                                      padding_mode: str = "constant", cval: float = 0.0, sw_device=None, device=None,
                                      *args: Any, shard_group=None, **kwargs: Any):
    """One-head variant (trainer_CUNet.py:268-400): blends predictor(...)[0]."""
    return _sliding_window(inputs, roi_size, sw_batch_size, predictor, overlap, mode, sigma_scale, padding_mode, cval,
                           sw_device, device, False, shard_group, args, kwargs)
