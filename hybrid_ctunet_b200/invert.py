"""`Invertd` of the reference's evaluation scripts on the device.

test_CTUNet.py:162-199 / test_CTUNet_final.py:470-505 carry the blended logits back to the voxel grid of the file with
`transforms.Invertd(keys="pred*", transform=invert_transform, orig_keys="image", nearest_interp=False)`, i.e. the inverse
of utils/data_utils.py:103-116 applied last-to-first on the host: CropForegroundd (zero pad), ScaleIntensityRanged (not
invertible: skipped), Spacingd (trilinear resample back to the file's voxel size: border padding, align_corners=False,
float64 arithmetic, float32 result), Orientationd (flips / transposes back to the file's axis order).  Every step is an
index map, so `InvertGeometry` folds the chain into ONE 3x4 matrix and `invert_pred` / `invert_ensemble_masks` gather
from the cropped prediction in a single kernel (ctu_invert_resample / ctu_invert_ensemble_argmax, csrc/invert.cu).

The metadata arithmetic below restates monai==0.7.0 (`Orientation`, `Spacing`, `zoom_affine`, `compute_shape_offset`,
`CropForegroundd.inverse`) and nibabel==3.1.1 (`io_orientation`, `ornt_transform`) — requirements.txt:1-2; neither is
vendored by the reference nor installed here.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import lib as _lib

__all__ = ["InvertGeometry", "invert_pred", "invert_ensemble_masks"]

_CODES = (("L", "R"), ("P", "A"), ("I", "S"))


def _axis_signs(affine: np.ndarray):
    """For every voxel axis of `affine`: (closest world axis, +1 / -1) — nibabel's io_orientation: the rotation part is
    made orthonormal by an SVD (so shears do not bias the choice) and world axes are handed out greedily."""
    rzs = np.asarray(affine, dtype=np.float64)[:3, :3]
    norms = np.sqrt((rzs * rzs).sum(axis=0))
    norms[norms == 0] = 1
    u, s, vt = np.linalg.svd(rzs / norms, full_matrices=False)
    keep = s > s.max() * 3 * np.finfo(np.float64).eps
    r = u[:, keep] @ vt[keep]
    out = []
    for ax in range(3):
        col = r[:, ax]
        if np.allclose(col, 0):
            raise ValueError("degenerate affine: a voxel axis has no direction")
        w = int(np.argmax(np.abs(col)))
        out.append((w, -1 if col[w] < 0 else 1))
        r[w, :] = 0
    return out


def _axcodes(affine: np.ndarray) -> Tuple[str, ...]:
    return tuple(_CODES[w][0 if sgn < 0 else 1] for w, sgn in _axis_signs(affine))


def _reorient(affine: np.ndarray, shape: Sequence[int], axcodes: Sequence[str]):
    """Index map of `Orientation(axcodes)`: returns (G, new_shape) with old_index = G @ (new_index, 1); the new affine is
    affine @ G.  New axis k is the old axis that points along axcodes[k]'s world axis, reversed when the signs differ."""
    have = _axis_signs(affine)
    G = np.zeros((4, 4))
    G[3, 3] = 1.0
    new_shape = []
    for k, code in enumerate(axcodes[:3]):
        w = next(i for i, pair in enumerate(_CODES) if code in pair)
        want = -1 if code == _CODES[w][0] else 1
        a = next(i for i, (hw, _) in enumerate(have) if hw == w)
        if have[a][1] == want:
            G[a, k] = 1.0
        else:
            G[a, k] = -1.0
            G[a, 3] = shape[a] - 1
        new_shape.append(int(shape[a]))
    return G, tuple(new_shape)


def _zoom_affine(affine: np.ndarray, pixdim: Sequence[float]) -> np.ndarray:
    """Same axes directions, voxel sizes `pixdim`, zero origin (monai zoom_affine, diagonal=False: the polar factor of the
    3x3 block through a Cholesky factorisation of its Gram matrix)."""
    rzs = np.asarray(affine, dtype=np.float64)[:3, :3]
    upper = np.linalg.cholesky(rzs.T @ rzs).T
    scale = np.asarray(pixdim, dtype=np.float64)[:3].copy()
    scale[scale == 0] = 1.0
    out = np.eye(4)
    out[:3, :3] = rzs @ np.linalg.inv(upper) @ np.diag(np.sign(np.diag(upper)) * np.abs(scale))
    return out


def _respace(affine: np.ndarray, shape: Sequence[int], pixdim: Sequence[float]):
    """`Spacing(pixdim)` on a grid: returns (T, new_affine, new_shape) with old_index = T @ (new_index, 1).  The new grid
    keeps the origin (zooming never changes the axis directions, so compute_shape_offset's same-orientation branch
    applies) and covers the corners of the old one; |T - I| <= 1e-3 everywhere means MONAI copies instead of resampling."""
    affine = np.asarray(affine, dtype=np.float64)
    new_affine = _zoom_affine(affine, pixdim)
    corners = np.array([[i, j, k, 1.0] for i in (0.0, shape[0] - 1.0) for j in (0.0, shape[1] - 1.0)
                        for k in (0.0, shape[2] - 1.0)]).T
    world = affine @ corners
    span = np.linalg.inv(new_affine) @ world
    span = span[:3] / span[3]
    new_shape = tuple(int(v) for v in np.round(span.max(axis=1) - span.min(axis=1) + 1.0))
    new_affine[:3, 3] = affine[:3, 3] / affine[3, 3]
    T = np.linalg.inv(affine) @ new_affine
    identity = bool(np.allclose(T, np.eye(4), atol=1e-3))
    return T, new_affine, new_shape, identity


@dataclass
class InvertGeometry:
    """The composite index map of the inverse chain.  Build it with `from_file` (file affine + shape + the loader's
    settings; the crop box is what CropForegroundd recorded) or `from_trace` (MONAI's `image_transforms` entries)."""
    m: np.ndarray                      # 3x4 float64: output index -> fractional index in the padded grid
    out_size: Tuple[int, int, int]     # the file's grid
    pad_size: Tuple[int, int, int]     # the grid before CropForegroundd
    crop_start: Tuple[int, int, int]   # where the (margin-trimmed) prediction starts in the padded grid
    pred_size: Tuple[int, int, int]    # spatial shape of the (margin-trimmed) prediction
    roi_start: Tuple[int, int, int] = (0, 0, 0)   # what to trim from the network's grid first (margins outside the image)
    affine: np.ndarray = field(default_factory=lambda: np.eye(4))   # affine MONAI leaves in the meta dict

    @classmethod
    def from_parts(cls, orient_old_affine, orient_orig_size, spacing_old_affine, spacing_orig_size, cur_affine,
                   crop_orig_size, box_start, box_end, pred_size=None) -> "InvertGeometry":
        box_start, box_end = np.asarray(box_start, dtype=np.int64), np.asarray(box_end, dtype=np.int64)
        crop_orig = np.asarray(crop_orig_size, dtype=np.int64)
        cur = box_end - box_start if pred_size is None else np.asarray(pred_size, dtype=np.int64)
        # CropForegroundd.inverse: trim what a margin padded outside the image, pad the rest back with zeros
        roi_start = np.maximum(-box_start, 0)
        roi_end = cur - np.maximum(box_end - crop_orig, 0)
        pad_to_start = np.maximum(box_start, 0)
        # Spacingd.inverse: Spacing(voxel sizes of the affine it saw) from the current affine, output forced to its input size
        old = np.asarray(spacing_old_affine, dtype=np.float64)
        orig_pixdim = np.sqrt((old * old).sum(axis=0))[:3]
        T, aff, shape, identity = _respace(np.asarray(cur_affine, dtype=np.float64), tuple(crop_orig), orig_pixdim)
        if identity:   # MONAI returns the (padded) array as it is
            T, shape = np.eye(4), tuple(int(v) for v in crop_orig)
        else:
            shape = tuple(int(v) for v in spacing_orig_size)
        # Orientationd.inverse: Orientation(axis codes of the file's affine)
        G, out_shape = _reorient(aff, shape, _axcodes(np.asarray(orient_old_affine, dtype=np.float64)))
        if tuple(out_shape) != tuple(int(v) for v in orient_orig_size):
            raise ValueError(f"trace is inconsistent: inverse orientation gives {out_shape}, the file has {tuple(orient_orig_size)}")
        return cls(m=(T @ G)[:3].copy(), out_size=tuple(out_shape), pad_size=tuple(int(v) for v in crop_orig),
                   crop_start=tuple(int(v) for v in pad_to_start), pred_size=tuple(int(v) for v in (roi_end - roi_start)),
                   roi_start=tuple(int(v) for v in roi_start), affine=aff @ G)

    @classmethod
    def from_file(cls, affine, shape, pixdim, box_start, box_end, axcodes: str = "RAS") -> "InvertGeometry":
        """affine / shape: the file as loaded; pixdim / axcodes: the loader's Spacingd / Orientationd settings
        (utils/data_utils.py:107-108); box_start / box_end: CropForegroundd's bounding box on the resampled grid."""
        a0 = np.asarray(affine, dtype=np.float64)
        G, s1 = _reorient(a0, shape, tuple(axcodes))
        a1 = a0 @ G
        _, a2, s2, identity = _respace(a1, s1, pixdim)
        if identity:
            s2 = s1
        return cls.from_parts(a0, tuple(shape), a1, s1, a2, s2, box_start, box_end)

    @classmethod
    def from_trace(cls, transforms: Sequence[Dict], affine) -> "InvertGeometry":
        """transforms: MONAI's `<key>_transforms` list (dicts with "class", "orig_size", "extra_info"); affine: the
        `<key>_meta_dict["affine"]` left by the forward pass."""
        by = {t["class"]: t for t in transforms}
        o, s, c = by["Orientationd"], by["Spacingd"], by["CropForegroundd"]
        if s["extra_info"].get("padding_mode", "border") != "border":
            raise NotImplementedError("only padding_mode='border' (Spacingd's default, what the loader uses) is implemented")
        if s["extra_info"].get("align_corners", "none") not in ("none", False):
            raise NotImplementedError("only align_corners=False is implemented")
        return cls.from_parts(o["extra_info"]["old_affine"], o["orig_size"], s["extra_info"]["old_affine"], s["orig_size"],
                              affine, c["orig_size"], c["extra_info"]["box_start"], c["extra_info"]["box_end"])

    def c_struct(self, mode: int) -> "_lib.InvertGeom":
        g = _lib.InvertGeom()
        g.m[:] = [float(v) for v in np.asarray(self.m, dtype=np.float64).reshape(-1)]
        g.out_size[:] = self.out_size
        g.pad_size[:] = self.pad_size
        g.crop_start[:] = self.crop_start
        g.pred_size[:] = self.pred_size
        g.mode = mode
        return g


def _trimmed(pred: torch.Tensor, geom: InvertGeometry) -> torch.Tensor:
    p = pred.reshape(pred.shape[-4:])
    if not p.is_cuda:
        raise ValueError("pred must be a CUDA tensor [C, x, y, z]")
    r, n = geom.roi_start, geom.pred_size
    p = p[:, r[0]:r[0] + n[0], r[1]:r[1] + n[1], r[2]:r[2] + n[2]]
    if tuple(p.shape[1:]) != tuple(n):
        raise ValueError(f"pred has spatial shape {tuple(pred.shape[-3:])}, the trace describes {tuple(n)} (+ margins)")
    return p.float().contiguous()


def invert_pred(pred: torch.Tensor, geom: InvertGeometry, nearest_interp: bool = False) -> torch.Tensor:
    """What `Invertd(..., to_tensor=True)` leaves in `batch["pred"]`: fp32 [C, X0, Y0, Z0] on the file's grid."""
    lib = _lib.require_device()
    p = _trimmed(pred, geom)
    out = torch.empty((p.shape[0],) + tuple(geom.out_size), dtype=torch.float32, device=p.device)
    g = geom.c_struct(0 if nearest_interp else 1)
    _lib.check(lib.ctu_invert_resample(p.data_ptr(), p.shape[0], ctypes.addressof(g), out.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream), "ctu_invert_resample")
    return out


def invert_ensemble_masks(pred1: torch.Tensor, pred2: torch.Tensor, geom: InvertGeometry,
                          labels: Optional[torch.Tensor] = None, nearest_interp: bool = False) -> Dict[str, torch.Tensor]:
    """test_CTUNet.py:222-251 in one kernel: Invertd of both models' blended logits, softmax of each, their mean, the three
    argmax masks and — with `labels` on the file's grid — the per-class Dice of each mask.  Same result dict as
    `ensemble.ensemble_masks(invert_pred(pred1), invert_pred(pred2), labels)`, without the two inverted volumes."""
    lib = _lib.require_device()
    p1, p2 = _trimmed(pred1, geom), _trimmed(pred2, geom)
    if p1.shape != p2.shape:
        raise ValueError("pred1 / pred2 must have the same shape")
    C = p1.shape[0]
    out = {k: torch.empty(tuple(geom.out_size), dtype=torch.uint8, device=p1.device) for k in ("ensemble", "head1", "head2")}
    lab = counts = None
    if labels is not None:
        lab = labels.reshape(tuple(geom.out_size)).to(device=p1.device, dtype=torch.float32).contiguous()
        counts = torch.zeros(3, C, 3, dtype=torch.int64, device=p1.device)
    g = geom.c_struct(0 if nearest_interp else 1)
    _lib.check(lib.ctu_invert_ensemble_argmax(p1.data_ptr(), p2.data_ptr(), C, ctypes.addressof(g), out["ensemble"].data_ptr(),
                                              out["head1"].data_ptr(), out["head2"].data_ptr(),
                                              None if lab is None else lab.data_ptr(),
                                              None if counts is None else counts.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream), "ctu_invert_ensemble_argmax")
    if counts is not None:
        c = counts.double()
        out["dice"] = torch.where(c[..., 2] > 0, 2.0 * c[..., 0] / (c[..., 1] + c[..., 2]).clamp_min(1.0), torch.zeros_like(c[..., 0]))
        out["counts"] = counts
    return out
