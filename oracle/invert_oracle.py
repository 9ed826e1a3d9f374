"""CPU restatement of the `Invertd` step of the reference's evaluation scripts (TEST INFRASTRUCTURE ONLY).

test_CTUNet.py:162-199 / test_CTUNet_final.py:470-505 wrap the loader's `invert_transform` (utils/data_utils.py:103-116:
LoadImaged, AddChanneld, Orientationd("RAS"), Spacingd(pixdim, "bilinear"), ScaleIntensityRanged, CropForegroundd,
ToTensord) in `transforms.Invertd(keys="pred*", orig_keys="image", nearest_interp=False)`: the blended logits
[14, x, y, z] of the cropped, resampled, RAS-oriented grid are carried back to the grid of the file — CropForegroundd is
undone by zero padding, Spacingd by a trilinear resample (border padding, align_corners=False, computed in float64,
stored as float32), Orientationd by flips and transposes; ScaleIntensityRanged has no inverse and is skipped.

PARITY UNPINNED for the transform metadata: the algorithm lives in third-party dependencies that are absent from
/root/reference and from this image — `monai==0.7.0` and `nibabel==3.1.1` (requirements.txt:1-2).  Their published
algorithms are restated here function by function (names kept); the resampling itself is executed by torch's own
`affine_grid` / `grid_sample` on the CPU in float64, which is what MONAI 0.7.0's `AffineTransform` layer calls.  The
stepwise structure (pad -> resample -> flip/transpose on real arrays) is deliberately different from the product's single
composite index map, so the two check each other.
"""
from __future__ import annotations

import itertools
from typing import Dict, Sequence, Tuple

import numpy as np
import torch

LABELS = (("L", "R"), ("P", "A"), ("I", "S"))


# ----------------------------------------------------------------------------------------------- nibabel.orientations
def io_orientation(affine: np.ndarray) -> np.ndarray:
    """nibabel.orientations.io_orientation: for every input axis, the output (world) axis it is closest to and its sign."""
    affine = np.asarray(affine, dtype=np.float64)
    q, p = affine.shape[0] - 1, affine.shape[1] - 1
    rzs = affine[:q, :p]
    zooms = np.sqrt(np.sum(rzs * rzs, axis=0))
    zooms[zooms == 0] = 1
    rs = rzs / zooms
    P, S, Qs = np.linalg.svd(rs, full_matrices=False)
    tol = S.max() * max(rs.shape) * np.finfo(S.dtype).eps
    keep = S > tol
    R = np.dot(P[:, keep], Qs[keep])
    ornt = np.ones((p, 2), dtype=np.float64) * np.nan
    for in_ax in range(p):
        col = R[:, in_ax]
        if not np.allclose(col, 0):
            out_ax = int(np.argmax(np.abs(col)))
            ornt[in_ax, 0] = out_ax
            ornt[in_ax, 1] = -1 if col[out_ax] < 0 else 1
            R[out_ax, :] = 0
    return ornt


def axcodes2ornt(axcodes: Sequence[str]) -> np.ndarray:
    ornt = np.ones((len(axcodes), 2), dtype=np.float64) * np.nan
    for code_idx, code in enumerate(axcodes):
        for label_idx, codes in enumerate(LABELS):
            if code in codes:
                ornt[code_idx, :] = [label_idx, -1 if code == codes[0] else 1]
                break
    return ornt


def ornt2axcodes(ornt: np.ndarray) -> Tuple[str, ...]:
    return tuple(LABELS[int(axno)][0 if direction == -1 else 1] for axno, direction in ornt)


def aff2axcodes(aff: np.ndarray) -> Tuple[str, ...]:
    return ornt2axcodes(io_orientation(aff))


def ornt_transform(start_ornt: np.ndarray, end_ornt: np.ndarray) -> np.ndarray:
    result = np.empty_like(start_ornt)
    for end_in_idx, (end_out_idx, end_flip) in enumerate(end_ornt):
        for start_in_idx, (start_out_idx, start_flip) in enumerate(start_ornt):
            if end_out_idx == start_out_idx:
                result[start_in_idx, :] = [end_in_idx, 1 if start_flip == end_flip else -1]
                break
        else:
            raise ValueError("Unable to find out axis %d in start_ornt" % end_out_idx)
    return result


def apply_orientation(arr: np.ndarray, ornt: np.ndarray) -> np.ndarray:
    t_arr = np.asarray(arr)
    n = ornt.shape[0]
    for ax, flip in enumerate(ornt[:, 1]):
        if flip == -1:
            t_arr = np.flip(t_arr, axis=ax)
    full_transpose = np.arange(t_arr.ndim)
    full_transpose[:n] = np.argsort(ornt[:, 0])
    return t_arr.transpose(full_transpose)


def inv_ornt_aff(ornt: np.ndarray, shape: Sequence[int]) -> np.ndarray:
    p = ornt.shape[0]
    shape = np.array(shape)[:p]
    axis_transpose = [int(v) for v in ornt[:, 0]]
    undo_reorder = np.eye(p + 1)[axis_transpose + [p], :]
    undo_flip = np.diag(list(ornt[:, 1]) + [1.0])
    center_trans = -(shape - 1) / 2.0
    undo_flip[:p, p] = (ornt[:, 1] * center_trans) - center_trans
    return np.dot(undo_flip, undo_reorder)


# ------------------------------------------------------------------------------------------------ monai.data.utils
def zoom_affine(affine: np.ndarray, scale: Sequence[float]) -> np.ndarray:
    """monai.data.utils.zoom_affine(diagonal=False): same rotation, new voxel sizes, zero translation."""
    affine = np.array(affine, dtype=float, copy=True)
    d = len(affine) - 1
    scale_np = np.array(scale, dtype=float)[:d]
    scale_np[scale_np == 0] = 1.0
    rzs = affine[:-1, :-1]
    zs = np.linalg.cholesky(rzs.T @ rzs).T
    rotation = rzs @ np.linalg.inv(zs)
    s = np.sign(np.diag(zs)) * np.abs(scale_np)
    new_affine = np.eye(len(affine))
    new_affine[:-1, :-1] = rotation @ np.diag(s)
    return new_affine


def compute_shape_offset(spatial_shape, in_affine: np.ndarray, out_affine: np.ndarray):
    """monai.data.utils.compute_shape_offset: the output grid that covers the corners of the input grid."""
    shape = np.array(spatial_shape, copy=True, dtype=float)
    sr = len(shape)
    in_coords = [(0.0, dim - 1.0) for dim in shape]
    corners = np.asarray(np.meshgrid(*in_coords, indexing="ij")).reshape((len(shape), -1))
    corners = np.concatenate((corners, np.ones_like(corners[:1])))
    corners = in_affine @ corners
    corners_out = np.linalg.inv(out_affine) @ corners
    corners_out = corners_out[:-1] / corners_out[-1]
    out_shape = np.round(np.ptp(corners_out, axis=1) + 1.0)
    if np.allclose(io_orientation(in_affine), io_orientation(out_affine)):
        offset = in_affine @ ([0] * sr + [1])
        offset = offset[:-1] / offset[-1]
    else:
        corners = corners[:-1] / corners[-1]
        offset = np.min(corners, 1)
    return out_shape.astype(int), offset


# ------------------------------------------------------------------------ monai.networks.layers.AffineTransform
def normalize_transform(shape, align_corners: bool = False) -> torch.Tensor:
    norm = torch.tensor(shape, dtype=torch.float64)
    if align_corners:
        norm[norm <= 1.0] = 2.0
        norm = 2.0 / (norm - 1.0)
        norm = torch.diag(torch.cat((norm, torch.ones((1,), dtype=torch.float64))))
        norm[:-1, -1] = -1.0
    else:
        norm[norm <= 0.0] = 2.0
        norm = 2.0 / norm
        norm = torch.diag(torch.cat((norm, torch.ones((1,), dtype=torch.float64))))
        norm[:-1, -1] = 1.0 / torch.tensor(shape, dtype=torch.float64) - 1.0
    return norm.unsqueeze(0)


def affine_transform(src: torch.Tensor, theta: torch.Tensor, spatial_size, mode: str, padding_mode: str,
                     align_corners: bool = False) -> torch.Tensor:
    """AffineTransform(normalized=False, reverse_indexing=True).forward: `theta` maps output voxel indices to input voxel
    indices; it is conjugated into torch's normalised coordinates, the axes are reversed (torch's grids are x-fastest) and
    torch.nn.functional.affine_grid / grid_sample do the work."""
    sr = src.dim() - 2
    theta = theta.clone().to(src.dtype)
    if theta.dim() == 2:
        theta = theta[None]
    src_size = tuple(src.shape)
    dst_size = src_size[:2] + tuple(int(s) for s in spatial_size)
    src_xform = normalize_transform(src_size[2:], align_corners).to(theta.dtype)
    dst_xform = normalize_transform(dst_size[2:], align_corners).to(theta.dtype)
    theta = src_xform @ theta @ torch.inverse(dst_xform)
    rev_idx = torch.as_tensor(range(sr - 1, -1, -1))
    theta[:, :sr] = theta[:, rev_idx]
    theta[:, :, :sr] = theta[:, :, rev_idx]
    grid = torch.nn.functional.affine_grid(theta=theta[:, :sr], size=list(dst_size), align_corners=align_corners)
    return torch.nn.functional.grid_sample(input=src.contiguous(), grid=grid, mode=mode, padding_mode=padding_mode,
                                           align_corners=align_corners)


# ------------------------------------------------------------------------------- forward trace (what the loader records)
def orientation(data: np.ndarray, affine: np.ndarray, axcodes: Sequence[str]):
    """monai.transforms.Orientation.__call__ on a channel-first array: returns (data, old_affine, new_affine)."""
    sr = data.ndim - 1
    src = io_orientation(affine)
    dst = axcodes2ornt(axcodes[:sr])
    spatial_ornt = ornt_transform(src, dst)
    ornt = spatial_ornt.copy()
    ornt[:, 0] += 1
    ornt = np.concatenate([np.array([[0, 1]]), ornt])
    shape = data.shape[1:]
    out = np.ascontiguousarray(apply_orientation(data, ornt))
    new_affine = affine @ inv_ornt_aff(spatial_ornt, shape)
    return out, affine, new_affine


def spacing(data: np.ndarray, affine: np.ndarray, pixdim, mode: str = "bilinear", padding_mode: str = "border",
            output_spatial_shape=None):
    """monai.transforms.Spacing.__call__(diagonal=False, align_corners=False, dtype=float64): (data f32, affine, new_affine)."""
    sr = data.ndim - 1
    out_d = np.asarray(pixdim, dtype=float)[:sr]
    new_affine = zoom_affine(affine, out_d)
    output_shape, offset = compute_shape_offset(data.shape[1:], affine, new_affine)
    new_affine[:sr, -1] = offset[:sr]
    transform = np.linalg.inv(affine) @ new_affine
    if np.allclose(transform, np.diag(np.ones(len(transform))), atol=1e-3):
        return data.copy().astype(np.float32), affine, new_affine
    out = affine_transform(torch.as_tensor(np.ascontiguousarray(data).astype(np.float64)).unsqueeze(0),
                           torch.as_tensor(np.ascontiguousarray(transform).astype(np.float64)),
                           spatial_size=output_shape if output_spatial_shape is None else output_spatial_shape,
                           mode=mode, padding_mode=padding_mode)
    return np.asarray(out.squeeze(0).numpy(), dtype=np.float32), affine, new_affine


def generate_spatial_bounding_box(img: np.ndarray):
    """monai.transforms.utils.generate_spatial_bounding_box(select_fn = x > 0, margin 0)."""
    data = np.any(img > 0, axis=0)
    ndim = data.ndim
    box_start, box_end = [0] * ndim, [0] * ndim
    for di, ax in enumerate(itertools.combinations(reversed(range(ndim)), ndim - 1)):
        dt = data.any(axis=ax)
        if not np.any(dt):
            return [0] * ndim, [0] * ndim
        min_d = max(int(np.argmax(dt)), 0)
        max_d = max(data.shape[di] - max(int(np.argmax(dt[::-1])), 0), min_d + 1)
        box_start[di], box_end[di] = min_d, max_d
    return box_start, box_end


def forward_trace(image: np.ndarray, affine: np.ndarray, pixdim, axcodes: str = "RAS") -> Dict:
    """Runs the geometric part of utils/data_utils.py:103-116 on `image` [1, X, Y, Z] (values in [0, 1] stand for the
    intensity-scaled scan: CropForegroundd selects > 0) and returns the array the network would see plus the entries
    MONAI pushes on `image_transforms` / keeps in `image_meta_dict`."""
    a0 = np.asarray(affine, dtype=np.float64)
    d1, old1, a1 = orientation(image, a0, axcodes)
    d2, old2, a2 = spacing(d1, a1, pixdim)
    box_start, box_end = generate_spatial_bounding_box(d2)
    sl = tuple(slice(s, e) for s, e in zip(box_start, box_end))
    d3 = d2[(slice(None),) + sl]
    return {"image": d3, "affine": a2,
            "orientation": {"orig_size": image.shape[1:], "old_affine": old1},
            "spacing": {"orig_size": d1.shape[1:], "old_affine": old2, "mode": "bilinear", "padding_mode": "border"},
            "crop": {"orig_size": d2.shape[1:], "box_start": np.asarray(box_start), "box_end": np.asarray(box_end)}}


# ---------------------------------------------------------------------------------------------------------- the inverse
def invertd(pred: np.ndarray, trace: Dict, nearest_interp: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """Compose.inverse over the trace, last transform first, on `pred` [C, x, y, z] (float32); returns the array on the
    grid of the file and the affine MONAI leaves in the meta dict."""
    # CropForegroundd.inverse: crop what a margin had padded, then BorderPad (constant 0) back to the pre-crop size
    crop = trace["crop"]
    orig_size = np.asarray(crop["orig_size"])
    cur_size = np.asarray(pred.shape[1:])
    box_start, box_end = np.asarray(crop["box_start"]), np.asarray(crop["box_end"])
    roi_start = np.maximum(-box_start, 0)
    roi_end = cur_size - np.maximum(box_end - orig_size, 0)
    d = pred[(slice(None),) + tuple(slice(int(s), int(e)) for s, e in zip(roi_start, roi_end))]
    pad_to_start = np.maximum(box_start, 0)
    pad_to_end = orig_size - np.minimum(box_end, orig_size)
    d = np.pad(d, [(0, 0)] + [(int(s), int(e)) for s, e in zip(pad_to_start, pad_to_end)], mode="constant")
    # Spacingd.inverse: Spacing(orig_pixdim) from the current affine, forced to the recorded size
    sp = trace["spacing"]
    old_affine = np.asarray(sp["old_affine"])
    orig_pixdim = np.sqrt(np.sum(np.square(old_affine), 0))[:-1]
    mode = "nearest" if nearest_interp else sp["mode"]
    d, _, affine = spacing(d, np.asarray(trace["affine"]), orig_pixdim, mode=mode, padding_mode=sp["padding_mode"],
                           output_spatial_shape=sp["orig_size"])
    # Orientationd.inverse: Orientation(axcodes of the file's affine)
    orig_axcodes = aff2axcodes(np.asarray(trace["orientation"]["old_affine"]))
    d, _, affine = orientation(d, affine, orig_axcodes)
    return np.ascontiguousarray(d), affine


def make_case(shape=(40, 36, 22), spacing_mm=(0.8, 0.75, 3.0), axcodes: str = "LAS", pixdim=(1.5, 1.5, 2.0), seed: int = 0,
              oblique: float = 0.0):
    """A synthetic scan: a soft blob with an empty rim (so CropForegroundd crops on every side), voxel axes in `axcodes`
    order with `spacing_mm`, an off-centre origin and (optionally) a small oblique rotation."""
    rng = np.random.default_rng(seed)
    grids = np.meshgrid(*[np.linspace(-1, 1, n) for n in shape], indexing="ij")
    r = sum(((g - c) / w) ** 2 for g, c, w in zip(grids, (0.1, -0.05, 0.0), (0.7, 0.75, 0.8)))
    img = np.clip(1.0 - r, 0.0, 1.0) * (0.5 + 0.5 * rng.random(shape))
    ornt = axcodes2ornt(tuple(axcodes))
    aff = np.zeros((4, 4))
    for in_ax, (out_ax, sign) in enumerate(ornt):
        aff[int(out_ax), in_ax] = sign * spacing_mm[in_ax]
    aff[:3, 3] = (-37.5, 12.25, -80.0)
    aff[3, 3] = 1.0
    if oblique:
        c, s = np.cos(oblique), np.sin(oblique)
        rot = np.array([[c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
        aff = rot @ aff
    return img[None].astype(np.float32), aff
