"""Connected-component post-processing of the predicted label volumes on the device.

Drop-in for `remove_all_but_the_largest_connected_component` of the reference's evaluation script
(test_CTUNet_final.py:132-190): same arguments, same return triple `(image, largest_removed, kept_size)`, same class /
class-group semantics (entries of `for_which_classes` are processed in order, each on the volume the previous ones left),
same float64 size arithmetic (`voxel count * volume_per_voxel`).  The reference labels every mask with scipy.ndimage.label
and sizes each object with a full-volume pass on the host; here one call of ctu_cc_filter_largest per class does the
labelling (lock-free union-find), the size histogram and the removal on the GPU, and three integers come back.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import lib as _lib

__all__ = ["remove_all_but_the_largest_connected_component", "determine_postprocessing", "com_dice", "dice"]


def _key(c):
    return tuple(c) if isinstance(c, (list, tuple)) else c


def remove_all_but_the_largest_connected_component(image_in, for_which_classes, volume_per_voxel: float,
                                                   minimum_valid_object_size: Optional[dict] = None):
    """image_in: integer label volume [X, Y, Z] (numpy array or torch tensor, labels 0..255).  Returns the filtered volume
    (same kind and dtype as the input; a CUDA tensor stays on the device), `largest_removed` and `kept_size` keyed by
    class (tuples for class groups), values float or None — as test_CTUNet_final.py:132-190."""
    lib = _lib.require_device()
    is_np = isinstance(image_in, np.ndarray)
    src = torch.from_numpy(np.ascontiguousarray(image_in)) if is_np else image_in
    if src.dim() != 3:
        raise ValueError("expected a [X, Y, Z] label volume")
    dev = src.device if src.is_cuda else torch.device("cuda", torch.cuda.current_device())
    img = src.to(device=dev, dtype=torch.uint8, copy=True).contiguous()
    X, Y, Z = (int(v) for v in img.shape)
    if for_which_classes is None:
        present = torch.unique(img).tolist()
        for_which_classes = [c for c in present if c > 0]
    classes_flat = [cl for c in for_which_classes for cl in (c if isinstance(c, (list, tuple)) else (c,))]
    assert 0 not in classes_flat, "cannot remove background"
    parent = torch.empty(img.numel(), dtype=torch.int32, device=dev)
    sizes = torch.empty(img.numel(), dtype=torch.int32, device=dev)
    summary = torch.empty(4, dtype=torch.int32, device=dev)
    member = torch.empty(256, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    largest_removed: Dict = {}
    kept_size: Dict = {}
    vpv = float(volume_per_voxel)
    for c in for_which_classes:
        key = _key(c)
        m = np.zeros(256, dtype=np.uint8)
        for cl in (key if isinstance(key, tuple) else (key,)):
            if not 0 <= int(cl) < 256:
                raise ValueError("labels must be in 0..255")
            m[int(cl)] = 1
        member.copy_(torch.from_numpy(m))
        has_min = minimum_valid_object_size is not None
        min_valid = float(minimum_valid_object_size[key]) if has_min else 0.0
        _lib.check(lib.ctu_cc_filter_largest(img.data_ptr(), member.data_ptr(), X, Y, Z, vpv, int(has_min), min_valid,
                                             parent.data_ptr(), sizes.data_ptr(), summary.data_ptr(), stream),
                   "ctu_cc_filter_largest")
        n_obj, mx, rem, _ = summary.tolist()
        largest_removed[key] = None
        kept_size[key] = None
        if n_obj > 0:
            kept_size[key] = np.int64(mx) * vpv
            if rem > 0:
                largest_removed[key] = np.int64(rem) * vpv
    out = img.to(src.dtype)
    if is_np:
        return out.cpu().numpy().astype(image_in.dtype, copy=False), largest_removed, kept_size
    return (out if src.is_cuda else out.cpu()), largest_removed, kept_size


def dice(x, y) -> float:
    """utils/utils.py:16-22 on boolean masks (numpy arrays or torch tensors): 2 |x y| / (|x| + |y|), 0.0 for an empty
    label mask.  Integer counts, the reference's float64 division."""
    if isinstance(x, torch.Tensor) or isinstance(y, torch.Tensor):
        x = torch.as_tensor(x)
        y = torch.as_tensor(y, device=x.device)
        inter, xs, ys = int((x & y).sum()), int(x.sum()), int(y.sum())
    else:
        inter, xs, ys = int(np.sum(x * y)), int(np.sum(x)), int(np.sum(y))
    if ys == 0:
        return 0.0
    return 2 * inter / (xs + ys)


def com_dice(infers, labels, classes: Sequence[int] = tuple(range(1, 14))) -> np.ndarray:
    """test_CTUNet_final.py:106-117: per-class Dice averaged over the cases."""
    rows = [[dice(infers[i] == j, labels[i] == j) for j in classes] for i in range(len(labels))]
    return np.mean(rows, 0)


def determine_postprocessing(infers, labels, volume_per_voxel, dice_threshold: float = 0.0, processes: int = 8,
                             advanced_postprocessing: bool = False, _remove=None):
    """test_CTUNet_final.py:192-401: decide, on a set of cases, for which classes "keep only the largest connected
    component" improves the mean Dice (first all foreground classes as one region, then every class on its own, optionally
    with size thresholds learnt from the cases), then apply that rule to every case.  Returns the list of post-processed
    volumes like the reference (`processes` is accepted and ignored: the per-case work runs on the GPU, not in a host pool).
    `_remove`: the component filter to use (tests inject the host oracle)."""
    remove = _remove or remove_all_but_the_largest_connected_component
    classes = list(range(1, 14))
    n = len(labels)
    for_which_classes: list = []
    min_valid_object_sizes: Optional[dict] = {}

    def learn_min_sizes(sources, which):
        kept_min: dict = {}
        for i in range(n):
            _, _, kept = remove(sources[i], which, volume_per_voxel[i], None)
            for k, v in kept.items():
                if v is not None:
                    kept_min[k] = v if kept_min.get(k) is None else min(kept_min[k], v)
        return kept_min

    # 1. all foreground classes as one region (:208-281)
    min_size_kept = learn_min_sizes(infers, (classes,)) if advanced_postprocessing else None
    infers_pp = [remove(infers[i], (classes,), volume_per_voxel[i], min_size_kept)[0] for i in range(n)]
    raw = com_dice(infers, labels)
    pp_all = com_dice(infers_pp, labels)
    do_fg_cc = False
    if any(pp_all[i] > raw[i] + dice_threshold for i in range(len(classes))):
        if not any(pp_all[i] < raw[i] for i in range(len(classes))):
            for_which_classes.append(classes)
            if min_size_kept is not None:
                min_valid_object_sizes.update(dict(min_size_kept))
            do_fg_cc = True
    # 2. every class on its own, on top of step 1 if that was adopted (:296-371)
    source = infers_pp if do_fg_cc else infers
    min_size_kept = learn_min_sizes(source, classes) if advanced_postprocessing else None
    infers_pp_new = [remove(source[i], classes, volume_per_voxel[i], min_size_kept)[0] for i in range(n)]
    old_res = pp_all if do_fg_cc else raw
    pp_cls = com_dice(infers_pp_new, labels)
    for i, cl in enumerate(classes):
        if pp_cls[i] > old_res[i] + dice_threshold:
            for_which_classes.append(int(cl))
            if min_size_kept is not None:
                min_valid_object_sizes.update({cl: min_size_kept[cl]})
    if not advanced_postprocessing:
        min_valid_object_sizes = None
    # 3. apply the rule (:387-399)
    return [remove(infers[i], for_which_classes, volume_per_voxel[i], min_valid_object_sizes)[0] for i in range(n)]
