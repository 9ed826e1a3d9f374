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

__all__ = ["remove_all_but_the_largest_connected_component"]


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
