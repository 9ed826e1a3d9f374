"""CPU restatement of the reference's connected-component post-processing (TEST INFRASTRUCTURE ONLY).

Follows test_CTUNet_final.py:132-190 (`remove_all_but_the_largest_connected_component`) statement by statement with
scipy.ndimage.label (the reference's own dependency, default structure = 6-connectivity); pinned against the reference's
function executed unmodified (AST-extracted, tests/test_postprocess_cpu.py)."""
from __future__ import annotations

from copy import deepcopy

import numpy as np
from scipy.ndimage import label


def remove_all_but_the_largest_connected_component(image_in: np.ndarray, for_which_classes, volume_per_voxel: float,
                                                   minimum_valid_object_size: dict = None):
    image = deepcopy(image_in)                                                    # :144
    if for_which_classes is None:                                                 # :145-147
        for_which_classes = np.unique(image)
        for_which_classes = for_which_classes[for_which_classes > 0]
    assert 0 not in for_which_classes, "cannot remove background"                 # :149
    largest_removed, kept_size = {}, {}
    for c in for_which_classes:                                                   # :152
        if isinstance(c, (list, tuple)):                                          # :153-157
            c = tuple(c)
            mask = np.zeros_like(image, dtype=bool)
            for cl in c:
                mask[image == cl] = True
        else:
            mask = image == c                                                     # :159
        lmap, num_objects = label(mask.astype(int))                               # :161
        object_sizes = {i: (lmap == i).sum() * volume_per_voxel for i in range(1, num_objects + 1)}   # :164-166
        largest_removed[c] = None
        kept_size[c] = None
        if num_objects > 0:                                                       # :171
            maximum_size = max(object_sizes.values())
            kept_size[c] = maximum_size
            for object_id in range(1, num_objects + 1):
                if object_sizes[object_id] != maximum_size:                       # :179
                    remove = True
                    if minimum_valid_object_size is not None:
                        remove = object_sizes[object_id] < minimum_valid_object_size[c]   # :183
                    if remove:
                        image[(lmap == object_id) & mask] = 0                     # :185
                        if largest_removed[c] is None:
                            largest_removed[c] = object_sizes[object_id]
                        else:
                            largest_removed[c] = max(largest_removed[c], object_sizes[object_id])
    return image, largest_removed, kept_size


def blob_volume(shape, n_classes: int = 5, seed: int = 0, density: float = 0.55) -> np.ndarray:
    """Synthetic label volume with many small and a few large components per class (smoothed noise, thresholded)."""
    from scipy.ndimage import uniform_filter
    rng = np.random.default_rng(seed)
    out = np.zeros(shape, dtype=np.int64)
    for c in range(1, n_classes + 1):
        f = uniform_filter(rng.random(shape), size=3)
        out[(f > density) & (out == 0)] = c
    return out
