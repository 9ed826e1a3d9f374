"""The post-processing oracle (oracle/postprocess_oracle.py) against the reference's own function, executed unmodified
(AST-extracted from /root/reference/test_CTUNet_final.py:132-190; build container only), and against the committed
fixtures the GPU test uses."""
import os

import numpy as np
import pytest

from oracle import postprocess_oracle as PO
from oracle import ref_exec

GOLD = os.path.join(os.path.dirname(__file__), "golden", "postprocess_ref.npz")
CASES = [  # (shape, classes argument, volume_per_voxel, minimum sizes)
    ((24, 20, 28), [1, 2, 3, 4, 5], 1.5, None),
    ((24, 20, 28), [(1, 2, 3, 4, 5)], 2.25, None),                       # all foreground classes as one region
    ((17, 31, 13), [(1, 2), 3, 5], 0.75, {(1, 2): 30.0, 3: 12.0, 5: 1e9}),   # size thresholds, a group first
    ((16, 16, 16), None, 1.0, None),                                      # classes taken from the volume
]


def _same(a, b):
    assert a[0].dtype == b[0].dtype and np.array_equal(a[0], b[0])
    assert a[1] == b[1] and a[2] == b[2]


@pytest.mark.skipif(not ref_exec.available(), reason="/root/reference only exists in the build container")
def test_oracle_equals_the_reference_function():
    from copy import deepcopy
    from scipy.ndimage import label
    ref = ref_exec.extract("test_CTUNet_final.py", ["remove_all_but_the_largest_connected_component"],
                           extra_globals=dict(deepcopy=deepcopy, label=label))["remove_all_but_the_largest_connected_component"]
    for i, (shape, classes, vpv, mins) in enumerate(CASES):
        img = PO.blob_volume(shape, seed=i)
        _same(PO.remove_all_but_the_largest_connected_component(img, classes, vpv, mins), ref(img, classes, vpv, mins))


def test_oracle_reproduces_the_committed_reference_outputs():
    g = np.load(GOLD, allow_pickle=True)
    for i, (shape, classes, vpv, mins) in enumerate(CASES):
        img = g[f"in{i}"]
        assert np.array_equal(img, PO.blob_volume(shape, seed=i))
        out, removed, kept = PO.remove_all_but_the_largest_connected_component(img, classes, vpv, mins)
        assert np.array_equal(out, g[f"out{i}"])
        assert removed == g[f"removed{i}"].item() and kept == g[f"kept{i}"].item()


def test_every_largest_object_is_kept_and_nothing_else_changes():
    img = np.zeros((6, 6, 6), dtype=np.int64)
    img[0, 0, :3] = 1          # two objects of the same (largest) size: both stay (test_CTUNet_final.py:179)
    img[5, 5, 3:] = 1
    img[3, 3, 3] = 1           # a single voxel: removed
    img[2, :, 0] = 2           # another class: untouched when only class 1 is processed
    out, removed, kept = PO.remove_all_but_the_largest_connected_component(img, [1], 2.0)
    assert out[3, 3, 3] == 0 and (out[0, 0, :3] == 1).all() and (out[5, 5, 3:] == 1).all() and (out[2, :, 0] == 2).all()
    assert removed == {1: 2.0} and kept == {1: 6.0}
