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


def _pp_cases():
    """Three "cases" (prediction, label, voxel volume): the label plus spurious blobs / holes, so that the largest-component
    rule improves some classes and not others."""
    cases = []
    for seed, shape, vpv in ((11, (20, 22, 18), 1.0), (12, (18, 18, 24), 2.0), (13, (22, 16, 20), 0.5)):
        lab = np.zeros(shape, dtype=np.int64)
        rng = np.random.default_rng(seed)
        for c in range(1, 14):                      # one box per class
            lo = rng.integers(0, np.array(shape) - 6)
            lab[lo[0]:lo[0] + 5, lo[1]:lo[1] + 5, lo[2]:lo[2] + 5] = c
        inf = lab.copy()
        noise = PO.blob_volume(shape, n_classes=13, seed=seed + 100, density=0.62)
        m = (inf == 0) & (noise > 0) & (rng.random(shape) < 0.35)
        inf[m] = noise[m]
        cases.append((inf, lab, vpv))
    return [c[0] for c in cases], [c[1] for c in cases], [c[2] for c in cases]


class _SyncPool:
    """multiprocessing.pool.Pool stand-in for the reference's determine_postprocessing (test_CTUNet_final.py:197)."""
    def __init__(self, *_a, **_k):
        pass

    def starmap_async(self, fn, iterable):
        res = [fn(*args) for args in iterable]

        class R:
            def get(self_inner):
                return res
        return R()

    def close(self):
        pass

    def join(self):
        pass


@pytest.mark.skipif(not ref_exec.available(), reason="/root/reference only exists in the build container")
@pytest.mark.parametrize("advanced", [False, True])
def test_determine_postprocessing_equals_the_reference(advanced):
    """hybrid_ctunet_b200.postprocess.determine_postprocessing (control flow restated, component filter injected = the
    host oracle) against the reference's function executed unmodified with a synchronous pool."""
    import contextlib
    import io
    from copy import deepcopy
    from scipy.ndimage import label
    from hybrid_ctunet_b200 import postprocess as P
    ns = dict(deepcopy=deepcopy, label=label, Pool=_SyncPool, dice=P.dice)
    fns = ref_exec.extract("test_CTUNet_final.py", ["remove_all_but_the_largest_connected_component", "determine_postprocessing",
                                                    "com_dice"], extra_globals=ns)
    infers, labels, vpv = _pp_cases()
    with contextlib.redirect_stdout(io.StringIO()):
        ref = fns["determine_postprocessing"](infers, labels, vpv, 0.0, 2, advanced)
    ours = P.determine_postprocessing(infers, labels, vpv, 0.0, 2, advanced,
                                      _remove=PO.remove_all_but_the_largest_connected_component)
    assert len(ref) == len(ours) and any((r != i).any() for r, i in zip(ref, infers))
    for r, o in zip(ref, ours):
        assert np.array_equal(r, o)
