"""Connected-component post-processing on the device (ctu_cc_filter_largest through
hybrid_ctunet_b200.postprocess) against the committed outputs of the reference's own function
(test_CTUNet_final.py:132-190) and against the scipy oracle on larger volumes — integer work: bit-exact."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "postprocess_ref.npz")


def _same(ours, ref):
    assert ours[0].dtype == ref[0].dtype and np.array_equal(ours[0], ref[0])
    assert ours[1] == ref[1] and ours[2] == ref[2]


def test_matches_the_reference_outputs():
    from hybrid_ctunet_b200.postprocess import remove_all_but_the_largest_connected_component as ours
    from test_postprocess_cpu import CASES
    g = np.load(GOLD, allow_pickle=True)
    for i, (shape, classes, vpv, mins) in enumerate(CASES):
        out, removed, kept = ours(g[f"in{i}"], classes, vpv, mins)
        assert out.dtype == g[f"out{i}"].dtype and np.array_equal(out, g[f"out{i}"])
        assert removed == g[f"removed{i}"].item() and kept == g[f"kept{i}"].item()


@pytest.mark.parametrize("shape,classes,mins", [((96, 80, 64), list(range(1, 14)), None),
                                                ((64, 64, 96), [tuple(range(1, 14))], None),
                                                ((50, 70, 33), [(1, 2, 3), 4, 5, 6], {(1, 2, 3): 200.0, 4: 50.0, 5: 5.0, 6: 1e12})])
def test_matches_the_scipy_oracle(shape, classes, mins):
    from hybrid_ctunet_b200.postprocess import remove_all_but_the_largest_connected_component as ours
    from oracle import postprocess_oracle as PO
    img = PO.blob_volume(shape, n_classes=13, seed=sum(shape), density=0.52)
    _same(ours(img, classes, 0.8, mins), PO.remove_all_but_the_largest_connected_component(img, classes, 0.8, mins))


def test_full_volume_properties():
    """512 x 512 x 256 (BASELINE geometry; too slow for the host oracle): idempotent, only removes, keeps exactly one largest
    foreground object, the kept size is the number of surviving voxels, and a CUDA tensor stays on the device."""
    from hybrid_ctunet_b200.postprocess import remove_all_but_the_largest_connected_component as ours
    g = torch.Generator(device="cuda").manual_seed(0)
    noise = torch.rand(1, 1, 512, 512, 256, device="cuda", generator=g)
    smooth = torch.nn.functional.avg_pool3d(noise, 3, stride=1, padding=1)[0, 0]
    img = (smooth > 0.5).to(torch.int64) * (1 + (torch.arange(256, device="cuda") // 64))[None, None, :]   # classes 1..4 by z slab
    fg = tuple(range(1, 5))
    out, removed, kept = ours(img, [fg], 1.0)
    assert out.is_cuda and out.dtype == img.dtype and out.shape == img.shape
    changed = out != img
    assert bool((out[changed] == 0).all()) and int(changed.sum()) > 0
    assert kept[fg] == float((out > 0).sum()) and removed[fg] is not None and removed[fg] < kept[fg]
    out2, removed2, kept2 = ours(out, [fg], 1.0)
    assert torch.equal(out2, out) and removed2[fg] is None and kept2[fg] == kept[fg]


@pytest.mark.parametrize("advanced", [False, True])
def test_determine_postprocessing_on_the_device_equals_the_host_flow(advanced):
    """determine_postprocessing with the CUDA component filter == the same flow with the scipy oracle injected (which
    tests/test_postprocess_cpu.py pins against the reference's own function, test_CTUNet_final.py:192-401)."""
    from hybrid_ctunet_b200 import postprocess as P
    from oracle import postprocess_oracle as PO
    from test_postprocess_cpu import _pp_cases
    infers, labels, vpv = _pp_cases()
    ours = P.determine_postprocessing(infers, labels, vpv, 0.0, 8, advanced)
    host = P.determine_postprocessing(infers, labels, vpv, 0.0, 8, advanced,
                                      _remove=PO.remove_all_but_the_largest_connected_component)
    assert any((o != i).any() for o, i in zip(ours, infers))
    for o, h in zip(ours, host):
        assert o.dtype == h.dtype and np.array_equal(o, h)
