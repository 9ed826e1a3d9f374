"""Host logic of the sliding-window path and its oracle against the known-answer values of SURVEY 8c
(window starts, importance map) — bit-exact integer / fp32 work — plus the reference's error behaviour."""
import hashlib

import numpy as np
import pytest
import torch

from hybrid_ctunet_b200 import sliding_window as S
from oracle import sliding_window_oracle as O


def test_gaussian_kernel_known_answers():
    k = O.gaussian_1d(12.0)
    assert k.numel() == 97
    assert k[0].item() == 1.1205673217773438e-05 and k[96].item() == 1.1205673217773438e-05
    assert k[48].item() == 0.0332355722784996


def test_importance_map_known_answers():
    m = S.compute_importance_map((96, 96, 96), "gaussian", 0.125)
    assert m.dtype == torch.float32 and m.shape == (96, 96, 96)
    assert m.max().item() == 1.0 and m[48, 48, 48].item() == 1.0
    assert m[0, 0, 0].item() == 3.8326959661549864e-11 == m.min().item()
    assert m[95, 95, 95].item() == 1.0314445825221341e-10
    assert m.double().sum().item() == 27233.70981889281
    assert hashlib.sha256(m.numpy().tobytes()).hexdigest()[:16] == "370ebbcd849595b9"


@pytest.mark.parametrize("patch", [(96, 96, 96), (64, 48, 32), (5, 7, 9), (96, 96, 16), (3, 3, 3)])
def test_importance_map_product_form_equals_filter_chain(patch):
    a = S.compute_importance_map(patch, "gaussian", 0.125)
    b = O.compute_importance_map(patch, "gaussian", 0.125)   # zero-padded separable conv of an impulse (MONAI)
    assert torch.equal(a, b)
    assert torch.equal(S.compute_importance_map(patch, "constant"), torch.ones(patch))


def _starts(img, overlap):
    return S.dense_patch_starts(img, (96, 96, 96), S.get_scan_interval(img, (96, 96, 96), 3, overlap))


def test_window_starts_known_answers():
    st = _starts((512, 512, 256), 0.5)
    assert len(st) == 500 and tuple(st[0]) == (0, 0, 0) and tuple(st[-1]) == (416, 416, 160)
    assert sorted(set(st[:, 0])) == [0, 48, 96, 144, 192, 240, 288, 336, 384, 416]
    assert sorted(set(st[:, 2])) == [0, 48, 96, 144, 160]
    assert tuple(st[1]) == (0, 0, 48)  # C-order: z fastest
    assert len(_starts((512, 512, 256), 0.7)) == 1792
    st = _starts((200, 180, 150), 0.5)
    assert len(st) == 36
    assert sorted(set(st[:, 0])) == [0, 48, 96, 104] and sorted(set(st[:, 1])) == [0, 48, 84]
    assert sorted(set(st[:, 2])) == [0, 48, 54]
    assert len(_starts((96, 96, 96), 0.5)) == 1
    assert S.get_scan_interval((96, 96, 96), (96, 96, 96), 3, 0.5) == (96, 96, 96)
    assert S.get_scan_interval((512, 512, 256), (96, 96, 96), 3, 0.7) == (28, 28, 28)
    assert S.get_scan_interval((100, 100, 100), (2, 2, 2), 3, 0.9) == (1, 1, 1)


def test_starts_match_oracle_slices():
    for img, ov in [((130, 97, 200), 0.5), ((96, 300, 96), 0.25), ((512, 512, 256), 0.7)]:
        iv = S.get_scan_interval(img, (96, 96, 96), 3, ov)
        assert iv == O.get_scan_interval(img, (96, 96, 96), 3, ov)
        sl = O.dense_patch_slices(img, (96, 96, 96), iv)
        st = S.dense_patch_starts(img, (96, 96, 96), iv)
        assert [tuple(s.start for s in w) for w in sl] == [tuple(int(v) for v in r) for r in st]


def test_count_map_range_full_volume():
    """min 3.83e-11 at the corners, max 1.7063407897949219 => fp32 accumulators, safe divide."""
    m = S.compute_importance_map((96, 96, 96), "gaussian", 0.125)
    st = _starts((512, 512, 256), 0.5)
    # separable: the count map is the outer product structure only per axis sums -> evaluate along one line per axis
    cnt = torch.zeros(512, 512, 256)
    for s in st[:5]:  # first x/y start, all z starts: a z-line through the centre of the first window column
        cnt[s[0]:s[0] + 96, s[1]:s[1] + 96, s[2]:s[2] + 96] += m
    assert cnt[0, 0, 0].item() == 3.8326959661549864e-11
    assert cnt[48, 48].max().item() > 1.0


def test_error_behaviour():
    x = torch.zeros(1, 1, 8, 8, 8)
    with pytest.raises(AssertionError):
        S.sliding_window_inference(x, 4, 1, lambda t: t, overlap=1.0)
    with pytest.raises(AssertionError):
        O.sliding_window_inference(x, 4, 1, lambda t: t, overlap=-0.1)
    with pytest.raises(ValueError):
        S.get_scan_interval((8, 8), (4, 4, 4), 3, 0.5)
    with pytest.raises(ValueError):
        S.get_scan_interval((8, 8, 8), (4, 4), 3, 0.5)
    with pytest.raises(RuntimeError):  # no CPU fallback for the blend
        S.sliding_window_inference(x, 4, 1, lambda t: ((t,), (t,)), overlap=0.5)
    assert S.fall_back_tuple((96, -1, None), (10, 20, 30)) == (96, 20, 30)
    assert S.fall_back_tuple(96, (10, 20, 30)) == (96, 96, 96)
    assert S.get_valid_patch_size((50, 200, 96), (96, 96, 96)) == (50, 96, 96)


def test_oracle_blend_on_cpu_small():
    """The oracle itself: a constant predictor must come back unchanged through blend + normalise."""
    x = torch.rand(1, 1, 20, 17, 23)
    pred = lambda w: ((w.repeat(1, 3, 1, 1, 1),), (2 * w.repeat(1, 3, 1, 1, 1),))
    a, b = O.sliding_window_inference(x, (8, 8, 8), 4, pred, overlap=0.5, mode="gaussian")
    assert a.shape == (1, 3, 20, 17, 23)
    assert torch.allclose(a, x.repeat(1, 3, 1, 1, 1), atol=1e-6) and torch.allclose(b, 2 * a, atol=1e-6)
    c = O.sliding_window_inference(x, (32, 32, 32), 2, lambda w: (w,), overlap=0.25, two_heads=False)  # roi > image: pad+crop
    assert c.shape == x.shape and torch.allclose(c, x, atol=1e-6)


def test_shard_window_range_partitions_in_whole_calls():
    """Sharded sliding window: contiguous, disjoint chunks that cover the window list, every chunk but the last non-empty one a
    whole number of sw_batch calls; ranks beyond the work get an empty range."""
    from hybrid_ctunet_b200.sliding_window import shard_window_range
    for total, sw, world in ((500, 4, 8), (500, 4, 2), (500, 4, 4), (1792, 4, 8), (1, 4, 2), (7, 4, 8), (64, 4, 3), (10, 1, 4), (9, 4, 1)):
        ranges = [shard_window_range(total, sw, world, r) for r in range(world)]
        assert ranges[0][0] == 0 and max(hi for _, hi in ranges) == total
        for (lo, hi), (lo2, _) in zip(ranges, ranges[1:]):
            assert lo <= hi and hi == lo2
        nonempty = [(lo, hi) for lo, hi in ranges if hi > lo]
        assert all((hi - lo) % sw == 0 for lo, hi in nonempty[:-1])
        calls = [-(-(hi - lo) // sw) for lo, hi in ranges]
        assert max(calls) == -(-(-(-total // sw)) // world)          # the slowest rank runs ceil(calls / world) calls
    assert [shard_window_range(500, 4, 8, r) for r in (0, 6, 7)] == [(0, 64), (384, 448), (448, 500)]
