"""Host geometry of the device `Invertd` (hybrid_ctunet_b200/invert.py) against the stepwise MONAI restatement
(oracle/invert_oracle.py): the composite 3x4 index map, applied by a plain numpy gather with the kernel's semantics, must
reproduce pad -> resample (torch grid_sample, float64) -> flip / transpose.  Floating point: the two sides compute the
source coordinates along different routes (voxel-space matrix vs torch's normalised grid), both in float64, and round to
float32 — tolerance 2e-6 relative to the value range (a float32 ulp or two at coordinates that differ by ~1e-13 voxels)."""
import numpy as np
import pytest

from hybrid_ctunet_b200.invert import InvertGeometry
from oracle import invert_oracle as IO

CASES = [
    dict(axcodes="LAS"),                                                    # the usual abdominal CT file
    dict(axcodes="RAS", shape=(33, 40, 19), spacing_mm=(0.9, 0.7, 2.5)),
    dict(axcodes="PIL", oblique=0.03),                                      # permuted axes + a slightly oblique affine
    dict(axcodes="ASR", shape=(21, 30, 34), spacing_mm=(3.0, 0.8, 0.8)),
    dict(axcodes="LPS", pixdim=(1.0, 1.0, 1.0), spacing_mm=(1.0, 1.0, 1.0)),  # Spacing is the identity: MONAI copies
    dict(axcodes="RPI", pixdim=(2.0, 0.6, 1.3)),
]


def apply_geometry(pred: np.ndarray, g: InvertGeometry, mode: int) -> np.ndarray:
    """numpy statement of csrc/invert.cu: clamp to the padded grid, 8 corners in grid_sample's order, corners outside the
    crop (or the padded grid) add nothing, float64 accumulation, float32 result."""
    r, n = g.roi_start, g.pred_size
    pred = pred[:, r[0]:r[0] + n[0], r[1]:r[1] + n[1], r[2]:r[2] + n[2]]
    idx = np.stack(np.meshgrid(*[np.arange(s, dtype=np.float64) for s in g.out_size], indexing="ij"), axis=0)
    c = np.tensordot(g.m[:, :3], idx, axes=1) + g.m[:, 3].reshape(3, 1, 1, 1)
    for a in range(3):
        c[a] = np.clip(c[a], 0.0, g.pad_size[a] - 1.0)
    out = np.zeros((pred.shape[0],) + tuple(g.out_size), dtype=np.float64)

    def corner(i, w):
        ok = np.ones(i[0].shape, dtype=bool)
        p = []
        for a in range(3):
            ok &= (i[a] >= 0) & (i[a] < g.pad_size[a])
            pa = i[a] - g.crop_start[a]
            ok &= (pa >= 0) & (pa < g.pred_size[a])
            p.append(np.clip(pa, 0, g.pred_size[a] - 1))
        v = pred[:, p[0], p[1], p[2]].astype(np.float64)
        return np.where(ok, v * w, 0.0)

    if mode == 0:
        i = [np.rint(c[a]).astype(np.int64) for a in range(3)]
        return corner(i, 1.0).astype(np.float32)
    fl = np.floor(c)
    f1, f0 = c - fl, (fl + 1.0) - c
    i0 = fl.astype(np.int64)
    for k in range(8):
        kx, ky, kz = (k >> 2) & 1, (k >> 1) & 1, k & 1
        w = ((f1[2] if kz else f0[2]) * (f1[1] if ky else f0[1])) * (f1[0] if kx else f0[0])
        out = out + corner([i0[0] + kx, i0[1] + ky, i0[2] + kz], w)
    return out.astype(np.float32)


def ties(g: InvertGeometry) -> np.ndarray:
    idx = np.stack(np.meshgrid(*[np.arange(s, dtype=np.float64) for s in g.out_size], indexing="ij"), axis=0)
    c = np.tensordot(g.m[:, :3], idx, axes=1) + g.m[:, 3].reshape(3, 1, 1, 1)
    return (np.abs(np.abs(c - np.floor(c)) - 0.5) < 1e-9).any(axis=0)


def _case(kw, seed=0, channels=3):
    kw = dict(kw)
    pixdim = kw.pop("pixdim", (1.5, 1.5, 2.0))
    img, aff = IO.make_case(seed=seed, **kw)
    trace = IO.forward_trace(img, aff, pixdim)
    rng = np.random.default_rng(seed + 1)
    pred = rng.standard_normal((channels,) + trace["image"].shape[1:]).astype(np.float32) * 3.0
    return img, aff, pixdim, trace, pred


def _geom(trace):
    o, s, c = trace["orientation"], trace["spacing"], trace["crop"]
    return InvertGeometry.from_parts(o["old_affine"], o["orig_size"], s["old_affine"], s["orig_size"], trace["affine"],
                                     c["orig_size"], c["box_start"], c["box_end"])


@pytest.mark.parametrize("kw", CASES, ids=[c["axcodes"] for c in CASES])
@pytest.mark.parametrize("nearest", [False, True], ids=["trilinear", "nearest"])
def test_composite_map_matches_stepwise_inverse(kw, nearest):
    img, aff, pixdim, trace, pred = _case(kw)
    ref, ref_affine = IO.invertd(pred, trace, nearest_interp=nearest)
    g = _geom(trace)
    assert tuple(g.out_size) == img.shape[1:] == ref.shape[1:]
    got = apply_geometry(pred, g, 0 if nearest else 1)
    if nearest:
        # a sample exactly half-way between two voxels (3.0 mm -> 2.0 mm puts every other z sample there) rounds by the last
        # bit of whichever route computed the coordinate: compare everywhere else
        clear = ~ties(g)
        assert clear.mean() > 0.2
        assert np.array_equal(got[:, clear], ref[:, clear])
    else:
        assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()
    assert np.allclose(g.affine, ref_affine, atol=1e-9)
    assert np.allclose(g.affine, aff, atol=1e-6)      # back on the file's grid


@pytest.mark.parametrize("kw", CASES, ids=[c["axcodes"] for c in CASES])
def test_from_file_reproduces_the_forward_metadata(kw):
    img, aff, pixdim, trace, _ = _case(kw, seed=3)
    c = trace["crop"]
    a = InvertGeometry.from_file(aff, img.shape[1:], pixdim, c["box_start"], c["box_end"])
    b = _geom(trace)
    assert a.out_size == b.out_size and a.pad_size == b.pad_size and a.crop_start == b.crop_start and a.pred_size == b.pred_size
    assert np.allclose(a.m, b.m, atol=1e-12)


def test_from_trace_reads_monai_style_entries_and_margins():
    img, aff, pixdim, trace, pred = _case(dict(axcodes="LAS"), seed=5)
    o, s, c = trace["orientation"], trace["spacing"], trace["crop"]
    entries = [
        {"class": "Orientationd", "orig_size": o["orig_size"], "extra_info": {"old_affine": o["old_affine"]}},
        {"class": "Spacingd", "orig_size": s["orig_size"],
         "extra_info": {"old_affine": s["old_affine"], "mode": "bilinear", "padding_mode": "border", "align_corners": "none"}},
        {"class": "CropForegroundd", "orig_size": c["orig_size"], "extra_info": {"box_start": c["box_start"], "box_end": c["box_end"]}},
    ]
    g = InvertGeometry.from_trace(entries, trace["affine"])
    assert np.array_equal(g.m, _geom(trace).m)
    # a margin that reaches outside the image: the forward transform pads, the inverse trims that rim first
    margin = 6
    t2 = {k: (dict(v) if isinstance(v, dict) else v) for k, v in trace.items()}
    t2["crop"]["box_start"] = np.asarray(c["box_start"]) - margin
    t2["crop"]["box_end"] = np.asarray(c["box_end"]) + margin
    size = tuple(int(e - b) for b, e in zip(t2["crop"]["box_start"], t2["crop"]["box_end"]))
    big = np.random.default_rng(9).standard_normal((2,) + size).astype(np.float32)
    ref, _ = IO.invertd(big, t2)
    g2 = _geom(t2)
    assert any(r > 0 for r in g2.roi_start)
    got = apply_geometry(big, g2, 1)
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()
    entries[1]["extra_info"]["padding_mode"] = "zeros"
    with pytest.raises(NotImplementedError):
        InvertGeometry.from_trace(entries, trace["affine"])


def test_inconsistent_trace_is_rejected():
    img, aff, pixdim, trace, _ = _case(dict(axcodes="LAS"))
    o, s, c = trace["orientation"], trace["spacing"], trace["crop"]
    with pytest.raises(ValueError):
        InvertGeometry.from_parts(o["old_affine"], (7, 7, 7), s["old_affine"], s["orig_size"], trace["affine"], c["orig_size"],
                                  c["box_start"], c["box_end"])


def test_random_orientations_spacings_and_sizes():
    """24 seeded random cases: every axis permutation / sign, voxel sizes 0.5-4 mm, target spacings 0.7-3 mm, small oblique
    rotations, odd sizes — the composite map must keep reproducing the stepwise inverse."""
    import itertools
    rng = np.random.default_rng(2024)
    perms = list(itertools.permutations("xyz"))
    letters = {"x": "LR", "y": "PA", "z": "IS"}
    for n in range(24):
        axcodes = "".join(letters[a][int(rng.integers(0, 2))] for a in perms[int(rng.integers(0, 6))])
        shape = tuple(int(v) for v in rng.integers(14, 40, 3))
        spacing_mm = tuple(float(v) for v in np.round(rng.uniform(0.5, 4.0, 3), 2))
        pixdim = tuple(float(v) for v in np.round(rng.uniform(0.7, 3.0, 3), 2))
        oblique = float(rng.uniform(-0.05, 0.05)) if n % 3 == 0 else 0.0
        img, aff = IO.make_case(shape=shape, spacing_mm=spacing_mm, axcodes=axcodes, seed=n, oblique=oblique)
        trace = IO.forward_trace(img, aff, pixdim)
        if min(trace["image"].shape[1:]) < 2:
            continue
        pred = rng.standard_normal((2,) + trace["image"].shape[1:]).astype(np.float32)
        ref, ref_affine = IO.invertd(pred, trace)
        g = _geom(trace)
        got = apply_geometry(pred, g, 1)
        assert got.shape == ref.shape, (n, axcodes, shape)
        assert np.abs(got - ref).max() <= 2e-6 * max(np.abs(ref).max(), 1.0), (n, axcodes, shape, spacing_mm, pixdim)
        assert np.allclose(g.affine, ref_affine, atol=1e-9)
        c = trace["crop"]
        f = InvertGeometry.from_file(aff, img.shape[1:], pixdim, c["box_start"], c["box_end"])
        assert f.out_size == g.out_size and f.pad_size == g.pad_size and np.allclose(f.m, g.m, atol=1e-12), (n, axcodes)
