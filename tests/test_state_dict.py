"""Drop-in boundary: constructors, state_dict keys / order / shapes and same-seed default init of the
drop-in modules against the reference (when /root/reference is present) and against the pinned hash."""
import hashlib

import pytest
import torch

from oracle import ref_import

KW = dict(in_channels=1, dim_conv_stem=64, out_channels=14, img_size=(96, 96), frames=96, patch_frame=8)
SURVEY_HASH = "344d074b632bcf07"  # sha256 of "key:shape" lines of the reference CTUNet(101, pf8), SURVEY 8b


def _hash(sd):
    return hashlib.sha256("\n".join(f"{k}:{tuple(v.shape)}" for k, v in sd.items()).encode()).hexdigest()[:16]


def test_ctunet_state_dict_hash_and_count():
    from hybrid_ctunet_b200.networks.hybrid_CTUNet import CTUNet
    m = CTUNet(model_depth=101, **KW)
    sd = m.state_dict()
    assert len(sd) == 412
    assert sum(p.numel() for p in m.parameters()) == 174_801_766
    assert _hash(sd) == SURVEY_HASH
    assert all(v.dtype == torch.float32 for v in sd.values())
    assert not any("rel_pos_indices" in k for k in sd)  # non-persistent buffer, hybrid_CTUNet.py:479


def test_constructor_argument_errors():
    from hybrid_ctunet_b200.networks import hybrid_CTUNet as H
    from hybrid_ctunet_b200.networks import resnet, vit
    with pytest.raises(AssertionError):
        resnet.generate_model(18)
    with pytest.raises(AssertionError):
        vit.ViT(image_size=(90, 96), image_patch_size=16, frames=96, frame_patch_size=8, dim=64, depth=1, heads=1, mlp_dim=64)
    with pytest.raises(AssertionError):
        resnet.get_padding(1, 4)
    assert resnet.get_padding(3, 1) == 1 and resnet.get_padding(7, (2, 2, 1)) == (3, 3, 3)
    assert resnet.get_padding(1, 2) == 0 and resnet.get_padding((2, 2, 1), (2, 2, 1)) == (0, 0, 0)
    assert resnet.get_output_padding(2, 2, 0) == 0


def test_no_cpu_fallback():
    from hybrid_ctunet_b200.lib import CtuError
    from hybrid_ctunet_b200.networks import hybrid_CTUNet as H
    m = H.ResBlock(3, 64, 64, 3, 1, "instance")
    with pytest.raises(CtuError):
        with torch.no_grad():
            m(torch.zeros(1, 64, 4, 4, 4))


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("which", ["ctunet", "tunet", "cunet"])
def test_matches_reference_keys_shapes_and_seeded_init(which):
    _, _, hyb = ref_import.load()
    from hybrid_ctunet_b200.networks import hybrid_CTUNet as H
    if which == "ctunet":
        mk = lambda mod: mod.CTUNet(model_depth=101, **KW)
    elif which == "tunet":
        mk = lambda mod: mod.TUNet(**KW)
    else:
        mk = lambda mod: mod.CUNet(14, 101)
    torch.manual_seed(0)
    r = mk(hyb).state_dict()
    torch.manual_seed(0)
    m = mk(H).state_dict()
    assert list(r.keys()) == list(m.keys())
    for k in r:
        assert r[k].shape == m[k].shape, k
        assert torch.equal(r[k], m[k]), k


def test_dropin_install_serves_the_reference_module_names():
    """`from networks.hybrid_CTUNet import CTUNet` (main_CTUNet.py:21) resolves to the drop-in after dropin.install()."""
    import importlib
    import sys
    import hybrid_ctunet_b200.dropin as dropin
    saved_path, saved_mods = list(sys.path), {k: v for k, v in sys.modules.items() if k == "networks" or k.startswith("networks.")}
    try:
        dropin.install()
        hyb = importlib.import_module("networks.hybrid_CTUNet")
        res = importlib.import_module("networks.resnet")
        vit = importlib.import_module("networks.vit")
        from hybrid_ctunet_b200.networks import hybrid_CTUNet as ours
        assert hyb.CTUNet is ours.CTUNet and hyb.TUNet is ours.TUNet and hyb.CUNet is ours.CUNet
        assert callable(res.generate_model) and hasattr(vit, "ViT")
        fake = type(sys)("trainer_CTUNet")
        fake.sliding_window_inference = lambda *a, **k: None
        sys.modules["trainer_CTUNet"] = fake
        assert dropin.patch_trainers() == ["trainer_CTUNet"]
        from hybrid_ctunet_b200.trainer_CTUNet import sliding_window_inference
        assert fake.sliding_window_inference is sliding_window_inference
    finally:
        sys.modules.pop("trainer_CTUNet", None)
        for k in [k for k in sys.modules if k == "networks" or k.startswith("networks.")]:
            del sys.modules[k]
        sys.modules.update(saved_mods)
        sys.path[:] = saved_path
