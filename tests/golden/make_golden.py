"""Generates tests/golden/*.npz from the UNMODIFIED reference modules (build container only).

    python tests/golden/make_golden.py

Imports /root/reference/networks/*.py through oracle/monai_stub (MONAI is not installable here), runs them on
CPU in fp32 with seeded default-initialised weights and seeded inputs, and stores inputs + outputs (+ the
state_dict for the small block cases).  The GPU box has no /root/reference: the `-m gpu` tests load these files.
Whole-network outputs are stored sub-sampled (every 8th voxel per axis) to keep the fixtures small.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402

resnet, vit, hyb = ref_import.load()
torch.set_grad_enabled(False)


def sd_np(mod, prefix="w:"):
    return {prefix + k: v.numpy() for k, v in mod.state_dict().items()}


def probe_np(mod):
    """For blocks too large to ship their weights: the drop-in module re-draws the identical default init from the
    same torch.manual_seed (same layer types, same construction order); the probe pins that mapping."""
    last = list(mod.state_dict().values())[-1]
    return {"probe": last.flatten()[:16].numpy()}


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in arrs.items()})
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def blocks():
    # Bottleneck, identity shortcut (resnet.py:82-126)
    torch.manual_seed(10)
    m = resnet.Bottleneck(128, 32).eval()
    x = torch.randn(1, 128, 8, 8, 16)
    save("bottleneck_128_32", x=x, y=m(x), **sd_np(m))
    # Bottleneck, stride-2 with downsample branch (resnet.py:188-199)
    torch.manual_seed(11)
    ds = torch.nn.Sequential(resnet.get_conv_layer(3, 128, 256, kernel_size=1, stride=(2, 2, 2), conv_only=True),
                             torch.nn.InstanceNorm3d(256))
    m = resnet.Bottleneck(128, 64, stride=(2, 2, 2), downsample=ds).eval()
    x = torch.randn(1, 128, 8, 12, 16)
    save("bottleneck_down_128_64", x=x, y=m(x), **sd_np(m))
    # ResBlock 64->64 and 128->64 (hybrid_CTUNet.py:29-105)
    torch.manual_seed(12)
    m = hyb.ResBlock(3, 64, 64, 3, 1, "instance").eval()
    x = torch.randn(2, 64, 8, 8, 16)
    save("resblock_64_64", x=x, y=m(x), **sd_np(m))
    torch.manual_seed(13)
    m = hyb.ResBlock(3, 128, 64, 3, 1, "instance").eval()
    x = torch.randn(1, 128, 6, 10, 12)
    save("resblock_128_64", x=x, y=m(x), **sd_np(m))
    # ResBlock 1->64 (vit_encoder0)
    torch.manual_seed(14)
    m = hyb.ResBlock(3, 1, 64, 3, 1, "instance").eval()
    x = torch.randn(1, 1, 12, 12, 16)
    save("resblock_1_64", x=x, y=m(x), **sd_np(m))
    # pixelweight_attention (hybrid_CTUNet.py:622-669)
    torch.manual_seed(15)
    m = hyb.pixelweight_attention(128).eval()
    x1, x2 = torch.randn(1, 128, 6, 6, 12), torch.randn(1, 128, 6, 6, 12)
    save("pwa_128", x1=x1, x2=x2, y=m(x1, x2), **sd_np(m))
    # Up_2Fusion_Block (hybrid_CTUNet.py:257-341)
    torch.manual_seed(16)
    m = hyb.Up_2Fusion_Block(3, 256, 128, 3, (2, 2, 2), "instance").eval()
    inp, sc, sv = torch.randn(1, 256, 3, 3, 6), torch.randn(1, 128, 6, 6, 12), torch.randn(1, 128, 6, 6, 12)
    save("up2fusion_256_128", inp=inp, skip_conv=sc, skip_vit=sv, y=m(inp, sc, sv), **probe_np(m))  # weights: seed 16
    # UpConvBlock with the anisotropic (2,2,1) transposed conv (hybrid_CTUNet.py:203-255)
    torch.manual_seed(17)
    m = hyb.UpConvBlock(3, 128, 64, 3, (2, 2, 1), "instance").eval()
    x = torch.randn(1, 128, 4, 6, 16)
    save("upconv_128_64", x=x, y=m(x), **sd_np(m))
    # PixelShuffle (hybrid_CTUNet.py:388-432)
    torch.manual_seed(18)
    m = hyb.PixelShuffle(3, (2, 2, 2), 512, 256).eval()
    x = torch.randn(1, 512, 3, 4, 6)
    save("pixelshuffle_512_256", x=x, y=m(x), **sd_np(m))
    torch.manual_seed(19)
    m = hyb.PixelShuffle(3, (2, 2, 1), 128, 64).eval()
    x = torch.randn(2, 128, 4, 4, 8)
    save("pixelshuffle_221_128_64", x=x, y=m(x), **sd_np(m))
    # ViT attention / transformer block (vit.py:46-96)
    torch.manual_seed(20)
    m = vit.TransformerBlock(768, 12, 64, 3072).eval()
    x = torch.randn(1, 432, 768)
    save("vit_block", x=x, y=m(x), **probe_np(m))  # weights: seed 20
    # heads
    torch.manual_seed(21)
    m = hyb.UnetOutBlock(spatial_dims=3, in_channels=64, out_channels=14).eval()
    x = torch.randn(1, 64, 8, 8, 16)
    save("outblock_64_14", x=x, y=m(x), **sd_np(m))


def vit_small():
    torch.manual_seed(30)
    m = vit.ViT(image_size=(32, 32), image_patch_size=16, frames=48, frame_patch_size=8, dim=128, depth=2, heads=2,
                mlp_dim=256).eval()
    x = torch.randn(2, 1, 32, 32, 48)
    save("vit_small", x=x, y=m(x), **sd_np(m))


def up_attention():
    """UpAttentionBlock on the smallest legal grid (6,6,6): outputs sub-sampled.  The 23 M weights are not
    stored: the drop-in module draws the identical default init from torch.manual_seed(31) (same layer types in
    the same construction order; tests/test_state_dict.py pins that property against the reference)."""
    torch.manual_seed(31)
    m = hyb.UpAttentionBlock(3, 768, dims=[128, 256, 512, 1024]).eval()
    x = torch.randn(1, 768, 6, 6, 6)
    ys = m(x)
    save("up_attention_666", x=x, y1=ys[1], y2=ys[2][:, :, ::2, ::2, ::2], y3=ys[3][:, :, ::4, ::4, ::4],
         y4=ys[4][:, :, ::8, ::8, ::4], **probe_np(m))


def whole_nets():
    """Full-size CTUNet(101, pf 8) / TUNet on one 96^3 patch: weights = torch.manual_seed(0) default init of the
    reference module (reproducible on the GPU box: the drop-in modules draw the same values in the same order),
    input = torch.manual_seed(1) randn.  Outputs stored at every 8th voxel."""
    kw = dict(in_channels=1, dim_conv_stem=64, out_channels=14, img_size=(96, 96), frames=96, patch_frame=8)
    torch.manual_seed(0)
    m = hyb.CTUNet(model_depth=101, **kw).eval()
    torch.manual_seed(1)
    x = torch.randn(1, 1, 96, 96, 96)
    (r0, r1, r2), (v0, v1) = m(x)
    s = (slice(None), slice(None), slice(None, None, 8), slice(None, None, 8), slice(None, None, 8))
    stats = {}
    for name, t in (("res_logits", r0), ("res_48", r1), ("res_24", r2), ("vit_logits", v0), ("vit_96", v1)):
        stats[name + "_sub"] = t[s]
        stats[name + "_norm"] = np.array([t.double().norm().item(), t.double().mean().item(), t.double().std().item()])
    save("ctunet_101_pf8_seed0_x1", **stats)
    torch.manual_seed(0)
    m = hyb.TUNet(**kw).eval()
    v0, v1 = m(x)
    save("tunet_pf8_seed0_x1", vit_logits_sub=v0[s], vit_96_sub=v1[s],
         vit_logits_norm=np.array([v0.double().norm().item(), v0.double().mean().item(), v0.double().std().item()]))


def sliding_window_ref():
    """Outputs of the reference's own two-head sliding_window_inference (trainer_CTUNet.py:417-557, executed through
    oracle/ref_exec.py) on the shared parity cases: what the CUDA blend must reproduce on the GPU box."""
    from oracle import ref_exec
    fn = ref_exec.sliding_window_two_heads()
    arrs = {}
    for i, (shape, roi, swb, overlap, mode) in enumerate(ref_exec.SW_CASES):
        torch.manual_seed(100 + i)
        vol = torch.rand(shape)
        h0, h1 = fn(vol, roi, swb, ref_exec.sw_case_predictor(), overlap=overlap, mode=mode)
        arrs[f"vol{i}"], arrs[f"head0_{i}"], arrs[f"head1_{i}"] = vol, h0, h1
    save("sliding_window_ref", **arrs)


def _bf16_bits(t):
    return t.to(torch.bfloat16).view(torch.int16).numpy()


def resnet_stages():
    """The reference ResNet-101 encoder (resnet.py:128-245, DS_stride of hybrid_CTUNet.py:728) on one 32^3 patch with
    TEACHER FORCING at stage granularity: the stem output and every stage output are rounded to bf16 before the next
    stage consumes them (forward hooks on the unmodified module), so the drop-in, fed the same rounded tensors, is
    compared stage by stage without inheriting upstream error (SURVEY 8c level 2).  Stored: the rounded stage inputs
    (bf16 bit patterns) and the reference's fp32 outputs at every 2nd voxel.  Weights: torch.manual_seed(40) default
    init (the drop-in draws the same values; the probe pins that)."""
    torch.manual_seed(40)
    m = resnet.generate_model(101, DS_stride=((2, 2, 1), (2, 2, 2), (2, 2, 2), (2, 2, 2))).eval()
    torch.manual_seed(41)
    x = torch.randn(1, 1, 32, 32, 32)
    raw = {}

    def hook(name):
        def fn(mod, inp, out):
            raw[name] = out.clone()
            return out.to(torch.bfloat16).float()
        return fn
    hs = [m.lrelu.register_forward_hook(hook("stem"))]
    for li in (1, 2, 3, 4):
        hs.append(getattr(m, f"layer{li}").register_forward_hook(hook(f"layer{li}")))
    m(x)
    for h in hs:
        h.remove()
    sub = (slice(None), slice(None), slice(None, None, 2), slice(None, None, 2), slice(None, None, 2))
    arrs = {"x": x}
    for name in ("stem", "layer1", "layer2", "layer3", "layer4"):
        arrs["out_" + name] = raw[name][sub]
        arrs["norm_" + name] = np.array([raw[name].double().norm().item()])
        if name != "layer4":
            arrs["teacher_" + name] = _bf16_bits(raw[name])
    save("resnet101_stages_32", **arrs, **probe_np(m))


def postprocess_ref():
    """tests/golden/postprocess_ref.npz: the reference's remove_all_but_the_largest_connected_component
    (test_CTUNet_final.py:132-190, executed unmodified through oracle/ref_exec.py) on the synthetic label volumes of
    tests/test_postprocess_cpu.py::CASES."""
    from copy import deepcopy
    from scipy.ndimage import label
    from oracle import postprocess_oracle as PO
    from oracle import ref_exec
    sys.path.insert(0, os.path.join(HERE, ".."))
    from test_postprocess_cpu import CASES
    ref = ref_exec.extract("test_CTUNet_final.py", ["remove_all_but_the_largest_connected_component"],
                           extra_globals=dict(deepcopy=deepcopy, label=label))["remove_all_but_the_largest_connected_component"]
    d = {}
    for i, (shape, classes, vpv, mins) in enumerate(CASES):
        img = PO.blob_volume(shape, seed=i)
        out, rem, kept = ref(img, classes, vpv, mins)
        d[f"in{i}"], d[f"out{i}"] = img, out
        d[f"removed{i}"], d[f"kept{i}"] = np.array(rem, dtype=object), np.array(kept, dtype=object)
    np.savez_compressed(os.path.join(HERE, "postprocess_ref.npz"), **d)


if __name__ == "__main__":
    postprocess_ref()
    sliding_window_ref()
    resnet_stages()
    if "--only-new" in sys.argv:
        sys.exit(0)
    blocks()
    vit_small()
    up_attention()
    whole_nets()
