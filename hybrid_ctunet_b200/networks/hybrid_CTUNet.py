"""Drop-in for the reference's networks/hybrid_CTUNet.py: CTUNet, CUNet, TUNet and their decoder blocks.

Constructors, forward signatures, return structures and state_dict keys/shapes follow the reference
(hybrid_CTUNet.py:29-1036).  The torch layers are parameter holders created in the reference's order (so the
default initialisation under a given seed is identical); all arithmetic runs on the sm_100a kernels via
hybrid_ctunet_b200.engine.Engine.  Dead code of the reference (PixelweightConvBlock, Up_2Fusion_Block.forward_)
is not reproduced.
"""
from __future__ import annotations

from typing import Sequence, Tuple, Union

import torch
import torch.nn as nn

from ._base import KernelModule, Program, from_cl, to_cl
from .resnet import ConvHolder, generate_model as resnet
from .resnet import get_conv_layer
from .vit import ViT, _no_dropout

DS_STRIDE = ((2, 2, 1), (2, 2, 2), (2, 2, 2), (2, 2, 2))


def _check_norm(norm_name):
    name = norm_name[0] if isinstance(norm_name, (tuple, list)) else norm_name
    if str(name).lower() != "instance":
        raise NotImplementedError("only norm_name='instance' (every README command) is implemented on the CUDA path")


class ResBlock(KernelModule):
    """hybrid_CTUNet.py:29-105.  conv3 always exists in the state_dict, used only if in != out channels."""

    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size, stride, norm_name, dropout=None):
        super().__init__()
        _check_norm(norm_name)
        if kernel_size != 3 or stride != 1:
            raise NotImplementedError("CTUNet uses ResBlock with kernel 3, stride 1 only")
        self.conv1 = get_conv_layer(spatial_dims, in_channels, out_channels, kernel_size=kernel_size, stride=stride)
        self.conv2 = get_conv_layer(spatial_dims, out_channels, out_channels, kernel_size=kernel_size, stride=1)
        self.conv3 = get_conv_layer(spatial_dims, in_channels, out_channels, kernel_size=1, stride=stride)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.downsample = in_channels != out_channels

    def _program(self, eng, inp):
        eng.stats.reset()
        x = self._input(inp)
        if self.in_channels == 1:
            return Program([None], [eng.res_block_cin1("", x)])
        a = to_cl(x)
        return Program([a], [eng.res_block("", a, self.in_channels, self.out_channels)])

    def forward(self, inp):
        return self._call(inp)[0]


class BasicConvBlock(KernelModule):
    """hybrid_CTUNet.py:107-146."""

    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size, stride, norm_name):
        super().__init__()
        self.layer = ResBlock(spatial_dims, in_channels, out_channels, kernel_size, stride, norm_name)

    def forward(self, inp):
        return self.layer(inp)


class UpCatConvBlock(KernelModule):
    """hybrid_CTUNet.py:148-201 (CUNet decoder level)."""

    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size, upsample_kernel_size, norm_name):
        super().__init__()
        self.transp_conv = get_conv_layer(spatial_dims, in_channels, out_channels, kernel_size=upsample_kernel_size,
                                          stride=upsample_kernel_size, is_transposed=True)
        self.conv_block = ResBlock(spatial_dims, out_channels + out_channels, out_channels, kernel_size, 1, norm_name)
        self.out_channels = out_channels

    def _program(self, eng, inp, skip):
        eng.stats.reset()
        a, b = to_cl(self._input(inp)), to_cl(self._input(skip))
        return Program([a, b], [eng.up_cat_conv("", a, b, self.out_channels)])

    def forward(self, inp, skip):
        return self._call(inp, skip)[0]


class UpConvBlock(KernelModule):
    """hybrid_CTUNet.py:203-255."""

    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size, upsample_kernel_size, norm_name):
        super().__init__()
        self.transp_conv = get_conv_layer(spatial_dims, in_channels, out_channels, kernel_size=upsample_kernel_size,
                                          stride=upsample_kernel_size, is_transposed=True)
        self.conv_block = ResBlock(spatial_dims, out_channels, out_channels, kernel_size, 1, norm_name)
        self.out_channels = out_channels

    def _program(self, eng, inp):
        eng.stats.reset()
        a = to_cl(self._input(inp))
        up = eng.up_gemm(a, "convt", "transp_conv.conv")
        return Program([a], [eng.res_block("conv_block", up, self.out_channels, self.out_channels)])

    def forward(self, inp):
        return self._call(inp)[0]


class pixelweight_attention(KernelModule):
    """hybrid_CTUNet.py:622-669 — binary cross-weight fusion."""

    def __init__(self, dim, dim_head=32, dropout=0.0):
        super().__init__()
        _no_dropout(dropout, "dropout")
        if dim_head != 32:
            raise NotImplementedError("fusion kernel is built for dim_head = 32")
        self.dim_head = dim_head
        self.heads = dim // dim_head
        self.scale = dim_head ** -0.5
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.to_qkv1 = nn.Linear(dim, dim * 3, bias=False)
        self.to_qkv2 = nn.Linear(dim, dim * 3, bias=False)
        self.attend = nn.Sequential(nn.Softmax(dim=-1), nn.Dropout(dropout))
        self.to_out = nn.Sequential(nn.Linear(dim, dim, bias=False), nn.Dropout(dropout))

    def _program(self, eng, x1, x2):
        a, b = to_cl(self._input(x1)), to_cl(self._input(x2))
        return Program([a, b], [eng.pixelweight_attention("", a, b)])

    def forward(self, x1, x2):
        return self._call(x1, x2)[0]


class Up_2Fusion_Block(KernelModule):
    """hybrid_CTUNet.py:257-341."""

    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size, upsample_kernel_size, norm_name):
        super().__init__()
        self.transp_conv = get_conv_layer(spatial_dims, in_channels, out_channels, kernel_size=upsample_kernel_size,
                                          stride=upsample_kernel_size, is_transposed=True)
        self.pixelweight_attention1 = pixelweight_attention(out_channels)
        self.pixelweight_attention2 = pixelweight_attention(out_channels)
        self.up_addconv_block1 = ResBlock(spatial_dims, out_channels, out_channels, kernel_size, 1, norm_name)
        self.up_addconv_block2 = ResBlock(spatial_dims, out_channels, out_channels, kernel_size, 1, norm_name)
        self.out_channels = out_channels

    def _program(self, eng, inp, skip_conv, skip_vit):
        eng.stats.reset()
        a, b, c = to_cl(self._input(inp)), to_cl(self._input(skip_conv)), to_cl(self._input(skip_vit))
        return Program([a, b, c], [eng.up_2fusion("", a, b, c, self.out_channels)])

    def forward(self, inp, skip_conv=None, skip_vit=None):
        if skip_vit is None:
            # the reference raises NameError here (hybrid_CTUNet.py:332-338: `skip` is unbound)
            raise NameError("name 'skip' is not defined")
        return self._call(inp, skip_conv, skip_vit)[0]


class PixelShuffle(KernelModule):
    """hybrid_CTUNet.py:388-432: 3-D pixel shuffle followed by a per-voxel Linear."""

    def __init__(self, spatial_dims, scale_factor, in_channels, out_channels):
        super().__init__()
        self.spatial_dims = spatial_dims
        self.scale_factor = scale_factor
        self.to_out = nn.Linear(in_channels // (scale_factor[0] * scale_factor[1] * scale_factor[2]), out_channels)

    def forward(self, x):
        div = self.scale_factor[0] * self.scale_factor[1] * self.scale_factor[2]
        if x.shape[1] % div != 0:
            raise ValueError(f"Number of input channels ({x.shape[1]}) must be evenly"
                             f"divisibel by scale_factor ** dimensions ({self.scale_factor}**{self.spatial_dims}={div}).")
        return self._call(x)[0]

    def _program(self, eng, x):
        a = to_cl(self._input(x))
        return Program([a], [eng.up_gemm(a, "ps", "to_out", extra=tuple(self.scale_factor))])


class Residual(nn.Module):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn


class MultiAxisAttention(KernelModule):
    """hybrid_CTUNet.py:442-511 (parameter holder; executed inside UpAttentionBlock)."""

    def __init__(self, dim, dim_head=32, dropout=0.0, window_size=7):
        super().__init__()
        assert (dim % dim_head) == 0, 'dimension must be divisible by the head dimension'
        _no_dropout(dropout, "dropout")
        self.heads = dim // dim_head
        self.scale = dim_head ** -0.5
        self.norm = nn.LayerNorm(dim)
        self.to_qkv = nn.Linear(dim, dim * 3, bias=False)
        self.attend = nn.Sequential(nn.Softmax(dim=-1), nn.Dropout(dropout))
        self.to_out = nn.Sequential(nn.Linear(dim, dim, bias=False), nn.Dropout(dropout))
        self.rel_pos_bias = nn.Embedding((2 * window_size - 1) ** 3, self.heads)
        from ..engine import rel_pos_index
        self.register_buffer('rel_pos_indices', rel_pos_index(window_size), persistent=False)


class FeedForward(KernelModule):
    """hybrid_CTUNet.py:513-526 (parameter holder; executed inside UpAttentionBlock)."""

    def __init__(self, dim, mult=4, dropout=0.0):
        super().__init__()
        _no_dropout(dropout, "dropout")
        inner_dim = int(dim * mult)
        self.net = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, inner_dim), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(inner_dim, dim), nn.Dropout(dropout))


class UpAttentionBlock(KernelModule):
    """hybrid_CTUNet.py:528-591."""

    def __init__(self, spatial_dims, in_channels, dims=(512, 256, 128, 64), DS_stride=DS_STRIDE, depth=(1, 1, 1, 1),
                 dropout=0.0):
        super().__init__()
        if tuple(depth) != (1, 1, 1, 1) or tuple(map(tuple, DS_stride)) != DS_STRIDE:
            raise NotImplementedError("CTUNet/TUNet build UpAttentionBlock with depth (1,1,1,1) and the default DS_stride")
        dims = (in_channels, *dims[::-1][1:], 64)
        if dims != (768, 512, 256, 128, 64):
            raise NotImplementedError("stage widths other than (768,512,256,128,64) are not wired to kernels")
        self.layers = nn.ModuleList([])
        w = 6
        for ind, (d_in, d_out) in enumerate(zip(dims[:-1], dims[1:])):
            f = DS_stride[::-1][ind]
            if ind <= 2:
                block = nn.Sequential(
                    nn.Identity(),  # Rearrange 'b c (h h1) (w w1) (f f1) -> b h w f h1 w1 f1 c'
                    Residual(MultiAxisAttention(dim=d_in, dim_head=32, dropout=dropout, window_size=w)),
                    Residual(FeedForward(d_in, dropout=dropout)),
                    nn.Identity(),  # back to 'b c (h h1) (w w1) (f f1)'
                    nn.Identity(),  # Rearrange 'b c (h1 h) (w1 w) (f1 f) -> b h w f h1 w1 f1 c'
                    Residual(MultiAxisAttention(dim=d_in, dim_head=32, dropout=dropout, window_size=w)),
                    Residual(FeedForward(d_in, dropout=dropout)),
                    nn.Identity(),
                    PixelShuffle(spatial_dims, f, d_in, d_out))
            else:
                block = nn.Sequential(nn.Identity(), Residual(FeedForward(d_in, dropout=dropout)),
                                      Residual(FeedForward(d_in, dropout=dropout)), nn.Identity(),
                                      PixelShuffle(spatial_dims, f, d_in, d_out))
            self.layers.append(nn.Sequential(block))

    def forward(self, x):
        """x: [B, 768, X, Y, Z] -> [x, 512@2x, 256@4x, 128@8x, 64@(16,16,8)x] like hybrid_CTUNet.py:585-591."""
        return [x] + list(self._call(x))

    def _program(self, eng, x):
        B, C, X, Y, Z = x.shape
        tokens = self._input(x).permute(0, 2, 3, 4, 1).reshape(B * X * Y * Z, C).contiguous()
        return Program([tokens.view(B, X, Y, Z, C)], eng.up_attention_block("", tokens, B, (X, Y, Z)))


class CatConvBlock(KernelModule):
    """hybrid_CTUNet.py:593-620."""

    def __init__(self, spatial_dims, in_channels, kernel_size, norm_name):
        super().__init__()
        self.conv_block = ResBlock(spatial_dims, in_channels + in_channels, in_channels, kernel_size, 1, norm_name)
        self.in_channels = in_channels

    def _program(self, eng, x, skip):
        eng.stats.reset()
        C = self.in_channels
        cat = torch.cat((to_cl(self._input(x)), to_cl(self._input(skip))), dim=-1)
        lo, hi = cat[..., :C], cat[..., C:]
        eng._alias(lo, cat, 0)
        eng._alias(hi, cat, C)
        return Program([lo, hi], [eng.res_block("conv_block", cat, 2 * C, C)])

    def forward(self, x, skip):
        return self._call(x, skip)[0]


class DecoderLinear(KernelModule):
    """hybrid_CTUNet.py:671-691 (patch_size 1: per-voxel Linear d_encoder -> n_cls)."""

    def __init__(self, n_cls, patch_size, d_encoder):
        super().__init__()
        self.d_encoder, self.patch_size, self.n_cls = d_encoder, patch_size, n_cls
        self.head = nn.Linear(self.d_encoder, n_cls)

    @torch.jit.ignore
    def no_weight_decay(self):
        return set()

    def forward(self, x, im_size):
        F_, H, W = im_size
        if self.patch_size != 1:
            raise NotImplementedError("CTUNet uses DecoderLinear with patch_size 1")
        object.__setattr__(self, "_im_size", (F_, H, W))
        return self._call(x)[0]

    def _program(self, eng, x):
        F_, H, W = self._im_size
        cl = self._input(x).to(torch.bfloat16).reshape(x.shape[0], F_, H, W, self.d_encoder)
        return Program([cl], [eng.head(cl, "lin", "head")], converted=False)


class UnetOutBlock(KernelModule):
    """MONAI 0.7 UnetOutBlock as the reference uses it (hybrid_CTUNet.py:781-783,810): 1x1x1 conv + bias."""

    def __init__(self, spatial_dims, in_channels, out_channels, dropout=None):
        super().__init__()
        self.conv = get_conv_layer(spatial_dims, in_channels, out_channels, kernel_size=1, stride=1, bias=True)

    def _program(self, eng, inp):
        a = to_cl(self._input(inp))
        return Program([a], [eng.head(a, "conv1", "conv.conv")], converted=False)

    def forward(self, inp):
        return self._call(inp)[0]


class _Net(KernelModule):
    """Whole-network boundary: fp32 NCDHW in, fp32 NCDHW logits out, optional CUDA-graph replay in eval mode."""

    def _run(self, eng, x):
        raise NotImplementedError

    def enable_cuda_graph(self, flag: bool = True):
        """Replay the whole forward as one CUDA graph per input shape (inference only).  The returned logits are
        the graph's static output buffers: consume them before the next call (sliding-window blending does)."""
        object.__setattr__(self, "_use_graph", bool(flag))
        object.__setattr__(self, "_graphs", {})
        return self

    def _graph_forward(self, eng, x):
        graphs = getattr(self, "_graphs", None)
        if graphs is None:
            graphs = {}
            object.__setattr__(self, "_graphs", graphs)
        # The graph reads the packed bf16 copies / tables of engine.WeightCache.  Those are refreshed IN PLACE (same
        # storage), so a captured graph stays valid across weight updates: when any parameter's version moved since the
        # last call they are re-packed with one launch and the graph is replayed.  (Optimizers that update parameters
        # without bumping Tensor._version are covered by the refresh every training forward and train()/eval() switch
        # does; hybrid_ctunet_b200.optim.AdamW bumps the versions itself.)  The graph is re-captured only when a
        # parameter's storage moved (.to(), load_state_dict(assign=True)) or the cache replaced a buffer.
        params = list(self.parameters())
        vers = tuple(p._version for p in params)
        if getattr(self, "_graph_vers", None) != vers:
            if getattr(self, "_graph_vers", None) is not None:
                eng.w.refresh_all()
            object.__setattr__(self, "_graph_vers", vers)
        tag = (tuple(p.data_ptr() for p in params), eng.w.storage_epoch)
        key = tuple(x.shape)
        ent = graphs.get(key)
        if ent is not None and ent["tag"] != tag:
            ent = None
        if ent is None:
            static_x = x.clone()
            self._run(eng, static_x)  # warm-up: packs weights, sets kernel attributes, primes the allocator
            torch.cuda.synchronize()
            eng.prepare_for_capture()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._run(eng, static_x)
            tag = (tag[0], eng.w.storage_epoch)   # the warm-up call may have (re)built cache entries
            ent = dict(graph=g, x=static_x, out=out, tag=tag)
            graphs[key] = ent
        ent["x"].copy_(x)
        ent["graph"].replay()
        return ent["out"]

    def _program(self, eng, x_in):
        outs = self._run(eng, self._input(x_in))
        flat = [t for grp in outs for t in (grp if isinstance(grp, (tuple, list)) else (grp,))]
        return Program([None], flat, converted=False)

    def _nest(self, flat):
        """Flat logits tuple -> the reference's return structure."""
        return tuple(flat)

    def forward(self, x_in):
        training = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if getattr(self, "_use_graph", False) and not training:
            eng = self._engine()
            eng.tape = None
            with torch.no_grad():
                return self._graph_forward(eng, self._input(x_in))
        return self._nest(self._call(x_in))


class CTUNet(_Net):
    """hybrid_CTUNet.py:694-857."""

    def __init__(self, in_channels: int, dim_conv_stem: int, out_channels: int, model_depth: int,
                 img_size: Tuple[int, int], frames: int, patch_frame: int, hidden_size: int = 768,
                 num_depths: int = 12, mlp_dim: int = 3072, num_heads: int = 12,
                 norm_name: Union[Tuple, str] = "instance", dropout_rate: float = 0.0) -> None:
        super().__init__()
        _check_norm(norm_name)
        if in_channels != 1 or dim_conv_stem != 64:
            raise NotImplementedError("vit_encoder0 kernels cover in_channels=1, dim_conv_stem=64 (README configuration)")
        self.patch_size = (16, 16, patch_frame)
        self.feat_size = (img_size[0] // 16, img_size[1] // 16, frames // patch_frame)
        if any(f % 6 for f in self.feat_size):
            raise ValueError("token grid must be divisible by the 6x6x6 attention window (hybrid_CTUNet.py:551)")
        self.hidden_size = hidden_size
        self.model_depth, self.patch_frame, self.num_depths, self.num_heads = model_depth, patch_frame, num_depths, num_heads
        dims = [int(4 * item) for item in [32, 64, 128, 256]]
        self.convnet = resnet(model_depth, DS_stride=DS_STRIDE)
        self.vit = ViT(image_size=img_size, image_patch_size=16, frames=frames, frame_patch_size=patch_frame,
                       dim=hidden_size, depth=num_depths, heads=num_heads, mlp_dim=mlp_dim, dropout=dropout_rate,
                       emb_dropout=dropout_rate, drop_path=dropout_rate)
        self.res_decoder3 = Up_2Fusion_Block(3, dims[3], dims[2], 3, DS_STRIDE[3], norm_name)
        self.res_decoder2 = Up_2Fusion_Block(3, dims[2], dims[1], 3, DS_STRIDE[2], norm_name)
        self.res_decoder1 = Up_2Fusion_Block(3, dims[1], dims[0], 3, DS_STRIDE[1], norm_name)
        self.res_decoder0 = UpConvBlock(3, dims[0], 64, 3, DS_STRIDE[0], norm_name)
        self.res_out = UnetOutBlock(spatial_dims=3, in_channels=64, out_channels=out_channels)
        self.res_out_48x48 = UnetOutBlock(spatial_dims=3, in_channels=dims[0], out_channels=out_channels)
        self.res_out_24x24 = UnetOutBlock(spatial_dims=3, in_channels=dims[1], out_channels=out_channels)
        self.vit_encoder0 = BasicConvBlock(3, in_channels, dim_conv_stem, 3, 1, norm_name)
        self.vit_encoder = UpAttentionBlock(3, hidden_size, dims=dims, DS_stride=DS_STRIDE, depth=(1, 1, 1, 1),
                                            dropout=dropout_rate)
        self.vit_decoder0 = CatConvBlock(3, dim_conv_stem, 3, norm_name)
        self.decoder_linear_96x96 = DecoderLinear(out_channels, 1, 64)
        self.vit_out = UnetOutBlock(spatial_dims=3, in_channels=dim_conv_stem, out_channels=out_channels)

    def proj_feat(self, x, hidden_size, feat_size):
        x = x.view(x.size(0), feat_size[0], feat_size[1], feat_size[2], hidden_size)
        return x.permute(0, 4, 1, 2, 3).contiguous()

    def _run(self, eng, x):
        return eng.ctunet(x, self.convnet.block_counts, self.patch_frame, self.num_depths, self.num_heads)

    def _nest(self, flat):
        return ((flat[0], flat[1], flat[2]), (flat[3], flat[4]))


class CUNet(_Net):
    """hybrid_CTUNet.py:859-937."""

    def __init__(self, out_channels: int, model_depth: int, norm_name: Union[Tuple, str] = "instance") -> None:
        super().__init__()
        _check_norm(norm_name)
        dims = [int(4 * item) for item in [32, 64, 128, 256]]
        self.convnet = resnet(model_depth, DS_stride=DS_STRIDE)
        self.res_decoder3 = UpCatConvBlock(3, dims[3], dims[2], 3, DS_STRIDE[3], norm_name)
        self.res_decoder2 = UpCatConvBlock(3, dims[2], dims[1], 3, DS_STRIDE[2], norm_name)
        self.res_decoder1 = UpCatConvBlock(3, dims[1], dims[0], 3, DS_STRIDE[1], norm_name)
        self.res_decoder0 = UpConvBlock(3, dims[0], 64, 3, DS_STRIDE[0], norm_name)
        self.res_out = UnetOutBlock(spatial_dims=3, in_channels=64, out_channels=out_channels)
        self.res_out_48x48 = UnetOutBlock(spatial_dims=3, in_channels=dims[0], out_channels=out_channels)
        self.res_out_24x24 = UnetOutBlock(spatial_dims=3, in_channels=dims[1], out_channels=out_channels)

    def _run(self, eng, x):
        return eng.cunet(x, self.convnet.block_counts)


class TUNet(_Net):
    """hybrid_CTUNet.py:939-1036."""

    def __init__(self, in_channels: int, dim_conv_stem: int, out_channels: int, img_size: Tuple[int, int], frames: int,
                 patch_frame: int, hidden_size: int = 768, num_depths: int = 12, mlp_dim: int = 3072,
                 num_heads: int = 12, norm_name: Union[Tuple, str] = "instance", dropout_rate: float = 0.0) -> None:
        super().__init__()
        _check_norm(norm_name)
        if in_channels != 1 or dim_conv_stem != 64:
            raise NotImplementedError("vit_encoder0 kernels cover in_channels=1, dim_conv_stem=64 (README configuration)")
        self.patch_size = (16, 16, patch_frame)
        self.feat_size = (img_size[0] // 16, img_size[1] // 16, frames // patch_frame)
        if any(f % 6 for f in self.feat_size):
            raise ValueError("token grid must be divisible by the 6x6x6 attention window (hybrid_CTUNet.py:551)")
        self.hidden_size = hidden_size
        self.patch_frame, self.num_depths, self.num_heads = patch_frame, num_depths, num_heads
        dims = [int(4 * item) for item in [32, 64, 128, 256]]
        self.vit = ViT(image_size=img_size, image_patch_size=16, frames=frames, frame_patch_size=patch_frame,
                       dim=hidden_size, depth=num_depths, heads=num_heads, mlp_dim=mlp_dim, dropout=dropout_rate,
                       emb_dropout=dropout_rate, drop_path=dropout_rate)
        self.vit_encoder0 = BasicConvBlock(3, in_channels, dim_conv_stem, 3, 1, norm_name)
        self.vit_encoder = UpAttentionBlock(3, hidden_size, dims=dims, DS_stride=DS_STRIDE, depth=(1, 1, 1, 1),
                                            dropout=dropout_rate)
        self.vit_decoder0 = CatConvBlock(3, dim_conv_stem, 3, norm_name)
        self.decoder_linear_96x96 = DecoderLinear(out_channels, 1, 64)
        self.vit_out = UnetOutBlock(spatial_dims=3, in_channels=dim_conv_stem, out_channels=out_channels)

    def proj_feat(self, x, hidden_size, feat_size):
        x = x.view(x.size(0), feat_size[0], feat_size[1], feat_size[2], hidden_size)
        return x.permute(0, 4, 1, 2, 3).contiguous()

    def _run(self, eng, x):
        return eng.tunet(x, self.patch_frame, self.num_depths, self.num_heads)
