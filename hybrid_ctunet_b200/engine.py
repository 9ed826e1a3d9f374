"""Forward engine of the CTUNet path on the sm_100a kernels.

Everything between the module boundary (fp32 NCDHW in, fp32 NCDHW logits out) runs here on channels-last bf16
activations; the einops rearranges of the reference (window / grid partition, proj_feat, pixel shuffle, token
<-> volume views) are folded into kernel indexing, so no permute copies exist on this path.

Reference call graph this engine restates (hybrid_CTUNet.py:817-857): vit -> vit_encoder0 -> vit_encoder ->
vit_decoder0 -> heads; convnet -> res_decoder3..0 -> heads.  Block functions cite the reference per function.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from .ops import ACT_GELU, OUT_BF16, OUT_F32, OUT_F32_CF, PackedWeight

DS_STRIDE = ((2, 2, 1), (2, 2, 2), (2, 2, 2), (2, 2, 2))
BF16 = torch.bfloat16


def _pad_to(v: int, m: int) -> int:
    return -(-v // m) * m


def rel_pos_index(w: int) -> torch.Tensor:
    """Index table of MultiAxisAttention (hybrid_CTUNet.py:472-477): [w^3, w^3] into the (2w-1)^3 embedding."""
    pos = torch.arange(w)
    g = torch.stack(torch.meshgrid(pos, pos, pos, indexing="ij")).reshape(3, -1).t()
    rel = g[:, None, :] - g[None, :, :] + (w - 1)
    return (rel * torch.tensor([(2 * w - 1) ** 2, 2 * w - 1, 1])).sum(-1)


class WeightCache:
    """bf16 kernel-layout copies of the module's fp32 parameters, refreshed when a parameter changes."""

    def __init__(self, params: Dict[str, torch.Tensor]):
        self.params = params
        self._cache: Dict[str, Tuple[tuple, object]] = {}

    def _get(self, key: str, names, build):
        ps = [self.params[n.lstrip('.')] for n in names]
        tag = tuple((p.data_ptr(), p._version, str(p.device)) for p in ps)
        hit = self._cache.get(key)
        if hit is not None and hit[0] == tag:
            return hit[1]
        with torch.no_grad():
            val = build(*[p.detach() for p in ps])
        self._cache[key] = (tag, val)
        return val

    # -- nn.Linear [N, K] (+bias)
    def linear(self, name: str, bias: bool = True, block_n: Optional[int] = None) -> PackedWeight:
        names = [name + ".weight"] + ([name + ".bias"] if bias else [])
        return self._get("lin:" + name, names,
                         lambda w, b=None: ops.pack_matrix(w, bias=b, block_n=block_n))

    # -- Conv3d 1x1x1 [Cout, Cin, 1,1,1]; channel counts below 64 are zero-padded to 64
    def conv1(self, name: str, bias: bool = False) -> PackedWeight:
        names = [name + ".weight"] + ([name + ".bias"] if bias else [])

        def build(w, b=None):
            co, ci = w.shape[:2]
            w2 = w.reshape(co, ci)
            if b is None:  # feature convs: pad to the 64-channel granularity of the activation buffers
                cop, cip = max(co, 64), max(ci, 64)
                if (cop, cip) != (co, ci):
                    wp = torch.zeros(cop, cip, device=w.device, dtype=w.dtype)
                    wp[:co, :ci] = w2
                    w2 = wp
            return ops.pack_matrix(w2, bias=b)
        return self._get("c1:" + name, names, build)

    # -- Conv3d 3x3x3 [Cout, Cin, 3,3,3] -> [Cout, 27*Cin] tap-major
    def conv3(self, name: str) -> PackedWeight:
        def build(w):
            co, ci = w.shape[:2]
            cop, cip = max(co, 64), max(ci, 64)
            wp = torch.zeros(cop, 3, 3, 3, cip, device=w.device, dtype=w.dtype)
            wp[:co, ..., :ci] = w.permute(0, 2, 3, 4, 1)
            return ops.pack_matrix(wp.reshape(cop, 27 * cip), ksize=3, a_c=cip)
        return self._get("c3:" + name, [name + ".weight"], build)

    # -- ConvTranspose3d kernel == stride [Cin, Cout, kX, kY, kZ]
    def convt(self, name: str) -> PackedWeight:
        def build(w):
            ci, co, kx, ky, kz = w.shape
            w2 = w.permute(2, 3, 4, 1, 0).reshape(kx * ky * kz * co, ci)
            return ops.pack_matrix(w2, block_n=64 if co % 128 else 128, convt=(co, kz, ky, kx))
        return self._get("ct:" + name, [name + ".weight"], build)

    # -- PixelShuffle + Linear (hybrid_CTUNet.py:404-432) as a transposed-conv-shaped GEMM
    def pixel_shuffle(self, name: str, factor) -> PackedWeight:
        def build(w, b):
            co, corg = w.shape
            fx, fy, fz = factor
            k3 = fx * fy * fz
            big = torch.zeros(k3, co, corg, k3, device=w.device, dtype=w.dtype)
            for s in range(k3):
                big[s, :, :, s] = w
            big = big.reshape(k3 * co, corg * k3)
            return ops.pack_matrix(big, bias=b.repeat(k3), block_n=64 if co % 128 else 128, convt=(co, fz, fy, fx))
        return self._get("ps:" + name, [name + ".weight", name + ".bias"], build)

    # -- single-input-channel convs on CUDA cores: fp32 [taps, 64]
    def conv_cin1(self, name: str) -> torch.Tensor:
        return self._get("d1:" + name, [name + ".weight"],
                         lambda w: w.reshape(w.shape[0], -1).t().contiguous().float())

    def rel_bias(self, name: str, w: int = 6) -> torch.Tensor:
        def build(emb):
            idx = rel_pos_index(w).to(emb.device)
            return emb[idx].permute(2, 0, 1).contiguous().float()
        return self._get("rb:" + name, [name + ".weight"], build)

    def f32(self, name: str) -> torch.Tensor:
        return self._get("f:" + name, [name], lambda p: p.float().contiguous())


class StatsArena:
    """fp64 (sum, sumsq) accumulators for every InstanceNorm of one forward, zeroed with a single memset."""

    def __init__(self, device, capacity: int = 1 << 18):
        self.buf = torch.zeros(capacity, dtype=torch.float64, device=device)
        self.off = 0

    def reset(self):
        self.buf.zero_()
        self.off = 0

    def take(self, B: int, C: int) -> torch.Tensor:
        n = B * C * 2
        if self.off + n > self.buf.numel():
            raise RuntimeError("InstanceNorm statistics arena exhausted")
        t = self.buf[self.off:self.off + n].view(B, C, 2)
        self.off += n
        return t


class Engine:
    def __init__(self, params: Dict[str, torch.Tensor], device):
        self.w = WeightCache(params)
        self.dev = device
        self.stats = StatsArena(device)

    # ------------------------------------------------------------------ helpers
    def _empty(self, *shape, dtype=BF16):
        return torch.empty(shape, dtype=dtype, device=self.dev)

    @staticmethod
    def _dims(x):  # channels-last [B, X, Y, Z, C] -> (d1, d2, d3, d4)
        B, X, Y, Z, _ = x.shape
        return (Z, Y, X, B)

    @staticmethod
    def _flat_dims(x):  # per-batch token GEMM dims
        B, X, Y, Z, _ = x.shape
        return (X * Y * Z, 1, 1, B)

    def conv3x3(self, x, pw: PackedWeight, stats=None, out=None):
        B, X, Y, Z, _ = x.shape
        if out is None:
            out = self._empty(B, X, Y, Z, pw.n_real)
        ops.gemm(x, pw, out, dims=self._dims(x), stats=stats, a_c=pw.a_c)
        return out

    def conv1x1(self, x, pw: PackedWeight, stats=None, out=None):
        B, X, Y, Z, _ = x.shape
        if out is None:
            out = self._empty(B, X, Y, Z, pw.n_real)
        ops.gemm(x, pw, out, dims=self._flat_dims(x), stats=stats, a_c=pw.a_c)
        return out

    def up_gemm(self, x, pw: PackedWeight, out=None, a_c=None):
        """ConvTranspose3d(k=s) / pixel-shuffle+Linear: [B,X,Y,Z,Cin] -> [B,X*ux,Y*uy,Z*uz,Cout]."""
        B, X, Y, Z, _ = x.shape
        co, uz, uy, ux = pw.convt
        if out is None:
            out = self._empty(B, X * ux, Y * uy, Z * uz, co)
        ops.gemm(x, pw, out, dims=self._dims(x), a_c=a_c)
        return out

    def head(self, x, pw: PackedWeight, a_c=None):
        """UnetOutBlock / DecoderLinear: per-voxel C -> n_cls with bias, fp32 NCDHW output."""
        B, X, Y, Z, _ = x.shape
        out = self._empty(B, pw.n_real, X, Y, Z, dtype=torch.float32)
        ops.gemm(x, pw, out, dims=self._flat_dims(x), out_mode=OUT_F32_CF, a_c=a_c)
        return out

    # ------------------------------------------------------------------ networks/resnet.py
    def bottleneck(self, pre: str, x, stride, has_down: bool):
        """resnet.py:106-126: 1x1 -> IN -> lrelu -> 3x3x3(stride) -> IN -> lrelu -> 1x1 -> IN (+res) -> lrelu."""
        B = x.shape[0]
        w1, w2, w3 = self.w.conv1(pre + ".conv1.conv"), self.w.conv3(pre + ".conv2.conv"), self.w.conv1(pre + ".conv3.conv")
        st1 = self.stats.take(B, w1.n_real)
        c1 = self.conv1x1(x, w1, st1)
        ops.in_apply(c1, st1, c1, act=True)
        st2 = self.stats.take(B, w2.n_real)
        strided = tuple(stride) != (1, 1, 1)
        if not strided:
            c2 = self.conv3x3(c1, w2, st2)
        else:
            # stride-s 3x3x3, pad 1 == the stride-1 result sampled at multiples of s
            full = self.conv3x3(c1, w2)
            _, X, Y, Z, C = full.shape
            c2 = self._empty(B, -(-X // stride[0]), -(-Y // stride[1]), -(-Z // stride[2]), C)
            ops.subsample(full, c2, stride)
            ops.in_stats(c2, st2)
        ops.in_apply(c2, st2, c2, act=True)
        st3 = self.stats.take(B, w3.n_real)
        c3 = self.conv1x1(c2, w3, st3)
        if has_down:
            wd = self.w.conv1(pre + ".downsample.0.conv")
            xs = x
            if strided:
                _, X, Y, Z, C = x.shape
                xs = self._empty(B, -(-X // stride[0]), -(-Y // stride[1]), -(-Z // stride[2]), C)
                ops.subsample(x, xs, stride)
            std = self.stats.take(B, wd.n_real)
            r = self.conv1x1(xs, wd, std)
            ops.in_apply(c3, st3, c3, res=r, rstats=std, act=True)
        else:
            ops.in_apply(c3, st3, c3, res=x, act=True)
        return c3

    def resnet(self, pre: str, x_in, layers: List[int]):
        """resnet.py:213-230 (no max pool): stem k7 s(2,2,1) -> IN -> lrelu -> 4 stages; returns 4 feature maps."""
        B, _, X, Y, Z = x_in.shape
        s0 = DS_STRIDE[0]
        wst = self.w.conv_cin1(pre + "conv1.conv")
        x = self._empty(B, (X + 6 - 7) // s0[0] + 1, (Y + 6 - 7) // s0[1] + 1, (Z + 6 - 7) // s0[2] + 1, 64)
        ops.conv_cin1(x_in, wst, x, k=(7, 7, 7), s=s0, p=(3, 3, 3))
        st = self.stats.take(B, 64)
        ops.in_stats(x, st)
        ops.in_apply(x, st, x, act=True)
        feats = []
        strides = [(1, 1, 1), DS_STRIDE[1], DS_STRIDE[2], DS_STRIDE[3]]
        for li, nb in enumerate(layers):
            for bi in range(nb):
                x = self.bottleneck(f"{pre}layer{li + 1}.{bi}", x, strides[li] if bi == 0 else (1, 1, 1), bi == 0)
            feats.append(x)
        return feats

    # ------------------------------------------------------------------ networks/vit.py
    def ffn(self, pre: str, x, out=None):
        """LN -> Linear -> GELU -> Linear, + x (vit.py:34-44,95; hybrid_CTUNet.py:517-526 inside Residual)."""
        M, D = x.shape
        h = self._empty(M, D)
        ops.layernorm(x, self.w.f32(pre + ".net.0.weight"), self.w.f32(pre + ".net.0.bias"), h)
        w1, w2 = self.w.linear(pre + ".net.1"), self.w.linear(pre + ".net.4")
        f = self._empty(M, w1.n_real)
        ops.gemm(h, w1, f, dims=(M, 1, 1, 1), act=ACT_GELU)
        out = x if out is None else out
        ops.gemm(f, w2, out, dims=(M, 1, 1, 1), out_mode=OUT_F32 if out.dtype == torch.float32 else OUT_BF16, residual=x)
        return out

    def vit_attention(self, pre: str, x, B: int, n: int, heads: int):
        """vit.py:66-78 + residual (vit.py:94); x: fp32 [B*n, D] updated in place."""
        M, D = x.shape
        h = self._empty(M, D)
        ops.layernorm(x, self.w.f32(pre + ".norm.weight"), self.w.f32(pre + ".norm.bias"), h)
        wq, wo = self.w.linear(pre + ".to_qkv", bias=False), self.w.linear(pre + ".to_out.0")
        qkv = self._empty(M, 3 * D)
        ops.gemm(h, wq, qkv, dims=(M, 1, 1, 1))
        a = self._empty(M, D)
        ops.attention(qkv, a, dim_head=D // heads, n=n, windows=B, mode=0)
        ops.gemm(a, wo, x, dims=(M, 1, 1, 1), out_mode=OUT_F32, residual=x)
        return x

    def vit(self, pre: str, x_in, pf: int, depth: int, heads: int):
        """vit.py:130-139: returns the fp32 token stream [B*n, dim]."""
        B, _, X, Y, Z = x_in.shape
        n = (X // 16) * (Y // 16) * (Z // pf)
        e = pre + "to_patch_embedding"
        tok = self._empty(B * n, 256 * pf)
        ops.patchify_ln(x_in, pf, self.w.f32(e + ".1.weight"), self.w.f32(e + ".1.bias"), tok)
        wemb = self.w.linear(e + ".2")
        dim = wemb.n_real
        emb = self._empty(B * n, dim, dtype=torch.float32)
        ops.gemm(tok, wemb, emb, dims=(B * n, 1, 1, 1), out_mode=OUT_F32)
        x = self._empty(B * n, dim, dtype=torch.float32)
        ops.layernorm(emb, self.w.f32(e + ".3.weight"), self.w.f32(e + ".3.bias"), x, add=self.w.f32(pre + "pos_embedding"))
        for i in range(depth):
            t = f"{pre}transformer.{i}"
            self.vit_attention(t + ".attn", x, B, n, heads)
            self.ffn(t + ".ff", x)
        return x, n

    # ------------------------------------------------------------------ networks/hybrid_CTUNet.py
    def res_block(self, pre: str, x, cin: int, cout: int, out=None):
        """hybrid_CTUNet.py:93-105 (k3, stride 1): conv-IN-lrelu-conv-IN, + (conv1x1-IN)(x) or x, lrelu."""
        B = x.shape[0]
        w1, w2 = self.w.conv3(pre + ".conv1.conv"), self.w.conv3(pre + ".conv2.conv")
        st1 = self.stats.take(B, w1.n_real)
        c1 = self.conv3x3(x, w1, st1)
        ops.in_apply(c1, st1, c1, act=True)
        st2 = self.stats.take(B, w2.n_real)
        c2 = self.conv3x3(c1, w2, st2)
        out = c2 if out is None else out
        if cin != cout:
            w3 = self.w.conv1(pre + ".conv3.conv")
            st3 = self.stats.take(B, w3.n_real)
            r = self.conv1x1(x, w3, st3)
            ops.in_apply(c2, st2, out, res=r, rstats=st3, act=True)
        else:
            ops.in_apply(c2, st2, out, res=x, act=True)
        return out

    def res_block_cin1(self, pre: str, x_in, out=None):
        """ResBlock(1 -> 64) of vit_encoder0 (hybrid_CTUNet.py:786-793): conv1/conv3 have one input channel."""
        B, _, X, Y, Z = x_in.shape
        c1 = self._empty(B, X, Y, Z, 64)
        ops.conv_cin1(x_in, self.w.conv_cin1(pre + ".conv1.conv"), c1, k=(3, 3, 3), s=(1, 1, 1), p=(1, 1, 1))
        st1 = self.stats.take(B, 64)
        ops.in_stats(c1, st1)
        ops.in_apply(c1, st1, c1, act=True)
        w2 = self.w.conv3(pre + ".conv2.conv")
        st2 = self.stats.take(B, 64)
        c2 = self.conv3x3(c1, w2, st2)
        ops.conv_cin1(x_in, self.w.conv_cin1(pre + ".conv3.conv"), c1, k=(1, 1, 1), s=(1, 1, 1), p=(0, 0, 0))
        st3 = self.stats.take(B, 64)
        ops.in_stats(c1, st3)
        out = c2 if out is None else out
        ops.in_apply(c2, st2, out, res=c1, rstats=st3, act=True)
        return out

    def pixelweight_attention(self, pre: str, x1, x2):
        """hybrid_CTUNet.py:645-669, the binary cross-weight fusion of two [B,X,Y,Z,C] maps."""
        B, X, Y, Z, C = x1.shape
        T = B * X * Y * Z
        h = self._empty(T, C)
        q = []
        for x, nrm, lin in ((x1, ".norm1", ".to_qkv1"), (x2, ".norm2", ".to_qkv2")):
            ops.layernorm(x.reshape(T, C), self.w.f32(pre + nrm + ".weight"), self.w.f32(pre + nrm + ".bias"), h)
            qkv = self._empty(T, 3 * C)
            ops.gemm(h, self.w.linear(pre + lin, bias=False), qkv, dims=(T, 1, 1, 1))
            q.append(qkv)
        ops.pwa_fuse(q[0], q[1], h)
        out = self._empty(B, X, Y, Z, C)
        ops.gemm(h, self.w.linear(pre + ".to_out.0", bias=False), out, dims=(T, 1, 1, 1))
        return out

    def up_2fusion(self, pre: str, inp, skip_conv, skip_vit, cout: int):
        """hybrid_CTUNet.py:329-341."""
        skip = self.pixelweight_attention(pre + ".pixelweight_attention1", skip_conv, skip_vit)
        skip = self.res_block(pre + ".up_addconv_block1", skip, cout, cout)
        up = self.up_gemm(inp, self.w.convt(pre + ".transp_conv.conv"))
        fused = self.pixelweight_attention(pre + ".pixelweight_attention2", up, skip)
        return self.res_block(pre + ".up_addconv_block2", fused, cout, cout)

    def window_attention(self, pre: str, x, grid, mode: int, out=None):
        """Residual(MultiAxisAttention) (hybrid_CTUNet.py:481-511); x: [T, D] residual stream, updated in place
        unless `out` is given."""
        T, D = x.shape
        out = x if out is None else out
        h = self._empty(T, D)
        ops.layernorm(x, self.w.f32(pre + ".norm.weight"), self.w.f32(pre + ".norm.bias"), h)
        qkv = self._empty(T, 3 * D)
        ops.gemm(h, self.w.linear(pre + ".to_qkv", bias=False), qkv, dims=(T, 1, 1, 1))
        ops.attention(qkv, h, dim_head=32, n=216, mode=mode, bias=self.w.rel_bias(pre + ".rel_pos_bias"), grid=grid, w=6)
        ops.gemm(h, self.w.linear(pre + ".to_out.0", bias=False), out, dims=(T, 1, 1, 1),
                 out_mode=OUT_F32 if out.dtype == torch.float32 else OUT_BF16, residual=x)
        return out

    def up_attention_block(self, pre: str, tokens, B: int, grid0, out_last=None):
        """UpAttentionBlock (hybrid_CTUNet.py:554-591); tokens: [B*X*Y*Z, 768] in (x,y,z) order = proj_feat view.
        Returns the four up-sampled stage outputs as channels-last bf16 maps."""
        feats = []
        x = tokens
        X, Y, Z = grid0
        for ind in range(4):
            p = f"{pre}layers.{ind}.0"
            f = DS_STRIDE[::-1][ind]
            T, D = x.shape
            xb = self._empty(T, D)  # bf16 stage output feeding the pixel-shuffle GEMM
            # the stage input is also a returned feature map (or the caller's tokens): the first residual update
            # goes to a fresh buffer, later ones are in place
            fresh = self._empty(T, D, dtype=x.dtype)
            if ind <= 2:
                x = self.window_attention(p + ".1.fn", x, (B, X, Y, Z), 1, out=fresh)
                self.ffn(p + ".2.fn", x)
                self.window_attention(p + ".5.fn", x, (B, X, Y, Z), 2)
                self.ffn(p + ".6.fn", x, out=xb)
                ps = self.w.pixel_shuffle(p + ".8.to_out", f)
            else:
                x = self.ffn(p + ".1.fn", x, out=fresh)
                self.ffn(p + ".2.fn", x, out=xb)
                ps = self.w.pixel_shuffle(p + ".4.to_out", f)
            out = out_last if (ind == 3 and out_last is not None) else None
            y = self.up_gemm(xb.view(B, X, Y, Z, D), ps, out=out)
            feats.append(y if out is None else out)
            X, Y, Z = X * f[0], Y * f[1], Z * f[2]
            x = y.reshape(B * X * Y * Z, -1) if out is None else None
        return feats

    # ------------------------------------------------------------------ whole networks
    def _vit_branch(self, x_in, pf: int, depth: int, heads: int):
        B, _, X, Y, Z = x_in.shape
        tokens, n = self.vit("vit.", x_in, pf, depth, heads)
        cat = self._empty(B, X, Y, Z, 128)  # torch.cat((vit_enc_96x96, vit_enc0), dim=1) built in place
        self.res_block_cin1("vit_encoder0.layer", x_in, out=cat[..., 64:])
        enc = self.up_attention_block("vit_encoder.", tokens, B, (X // 16, Y // 16, Z // pf), out_last=cat[..., :64])
        vit_out = self.res_block("vit_decoder0.conv_block", cat, 128, 64)
        vit_logits = self.head(vit_out, self.w.conv1("vit_out.conv.conv", bias=True))
        vit_96 = self.head(cat[..., :64], self.w.linear("decoder_linear_96x96.head"), a_c=64)
        return enc, vit_logits, vit_96

    def ctunet(self, x_in, layers, pf: int, depth: int = 12, heads: int = 12):
        """CTUNet.forward (hybrid_CTUNet.py:817-857)."""
        self.stats.reset()
        enc, vit_logits, vit_96 = self._vit_branch(x_in, pf, depth, heads)
        res = self.resnet("convnet.", x_in, layers)
        dec3 = self.up_2fusion("res_decoder3", res[3], res[2], enc[0], 512)
        dec2 = self.up_2fusion("res_decoder2", dec3, res[1], enc[1], 256)
        dec1 = self.up_2fusion("res_decoder1", dec2, res[0], enc[2], 128)
        up0 = self.up_gemm(dec1, self.w.convt("res_decoder0.transp_conv.conv"))
        res_out = self.res_block("res_decoder0.conv_block", up0, 64, 64)
        res_logits = self.head(res_out, self.w.conv1("res_out.conv.conv", bias=True))
        res_48 = self.head(dec1, self.w.conv1("res_out_48x48.conv.conv", bias=True))
        res_24 = self.head(dec2, self.w.conv1("res_out_24x24.conv.conv", bias=True))
        return ((res_logits, res_48, res_24), (vit_logits, vit_96))

    def tunet(self, x_in, pf: int, depth: int = 12, heads: int = 12):
        """TUNet.forward (hybrid_CTUNet.py:1021-1036)."""
        self.stats.reset()
        _, vit_logits, vit_96 = self._vit_branch(x_in, pf, depth, heads)
        return (vit_logits, vit_96)

    def up_cat_conv(self, pre: str, inp, skip, cout: int):
        """UpCatConvBlock (hybrid_CTUNet.py:196-201): ConvT -> cat(skip) -> ResBlock(2C -> C)."""
        B, X, Y, Z, _ = skip.shape
        cat = self._empty(B, X, Y, Z, 2 * cout)
        self.up_gemm(inp, self.w.convt(pre + ".transp_conv.conv"), out=cat[..., :cout])
        ops.subsample(skip, cat[..., cout:], (1, 1, 1))  # stride-1 gather == strided copy into the concat buffer
        return self.res_block(pre + ".conv_block", cat, 2 * cout, cout)

    def cunet(self, x_in, layers):
        """CUNet.forward (hybrid_CTUNet.py:919-937)."""
        self.stats.reset()
        res = self.resnet("convnet.", x_in, layers)
        dec3 = self.up_cat_conv("res_decoder3", res[3], res[2], 512)
        dec2 = self.up_cat_conv("res_decoder2", dec3, res[1], 256)
        dec1 = self.up_cat_conv("res_decoder1", dec2, res[0], 128)
        up0 = self.up_gemm(dec1, self.w.convt("res_decoder0.transp_conv.conv"))
        res_out = self.res_block("res_decoder0.conv_block", up0, 64, 64)
        return (self.head(res_out, self.w.conv1("res_out.conv.conv", bias=True)),
                self.head(dec1, self.w.conv1("res_out_48x48.conv.conv", bias=True)),
                self.head(dec2, self.w.conv1("res_out_24x24.conv.conv", bias=True)))
