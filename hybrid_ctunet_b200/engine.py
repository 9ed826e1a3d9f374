"""Forward / backward engine of the CTUNet path on the sm_100a kernels.

Everything between the module boundary (fp32 NCDHW in, fp32 NCDHW logits out) runs here on channels-last bf16
activations; the einops rearranges of the reference (window / grid partition, proj_feat, pixel shuffle, token
<-> volume views) are folded into kernel indexing, so no permute copies exist on this path.

Reference call graph this engine restates (hybrid_CTUNet.py:817-857): vit -> vit_encoder0 -> vit_encoder ->
vit_decoder0 -> heads; convnet -> res_decoder3..0 -> heads.  Block functions cite the reference per function.

Training (trainer_CTUNet.py:87-109): with `Engine.tape` set, every primitive records a closure that computes its
input / parameter gradients with the backward kernels (ctu_umma_wgrad, dgrad through ctu_umma_gemm with transposed
weights, ctu_in_bwd_*, ctu_layernorm_bwd, ctu_attention_bwd, ...); `Engine.backward` replays the closures in
reverse.  Activation gradients are bf16 channels-last (fp32 along the fp32 token residual streams), weight
gradients accumulate in fp32.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch

from . import ops
from .ops import ACT_GELU, ACT_NONE, OUT_BF16, OUT_F32, OUT_F32_CF, PackedWeight

_FUSED_FFN = os.environ.get("CTU_FUSED_FFN", "1") != "0"   # (0: the two-GEMM inference FFN, for A/B comparisons)
DS_STRIDE = ((2, 2, 1), (2, 2, 2), (2, 2, 2), (2, 2, 2))
BF16 = torch.bfloat16
F32 = torch.float32


def rel_pos_index(w: int) -> torch.Tensor:
    """Index table of MultiAxisAttention (hybrid_CTUNet.py:472-477): [w^3, w^3] into the (2w-1)^3 embedding."""
    pos = torch.arange(w)
    g = torch.stack(torch.meshgrid(pos, pos, pos, indexing="ij")).reshape(3, -1).t()
    rel = g[:, None, :] - g[None, :, :] + (w - 1)
    return (rel * torch.tensor([(2 * w - 1) ** 2, 2 * w - 1, 1])).sum(-1)


_REL_INDEX_CACHE: Dict[tuple, torch.Tensor] = {}


def rel_pos_index_on(w: int, device) -> torch.Tensor:
    """rel_pos_index(w) resident on `device` (cached: no host-to-device copy inside a CUDA-graph capture)."""
    key = (w, str(device))
    t = _REL_INDEX_CACHE.get(key)
    if t is None:
        t = rel_pos_index(w).to(device)
        _REL_INDEX_CACHE[key] = t
    return t


def _pad64(v: int) -> int:
    return max(v, 64)


PACK_LIN, PACK_LIN_T, PACK_CONV3, PACK_CONV3_T, PACK_CONVT, PACK_CONVT_T, PACK_PS, PACK_PS_T, PACK_CIN1, PACK_PS_BIAS, \
    PACK_VEC, PACK_PAIR_LIN, PACK_PAIR_LIN_T, PACK_PAIR_CONV3, PACK_PAIR_CONV3_T, PACK_X3_FROM_PACKED = range(16)
# ResNet layer 1 (planes = 32, resnet.py:181-186) on "paired" rows: two z-neighbouring voxels per dense 64-channel row
# instead of 32 live + 32 zero-padded channels per voxel (0: the zero-padded path, for A/B comparisons)
_PAIR_L1 = os.environ.get("CTU_PAIR_L1", "1") != "0"
# 64-output-channel 3x3x3 convolutions on the kernel that computes two x-planes per tile (needs a re-laid weight copy)
_HALO_X2 = os.environ.get("CTU_CONV_HALO_X2", "1") != "0"
_TUNET_LANES = os.environ.get("CTU_TUNET_LANES", "1") != "0"
_CIN1_TC = os.environ.get("CTU_CIN1_TC", "1") != "0"   # vit_encoder0 conv1 (1 -> 64, k3) on the tensor cores via im2col
_ITEM_DTYPE = np.dtype([("src", "u8"), ("dst", "u8"), ("kind", "i4"), ("rows", "i4"), ("cols", "i4"), ("a", "i4"),
                        ("b", "i4"), ("c", "i4"), ("unit0", "i8")])


class ItemTable:
    """Device-resident table of ctu_pack_item for the multi-tensor pack / unpack kernels (include/ctunet_b200.h)."""

    def __init__(self, device, unpack: bool = False):
        self.dev = device
        self.unpack = unpack
        self.rows: List[tuple] = []          # (src_ptr, dst_ptr, kind, rows, cols, a, b, c, n_tasks)
        self.table: Optional[torch.Tensor] = None
        self.units = 0

    def add(self, src_ptr, dst_ptr, kind, rows, cols, a, b, c) -> int:
        n_tasks = int(ops._lib.load().ctu_pack_item_tasks(int(self.unpack), int(kind), int(rows), int(cols), int(a), int(b), int(c)))
        self.rows.append((int(src_ptr), int(dst_ptr), int(kind), int(rows), int(cols), int(a), int(b), int(c), n_tasks))
        self.table = None
        return len(self.rows) - 1

    def _build(self, rows):
        arr = np.zeros(len(rows), dtype=_ITEM_DTYPE)
        unit = 0
        for i, (src, dst, kind, r, c_, a, b, c, n) in enumerate(rows):
            arr[i] = (src, dst, kind, r, c_, a, b, c, unit)
            unit += -(-n // 256)
        t = torch.from_numpy(arr.view(np.uint8).copy()).to(self.dev)
        return t, unit

    def device_table(self):
        if self.table is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("the weight / gradient item table changed during CUDA-graph capture; run one eager "
                                   "warm-up step first")
            self.table, self.units = self._build(self.rows)
        return self.table, self.units

    def run(self, entry: str, only: Optional[int] = None):
        lib = ops._lib.require_device()
        if only is not None:
            t, units = self._build([self.rows[only]])
            n = 1
        else:
            if not self.rows:
                return
            t, units = self.device_table()
            n = len(self.rows)
        ops.check(getattr(lib, entry)(t.data_ptr(), n, units, torch.cuda.current_stream().cuda_stream), entry)


class WeightCache:
    """bf16 kernel-layout copies of the module's fp32 parameters (forward and transposed / tap-flipped for the
    input-gradient GEMMs).  The copies live in persistent buffers filled by ONE multi-tensor kernel launch
    (ctu_pack_weights): `refresh_all()` re-packs everything (every training forward, every train()/eval() switch —
    fused optimizers update parameters without bumping Tensor._version), a single stale entry is re-packed when its
    parameter's version / storage changed."""

    def __init__(self, params: Dict[str, torch.Tensor]):
        self.params = params
        self._cache: Dict[str, list] = {}      # key -> [tag, value, item index or None, names, bias repeat(, build)]
        self.items: Optional[ItemTable] = None
        self.items2: Optional[ItemTable] = None   # second pass: re-laid copies of packed weights (CTU_PACK_X3_FROM_PACKED)
        self.storage_epoch = 0                  # bumped whenever a cached buffer's storage is replaced: CUDA graphs
                                                # captured over the old buffers are stale from then on

    def _tag(self, names):
        ps = [self.params[n.lstrip('.')] for n in names]
        return ps, tuple((p.data_ptr(), p._version, str(p.device)) for p in ps)

    def _get(self, key: str, names, build):
        """torch-built entries (a handful of small tables): rebuilt when a parameter changes."""
        ps, tag = self._tag(names)
        hit = self._cache.get(key)
        if hit is not None and hit[0] == tag:
            return hit[1]
        with torch.no_grad():
            val = build(*[p.detach() for p in ps])
            if hit is not None and hit[1].shape == val.shape and hit[1].dtype == val.dtype and hit[1].device == val.device \
                    and hit[1].data_ptr() != val.data_ptr():
                # keep the buffer a captured CUDA graph may hold the address of: new values, same storage
                hit[1].copy_(val)
                val = hit[1]
        self._cache[key] = [tag, val, None, list(names), 1, build]
        return val

    def _packed(self, key: str, wname: str, kind: int, n: int, k: int, *, a: int, b: int, c: int = 0, ksize: int = 1,
                a_c: Optional[int] = None, block_n: Optional[int] = None, convt=None, bias_name: Optional[str] = None,
                bias_repeat: int = 1) -> PackedWeight:
        """bf16 [n_pad, k_pad] matrix produced by the pack kernel from parameter `wname` with layout map `kind`."""
        names = [wname] + ([bias_name] if bias_name else [])
        ps, tag = self._tag(names)
        hit = self._cache.get(key)
        if hit is not None and hit[0] == tag:
            return hit[1]
        w = ps[0].detach()
        if self.items is None:
            self.items = ItemTable(w.device)
        if hit is not None and hit[0][0][0] == tag[0][0] and hit[1].w.device == w.device:
            pw, idx = hit[1], hit[2]          # same storage, new version: re-pack in place
        else:
            if hit is not None:
                self.storage_epoch += 1
            bn = block_n or ops.pick_block_n(n)
            n_pad, k_pad = -(-n // bn) * bn, -(-k // 8) * 8
            buf = torch.empty((n_pad, k_pad), dtype=BF16, device=w.device)
            bias = None
            if bias_name:
                bias = ps[1].detach().to(F32).contiguous() if bias_repeat == 1 else ps[1].detach().to(F32).repeat(bias_repeat)
            pw = PackedWeight(buf, n, a_c if a_c is not None else k_pad // (ksize ** 3), ksize, bn, convt, bias)
            # algorithmic work of one GEMM row with the TRUE channel counts (bench.py's per-class roofline)
            pw.alg_flops_per_row = 2.0 * a * b * (27 if kind in (PACK_CONV3, PACK_CONV3_T) else max(c, 1))
            idx = self.items.add(w.data_ptr(), buf.data_ptr(), kind, n_pad, k_pad, a, b, c)
            if _HALO_X2 and ksize == 3 and bn == 64 and pw.a_c % 64 == 0 and k_pad == 27 * pw.a_c:
                # 64-output-channel 3x3x3 layers: the copy the two-plane kernel reads (x-taps of a (y,z)-tap adjacent)
                if self.items2 is None:
                    self.items2 = ItemTable(w.device)
                rows3 = 9 * (n_pad // 64) * 192
                pw.x3 = torch.empty((rows3, pw.a_c), dtype=BF16, device=w.device)
                pw.x3_item = self.items2.add(buf.data_ptr(), pw.x3.data_ptr(), PACK_X3_FROM_PACKED, rows3, pw.a_c, n_pad,
                                             pw.a_c, 0)
        if bias_name and bias_repeat != 1:
            pw.bias.copy_(ps[1].detach().to(F32).repeat(bias_repeat))
        self.items.run("ctu_pack_weights", only=idx)
        if pw.x3 is not None:
            self.items2.run("ctu_pack_weights", only=pw.x3_item)
        self._cache[key] = [tag, pw, idx, names, bias_repeat]
        return pw

    def refresh_all(self):
        """Re-pack every registered weight from the parameters' current values with one launch.  Every packed copy and
        table keeps its storage (CUDA graphs captured over them stay valid and see the new values)."""
        for ent in self._cache.values():
            if ent[2] is not None and self._tag(ent[3])[1][0][0] != self.items.rows[ent[2]][0]:
                self.clear()  # a parameter's storage was replaced (e.g. module.to()): start over
                return
        if self.items is not None:
            self.items.run("ctu_pack_weights")
        if self.items2 is not None:
            self.items2.run("ctu_pack_weights")
        for key in list(self._cache):
            ent = self._cache[key]
            if ent[2] is None:
                # small torch-built tables (relative-position bias, Cin = 1 conv weights): rebuilt IN PLACE — an inference
                # CUDA graph captured earlier holds their addresses, so the storage must never be replaced
                if key.startswith(("d1:", "rb:", "rbT:")):
                    ps, tag = self._tag(ent[3])
                    with torch.no_grad():
                        val = ent[5](*[p.detach() for p in ps])
                        if val.data_ptr() != ent[1].data_ptr():
                            ent[1].copy_(val)
                    ent[0] = tag
                continue
            if ent[4] != 1:  # pixel-shuffle bias, repeated per sub-voxel
                ent[1].bias.copy_(self._p(ent[3][1]).detach().to(F32).repeat(ent[4]))
            ent[0] = self._tag(ent[3])[1]

    def clear(self):
        self._cache.clear()
        self.items = None
        self.items2 = None
        self.storage_epoch += 1

    # -- generic access by (kind, name, extra): forward packing and the packing of the dgrad GEMM
    def get(self, kind: str, name: str, extra=None) -> PackedWeight:
        if kind == "lin":
            return self.linear(name, bias=bool(extra))
        if kind == "conv1":
            return self.conv1(name, bias=bool(extra))
        if kind == "conv3":
            return self.conv3(name)
        if kind == "convt":
            return self.convt(name)
        if kind == "ps":
            return self.pixel_shuffle(name, extra)
        if kind == "cin1":
            return self.conv_cin1_tc(name)
        if kind == "pconv1":
            return self.pair_conv1(name)
        if kind == "pconv3":
            return self.pair_conv3(name)
        raise KeyError(kind)

    def _p(self, name: str) -> torch.Tensor:
        return self.params[name.lstrip(".")]

    def get_t(self, kind: str, name: str, extra=None) -> PackedWeight:
        """Weight of the input-gradient contraction dA = dOut (*) W^T."""
        w = self._p(name + ".weight")
        if kind == "lin":
            n, k = w.shape
            ncols = -(-n // 16) * 16 if (extra and n < 64) else n   # heads: 14 logits -> the 16-channel padded gradient
            return self._packed("linT:" + name, name + ".weight", PACK_LIN_T, k, ncols, a=n, b=k)
        if kind == "conv1":
            co, ci = w.shape[:2]
            kcols = -(-co // 16) * 16 if (extra and co < 64) else _pad64(co)   # biased 1x1x1 conv = logits head
            return self._packed("c1T:" + name, name + ".weight", PACK_LIN_T, _pad64(ci), kcols, a=co, b=ci)
        if kind == "conv3":
            co, ci = w.shape[:2]
            cop, cip = _pad64(co), _pad64(ci)
            pw = self._packed("c3T:" + name, name + ".weight", PACK_CONV3_T, cip, 27 * cop, a=co, b=ci, ksize=3, a_c=cop)
            pw.a_c_live = -(-co // 16) * 16   # the gradient rows are zero beyond the layer's true output channels
            return pw
        if kind == "convt":
            ci, co, kx, ky, kz = w.shape
            return self._packed("ctT:" + name, name + ".weight", PACK_CONVT_T, ci, kx * ky * kz * co, a=ci, b=co, c=kx * ky * kz)
        if kind == "ps":
            co, corg = w.shape
            k3 = extra[0] * extra[1] * extra[2]
            return self._packed("psT:" + name, name + ".weight", PACK_PS_T, corg * k3, k3 * co, a=co, b=corg, c=k3)
        if kind == "pconv1":
            co, ci = w.shape[:2]
            pw = self._packed("p1T:" + name, name + ".weight", PACK_PAIR_LIN_T, 2 * ci, 2 * co, a=co, b=ci)
            pw.alg_flops_per_row = 4.0 * co * ci
            return pw
        if kind == "pconv3":
            co, ci = w.shape[:2]
            pw = self._packed("p3T:" + name, name + ".weight", PACK_PAIR_CONV3_T, 2 * ci, 27 * 2 * co, a=co, b=ci, ksize=3,
                              a_c=2 * co)
            pw.alg_flops_per_row = 4.0 * co * ci * 27
            return pw
        raise KeyError(kind)

    # -- nn.Linear [N, K] (+bias)
    def linear(self, name: str, bias: bool = True, block_n: Optional[int] = None) -> PackedWeight:
        n, k = self._p(name + ".weight").shape
        return self._packed("lin:" + name, name + ".weight", PACK_LIN, n, k, a=n, b=k, block_n=block_n,
                            bias_name=(name + ".bias") if bias else None)

    # -- Conv3d 1x1x1 [Cout, Cin, 1,1,1]; channel counts below 64 are zero-padded to 64
    def conv1(self, name: str, bias: bool = False) -> PackedWeight:
        co, ci = self._p(name + ".weight").shape[:2]
        if bias:   # logits heads: no padding of the 14 output channels
            return self._packed("c1:" + name, name + ".weight", PACK_LIN, co, ci, a=co, b=ci, bias_name=name + ".bias")
        # feature convs: pad to the 64-channel granularity of the activation buffers
        return self._packed("c1:" + name, name + ".weight", PACK_LIN, _pad64(co), _pad64(ci), a=co, b=ci)

    # -- Conv3d 3x3x3 [Cout, Cin, 3,3,3] -> [Cout, 27*Cin] tap-major
    def conv3(self, name: str) -> PackedWeight:
        co, ci = self._p(name + ".weight").shape[:2]
        cop, cip = _pad64(co), _pad64(ci)
        pw = self._packed("c3:" + name, name + ".weight", PACK_CONV3, cop, 27 * cip, a=co, b=ci, ksize=3, a_c=cip)
        pw.a_c_live = -(-ci // 16) * 16       # activation rows are zero beyond the layer's true input channels
        return pw

    # -- the same two layers on paired rows (two z-neighbours per row): block-diagonal [2 Cout, 2 Cin] and the pair-tap
    #    3x3x3 [2 Cout, 27 * 2 Cin] (CTU_PACK_PAIR_* in include/ctunet_b200.h); FLOPs per (paired) row = two voxels
    def pair_conv1(self, name: str) -> PackedWeight:
        co, ci = self._p(name + ".weight").shape[:2]
        pw = self._packed("p1:" + name, name + ".weight", PACK_PAIR_LIN, 2 * co, 2 * ci, a=co, b=ci)
        pw.alg_flops_per_row = 4.0 * co * ci
        return pw

    def pair_conv3(self, name: str) -> PackedWeight:
        co, ci = self._p(name + ".weight").shape[:2]
        pw = self._packed("p3:" + name, name + ".weight", PACK_PAIR_CONV3, 2 * co, 27 * 2 * ci, a=co, b=ci, ksize=3, a_c=2 * ci)
        pw.alg_flops_per_row = 4.0 * co * ci * 27
        return pw

    # -- ConvTranspose3d kernel == stride [Cin, Cout, kX, kY, kZ]
    def convt(self, name: str) -> PackedWeight:
        ci, co, kx, ky, kz = self._p(name + ".weight").shape
        return self._packed("ct:" + name, name + ".weight", PACK_CONVT, kx * ky * kz * co, ci, a=ci, b=co, c=kx * ky * kz,
                            block_n=64 if co % 128 else 128, convt=(co, kz, ky, kx))

    # -- PixelShuffle + Linear (hybrid_CTUNet.py:404-432) as a transposed-conv-shaped GEMM (block-diagonal weight)
    def pixel_shuffle(self, name: str, factor) -> PackedWeight:
        co, corg = self._p(name + ".weight").shape
        fx, fy, fz = factor
        k3 = fx * fy * fz
        return self._packed("ps:" + name, name + ".weight", PACK_PS, k3 * co, corg * k3, a=co, b=corg, c=k3,
                            block_n=64 if co % 128 else 128, convt=(co, fz, fy, fx), bias_name=name + ".bias",
                            bias_repeat=k3)

    # -- single-input-channel convs on CUDA cores: fp32 [taps, 64]
    def conv_cin1(self, name: str) -> torch.Tensor:
        return self._get("d1:" + name, [name + ".weight"],
                         lambda w: w.reshape(w.shape[0], -1).t().contiguous().float())

    # -- the same weights as a [Cout, taps padded to 64n] bf16 matrix for the tensor-core path over an im2col operand
    def conv_cin1_tc(self, name: str) -> PackedWeight:
        w = self._p(name + ".weight")
        co, taps = w.shape[0], w[0].numel()
        return self._packed("d1tc:" + name, name + ".weight", PACK_CIN1, co, -(-taps // 64) * 64, a=co, b=taps)

    def rel_bias(self, name: str, w: int = 6) -> torch.Tensor:
        def build(emb):
            idx = rel_pos_index_on(w, emb.device)
            return emb[idx].permute(2, 0, 1).contiguous().float()
        return self._get("rb:" + name, [name + ".weight"], build)

    def rel_bias_t(self, name: str, w: int = 6) -> torch.Tensor:
        """[heads][key][query] copy for the attention backward kernel."""
        return self._get("rbT:" + name, [name + ".weight"],
                         lambda emb: emb[rel_pos_index_on(w, emb.device)].permute(2, 1, 0).contiguous().float())

    def f32(self, name: str) -> torch.Tensor:
        return self._get("f:" + name, [name], lambda p: p.float().contiguous())


class StatsArena:
    """fp64 accumulators (InstanceNorm statistics of one forward / reduction sums of one backward), zeroed with a
    single memset."""

    def __init__(self, device, capacity: int = 1 << 18):
        self.buf = torch.zeros(capacity, dtype=torch.float64, device=device)
        self.off = 0

    def reset(self):
        self.buf.zero_()
        self.off = 0

    def take(self, B: int, C: int, width: int = 2) -> torch.Tensor:
        n = B * C * width
        if self.off + n > self.buf.numel():
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("InstanceNorm statistics arena exhausted during CUDA-graph capture (the eager warm-up "
                                   "call sizes it: run it with the same batch size)")
            # grow: slices handed out so far keep the old storage alive (it was zeroed by reset()); from the next
            # reset() on, everything comes from the larger buffer
            self.buf = torch.zeros(max(2 * self.buf.numel(), 2 * n), dtype=torch.float64, device=self.buf.device)
            self.off = 0
        t = self.buf[self.off:self.off + n].view(B, C, width)
        self.off += n
        return t


class GradArena:
    """fp32 accumulators for every parameter gradient of one backward pass (the tensor-core wgrad kernel and the
    column-sum kernel accumulate with reductions), zeroed with a single memset."""

    def __init__(self, device, capacity: int):
        self.buf = torch.zeros(capacity, dtype=F32, device=device)
        self.off = 0

    def reset(self):
        self.buf.zero_()
        self.off = 0

    def take(self, *shape) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= int(s)
        n_al = -(-n // 64) * 64
        if self.off + n_al > self.buf.numel():
            raise RuntimeError("parameter-gradient arena exhausted")
        t = self.buf[self.off:self.off + n].view(*shape)
        self.off += n_al
        return t


def _key(t: torch.Tensor):
    return (t.data_ptr(), int(t.shape[-1]), int(t.stride(-2)) if t.dim() >= 2 else 0, t.numel())


class Tape:
    """Closures of one training forward, gradients of its activations, and the parameter-gradient records."""

    def __init__(self):
        self.fns: List[Callable[[], None]] = []
        self.grads: Dict[tuple, torch.Tensor] = {}
        self.alias: Dict[tuple, Tuple[torch.Tensor, int]] = {}
        self.reshaped: Dict[tuple, torch.Tensor] = {}   # key of a reshaped view -> the tensor that owns the gradient
        self.wrecs: List[tuple] = []  # (kind, name, buffer, meta)


class Engine:
    def __init__(self, params: Dict[str, torch.Tensor], device):
        self.w = WeightCache(params)
        self.dev = device
        self.stats = StatsArena(device)
        self.tape: Optional[Tape] = None
        self.bsums: Optional[StatsArena] = None
        self.garena: Optional[GradArena] = None
        # CTUNet's ViT branch and ResNet encoder are independent until res_decoder3 (hybrid_CTUNet.py:821-838): lane 1
        # (a side stream) runs the ViT branch while lane 0 (the caller's stream) runs the encoder — both contain long
        # runs of small kernels that leave most SMs idle on their own
        self.side: Optional[torch.cuda.Stream] = None
        self.lane = 0
        self.two_lanes = os.environ.get("CTU_TWO_LANES", "1") != "0"
        self.three_lanes = os.environ.get("CTU_THREE_LANES", "0") != "0"   # vit_encoder0 on a stream of its own
        self.side2: Optional[torch.cuda.Stream] = None
        self.wg_stream: Optional[List[torch.cuda.Stream]] = None   # parameter-gradient kernels of the small GEMMs (_off_path)
        self._wg_used = False
        self._wg_next = 0
        self.off_path_streams = max(1, int(os.environ.get("CTU_OFF_PATH_STREAMS", "1")))
        self.off_path_bytes = int(os.environ.get("CTU_OFF_PATH_MB", "16")) << 20
        # SMs a persistent tensor-core kernel may take while the two lanes run side by side (0: all of them)
        self.lane_sms = int(os.environ.get("CTU_LANE_SMS", "0"))

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

    # ------------------------------------------------------------------ tape plumbing (training only)
    def begin_training_forward(self):
        # a training forward always re-packs the weights: optimizers that update parameters through fused multi-tensor
        # kernels (torch.optim.AdamW(fused=True)) do not bump Tensor._version, which is what WeightCache keys on
        self.w.refresh_all()
        self.tape = Tape()
        self.stats = StatsArena(self.dev, 1 << 20)  # owned by this tape: the backward reads the forward's statistics
        if self.bsums is None:
            self.bsums = StatsArena(self.dev, 1 << 21)
        if self.garena is None:
            n = sum(p.numel() for p in self.w.params.values())
            self.garena = GradArena(self.dev, int(n * 1.15) + (32 << 20))

    def prepare_for_capture(self):
        """Build the device item tables of the pack / unpack kernels now (a host-to-device copy is not allowed while a
        CUDA graph is being captured)."""
        if self.w.items is not None:
            self.w.items.device_table()
        if self.w.items2 is not None:
            self.w.items2.device_table()
        if getattr(self, "_gtable", None) is not None:
            self._gtable.device_table()

    def _alias(self, view: torch.Tensor, base: torch.Tensor, c0: int):
        if self.tape is not None:
            self.tape.alias[_key(view)] = (base, c0)

    def _reshaped(self, view: torch.Tensor, base: torch.Tensor):
        """`view` is a contiguous reshape of `base` (same memory, other row width): their gradient is ONE buffer, kept
        under `base`."""
        if self.tape is not None:
            self.tape.reshaped[_key(view)] = base

    def _g(self, t: torch.Tensor) -> Optional[torch.Tensor]:
        """Gradient that has arrived for activation `t` (None if no consumer produced one)."""
        k = _key(t)
        rb = self.tape.reshaped.get(k)
        if rb is not None:
            g = self._g(rb)
            return None if g is None else g.view(t.shape)
        al = self.tape.alias.get(k)
        if al is not None:
            base, c0 = al
            bg = self.tape.grads.get(_key(base))
            return None if bg is None else bg.view(base.shape)[..., c0:c0 + t.shape[-1]]
        g = self.tape.grads.get(k)
        if g is not None and g.shape != t.shape and g.is_contiguous():
            g = g.view(t.shape)
        return g

    def _g16(self, t: torch.Tensor) -> Optional[torch.Tensor]:
        g = self._g(t)
        if g is not None and g.dtype != BF16:
            g16 = self._empty(*g.shape)
            ops.cast_f32_bf16(g, g16)
            return g16
        return g

    def _acc(self, t: torch.Tensor, g: torch.Tensor):
        """Add gradient `g` (freshly computed, ownership passes to the tape) to activation `t`."""
        k = _key(t)
        rb = self.tape.reshaped.get(k)
        if rb is not None:
            self._acc(rb, g.view(rb.shape))
            return
        al = self.tape.alias.get(k)
        if al is not None:
            base, c0 = al
            bk = _key(base)
            bg = self.tape.grads.get(bk)
            if bg is None:
                bg = torch.zeros(base.shape, dtype=base.dtype, device=self.dev)
                self.tape.grads[bk] = bg
            ops.accumulate(g, bg.view(base.shape)[..., c0:c0 + t.shape[-1]])
            return
        cur = self.tape.grads.get(k)
        if cur is None:
            if g.dtype != t.dtype or not g.is_contiguous():
                buf = torch.zeros(t.shape, dtype=t.dtype, device=self.dev)
                ops.accumulate(g, buf)
                g = buf
            self.tape.grads[k] = g
        else:
            ops.accumulate(g, cur)

    def _done(self, t: torch.Tensor):
        if _key(t) in self.tape.reshaped:
            return
        if _key(t) not in self.tape.alias:
            self.tape.grads.pop(_key(t), None)

    def _rec(self, fn):
        if self.tape is not None:
            self.tape.fns.append((self.lane, fn))

    def _off_path(self, fn, tensors):
        """Run `fn` (parameter-gradient kernels: nothing downstream of them until the gradients are unpacked) off the
        dependency chain of the backward pass.  While a CUDA graph is being captured and every tensor involved is small
        (kernels that cannot fill the GPU: the ViT at 864 tokens, the deep encoder stages), they go to a third stream
        forked from the current one and joined before the gradients are unpacked; otherwise `fn` runs in place."""
        small = all(t.numel() * t.element_size() <= self.off_path_bytes for t in tensors)
        if not (self.two_lanes and small and torch.cuda.is_current_stream_capturing()):
            fn()
            return
        if self.wg_stream is None:
            self.wg_stream = [torch.cuda.Stream(device=self.dev) for _ in range(self.off_path_streams)]
        wg = self.wg_stream[self._wg_next % len(self.wg_stream)]
        self._wg_next += 1
        cur = torch.cuda.current_stream()
        wg.wait_stream(cur)
        for t in tensors:  # not recycled by the capture's allocator before the join
            t.record_stream(wg)
        with torch.cuda.stream(wg):
            fn()
        self._wg_used = True

    def _set_lane_sms(self, on: bool):
        if self.lane_sms > 0:
            ops._lib.require_device().ctu_set_persistent_sm_limit(self.lane_sms if on else 0)

    def _side2_stream(self) -> torch.cuda.Stream:
        if self.side2 is None:
            self.side2 = torch.cuda.Stream(device=self.dev)
        return self.side2

    def _side_stream(self) -> torch.cuda.Stream:
        if self.side is None:
            # (a higher priority for this, the longer lane, was measured and does not help: 58.4 vs 57.2 ms per step)
            prio = int(os.environ.get("CTU_SIDE_PRIORITY", "0"))
            self.side = torch.cuda.Stream(device=self.dev, priority=prio)
        return self.side

    def backward(self, out_grads: List[Tuple[torch.Tensor, Optional[torch.Tensor]]], want=()):
        """out_grads: (forward output tensor, its gradient or None).  Returns ({parameter name: fp32 gradient},
        [gradient of each activation in `want`])."""
        tape = self.tape
        if tape is None:
            raise RuntimeError("backward() without a recorded training forward")
        self.bsums.reset()
        self.garena.reset()
        for t, g in out_grads:
            if g is not None:
                tape.grads[_key(t)] = g.contiguous()
        main = torch.cuda.current_stream()
        used_side = False
        used_side2 = False
        for lane, fn in reversed(tape.fns):
            if fn is None and lane == 0:  # forward join point: from here back, lane-1 closures run on the side stream
                side = self._side_stream()
                side.wait_stream(main)
                for g in tape.grads.values():  # gradients produced on `main` that lane 1 will read (and free)
                    g.record_stream(side)
                used_side = True
                self._set_lane_sms(True)
            elif fn is None:  # lane 2 joined lane 1 here in the forward: its closures fork off lane 1 from here back
                if used_side:
                    side2 = self._side2_stream()
                    side2.wait_stream(self.side)
                    for g in tape.grads.values():
                        g.record_stream(side2)
                    used_side2 = True
            elif lane == 2 and used_side2:
                with torch.cuda.stream(self.side2):
                    fn()
            elif lane in (1, 2) and used_side:
                with torch.cuda.stream(self.side):
                    fn()
            else:
                fn()
        if used_side2:
            main.wait_stream(self.side2)
        if used_side:
            main.wait_stream(self.side)
            self._set_lane_sms(False)
        if self._wg_used:
            for wg in self.wg_stream:
                main.wait_stream(wg)
            self._wg_used = False
        igrads = [None if a is None else self._g(a) for a in want]
        grads = self._finalize_param_grads(tape)
        self.tape = None
        return grads, igrads

    def _unpack_torch(self, kind, name, buf, meta) -> Tuple[str, torch.Tensor]:
        """Gradient accumulator -> parameter layout with torch ops (relative-position tables and repeated parameters;
        everything else goes through the multi-tensor kernel)."""
        P = lambda n: self.w.params[n.lstrip(".")]
        if kind in ("lin", "conv1"):
            p = P(name + ".weight")
            return name + ".weight", buf[:p[0].numel(), :p.shape[0]].t().reshape(p.shape)
        if kind == "conv3":
            co, ci = P(name + ".weight").shape[:2]
            return name + ".weight", buf.view(3, 3, 3, buf.shape[0] // 27, -1)[..., :ci, :co].permute(4, 3, 0, 1, 2).contiguous()
        if kind == "convt":
            ci, co, kx, ky, kz = P(name + ".weight").shape
            return name + ".weight", buf.view(ci, kx, ky, kz, co).permute(0, 4, 1, 2, 3).contiguous()
        if kind == "ps":
            co, corg = P(name + ".weight").shape
            k3 = buf.shape[0] // corg
            return name + ".weight", torch.einsum("csso->oc", buf.view(corg, k3, k3, co)).contiguous()
        if kind == "ps_bias":
            co = P(name + ".bias").shape[0]
            return name + ".bias", buf.view(-1, co).sum(0)
        if kind == "cin1":
            p = P(name + ".weight")
            return name + ".weight", buf[:p[0].numel(), :p.shape[0]].t().reshape(p.shape)
        if kind == "vec":
            p = P(name)
            return name, buf.reshape(-1)[:p.numel()].reshape(p.shape).clone()
        if kind == "relbias":  # buf [heads, key, query] -> embedding [(2w-1)^3, heads]
            p = P(name + ".weight")
            idx = rel_pos_index_on(meta, self.dev).reshape(-1)
            g = torch.zeros(p.shape, dtype=F32, device=self.dev)
            g.index_add_(0, idx, buf.permute(2, 1, 0).reshape(-1, p.shape[1]))
            return name + ".weight", g
        raise KeyError(kind)

    def _finalize_param_grads(self, tape: Tape) -> Dict[str, torch.Tensor]:
        """fp32 gradient accumulators (transposed-packed layouts of the wgrad / colsum kernels) -> gradients in the
        parameters' own layouts: one ctu_unpack_grads launch over a cached device item table."""
        P = lambda n: self.w.params[n.lstrip(".")]
        recs, slow, seen = [], [], set()
        for kind, name, buf, meta in tape.wrecs:
            pname = (name if kind == "vec" else name + (".bias" if kind == "ps_bias" else ".weight")).lstrip(".")
            if kind == "relbias" or pname in seen:
                slow.append((kind, name, buf, meta))
                continue
            seen.add(pname)
            p = P(pname)
            ld = int(buf.shape[-1])
            if kind in ("lin", "conv1"):
                item = (PACK_LIN, p.shape[0], p[0].numel(), 0)
            elif kind == "conv3":
                item = (PACK_CONV3, p.shape[0], p.shape[1], buf.shape[0] // 27)
            elif kind == "convt":
                item = (PACK_CONVT, p.shape[0], p.shape[1], p[0, 0].numel())
            elif kind == "ps":
                item = (PACK_PS, p.shape[0], p.shape[1], buf.shape[0] // p.shape[1])
            elif kind == "ps_bias":
                item = (PACK_PS_BIAS, p.shape[0], 0, buf.numel() // p.shape[0])
            elif kind == "cin1":
                item = (PACK_CIN1, p.shape[0], p[0].numel(), 0)
            elif kind == "pconv1":
                item = (PACK_PAIR_LIN, p.shape[0], p.shape[1], 0)
            elif kind == "pconv3":
                item = (PACK_PAIR_CONV3, p.shape[0], p.shape[1], 0)
            else:
                item = (PACK_VEC, 0, 0, 0)
            recs.append((pname, buf, ld, item))
        sig = tuple((pn, buf.data_ptr(), it) for pn, buf, _, it in recs)
        gflat = getattr(self, "_gflat", None)
        if gflat is not None and not torch.cuda.is_current_stream_capturing():
            # The gradients handed to autograd are views of the persistent flat buffer, and AccumulateGrad keeps such a
            # view as `.grad` without copying.  If a parameter still holds one (gradient accumulation, zero_grad(
            # set_to_none=False), a module called twice in one graph), unpacking into the same buffer would overwrite
            # the accumulated value before autograd adds to it: leave that buffer to the parameters, take a fresh one.
            base = gflat.untyped_storage().data_ptr()
            for prm in self.w.params.values():
                g = getattr(prm, "grad", None)
                if g is not None and g.untyped_storage().data_ptr() == base:
                    self._gsig = None
                    break
        if getattr(self, "_gsig", None) != sig:
            # every gradient starts on a 16-byte boundary of the flat buffer (vector stores in ctu_unpack_grads,
            # vector loads in ctu_adamw_step)
            total = sum(-(-P(pn).numel() // 4) * 4 for pn, _, _, _ in recs)
            self._gflat = torch.empty(total, dtype=F32, device=self.dev)
            self._gtable = ItemTable(self.dev, unpack=True)
            self._gviews = []
            off = 0
            for pn, buf, ld, (code, a_, b_, c_) in recs:
                n = P(pn).numel()
                self._gtable.add(buf.data_ptr(), self._gflat.data_ptr() + 4 * off, code, n, ld, a_, b_, c_)
                self._gviews.append((pn, off, n))
                off += -(-n // 4) * 4
            self._gsig = sig
        out: Dict[str, torch.Tensor] = {}
        if recs:
            self._gtable.run("ctu_unpack_grads")
            for pn, off, n in self._gviews:
                out[pn] = self._gflat[off:off + n].view(P(pn).shape)
        for kind, name, buf, meta in slow:
            pn, g = self._unpack_torch(kind, name, buf, meta)
            pn = pn.lstrip(".")
            out[pn] = g if pn not in out else out[pn] + g
        return out

    # ------------------------------------------------------------------ primitives (forward + recorded backward)
    def gemm(self, a, kind: str, name: str, out, *, dims, extra=None, stats=None, act=ACT_NONE, residual=None,
             out_mode=OUT_BF16, a_c=None, a_needs_grad: bool = True, a_gelu_of=None):
        """Tensor-core contraction through ctu_umma_gemm with the packed weight (kind, name).  `a_gelu_of`: `a` is
        gelu(a_gelu_of) computed without a tape record; the input-gradient GEMM then applies the GELU derivative in
        its epilogue and hands the gradient to the pre-activation directly."""
        pw = self.w.get(kind, name, extra)
        ac = int(a_c if a_c is not None else pw.a_c)
        ops.gemm(a, pw, out, dims=dims, stats=stats, act=act, residual=residual, out_mode=out_mode, a_c=ac)
        if self.tape is None:
            return out
        assert act == ACT_NONE, "training keeps the pre-activation: use gelu()"

        def bw():
            g = self._g(out)
            if g is None:
                return
            if residual is not None:
                self._acc(residual, g)  # identity branch (g stays valid: nothing accumulates into it before we return)
            if out_mode == OUT_F32_CF:   # logits head: NCDHW fp32 -> channels-last bf16, zero-padded to 16 channels
                B, _, X, Y, Z = out.shape
                n_eff = -(-pw.n_real // 16) * 16   # 32-byte rows: the gradient GEMMs read 14 (+2) channels, not 64
                g16 = self._empty(B, X, Y, Z, n_eff)
                ops.cf_to_cl(g.contiguous(), g16, n_eff)
            elif pw.convt is not None:   # up-sampling GEMM: gather the sub-voxels back into GEMM columns
                co, u1, u2, u3 = pw.convt
                B, Xo, Yo, Zo, _ = out.shape
                gv = g if g.dtype == BF16 else self._g16(out)
                g16 = self._empty(B, Xo // u3, Yo // u2, Zo // u1, u1 * u2 * u3 * co)
                ops.space_to_depth(gv, g16, (u3, u2, u1))
                n_eff = pw.n_real
            else:
                g16 = g if g.dtype == BF16 else self._g16(out)
                n_eff = pw.n_real
            ksize = pw.ksize
            # parameter gradients
            dw = self.garena.take(ksize ** 3 * ac, n_eff)
            db = self.garena.take(n_eff) if pw.bias is not None else None

            def param_grads():
                ops.wgrad(a, g16, dw, dims=dims, ksize=ksize, x_c=ac, n=n_eff, alg_flops_per_row=pw.alg_flops_per_row)
                if db is not None:
                    ops.colsum(g16.reshape(-1, g16.shape[-1]) if g16.is_contiguous() else g16, db, n=n_eff)
            # `g` was handed to the residual branch above as ITS gradient buffer, which later contributions are
            # accumulated into in place: kernels reading it must stay on the dependency chain
            if residual is not None and g16 is g:
                param_grads()
            else:
                self._off_path(param_grads, (a, g16))
            self.tape.wrecs.append((kind, name, dw, None))
            if db is not None:
                self.tape.wrecs.append(("ps_bias" if kind == "ps" else "vec", name if kind == "ps" else name + ".bias", db, None))
            # input gradient
            if a_needs_grad:
                pt = self.w.get_t(kind, name, extra)
                cur = self._g(a)
                if a_gelu_of is not None:
                    da = self._empty(*a_gelu_of.shape)
                    ops.gemm(g16, pt, da, dims=dims, a_c=n_eff, gelu_bwd_of=a_gelu_of)
                    self._acc(a_gelu_of, da)
                elif cur is not None and cur.dtype == BF16 and pw.convt is None and cur.shape[-1] == ac:
                    # a gradient has already arrived for `a` (e.g. through the residual connection): let the GEMM
                    # epilogue add it instead of running a separate accumulation pass (in place, row for row)
                    ops.gemm(g16, pt, cur, dims=dims, a_c=n_eff, residual=cur)
                else:
                    da = self._empty(*a.shape[:-1], ac)
                    ops.gemm(g16, pt, da, dims=dims, a_c=n_eff)
                    self._acc(a, da)
            self._done(out)
        self._rec(bw)
        return out

    def in_apply(self, x, st, *, res=None, rstats=None, out=None, fold: int = 0):
        """out = lrelu(IN(x) [+ res | + IN(res)]) (resnet.py:110-124; hybrid_CTUNet.py:95-104).  fold: `x` holds paired
        rows whose column halves [0, fold) and [fold, 2 fold) are the same channels (`st` already folded by the caller)."""
        if out is None:
            out = self._empty(*x.shape)
        ops.in_apply(x, st, out, res=res, rstats=rstats, act=True)
        if self.tape is not None:
            def bw():
                g = self._g(out)
                if g is None:
                    return
                B, C = x.shape[0], x.shape[-1]
                dx = self._empty(*x.shape)
                dres = self._empty(*res.shape) if res is not None else None
                ops.in_backward(g, out, x if res is not None else None, st, dx, res=res, rstats=rstats, dres=dres,
                                sums=self.bsums.take(B, C, 4), fold=fold)
                self._acc(x, dx)
                if res is not None:
                    self._acc(res, dres)
                self._done(out)
            self._rec(bw)
        return out

    def layernorm(self, x, name: str, out, *, add_name: Optional[str] = None):
        gamma, beta = self.w.f32(name + ".weight"), self.w.f32(name + ".bias")
        add = self.w.f32(add_name) if add_name else None
        ops.layernorm(x, gamma, beta, out, add=add)
        if self.tape is not None:
            def bw():
                dy = self._g(out)
                if dy is None:
                    return
                C = x.shape[-1]
                if add is not None:  # pos_embedding: sum of the output gradient over the batch
                    dadd = self.garena.take(add.numel())
                    rows = x.numel() // add.numel()
                    dyf = dy if dy.is_contiguous() else dy.contiguous()
                    ops.colsum(dyf.view(rows, add.numel()), dadd)
                    self.tape.wrecs.append(("vec", add_name, dadd, None))
                dy16 = dy if dy.dtype == BF16 else self._g16(out)
                dg, db = self.garena.take(C), self.garena.take(C)
                gx = self._g(x)
                x2 = x.reshape(-1, C) if x.is_contiguous() else x
                if gx is None:
                    dx = torch.empty(x.shape, dtype=x.dtype, device=self.dev)
                    ops.layernorm_backward(x2, gamma, dy16.reshape(-1, C) if dy16.is_contiguous() else dy16, dg, db,
                                           dx_f32=dx.view(-1, C) if dx.dtype == F32 else None,
                                           dx_bf16=dx.view(-1, C) if dx.dtype == BF16 else None)
                    self._acc(x, dx)
                else:  # join the gradient already flowing along the residual stream, in place
                    gx2 = gx.reshape(-1, C) if gx.is_contiguous() else gx
                    ops.layernorm_backward(x2, gamma, dy16.reshape(-1, C) if dy16.is_contiguous() else dy16, dg, db,
                                           dx_in=gx2, dx_f32=gx2 if gx.dtype == F32 else None,
                                           dx_bf16=gx2 if gx.dtype == BF16 else None)
                self.tape.wrecs.append(("vec", name + ".weight", dg, None))
                self.tape.wrecs.append(("vec", name + ".bias", db, None))
                self._done(out)
            self._rec(bw)
        return out

    def gelu(self, x):
        y = self._empty(*x.shape)
        ops.gelu(x, y)

        def bw():
            g = self._g(y)
            if g is None:
                return
            dx = self._empty(*x.shape)
            ops.gelu_backward(x, g, dx)
            self._acc(x, dx)
            self._done(y)
        self._rec(bw)
        return y

    def attention(self, qkv, out, *, dim_head, n, windows=0, mode=0, bias_name=None, grid=(1, 1, 1, 1)):
        bias = self.w.rel_bias(bias_name) if bias_name else None
        if self.tape is None:
            ops.attention(qkv, out, dim_head=dim_head, n=n, windows=windows, mode=mode, bias=bias, grid=grid, w=6)
            return out
        C = out.shape[-1]
        heads = C // dim_head
        lse = self._empty(qkv.shape[0], heads, dtype=F32)
        ops.attention(qkv, out, dim_head=dim_head, n=n, windows=windows, mode=mode, bias=bias, grid=grid, w=6, lse=lse)

        def bw():
            g = self._g(out)
            if g is None:
                return
            rows = qkv.shape[0]
            dqkv = self._empty(rows, 3 * C)
            direct = dim_head == 32 and n <= 224  # window attention: dQ is written straight into dqkv
            dq = None if direct else torch.zeros(rows, C, dtype=F32, device=self.dev)
            nwin = windows if mode == 0 else rows // n
            ds = bias_t = None
            if bias_name:
                bias_t = self.w.rel_bias_t(bias_name)
                ds = self._empty(nwin, heads, n, n)
            ops.attention_backward(qkv, out, g, lse, dqkv, dq, dim_head=dim_head, n=n, windows=windows, mode=mode,
                                   bias_t=bias_t, ds_out=ds, grid=grid, w=6)
            if dq is not None:
                ops.cast_f32_bf16(dq, dqkv[:, :C])
            if bias_name:
                dbt = self.garena.take(heads, n, n)
                ops.colsum(ds.view(nwin, heads * n * n), dbt.view(-1))
                self.tape.wrecs.append(("relbias", bias_name, dbt, 6))
            self._acc(qkv, dqkv)
            self._done(out)
        self._rec(bw)
        return out

    def pwa_fuse(self, q1, q2, out):
        ops.pwa_fuse(q1, q2, out)
        if self.tape is not None:
            def bw():
                g = self._g(out)
                if g is None:
                    return
                d1, d2 = self._empty(*q1.shape), self._empty(*q2.shape)
                ops.pwa_fuse_backward(q1, q2, g, d1, d2)
                self._acc(q1, d1)
                self._acc(q2, d2)
                self._done(out)
            self._rec(bw)
        return out

    def subsample(self, x, out, stride):
        ops.subsample(x, out, stride)
        if self.tape is not None:
            def bw():
                g = self._g(out)
                if g is None:
                    return
                gx = self._g(x)
                if gx is None:
                    dx = self._empty(*x.shape)
                    ops.subsample_backward(g, dx, stride)
                    self._acc(x, dx)
                else:
                    ops.subsample_backward(g, gx, stride, accumulate=True)
                self._done(out)
            self._rec(bw)
        return out

    def conv_cin1(self, x_in, name: str, out, *, k, s, p):
        ops.conv_cin1(x_in, self.w.conv_cin1(name), out, k=k, s=s, p=p)
        if self.tape is not None:
            def bw():
                g = self._g(out)
                if g is None:
                    return
                taps = k[0] * k[1] * k[2]
                kpad = -(-taps // 64) * 64
                B, Xo, Yo, Zo, _ = out.shape
                col = self._empty(B, Xo, Yo, Zo, kpad)
                ops.im2col_cin1(x_in, col, k=k, s=s, p=p)
                dw = self.garena.take(kpad, 64)
                ops.wgrad(col, g, dw, dims=(Zo, Yo, Xo, B), x_c=kpad, n=64, alg_flops_per_row=2.0 * taps * 64)
                self.tape.wrecs.append(("cin1", name, dw, None))
                self._done(out)
            self._rec(bw)
        return out

    def patchify_ln(self, x_in, pf: int, name: str, tok):
        gamma, beta = self.w.f32(name + ".weight"), self.w.f32(name + ".bias")
        ops.patchify_ln(x_in, pf, gamma, beta, tok)
        if self.tape is not None:
            def bw():
                g = self._g16(tok)
                if g is None:
                    return
                dg, db = self.garena.take(gamma.numel()), self.garena.take(gamma.numel())
                ops.patchify_ln_backward(x_in, pf, g, dg, db)
                self.tape.wrecs.append(("vec", name + ".weight", dg, None))
                self.tape.wrecs.append(("vec", name + ".bias", db, None))
                self._done(tok)
            self._rec(bw)
        return tok

    # ------------------------------------------------------------------ contractions by role
    def conv3x3(self, x, name: str, stats=None, out=None):
        pw = self.w.conv3(name)
        B, X, Y, Z, _ = x.shape
        if out is None:
            out = self._empty(B, X, Y, Z, pw.n_real)
        return self.gemm(x, "conv3", name, out, dims=self._dims(x), stats=stats)

    def conv1x1(self, x, name: str, stats=None, out=None):
        pw = self.w.conv1(name)
        B, X, Y, Z, _ = x.shape
        if out is None:
            out = self._empty(B, X, Y, Z, pw.n_real)
        return self.gemm(x, "conv1", name, out, dims=self._flat_dims(x), stats=stats)

    def up_gemm(self, x, kind: str, name: str, extra=None, out=None):
        """ConvTranspose3d(k=s) / pixel-shuffle+Linear: [B,X,Y,Z,Cin] -> [B,X*ux,Y*uy,Z*uz,Cout]."""
        pw = self.w.get(kind, name, extra)
        B, X, Y, Z, _ = x.shape
        co, uz, uy, ux = pw.convt
        if out is None:
            out = self._empty(B, X * ux, Y * uy, Z * uz, co)
        return self.gemm(x, kind, name, out, dims=self._dims(x), extra=extra)

    def head(self, x, kind: str, name: str, a_c=None):
        """UnetOutBlock / DecoderLinear: per-voxel C -> n_cls with bias, fp32 NCDHW output."""
        pw = self.w.get(kind, name, True)
        B, X, Y, Z, _ = x.shape
        out = self._empty(B, pw.n_real, X, Y, Z, dtype=F32)
        ac = int(a_c if a_c is not None else pw.a_c)
        if ac not in (64, 128, 256) or pw.n_real > 16:   # shapes outside CTUNet's heads: generic GEMM backward
            return self.gemm(x, kind, name, out, dims=self._flat_dims(x), extra=True, out_mode=OUT_F32_CF, a_c=a_c)
        ops.gemm(x, pw, out, dims=self._flat_dims(x), out_mode=OUT_F32_CF, a_c=ac)
        if self.tape is not None:
            def bw():
                # K = N = 14 contractions cannot fill a tensor-core tile: input, weight and bias gradients come from ONE
                # CUDA-core pass over the fp32 NCDHW logit gradient (ctu_head_bwd) instead of four launches over
                # 64-channel zero-padded copies of it
                g = self._g(out)
                if g is None:
                    return
                wparam = self.w._p(name + ".weight").detach()
                dw, db = self.garena.take(ac, 16), self.garena.take(16)
                cur = self._g(x)
                if cur is not None and cur.dtype == BF16 and cur.shape[-1] == ac:
                    ops.head_backward(g.contiguous(), x, wparam, cur, dw, db, accumulate=True)   # joins in place
                else:
                    da = self._empty(*x.shape[:-1], ac)
                    ops.head_backward(g.contiguous(), x, wparam, da, dw, db)
                    self._acc(x, da)
                self.tape.wrecs.append((kind, name, dw, None))
                self.tape.wrecs.append(("vec", name + ".bias", db, None))
                self._done(out)
            self._rec(bw)
        return out

    # ------------------------------------------------------------------ networks/resnet.py
    def bottleneck(self, pre: str, x, stride, has_down: bool):
        """resnet.py:106-126: 1x1 -> IN -> lrelu -> 3x3x3(stride) -> IN -> lrelu -> 1x1 -> IN (+res) -> lrelu."""
        B = x.shape[0]
        n1, n2, n3 = pre + ".conv1.conv", pre + ".conv2.conv", pre + ".conv3.conv"
        planes = self.w._p(n1 + ".weight").shape[0]
        if (_PAIR_L1 and planes == 32 and tuple(stride) == (1, 1, 1) and x.shape[3] % 2 == 0 and x.is_contiguous()
                and self.w._p(n3 + ".weight").shape[0] % 64 == 0 and x.shape[-1] % 64 == 0):
            return self._bottleneck_paired(pre, x, has_down)
        st1 = self.stats.take(B, self.w.conv1(n1).n_real)
        a1 = self.in_apply(self.conv1x1(x, n1, st1), st1)
        st2 = self.stats.take(B, self.w.conv3(n2).n_real)
        strided = tuple(stride) != (1, 1, 1)
        if not strided:
            c2 = self.conv3x3(a1, n2, st2)
        else:
            # stride-s 3x3x3, pad 1 == the stride-1 result sampled at multiples of s
            full = self.conv3x3(a1, n2)
            _, X, Y, Z, C = full.shape
            c2 = self._empty(B, -(-X // stride[0]), -(-Y // stride[1]), -(-Z // stride[2]), C)
            self.subsample(full, c2, stride)
            ops.in_stats(c2, st2)
        a2 = self.in_apply(c2, st2)
        st3 = self.stats.take(B, self.w.conv1(n3).n_real)
        c3 = self.conv1x1(a2, n3, st3)
        if has_down:
            nd = pre + ".downsample.0.conv"
            xs = x
            if strided:
                _, X, Y, Z, C = x.shape
                xs = self._empty(B, -(-X // stride[0]), -(-Y // stride[1]), -(-Z // stride[2]), C)
                self.subsample(x, xs, stride)
            std = self.stats.take(B, self.w.conv1(nd).n_real)
            r = self.conv1x1(xs, nd, std)
            return self.in_apply(c3, st3, res=r, rstats=std)
        return self.in_apply(c3, st3, res=x)

    def _bottleneck_paired(self, pre: str, x, has_down: bool):
        """The same bottleneck for planes = 32 (ResNet layer 1) with the 32-channel tensors held as PAIRED rows: two
        z-neighbouring voxels per dense 64-channel row ([B, X, Y, Z/2, 64]) instead of 32 live + 32 zero channels per voxel
        — half the rows through the 3x3x3 convolution, its weight gradient and the InstanceNorm passes, no padding bytes.
        The 1x1x1 convolutions read / write the un-paired 64- / 128-channel tensors through [.., Z/2, 2 C] views of the
        same memory with block-diagonal weights; the 3x3x3 convolution runs over pairs (CTU_PACK_PAIR_* layouts).  The
        InstanceNorm sums of the two column halves are the same channels: ctu_stats_fold merges them."""
        B, X, Y, Z, Cin = x.shape
        n1, n2, n3 = pre + ".conv1.conv", pre + ".conv2.conv", pre + ".conv3.conv"
        cout = self.w._p(n3 + ".weight").shape[0]
        xp = x.view(B, X, Y, Z // 2, 2 * Cin)
        self._reshaped(xp, x)
        st1 = self.stats.take(B, 64)
        c1 = self.gemm(xp, "pconv1", n1, self._empty(B, X, Y, Z // 2, 64), dims=self._flat_dims(xp), stats=st1)
        ops.stats_fold(st1, 32, 0.5)
        a1 = self.in_apply(c1, st1, fold=32)
        st2 = self.stats.take(B, 64)
        c2 = self.gemm(a1, "pconv3", n2, self._empty(B, X, Y, Z // 2, 64), dims=self._dims(a1), stats=st2)
        ops.stats_fold(st2, 32, 0.5)
        a2 = self.in_apply(c2, st2, fold=32)
        st3 = self.stats.take(B, 2 * cout)
        c3p = self.gemm(a2, "pconv1", n3, self._empty(B, X, Y, Z // 2, 2 * cout), dims=self._flat_dims(a2), stats=st3)
        ops.stats_fold(st3, cout, 1.0)          # consumed as un-paired rows: plain sums over all voxels
        c3 = c3p.view(B, X, Y, Z, cout)
        self._reshaped(c3, c3p)
        if has_down:
            nd = pre + ".downsample.0.conv"
            std = self.stats.take(B, self.w.conv1(nd).n_real)
            r = self.conv1x1(x, nd, std)
            return self.in_apply(c3, st3, res=r, rstats=std)
        return self.in_apply(c3, st3, res=x)

    def resnet(self, pre: str, x_in, layers: List[int], probe: Optional[Callable] = None):
        """resnet.py:213-230 (no max pool): stem k7 s(2,2,1) -> IN -> lrelu -> 4 stages; returns 4 feature maps.
        `probe(name, x) -> x'` (inference only) sees the stem output ("stem") and every block output
        ("layer{L}.{i}") as channels-last bf16 and may substitute what the NEXT block consumes — the teacher-forced
        wiring test (SURVEY 8c level 2) drives the real stage loop (strides, down-sample placement, names) this way."""
        B, _, X, Y, Z = x_in.shape
        s0 = DS_STRIDE[0]
        Xo, Yo, Zo = (X + 6 - 7) // s0[0] + 1, (Y + 6 - 7) // s0[1] + 1, (Z + 6 - 7) // s0[2] + 1
        # stem k7 (343 taps) on the tensor cores: single-channel im2col (taps zero-padded to 384 columns) + GEMM with the
        # InstanceNorm statistics in the epilogue; the column matrix is kept for the weight gradient in training
        col = self._empty(B, Xo, Yo, Zo, 384)
        ops.im2col_cin1(x_in, col, k=(7, 7, 7), s=s0, p=(3, 3, 3))
        st = self.stats.take(B, 64)
        x = self.gemm(col, "cin1", pre + "conv1.conv", self._empty(B, Xo, Yo, Zo, 64), dims=(Zo, Yo, Xo, B), stats=st,
                      a_needs_grad=False)
        x = self.in_apply(x, st)
        if probe is not None:
            assert self.tape is None
            x = probe("stem", x)
        feats = []
        strides = [(1, 1, 1), DS_STRIDE[1], DS_STRIDE[2], DS_STRIDE[3]]
        for li, nb in enumerate(layers):
            for bi in range(nb):
                x = self.bottleneck(f"{pre}layer{li + 1}.{bi}", x, strides[li] if bi == 0 else (1, 1, 1), bi == 0)
                if probe is not None:
                    x = probe(f"layer{li + 1}.{bi}", x)
            feats.append(x)
        return feats

    # ------------------------------------------------------------------ networks/vit.py
    def ffn(self, pre: str, x, out_dtype=None):
        """LN -> Linear -> GELU -> Linear, + x (vit.py:34-44,95; hybrid_CTUNet.py:517-526 inside Residual).
        Returns a new tensor (the residual stream is never updated in place)."""
        M, D = x.shape
        h = self.layernorm(x, pre + ".net.0", self._empty(M, D))
        n1, n2 = pre + ".net.1", pre + ".net.4"
        hidden = self.w.linear(n1).n_real
        pre_act = None
        if (self.tape is None and _FUSED_FFN and x.dtype == torch.bfloat16 and out_dtype in (None, torch.bfloat16)
                and ops.ffn_fused_supported(M, D, hidden)):
            # inference, 128-channel stage: both GEMMs in one kernel, the [M, hidden] activation never reaches HBM
            w1, w2 = self.w.get("lin", n1, True), self.w.get("lin", n2, True)
            if w1.w.shape == (hidden, D) and w2.w.shape == (D, hidden):
                return ops.ffn_fused(h, w1, w2, x, self._empty(M, D))
        if self.tape is None:
            f = self.gemm(h, "lin", n1, self._empty(M, hidden), dims=(M, 1, 1, 1), extra=True, act=ACT_GELU)
        else:
            # training keeps the pre-activation; the GELU backward is fused into the epilogue of the down-projection's
            # input-gradient GEMM (no tape record for the activation itself)
            pre_act = self.gemm(h, "lin", n1, self._empty(M, hidden), dims=(M, 1, 1, 1), extra=True)
            f = ops.gelu(pre_act, self._empty(M, hidden))
        out = self._empty(M, D, dtype=out_dtype or x.dtype)
        return self.gemm(f, "lin", n2, out, dims=(M, 1, 1, 1), extra=True, residual=x,
                         out_mode=OUT_F32 if out.dtype == F32 else OUT_BF16, a_gelu_of=pre_act)

    def vit_attention(self, pre: str, x, B: int, n: int, heads: int):
        """vit.py:66-78 + residual (vit.py:94); x: fp32 [B*n, D]; returns the updated stream."""
        M, D = x.shape
        h = self.layernorm(x, pre + ".norm", self._empty(M, D))
        qkv = self.gemm(h, "lin", pre + ".to_qkv", self._empty(M, 3 * D), dims=(M, 1, 1, 1), extra=False)
        a = self.attention(qkv, self._empty(M, D), dim_head=D // heads, n=n, windows=B, mode=0)
        return self.gemm(a, "lin", pre + ".to_out.0", self._empty(M, D, dtype=F32), dims=(M, 1, 1, 1), extra=True,
                         out_mode=OUT_F32, residual=x)

    def vit(self, pre: str, x_in, pf: int, depth: int, heads: int):
        """vit.py:130-139: returns the fp32 token stream [B*n, dim]."""
        B, _, X, Y, Z = x_in.shape
        n = (X // 16) * (Y // 16) * (Z // pf)
        e = pre + "to_patch_embedding"
        tok = self.patchify_ln(x_in, pf, e + ".1", self._empty(B * n, 256 * pf))
        dim = self.w.linear(e + ".2").n_real
        emb = self.gemm(tok, "lin", e + ".2", self._empty(B * n, dim, dtype=F32), dims=(B * n, 1, 1, 1), extra=True,
                        out_mode=OUT_F32, a_needs_grad=True)
        x = self.layernorm(emb, e + ".3", self._empty(B * n, dim, dtype=F32), add_name=pre + "pos_embedding")
        for i in range(depth):
            t = f"{pre}transformer.{i}"
            x = self.vit_attention(t + ".attn", x, B, n, heads)
            x = self.ffn(t + ".ff", x)
        return x, n

    # ------------------------------------------------------------------ networks/hybrid_CTUNet.py
    def res_block(self, pre: str, x, cin: int, cout: int, out=None):
        """hybrid_CTUNet.py:93-105 (k3, stride 1): conv-IN-lrelu-conv-IN, + (conv1x1-IN)(x) or x, lrelu."""
        B = x.shape[0]
        n1, n2 = pre + ".conv1.conv", pre + ".conv2.conv"
        st1 = self.stats.take(B, self.w.conv3(n1).n_real)
        a1 = self.in_apply(self.conv3x3(x, n1, st1), st1)
        st2 = self.stats.take(B, self.w.conv3(n2).n_real)
        c2 = self.conv3x3(a1, n2, st2)
        if cin != cout:
            n3 = pre + ".conv3.conv"
            st3 = self.stats.take(B, self.w.conv1(n3).n_real)
            r = self.conv1x1(x, n3, st3)
            return self.in_apply(c2, st2, res=r, rstats=st3, out=out)
        return self.in_apply(c2, st2, res=x, out=out)

    def res_block_cin1(self, pre: str, x_in, out=None):
        """ResBlock(1 -> 64) of vit_encoder0 (hybrid_CTUNet.py:786-793): conv1/conv3 have one input channel."""
        B, _, X, Y, Z = x_in.shape
        st1 = self.stats.take(B, 64)
        if _CIN1_TC:
            # k3 (27 taps) like the stem: single-channel im2col (taps zero-padded to 64 columns) + tensor-core GEMM with the
            # InstanceNorm statistics in its epilogue — no separate statistics pass, and the column matrix is kept for the
            # weight gradient instead of being rebuilt in the backward pass
            col = self._empty(B, X, Y, Z, 64)
            ops.im2col_cin1(x_in, col, k=(3, 3, 3), s=(1, 1, 1), p=(1, 1, 1))
            c1 = self.gemm(col, "cin1", pre + ".conv1.conv", self._empty(B, X, Y, Z, 64), dims=(Z, Y, X, B), stats=st1,
                           a_needs_grad=False)
        else:
            c1 = self.conv_cin1(x_in, pre + ".conv1.conv", self._empty(B, X, Y, Z, 64), k=(3, 3, 3), s=(1, 1, 1), p=(1, 1, 1))
            ops.in_stats(c1, st1)
        a1 = self.in_apply(c1, st1)
        st2 = self.stats.take(B, 64)
        c2 = self.conv3x3(a1, pre + ".conv2.conv", st2)
        r = self.conv_cin1(x_in, pre + ".conv3.conv", self._empty(B, X, Y, Z, 64), k=(1, 1, 1), s=(1, 1, 1), p=(0, 0, 0))
        st3 = self.stats.take(B, 64)
        if _CIN1_TC and x_in.dtype == F32 and x_in.is_contiguous():
            # the statistics of a pointwise single-channel convolution follow from the moments of its input: no pass over r
            ops.cin1_k1_stats(x_in, self.w.conv_cin1(pre + ".conv3.conv"), self.stats.take(B, 1), st3)
        else:
            ops.in_stats(r, st3)
        return self.in_apply(c2, st2, res=r, rstats=st3, out=out)

    def pixelweight_attention(self, pre: str, x1, x2):
        """hybrid_CTUNet.py:645-669, the binary cross-weight fusion of two [B,X,Y,Z,C] maps."""
        B, X, Y, Z, C = x1.shape
        T = B * X * Y * Z
        q = []
        for x, nrm, lin in ((x1, ".norm1", ".to_qkv1"), (x2, ".norm2", ".to_qkv2")):
            h = self.layernorm(x.reshape(T, C), pre + nrm, self._empty(T, C))
            q.append(self.gemm(h, "lin", pre + lin, self._empty(T, 3 * C), dims=(T, 1, 1, 1), extra=False))
        f = self.pwa_fuse(q[0], q[1], self._empty(T, C))
        return self.gemm(f, "lin", pre + ".to_out.0", self._empty(B, X, Y, Z, C), dims=(T, 1, 1, 1), extra=False)

    def up_2fusion(self, pre: str, inp, skip_conv, skip_vit, cout: int):
        """hybrid_CTUNet.py:329-341."""
        skip = self.pixelweight_attention(pre + ".pixelweight_attention1", skip_conv, skip_vit)
        skip = self.res_block(pre + ".up_addconv_block1", skip, cout, cout)
        up = self.up_gemm(inp, "convt", pre + ".transp_conv.conv")
        fused = self.pixelweight_attention(pre + ".pixelweight_attention2", up, skip)
        return self.res_block(pre + ".up_addconv_block2", fused, cout, cout)

    def window_attention(self, pre: str, x, grid, mode: int):
        """Residual(MultiAxisAttention) (hybrid_CTUNet.py:481-511); x: [T, D] residual stream; returns the new stream."""
        T, D = x.shape
        h = self.layernorm(x, pre + ".norm", self._empty(T, D))
        qkv = self.gemm(h, "lin", pre + ".to_qkv", self._empty(T, 3 * D), dims=(T, 1, 1, 1), extra=False)
        a = self.attention(qkv, self._empty(T, D), dim_head=32, n=216, mode=mode, bias_name=pre + ".rel_pos_bias", grid=grid)
        return self.gemm(a, "lin", pre + ".to_out.0", self._empty(T, D, dtype=x.dtype), dims=(T, 1, 1, 1), extra=False,
                         out_mode=OUT_F32 if x.dtype == F32 else OUT_BF16, residual=x)

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
            if ind <= 2:
                x = self.window_attention(p + ".1.fn", x, (B, X, Y, Z), 1)
                x = self.ffn(p + ".2.fn", x)
                x = self.window_attention(p + ".5.fn", x, (B, X, Y, Z), 2)
                xb = self.ffn(p + ".6.fn", x, out_dtype=BF16)
                ps_name = p + ".8.to_out"
            else:
                x = self.ffn(p + ".1.fn", x)
                xb = self.ffn(p + ".2.fn", x, out_dtype=BF16)
                ps_name = p + ".4.to_out"
            out = out_last if ind == 3 else None
            y = self.up_gemm(xb.view(B, X, Y, Z, D), "ps", ps_name, extra=f, out=out)
            feats.append(y)
            X, Y, Z = X * f[0], Y * f[1], Z * f[2]
            x = y.reshape(B * X * Y * Z, -1) if out is None else None
        return feats

    # ------------------------------------------------------------------ whole networks
    def _vit_branch(self, x_in, pf: int, depth: int, heads: int, lane2: Optional[torch.cuda.Stream] = None):
        """ViT -> window-attention decoder -> concat with vit_encoder0 -> vit_decoder0 -> the two ViT-side heads.
        `lane2`: a third stream for vit_encoder0 (independent of the transformer until the concat): its GPU-filling 96^3
        kernels then run beside the transformer's long run of small ones, forward and backward."""
        B, _, X, Y, Z = x_in.shape
        cat = self._empty(B, X, Y, Z, 128)  # torch.cat((vit_enc_96x96, vit_enc0), dim=1) built in place
        lo, hi = cat[..., :64], cat[..., 64:]
        self._alias(lo, cat, 0)
        self._alias(hi, cat, 64)
        if lane2 is None:
            tokens, n = self.vit("vit.", x_in, pf, depth, heads)
            self.res_block_cin1("vit_encoder0.layer", x_in, out=hi)
        else:
            # same recording order as the single-stream path (the gradient arena hands out its slices in tape order and
            # the unpack table built by the eager warm-up step must stay valid), but lane 2 only waits for what precedes
            # the transformer on this stream
            fork = torch.cuda.Event()
            fork.record(torch.cuda.current_stream())   # `cat` was allocated on this lane's stream
            tokens, n = self.vit("vit.", x_in, pf, depth, heads)
            lane2.wait_event(fork)
            cat.record_stream(lane2)
            prev = self.lane
            self.lane = 2
            try:
                with torch.cuda.stream(lane2):
                    self.res_block_cin1("vit_encoder0.layer", x_in, out=hi)
            finally:
                self.lane = prev
        enc = self.up_attention_block("vit_encoder.", tokens, B, (X // 16, Y // 16, Z // pf), out_last=lo)
        if lane2 is not None:
            torch.cuda.current_stream().wait_stream(lane2)
            if self.tape is not None:
                self.tape.fns.append((2, None))  # lane 2 joins lane 1 here (forward); the backward forks here
        # the head on `lo` is recorded BEFORE the block that consumes the whole concat buffer: in the backward pass the
        # block's gradient then becomes the buffer's gradient as it is, and the head's input-gradient GEMM adds into its
        # `lo` columns in place (the other order zero-fills a 453 MB buffer and runs two accumulation passes)
        vit_96 = self.head(lo, "lin", "decoder_linear_96x96.head", a_c=64)
        vit_out = self.res_block("vit_decoder0.conv_block", cat, 128, 64)
        vit_logits = self.head(vit_out, "conv1", "vit_out.conv.conv")
        return enc, vit_logits, vit_96

    def ctunet(self, x_in, layers, pf: int, depth: int = 12, heads: int = 12):
        """CTUNet.forward (hybrid_CTUNet.py:817-857)."""
        self.stats.reset()
        # (only while a CUDA graph is being captured: in eager mode the caching allocator cannot recycle blocks that
        # cross streams until their events complete, which costs far more than the overlap gains — measured 97 vs 35 ms)
        if self.two_lanes and torch.cuda.is_current_stream_capturing():
            main, side = torch.cuda.current_stream(), self._side_stream()
            side.wait_stream(main)
            self.lane = 1
            self._set_lane_sms(True)
            try:
                with torch.cuda.stream(side):
                    enc, vit_logits, vit_96 = self._vit_branch(x_in, pf, depth, heads,
                                                               lane2=self._side2_stream() if self.three_lanes else None)
                self.lane = 0
                res = self.resnet("convnet.", x_in, layers)
            finally:
                self.lane = 0
                self._set_lane_sms(False)
            main.wait_stream(side)
            for t in (*enc, vit_logits, vit_96):
                t.record_stream(main)
            if self.tape is not None:
                self.tape.fns.append((0, None))  # join marker for the backward pass
        else:
            enc, vit_logits, vit_96 = self._vit_branch(x_in, pf, depth, heads)
            res = self.resnet("convnet.", x_in, layers)
        dec3 = self.up_2fusion("res_decoder3", res[3], res[2], enc[0], 512)
        dec2 = self.up_2fusion("res_decoder2", dec3, res[1], enc[1], 256)
        dec1 = self.up_2fusion("res_decoder1", dec2, res[0], enc[2], 128)
        up0 = self.up_gemm(dec1, "convt", "res_decoder0.transp_conv.conv")
        res_out = self.res_block("res_decoder0.conv_block", up0, 64, 64)
        res_logits = self.head(res_out, "conv1", "res_out.conv.conv")
        res_48 = self.head(dec1, "conv1", "res_out_48x48.conv.conv")
        res_24 = self.head(dec2, "conv1", "res_out_24x24.conv.conv")
        return ((res_logits, res_48, res_24), (vit_logits, vit_96))

    def tunet(self, x_in, pf: int, depth: int = 12, heads: int = 12):
        """TUNet.forward (hybrid_CTUNet.py:1021-1036)."""
        self.stats.reset()
        # inference under CUDA-graph capture: vit_encoder0 (GPU-filling 96^3 kernels, independent of the transformer until
        # the concat) on a second stream beside the transformer's long run of small kernels — TUNet has no ResNet lane to
        # fill the idle SMs otherwise (CTU_TUNET_LANES=0 switches it off)
        lane2 = None
        if self.two_lanes and _TUNET_LANES and self.tape is None and torch.cuda.is_current_stream_capturing():
            lane2 = self._side2_stream()
        _, vit_logits, vit_96 = self._vit_branch(x_in, pf, depth, heads, lane2=lane2)
        return (vit_logits, vit_96)

    def up_cat_conv(self, pre: str, inp, skip, cout: int):
        """UpCatConvBlock (hybrid_CTUNet.py:196-201): ConvT -> cat(skip) -> ResBlock(2C -> C)."""
        B, X, Y, Z, _ = skip.shape
        cat = self._empty(B, X, Y, Z, 2 * cout)
        lo, hi = cat[..., :cout], cat[..., cout:]
        self._alias(lo, cat, 0)
        self._alias(hi, cat, cout)
        self.up_gemm(inp, "convt", pre + ".transp_conv.conv", out=lo)
        self.subsample(skip, hi, (1, 1, 1))  # stride-1 gather == strided copy into the concat buffer
        return self.res_block(pre + ".conv_block", cat, 2 * cout, cout)

    def cunet(self, x_in, layers):
        """CUNet.forward (hybrid_CTUNet.py:919-937)."""
        self.stats.reset()
        res = self.resnet("convnet.", x_in, layers)
        dec3 = self.up_cat_conv("res_decoder3", res[3], res[2], 512)
        dec2 = self.up_cat_conv("res_decoder2", dec3, res[1], 256)
        dec1 = self.up_cat_conv("res_decoder1", dec2, res[0], 128)
        up0 = self.up_gemm(dec1, "convt", "res_decoder0.transp_conv.conv")
        res_out = self.res_block("res_decoder0.conv_block", up0, 64, 64)
        return (self.head(res_out, "conv1", "res_out.conv.conv"),
                self.head(dec1, "conv1", "res_out_48x48.conv.conv"),
                self.head(dec2, "conv1", "res_out_24x24.conv.conv"))
