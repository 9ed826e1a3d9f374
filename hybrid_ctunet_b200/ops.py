"""Thin Python wrappers over the C ABI: torch tensors in, raw device pointers out.

PyTorch is used only for device memory (torch.empty) and the current CUDA stream; all arithmetic happens in
libctunet_b200.so.  Activations are channels-last bf16: a reference [B,C,X,Y,Z] tensor is [B,X,Y,Z,C] here.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from functools import lru_cache
from typing import Optional, Sequence, Tuple

import torch

from . import lib as _lib
from .lib import GemmDesc, WgradDesc, check

OUT_BF16, OUT_F32, OUT_F32_CF = 0, 1, 2
ACT_NONE, ACT_GELU = 0, 1
RES_NONE, RES_BF16, RES_F32, RES_GELU_BWD = 0, 1, 2, 3


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


@lru_cache(maxsize=None)
def pick_box(d1: int, d2: int, d3: int) -> Tuple[int, int, int]:
    """128-voxel tile box (b1,b2,b3) for a [d3,d2,d1] grid: least padded volume, then widest contiguous run."""
    best = None
    for e1 in range(8):
        for e2 in range(8 - e1):
            e3 = 7 - e1 - e2
            b1, b2, b3 = 1 << e1, 1 << e2, 1 << e3
            padded = (-(-d1 // b1) * b1) * (-(-d2 // b2) * b2) * (-(-d3 // b3) * b3)
            key = (padded, -b1, -b2)
            if best is None or key < best[0]:
                best = (key, (b1, b2, b3))
    return best[1]


@dataclass
class PackedWeight:
    """bf16 [n_pad, k_total] K-contiguous weight matrix as the tensor-core kernel consumes it."""
    w: torch.Tensor
    n_real: int
    a_c: int          # channels per filter tap (K of a plain GEMM)
    ksize: int = 1    # 1 or 3
    block_n: int = 64
    convt: Optional[Tuple[int, int, int, int]] = None  # (cout, u1, u2, u3)
    bias: Optional[torch.Tensor] = None                 # fp32 [n_real]
    alg_flops_per_row: float = 0.0                      # algorithmic FLOPs per GEMM row (true channel counts, no padding)
    a_c_live: int = 0                                   # 3x3x3 only: input channels that can be non-zero (0 = all a_c)
    x3: Optional[torch.Tensor] = None                   # 3x3x3, block_n 64: re-laid copy for the two-plane kernel (w_x3)
    x3_item: int = -1


def pick_block_n(n: int) -> int:
    if n <= 16:
        return 16
    if n <= 32:
        return 32
    if n % 128 == 0:
        return 128
    return 64


def pack_matrix(w2d: torch.Tensor, *, ksize: int = 1, a_c: Optional[int] = None, bias=None, block_n=None,
                convt=None) -> PackedWeight:
    """w2d: float [N, K] (K ordered (tap, channel)) -> padded bf16 PackedWeight on the same device."""
    n, k = w2d.shape
    bn = block_n or pick_block_n(n)
    n_pad = -(-n // bn) * bn
    k_pad = -(-k // 8) * 8
    out = torch.zeros((n_pad, k_pad), dtype=torch.bfloat16, device=w2d.device)
    out[:n, :k] = w2d.to(torch.bfloat16)
    if a_c is None:
        a_c = k_pad // (ksize ** 3)
    b = None if bias is None else bias.detach().to(torch.float32).contiguous()
    pw = PackedWeight(out, n, a_c, ksize, bn, convt, b)
    if ksize == 3 and bn == 64 and a_c % 64 == 0 and k_pad == 27 * a_c:
        pw.x3 = x3_layout(out, a_c)
    return pw


def x3_layout(packed: torch.Tensor, a_c: int) -> torch.Tensor:
    """CTU_PACK_X3_FROM_PACKED with torch ops: [n_pad][27 * a_c] (tap = (t3 * 3 + t2) * 3 + t1 major) ->
    [(t2, t1)][n tile][x-tap in the order +1, 0, -1][64 rows] x a_c."""
    n_pad = packed.shape[0]
    v = packed.view(n_pad // 64, 64, 3, 9, a_c).flip(2).permute(3, 0, 2, 1, 4)
    return v.contiguous().view(9 * (n_pad // 64) * 192, a_c)


def gemm(a: torch.Tensor, w: PackedWeight, out: torch.Tensor, *, dims: Sequence[int], out_mode: int = OUT_BF16,
         act: int = ACT_NONE, residual: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None,
         out_col0: int = 0, a_c: Optional[int] = None, gelu_bwd_of: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = epilogue(A (*) W^T) on the tcgen05 kernel.

    a    : bf16, channels-last, last dim stride 1; its row stride (a.stride(-2)) is lda.
    dims : (d1, d2, d3, d4) spatial extents of `a` (d1 fastest, d4 batch).  For a flat token GEMM use
           (M, 1, 1, 1); for a per-batch GEMM that feeds InstanceNorm use (S, 1, 1, B).
    out  : bf16/fp32 rows with row stride out.stride(-2) (OUT_BF16 / OUT_F32) or contiguous NCDHW fp32 (OUT_F32_CF).
    gelu_bwd_of : bf16 pre-activation x shaped like `out`: out = (A (*) W^T) * gelu'(x) (instead of a residual).
    """
    lib = _lib.require_device()
    d1, d2, d3, d4 = (int(v) for v in dims)
    assert a.dtype == torch.bfloat16 and a.stride(-1) == 1
    lda = a.stride(-2) if a.dim() >= 2 else a.shape[-1]
    ac = int(a_c if a_c is not None else w.a_c)
    box = pick_box(d1, d2, d3)
    d = GemmDesc()
    d.a, d.w, d.out = a.data_ptr(), w.w.data_ptr(), out.data_ptr()
    d.bias = _ptr(w.bias)
    d.residual = _ptr(residual)
    d.stats = _ptr(stats)
    d.a_c, d.lda = ac, int(lda)
    d.d1, d.d2, d.d3, d.d4 = d1, d2, d3, d4
    d.b1, d.b2, d.b3 = box
    d.k1 = d.k2 = d.k3 = w.ksize
    d.n_pad, d.n_real, d.k_total = w.w.shape[0], w.n_real, w.w.shape[1]
    d.block_n = w.block_n
    d.out_mode = out_mode
    d.ldc = 0 if out_mode == OUT_F32_CF else int(out.stride(-2))
    d.act = act
    if gelu_bwd_of is not None:
        assert residual is None and gelu_bwd_of.dtype == torch.bfloat16 and gelu_bwd_of.stride(-1) == 1
        d.residual = gelu_bwd_of.data_ptr()
        d.res_mode, d.ldr = RES_GELU_BWD, int(gelu_bwd_of.stride(-2))
    elif residual is None:
        d.res_mode, d.ldr = RES_NONE, 0
    else:
        assert residual.stride(-1) == 1
        d.res_mode = RES_F32 if residual.dtype == torch.float32 else RES_BF16
        d.ldr = int(residual.stride(-2))
    if w.convt is not None:
        d.convt_cout, d.u1, d.u2, d.u3 = w.convt
    else:
        d.convt_cout, d.u1, d.u2, d.u3 = 0, 1, 1, 1
    d.stats_ld = 0 if stats is None else int(stats.shape[-2])
    d.out_col0 = out_col0
    d.a_c_live = w.a_c_live if (w.ksize == 3 and 0 < w.a_c_live < ac) else 0
    d.w_x3 = w.x3.data_ptr() if (w.x3 is not None and ac == w.a_c) else None
    check(lib.ctu_umma_gemm(C.byref(d), _stream()), "ctu_umma_gemm")
    return out


def ffn_fused(a: torch.Tensor, w1: PackedWeight, w2: PackedWeight, residual: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out = residual + W2 · GELU(W1 · a + b1) + b2 with the hidden activation kept on chip (ctu_ffn_fused; inference).
    a / residual / out: bf16 [M, 128] rows; w1: packed [hidden, 128] with bias, w2: packed [128, hidden] with bias."""
    lib = _lib.require_device()
    M, Cc = a.shape
    hidden = w1.n_real
    assert a.dtype == residual.dtype == out.dtype == torch.bfloat16 and a.stride(-1) == residual.stride(-1) == out.stride(-1) == 1
    assert w1.w.shape == (hidden, Cc) and w2.w.shape == (Cc, hidden) and w1.bias is not None and w2.bias is not None
    check(lib.ctu_ffn_fused(a.data_ptr(), int(a.stride(-2)), w1.w.data_ptr(), w1.bias.data_ptr(), w2.w.data_ptr(),
                            w2.bias.data_ptr(), residual.data_ptr(), int(residual.stride(-2)), out.data_ptr(),
                            int(out.stride(-2)), M, Cc, hidden, _stream()), "ctu_ffn_fused")
    return out


def ffn_fused_supported(M: int, Cc: int, hidden: int) -> bool:
    return Cc == 128 and hidden % 128 == 0 and hidden <= 512 and M >= 128


# ------------------------------------------------------------------------------------------------ HBM-bound ops
IN_EPS = 1e-5
LRELU_SLOPE = 0.01


def in_stats(x: torch.Tensor, stats: torch.Tensor) -> torch.Tensor:
    """x: bf16 [B, ..., C] channels-last (row stride x.stride(-2)); stats: fp64 [B, C', 2] accumulated in place."""
    lib = _lib.require_device()
    B, C = x.shape[0], x.shape[-1]
    S = 1
    for d in x.shape[1:-1]:
        S *= int(d)
    check(lib.ctu_in_stats(x.data_ptr(), int(x.stride(-2)), B, S, C, stats.data_ptr(), int(stats.shape[-2]), _stream()),
          "ctu_in_stats")
    return stats


def in_apply(x: torch.Tensor, xstats: torch.Tensor, out: torch.Tensor, *, res: Optional[torch.Tensor] = None,
             rstats: Optional[torch.Tensor] = None, act: bool = True) -> torch.Tensor:
    lib = _lib.require_device()
    B, C = x.shape[0], x.shape[-1]
    S = 1
    for d in x.shape[1:-1]:
        S *= int(d)
    check(lib.ctu_in_apply(x.data_ptr(), int(x.stride(-2)), xstats.data_ptr(), int(xstats.shape[-2]),
                           _ptr(res), 0 if res is None else int(res.stride(-2)),
                           _ptr(rstats), 0 if rstats is None else int(rstats.shape[-2]),
                           out.data_ptr(), int(out.stride(-2)), B, S, C, IN_EPS, 1 if act else 0, LRELU_SLOPE, _stream()),
          "ctu_in_apply")
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, out: torch.Tensor, *,
              add: Optional[torch.Tensor] = None, eps: float = 1e-5) -> torch.Tensor:
    """x: [M, C] fp32 or bf16 (row stride x.stride(-2)); out: [M, C] fp32 or bf16."""
    lib = _lib.require_device()
    C_ = x.shape[-1]
    M = x.numel() // C_
    add_rows = 0 if add is None else add.numel() // C_
    check(lib.ctu_layernorm(x.data_ptr(), int(x.dtype == torch.float32), int(x.stride(-2)), gamma.data_ptr(),
                            beta.data_ptr(), _ptr(add), add_rows, out.data_ptr(), int(out.dtype == torch.float32),
                            int(out.stride(-2)), M, C_, eps, _stream()), "ctu_layernorm")
    return out


def patchify_ln(img: torch.Tensor, pf: int, gamma: torch.Tensor, beta: torch.Tensor, out: torch.Tensor,
                eps: float = 1e-5) -> torch.Tensor:
    """img: fp32 [B, 1, X, Y, Z] contiguous; out: bf16 [B*tokens, 256*pf]."""
    lib = _lib.require_device()
    B, c, X, Y, Z = img.shape
    assert c == 1 and img.is_contiguous() and img.dtype == torch.float32
    check(lib.ctu_patchify_ln(img.data_ptr(), B, X, Y, Z, pf, gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), eps,
                              _stream()), "ctu_patchify_ln")
    return out


def pwa_fuse(qkv1: torch.Tensor, qkv2: torch.Tensor, out: torch.Tensor, dim_head: int = 32) -> torch.Tensor:
    lib = _lib.require_device()
    T, C3 = qkv1.shape
    assert qkv1.is_contiguous() and qkv2.is_contiguous() and out.is_contiguous()
    check(lib.ctu_pwa_fuse(qkv1.data_ptr(), qkv2.data_ptr(), out.data_ptr(), T, C3 // 3, dim_head, _stream()),
          "ctu_pwa_fuse")
    return out


def subsample(x: torch.Tensor, out: torch.Tensor, stride: Tuple[int, int, int]) -> torch.Tensor:
    """x: bf16 [B, X, Y, Z, C] -> out [B, ceil(X/sx), ceil(Y/sy), ceil(Z/sz), C]; stride = (sx, sy, sz)."""
    lib = _lib.require_device()
    B, X, Y, Z, C_ = x.shape
    sx, sy, sz = stride
    check(lib.ctu_subsample(x.data_ptr(), int(x.stride(-2)), Z, Y, X, out.data_ptr(), int(out.stride(-2)), sz, sy, sx,
                            C_, B, _stream()), "ctu_subsample")
    return out


def attention(qkv: torch.Tensor, out: torch.Tensor, *, dim_head: int, n: int, windows: int = 0, mode: int = 0,
              bias: Optional[torch.Tensor] = None, grid: Tuple[int, int, int, int] = (1, 1, 1, 1), w: int = 6,
              lse: Optional[torch.Tensor] = None):
    """qkv: bf16 [rows, 3C]; out: bf16 [rows, C].  grid = (batch, X, Y, Z) token grid for mode 1 (block) / 2 (grid).
    lse: optional fp32 [rows, heads] receiving the base-2 log-sum-exp per (row, head) for the backward pass."""
    lib = _lib.require_device()
    C_ = qkv.shape[-1] // 3
    b, X, Y, Z = grid
    check(lib.ctu_attention(qkv.data_ptr(), int(qkv.stride(-2)), C_, dim_head, out.data_ptr(), int(out.stride(-2)),
                            _ptr(bias), n, windows, mode, b, X, Y, Z, w, _ptr(lse), _stream()), "ctu_attention")
    return out


def conv_cin1(x: torch.Tensor, w_taps: torch.Tensor, out: torch.Tensor, *, k, s, p) -> torch.Tensor:
    """x: fp32 [B,1,X,Y,Z]; w_taps: fp32 [kx*ky*kz, 64]; out: bf16 [B,Xo,Yo,Zo,ldo>=64]."""
    lib = _lib.require_device()
    B, c, X, Y, Z = x.shape
    assert c == 1 and x.is_contiguous() and x.dtype == torch.float32
    check(lib.ctu_conv_cin1(x.data_ptr(), w_taps.data_ptr(), out.data_ptr(), int(out.stride(-2)), w_taps.shape[1], B, X,
                            Y, Z, k[0], k[1], k[2], s[0], s[1], s[2], p[0], p[1], p[2], _stream()), "ctu_conv_cin1")
    return out


def blend_accumulate(logits0, logits1, imp, acc0, acc1, start):
    """logits*: fp32 [C, r3, r2, r1] (one window); acc*: fp32 [C, X, Y, Z]; start = (x0, y0, z0)."""
    lib = _lib.require_device()
    C_, r3, r2, r1 = logits0.shape
    _, X, Y, Z = acc0.shape
    check(lib.ctu_blend_accumulate(logits0.data_ptr(), _ptr(logits1), imp.data_ptr(), acc0.data_ptr(), _ptr(acc1), C_,
                                   r3, r2, r1, X, Y, Z, start[0], start[1], start[2], _stream()), "ctu_blend_accumulate")


def blend_count(imp, cnt, start):
    lib = _lib.require_device()
    r3, r2, r1 = imp.shape
    X, Y, Z = cnt.shape
    check(lib.ctu_blend_count(imp.data_ptr(), cnt.data_ptr(), r3, r2, r1, X, Y, Z, start[0], start[1], start[2],
                              _stream()), "ctu_blend_count")


def blend_normalize(acc, cnt, out):
    lib = _lib.require_device()
    C_ = acc.shape[0]
    vox = cnt.numel()
    check(lib.ctu_blend_normalize(acc.data_ptr(), cnt.data_ptr(), out.data_ptr(), C_, vox, _stream()),
          "ctu_blend_normalize")
    return out


# ------------------------------------------------------------------------------------------------ backward pass
def pick_wgrad_block_n(n: int) -> int:
    if n % 256 == 0:
        return 256
    if n % 128 == 0:
        return 128
    return 64


def wgrad(x: torch.Tensor, dy: torch.Tensor, dw: torch.Tensor, *, dims: Sequence[int], ksize: int = 1,
          x_c: Optional[int] = None, n: Optional[int] = None, alg_flops_per_row: float = 0.0) -> torch.Tensor:
    """dw[(tap, ci), co] += sum_v x[v + tap - pad, ci] * dy[v, co] on the tcgen05 wgrad kernel.

    x  : bf16 channels-last rows (row stride x.stride(-2)), first x_c channels used (x_c % 64 == 0).
    dy : bf16 channels-last rows, first n channels used.
    dw : fp32 [ksize^3 * x_c, >= n] accumulated in place (zero it first).
    dims as for `gemm`: (d1, d2, d3, d4) with d1 fastest; (M, 1, 1, 1) for a flat token GEMM.
    alg_flops_per_row is bookkeeping for bench.py's per-class roofline (unused here).
    """
    lib = _lib.require_device()
    d1, d2, d3, d4 = (int(v) for v in dims)
    assert x.dtype == torch.bfloat16 and dy.dtype == torch.bfloat16 and dw.dtype == torch.float32
    assert x.stride(-1) == 1 and dy.stride(-1) == 1 and dw.stride(-1) == 1
    d = WgradDesc()
    d.x, d.dy, d.dw = x.data_ptr(), dy.data_ptr(), dw.data_ptr()
    d.x_c = int(x_c if x_c is not None else x.shape[-1])
    d.ldx = int(x.stride(-2))
    d.n = int(n if n is not None else dy.shape[-1])
    d.ldy = int(dy.stride(-2))
    d.ldw = int(dw.stride(-2))
    assert dw.shape[-2] == ksize ** 3 * d.x_c and dw.shape[-1] >= d.n
    d.d1, d.d2, d.d3, d.d4 = d1, d2, d3, d4
    d.b1, d.b2, d.b3 = pick_box(d1, d2, d3)
    d.k1 = d.k2 = d.k3 = ksize
    d.block_n = pick_wgrad_block_n(d.n)
    check(lib.ctu_umma_wgrad(C.byref(d), _stream()), "ctu_umma_wgrad")
    return dw


def _bsc(x: torch.Tensor):
    B, C_ = x.shape[0], x.shape[-1]
    S = 1
    for d in x.shape[1:-1]:
        S *= int(d)
    return B, S, C_


def cin1_k1_stats(x_in: torch.Tensor, w: torch.Tensor, mom: torch.Tensor, stats: torch.Tensor) -> torch.Tensor:
    """InstanceNorm sums of r[v][c] = x[v] * w[c] from the moments of x (ctu_cin1_k1_stats).  x_in: fp32 [B, 1, X, Y, Z]
    contiguous; w: fp32 [C] (or [1, C]); mom: fp64 [B, 1, 2] zeroed; stats: fp64 [B, ld, 2]."""
    lib = _lib.require_device()
    B = x_in.shape[0]
    S = x_in.numel() // B
    assert x_in.dtype == torch.float32 and x_in.is_contiguous() and w.dtype == torch.float32 and w.is_contiguous()
    check(lib.ctu_cin1_k1_stats(x_in.data_ptr(), w.data_ptr(), B, S, int(w.numel()), mom.data_ptr(), stats.data_ptr(),
                                int(stats.shape[-2]), _stream()), "ctu_cin1_k1_stats")
    return stats


def stats_fold(stats: torch.Tensor, half: int, scale: float) -> torch.Tensor:
    """stats: fp64 [B, ld, width] sums of a PAIRED tensor (columns c and c + half are one channel): both halves become
    (v[c] + v[c + half]) * scale (ctu_stats_fold)."""
    lib = _lib.require_device()
    B, ld, width = stats.shape
    check(lib.ctu_stats_fold(stats.data_ptr(), B, ld, half, width, float(scale), _stream()), "ctu_stats_fold")
    return stats


def in_backward(dout, out, x, xstats, dx, *, res=None, rstats=None, dres=None, sums=None, act: bool = True, fold: int = 0):
    """Backward of in_apply.  dout/out/x(/res): bf16 [B, ..., C]; writes dx (and dres when the forward had a residual:
    the raw gradient g for an identity residual, the InstanceNorm backward for a normalised one).  x may be None when
    the forward had no residual (xhat is recovered from `out`)."""
    lib = _lib.require_device()
    B, S, C_ = _bsc(out)
    if sums is None:
        sums = torch.zeros(B, C_, 4, dtype=torch.float64, device=out.device)
    res_mode = 0 if dres is None else (2 if rstats is not None else 1)
    r2 = res if res_mode == 2 else None
    ldx = 0 if x is None else int(x.stride(-2))
    check(lib.ctu_in_bwd_stats(dout.data_ptr(), int(dout.stride(-2)), out.data_ptr(), int(out.stride(-2)), _ptr(x),
                               ldx, xstats.data_ptr(), int(xstats.shape[-2]), _ptr(r2),
                               0 if r2 is None else int(r2.stride(-2)), _ptr(rstats) if res_mode == 2 else None,
                               0 if res_mode != 2 else int(rstats.shape[-2]), B, S, C_, IN_EPS, 1 if act else 0,
                               LRELU_SLOPE, sums.data_ptr(), _stream()), "ctu_in_bwd_stats")
    if fold:   # paired rows: the two column halves are the same channels
        stats_fold(sums, fold, 0.5)
    check(lib.ctu_in_bwd_apply(dout.data_ptr(), int(dout.stride(-2)), out.data_ptr(), int(out.stride(-2)), _ptr(x),
                               ldx, xstats.data_ptr(), int(xstats.shape[-2]), _ptr(r2),
                               0 if r2 is None else int(r2.stride(-2)), _ptr(rstats) if res_mode == 2 else None,
                               0 if res_mode != 2 else int(rstats.shape[-2]), res_mode, B, S, C_, IN_EPS,
                               1 if act else 0, LRELU_SLOPE, sums.data_ptr(), dx.data_ptr(), int(dx.stride(-2)),
                               _ptr(dres), 0 if dres is None else int(dres.stride(-2)), _stream()), "ctu_in_bwd_apply")
    return dx


def layernorm_backward(x, gamma, dy, dgamma, dbeta, *, dx_in=None, dx_f32=None, dx_bf16=None, eps: float = 1e-5):
    """x: [M, C] fp32/bf16; dy: bf16 [M, C]; dgamma/dbeta: fp32 [C] accumulated; dx (+ dx_in) -> dx_f32 and/or dx_bf16."""
    lib = _lib.require_device()
    C_ = x.shape[-1]
    M = x.numel() // C_ if x.is_contiguous() else x.shape[0]
    check(lib.ctu_layernorm_bwd(x.data_ptr(), int(x.dtype == torch.float32), int(x.stride(-2)), gamma.data_ptr(),
                                dy.data_ptr(), int(dy.stride(-2)), _ptr(dx_in),
                                0 if dx_in is None else int(dx_in.dtype == torch.float32),
                                0 if dx_in is None else int(dx_in.stride(-2)), _ptr(dx_f32),
                                0 if dx_f32 is None else int(dx_f32.stride(-2)), _ptr(dx_bf16),
                                0 if dx_bf16 is None else int(dx_bf16.stride(-2)), dgamma.data_ptr(), dbeta.data_ptr(), M,
                                C_, eps, _stream()), "ctu_layernorm_bwd")


def gelu(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    lib = _lib.require_device()
    assert x.is_contiguous() and y.is_contiguous()
    check(lib.ctu_gelu(x.data_ptr(), y.data_ptr(), x.numel(), _stream()), "ctu_gelu")
    return y


def gelu_backward(x: torch.Tensor, dy: torch.Tensor, dx: torch.Tensor) -> torch.Tensor:
    lib = _lib.require_device()
    assert x.is_contiguous() and dy.is_contiguous() and dx.is_contiguous()
    check(lib.ctu_gelu_bwd(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), x.numel(), _stream()), "ctu_gelu_bwd")
    return dx


def pwa_fuse_backward(qkv1, qkv2, dout, dqkv1, dqkv2, dim_head: int = 32):
    lib = _lib.require_device()
    T, C3 = qkv1.shape
    for t in (qkv1, qkv2, dout, dqkv1, dqkv2):
        assert t.is_contiguous()
    check(lib.ctu_pwa_fuse_bwd(qkv1.data_ptr(), qkv2.data_ptr(), dout.data_ptr(), dqkv1.data_ptr(), dqkv2.data_ptr(), T,
                               C3 // 3, dim_head, _stream()), "ctu_pwa_fuse_bwd")


def colsum(x: torch.Tensor, out: torch.Tensor, n: Optional[int] = None) -> torch.Tensor:
    """out[c] += sum_rows x[row, c]; x: [M, >= n] bf16/fp32 with row stride x.stride(-2); out fp32 [n]."""
    lib = _lib.require_device()
    N = int(n if n is not None else x.shape[-1])
    M = x.numel() // x.shape[-1] if x.is_contiguous() else x.shape[0]
    check(lib.ctu_colsum(x.data_ptr(), int(x.dtype == torch.float32), int(x.stride(-2)), M, N, out.data_ptr(), _stream()),
          "ctu_colsum")
    return out


def cf_to_cl(src: torch.Tensor, dst: torch.Tensor, cpad: int) -> torch.Tensor:
    """src: fp32 [B, C, X, Y, Z] contiguous -> dst: bf16 [B, X, Y, Z, ld] with channels [C, cpad) zeroed."""
    lib = _lib.require_device()
    assert src.is_contiguous() and src.dtype == torch.float32
    B, C_ = src.shape[:2]
    S = src.numel() // (B * C_)
    check(lib.ctu_cf_to_cl(src.data_ptr(), dst.data_ptr(), B, C_, S, int(dst.stride(-2)), cpad, _stream()), "ctu_cf_to_cl")
    return dst


def head_backward(g: torch.Tensor, a: torch.Tensor, w: torch.Tensor, da: torch.Tensor, dw: torch.Tensor,
                  db: torch.Tensor, accumulate: bool = False) -> None:
    """Backward of a C -> n_cls logits head in one launch (ctu_head_bwd).  g: fp32 NCDHW [B, ncls, X, Y, Z] contiguous;
    a / da: bf16 channels-last [B, X, Y, Z, C] (row strides a.stride(-2) / da.stride(-2), da may hold a gradient to add
    to); w: the fp32 parameter [ncls, C(,1,1,1)]; dw: fp32 [C, ld] accumulated; db: fp32 [>= ncls] accumulated."""
    lib = _lib.require_device()
    B, ncls = g.shape[:2]
    S = g.numel() // (B * ncls)
    C_ = a.shape[-1]
    assert g.is_contiguous() and g.dtype == torch.float32 and w.is_contiguous() and w.dtype == torch.float32
    assert a.dtype == torch.bfloat16 and da.dtype == torch.bfloat16 and a.stride(-1) == 1 and da.stride(-1) == 1
    assert w.numel() == ncls * C_ and dw.shape[-2] == C_ and dw.stride(-1) == 1
    check(lib.ctu_head_bwd(g.data_ptr(), a.data_ptr(), int(a.stride(-2)), w.data_ptr(), da.data_ptr(), int(da.stride(-2)),
                           1 if accumulate else 0, dw.data_ptr(), int(dw.stride(-2)), db.data_ptr(), B, S, C_, ncls,
                           _stream()), "ctu_head_bwd")


def space_to_depth(x: torch.Tensor, out: torch.Tensor, up: Tuple[int, int, int]) -> torch.Tensor:
    """x: bf16 [B, X*ux, Y*uy, Z*uz, C] -> out: contiguous [B, X, Y, Z, ux*uy*uz*C]; up = (ux, uy, uz)."""
    lib = _lib.require_device()
    B, Xu, Yu, Zu, C_ = x.shape
    ux, uy, uz = up
    assert out.is_contiguous()
    check(lib.ctu_space_to_depth(x.data_ptr(), int(x.stride(-2)), out.data_ptr(), B, Xu // ux, Yu // uy, Zu // uz, ux, uy,
                                 uz, C_, _stream()), "ctu_space_to_depth")
    return out


def subsample_backward(dsub: torch.Tensor, dfull: torch.Tensor, stride, accumulate: bool = False) -> torch.Tensor:
    lib = _lib.require_device()
    B, X, Y, Z, C_ = dfull.shape
    sx, sy, sz = stride
    check(lib.ctu_subsample_bwd(dsub.data_ptr(), int(dsub.stride(-2)), dfull.data_ptr(), int(dfull.stride(-2)), Z, Y, X,
                                sz, sy, sx, C_, B, 1 if accumulate else 0, _stream()), "ctu_subsample_bwd")
    return dfull


def im2col_cin1(x: torch.Tensor, out: torch.Tensor, *, k, s, p) -> torch.Tensor:
    """x: fp32 [B,1,X,Y,Z]; out: bf16 contiguous [B,Xo,Yo,Zo,kpad]."""
    lib = _lib.require_device()
    B, c, X, Y, Z = x.shape
    assert c == 1 and x.is_contiguous() and out.is_contiguous()
    check(lib.ctu_im2col_cin1(x.data_ptr(), out.data_ptr(), B, X, Y, Z, k[0], k[1], k[2], s[0], s[1], s[2], p[0], p[1],
                              p[2], out.shape[-1], _stream()), "ctu_im2col_cin1")
    return out


def accumulate(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """dst += src for [.., C] tensors with arbitrary row strides; src / dst independently bf16 or fp32."""
    lib = _lib.require_device()
    assert src.numel() == dst.numel() and src.shape[-1] == dst.shape[-1]
    C_ = src.shape[-1]
    M = src.numel() // C_
    check(lib.ctu_accumulate(src.data_ptr(), int(src.dtype == torch.float32), int(src.stride(-2)), dst.data_ptr(),
                             int(dst.dtype == torch.float32), int(dst.stride(-2)), M, C_, _stream()), "ctu_accumulate")
    return dst


def cast_f32_bf16(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    lib = _lib.require_device()
    C_ = src.shape[-1]
    M = src.numel() // C_
    check(lib.ctu_cast_f32_bf16(src.data_ptr(), int(src.stride(-2)), dst.data_ptr(), int(dst.stride(-2)), M, C_, _stream()),
          "ctu_cast_f32_bf16")
    return dst


def patchify_ln_backward(img, pf: int, dtok, dgamma, dbeta, eps: float = 1e-5):
    lib = _lib.require_device()
    B, c, X, Y, Z = img.shape
    check(lib.ctu_patchify_ln_bwd(img.data_ptr(), B, X, Y, Z, pf, dtok.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), eps,
                                  _stream()), "ctu_patchify_ln_bwd")


def attention_backward(qkv, out, dout, lse, dqkv, dq_f32, *, dim_head: int, n: int, windows: int = 0, mode: int = 0,
                       bias_t: Optional[torch.Tensor] = None, ds_out: Optional[torch.Tensor] = None,
                       grid: Tuple[int, int, int, int] = (1, 1, 1, 1), w: int = 6):
    """Backward of `attention`.  dK/dV -> dqkv[:, C:], dQ accumulated into dq_f32 (fp32 [rows, C], zeroed by the caller);
    for dim_head 32 and n <= 224 (the 6^3 windows) dQ goes straight into dqkv[:, :C] and dq_f32 may be None."""
    lib = _lib.require_device()
    C_ = qkv.shape[-1] // 3
    rows = qkv.shape[0]
    heads = C_ // dim_head
    delta = torch.empty(rows, heads, dtype=torch.float32, device=qkv.device)
    check(lib.ctu_attention_delta(out.data_ptr(), int(out.stride(-2)), dout.data_ptr(), int(dout.stride(-2)),
                                  delta.data_ptr(), rows, C_, dim_head, _stream()), "ctu_attention_delta")
    b, X, Y, Z = grid
    check(lib.ctu_attention_bwd(qkv.data_ptr(), int(qkv.stride(-2)), C_, dim_head, dout.data_ptr(), int(dout.stride(-2)),
                                lse.data_ptr(), delta.data_ptr(), _ptr(bias_t), dqkv.data_ptr(), int(dqkv.stride(-2)),
                                _ptr(dq_f32), _ptr(ds_out), n, windows, mode, b, X, Y, Z, w, _stream()),
          "ctu_attention_bwd")
