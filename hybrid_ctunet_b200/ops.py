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
from .lib import GemmDesc, check

OUT_BF16, OUT_F32, OUT_F32_CF = 0, 1, 2
ACT_NONE, ACT_GELU = 0, 1
RES_NONE, RES_BF16, RES_F32 = 0, 1, 2


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
    return PackedWeight(out, n, a_c, ksize, bn, convt, b)


def gemm(a: torch.Tensor, w: PackedWeight, out: torch.Tensor, *, dims: Sequence[int], out_mode: int = OUT_BF16,
         act: int = ACT_NONE, residual: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None,
         out_col0: int = 0, a_c: Optional[int] = None) -> torch.Tensor:
    """out = epilogue(A (*) W^T) on the tcgen05 kernel.

    a    : bf16, channels-last, last dim stride 1; its row stride (a.stride(-2)) is lda.
    dims : (d1, d2, d3, d4) spatial extents of `a` (d1 fastest, d4 batch).  For a flat token GEMM use
           (M, 1, 1, 1); for a per-batch GEMM that feeds InstanceNorm use (S, 1, 1, B).
    out  : bf16/fp32 rows with row stride out.stride(-2) (OUT_BF16 / OUT_F32) or contiguous NCDHW fp32 (OUT_F32_CF).
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
    if residual is None:
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
    check(lib.ctu_umma_gemm(C.byref(d), _stream()), "ctu_umma_gemm")
    return out


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
              bias: Optional[torch.Tensor] = None, grid: Tuple[int, int, int, int] = (1, 1, 1, 1), w: int = 6):
    """qkv: bf16 [rows, 3C]; out: bf16 [rows, C].  grid = (batch, X, Y, Z) token grid for mode 1 (block) / 2 (grid)."""
    lib = _lib.require_device()
    C_ = qkv.shape[-1] // 3
    b, X, Y, Z = grid
    check(lib.ctu_attention(qkv.data_ptr(), int(qkv.stride(-2)), C_, dim_head, out.data_ptr(), int(out.stride(-2)),
                            _ptr(bias), n, windows, mode, b, X, Y, Z, w, _stream()), "ctu_attention")
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
