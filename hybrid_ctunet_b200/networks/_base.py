"""Shared plumbing of the drop-in modules: engine construction and layout conversion at block boundaries."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..engine import Engine
from ..lib import CtuError


def to_cl(x: torch.Tensor) -> torch.Tensor:
    """Reference layout [B,C,X,Y,Z] (any float dtype) -> channels-last bf16 [B,X,Y,Z,C] (block-level entry only;
    whole networks consume the fp32 input directly in their first kernels)."""
    return x.permute(0, 2, 3, 4, 1).to(torch.bfloat16).contiguous()


def from_cl(x: torch.Tensor) -> torch.Tensor:
    """Channels-last bf16 -> the reference's NCDHW fp32 (block-level exit only)."""
    return x.permute(0, 4, 1, 2, 3).float().contiguous()


class KernelModule(nn.Module):
    """nn.Module whose forward runs on the sm_100a kernels.  There is no CPU / eager fallback: calling it
    without a CUDA tensor or without the built library raises."""

    def _engine(self) -> Engine:
        params = dict(self.named_parameters())
        if not params:
            raise CtuError("module has no parameters")
        dev = next(iter(params.values())).device
        if dev.type != "cuda":
            raise CtuError("ctunet_b200 modules run only on a CUDA (sm_100) device; move the module with .cuda()")
        eng = getattr(self, "_eng", None)
        if eng is None or eng.dev != dev or eng.w.params.keys() != params.keys():
            eng = Engine(params, dev)
            object.__setattr__(self, "_eng", eng)
        else:
            eng.w.params = params
        return eng

    def _input(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise CtuError("input must be a CUDA tensor (no CPU fallback)")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise CtuError("the backward path is not implemented yet: call under torch.no_grad() / model.eval() "
                           "with torch.inference_mode()")
        return x.float().contiguous()
