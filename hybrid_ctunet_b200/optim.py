"""Fused AdamW for the CTUNet training step.

The reference builds `torch.optim.AdamW(model.parameters(), lr=args.optim_lr, weight_decay=args.reg_weight)`
(main_CTUNet.py:190-193) and calls `optimizer.step()` after `loss.backward()` (trainer_CTUNet.py:106-109).  This class
is a drop-in for that optimizer: same constructor arguments, same update (decoupled weight decay, bias-corrected
moments, no amsgrad), same `state_dict()` layout (`step`, `exp_avg`, `exp_avg_sq` per parameter, so checkpoints move
between the two), LR schedulers work through `param_groups[i]["lr"]`.  The update itself is ONE launch of
`ctu_adamw_step` per parameter group over a device-resident item table; parameters whose `.grad` is None (CTUNet's seven
never-used conv3 weights) are skipped exactly as torch skips them.  CUDA fp32 parameters only; there is no CPU path.
"""
from __future__ import annotations

from typing import Iterable, Tuple

import numpy as np
import torch

from . import lib as _lib
from .lib import check

_ITEM = np.dtype([("param", "u8"), ("grad", "u8"), ("exp_avg", "u8"), ("exp_avg_sq", "u8"), ("numel", "i8"), ("unit0", "i8")])


class AdamW(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, amsgrad: bool = False):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if eps < 0.0:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 0: {betas[0]}")
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 1: {betas[1]}")
        if weight_decay < 0.0:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        if amsgrad:
            raise NotImplementedError("amsgrad is not part of the reference's training recipe")
        # the remaining keys are the ones torch.optim.AdamW keeps in its param_groups (with its defaults), so that a
        # state_dict written here configures torch's optimizer identically (decoupled_weight_decay above all)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False,
                                      maximize=False, foreach=None, capturable=False, differentiable=False, fused=None,
                                      decoupled_weight_decay=True))
        self._tables = {}  # group index -> (signature, device table, units)

    def _table(self, gi: int, entries, cache: bool = True):
        sig = tuple((p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel()) for p, g, m, v in entries)
        hit = self._tables.get(gi) if cache else None
        if hit is not None and hit[0] == sig:
            return hit[1], hit[2]
        arr = np.zeros(len(entries), dtype=_ITEM)
        unit = 0
        for i, (pp, gp, mp, vp, n) in enumerate(sig):
            arr[i] = (pp, gp, mp, vp, n, unit)
            unit += -(-n // 1024)
        table = torch.from_numpy(arr.view(np.uint8).copy()).to(entries[0][0].device)
        if cache:
            self._tables[gi] = (sig, table, unit)
        return table, unit

    @torch.no_grad()
    def step(self, closure=None, only=None, key=None):
        """`only` (optional): restrict the update to these parameters — the data-parallel step updates the parameters of
        one all-reduced gradient chunk while the next chunk is still on the wire (dp.GradientAllReduce.reduce_and_step);
        `key` names the subset for the cached device table."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.require_device()
        subset = None if only is None else {id(p) for p in only}
        for gi, group in enumerate(self.param_groups):
            buckets = {}  # step count -> [(p, grad, exp_avg, exp_avg_sq)]: one launch per distinct count (normally one)
            for p in group["params"]:
                if p.grad is None or (subset is not None and id(p) not in subset):
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                    raise RuntimeError("ctunet_b200 AdamW: CUDA fp32 dense parameters and gradients only (no CPU fallback)")
                if not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("ctunet_b200 AdamW: parameters and gradients must be contiguous")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)   # host tensor, as torch keeps it by default
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                buckets.setdefault(int(st["step"].item()), []).append((p, p.grad, st["exp_avg"], st["exp_avg_sq"]))
            if not buckets:
                continue
            b1, b2 = group["betas"]
            for s, ent in buckets.items():
                table, units = self._table((gi, key), ent, cache=len(buckets) == 1)
                stream = torch.cuda.current_stream(ent[0][0].device).cuda_stream
                check(lib.ctu_adamw_step(table.data_ptr(), len(ent), units, float(group["lr"]), float(b1), float(b2),
                                         float(group["eps"]), float(group["weight_decay"]), int(s), stream),
                      "ctu_adamw_step")
                # the kernel writes through raw pointers: tell autograd / the packed-weight caches that the parameters
                # changed (host-side counter bump, no kernel)
                torch.autograd.graph.increment_version([e[0] for e in ent])
        return loss
