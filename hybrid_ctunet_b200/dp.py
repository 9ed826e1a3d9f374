"""Data-parallel gradient exchange of the CTUNet training step.

The reference wraps the model in DistributedDataParallel(find_unused_parameters=True) (main_CTUNet.py:187-189),
which all-reduces the fp32 gradients in 25 MB buckets while the backward runs.  Here the backward hands back every
parameter gradient at once (one tape replay) as slices of ONE flat buffer, so the exchange is ONE fp32 all-reduce (mean)
of that buffer, in place, over NVLink per step (gradients that live elsewhere are gathered into a flat buffer first):
about 0.7 GB, ~2 ms at the measured all-reduce bandwidth against ~100 ms of compute.  Parameters whose gradient is
None (the seven never-used conv3 weights) are skipped and stay None on every rank, as under DDP, so AdamW keeps
skipping them.  Works over any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class GradientAllReduce:
    def __init__(self, params: Iterable[torch.nn.Parameter], group: Optional[dist.ProcessGroup] = None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self._flat: Optional[torch.Tensor] = None

    @staticmethod
    def _span_view(grads: List[torch.Tensor]) -> Optional[torch.Tensor]:
        """One flat fp32 view covering every gradient when they are slices of ONE buffer (the engine hands them out as
        16-byte-aligned slices of its flat unpack buffer), else None.  The view also covers the alignment gaps between
        slices (at most 3 floats each); whatever they hold is reduced along and never read."""
        st = grads[0].untyped_storage()
        base = st.data_ptr()
        lo, hi, covered = None, 0, 0
        for g in grads:
            if g.dtype != torch.float32 or not g.is_contiguous() or g.untyped_storage().data_ptr() != base:
                return None
            a = g.storage_offset()
            lo = a if lo is None else min(lo, a)
            hi = max(hi, a + g.numel())
            covered += g.numel()
        if hi - lo > covered + 4 * len(grads):   # slices of a larger tensor with real data in between: do not touch it
            return None
        return torch.empty(0, dtype=torch.float32, device=grads[0].device).set_(st, lo, (hi - lo,))

    def _all_reduce_mean(self, flat: torch.Tensor):
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.mul_(1.0 / dist.get_world_size(self.group))

    def reduce_and_step(self, optimizer, chunks: int = 4) -> int:
        """All-reduce (mean) + optimizer update, pipelined: the flat gradient buffer is exchanged in `chunks` pieces cut
        at parameter boundaries; all pieces are queued on NCCL's stream at once and the parameters of piece i are updated
        (optimizer.step(only=...), hybrid_ctunet_b200.optim.AdamW) as soon as piece i has arrived, while piece i+1 is
        still on the wire.  The update of a parameter reads nothing but its own gradient, so the result equals
        reduce() followed by optimizer.step().  Optimizers without `only=` get the plain sequence."""
        import inspect
        can_subset = "only" in inspect.signature(optimizer.step).parameters
        with_grad = [p for p in self.params if p.grad is not None]
        if not with_grad:
            return 0
        # the gradients of a graph-replayed step are the same tensors every step: plan the pieces once
        sig = (tuple(p.grad.data_ptr() for p in with_grad), chunks, id(optimizer))
        plan = getattr(self, "_plan", None)
        if plan is not None and plan[0] == sig:
            return self._run_plan(plan, optimizer)
        by_storage = {}
        for p in with_grad:
            by_storage.setdefault(p.grad.untyped_storage().data_ptr(), []).append(p)
        main = max(by_storage.values(), key=lambda ps: sum(p.grad.numel() for p in ps))
        span = self._span_view([p.grad for p in main]) if len(main) > 1 else None
        if span is None or not can_subset or chunks <= 1:
            n = self.reduce()
            optimizer.step()
            return n
        # pieces of the flat span cut at gradient boundaries, about equal in bytes
        main = sorted(main, key=lambda p: p.grad.storage_offset())
        lo = span.storage_offset()
        total = span.numel()
        pieces, cur, start = [], [], lo
        for p in main:
            cur.append(p)
            end = p.grad.storage_offset() + p.grad.numel()
            if end - start >= total / chunks and len(pieces) < chunks - 1:
                pieces.append((start, end, cur))
                cur, start = [], end
        if cur:
            pieces.append((start, lo + total, cur))
        st = span.untyped_storage()
        views = [torch.empty(0, dtype=torch.float32, device=span.device).set_(st, a, (b - a,)) for a, b, _ in pieces]
        main_ids = {id(p) for p in main}
        rest = [p for p in with_grad if id(p) not in main_ids]
        plan = (sig, views, [ps for _, _, ps in pieces], rest, sum(p.grad.numel() for p in with_grad))
        self._plan = plan
        return self._run_plan(plan, optimizer)

    def _run_plan(self, plan, optimizer) -> int:
        _, views, piece_params, rest, n = plan
        nccl = dist.get_backend(self.group) == "nccl"
        works = [dist.all_reduce(v, op=dist.ReduceOp.AVG if nccl else dist.ReduceOp.SUM, group=self.group, async_op=True)
                 for v in views]
        for i, (ps, wk) in enumerate(zip(piece_params, works)):
            wk.wait()                                   # the current stream waits for piece i only
            if not nccl:
                views[i].mul_(1.0 / dist.get_world_size(self.group))
            optimizer.step(only=ps, key=("dp", i, len(views)))
        if rest:                                        # gradients living outside the flat buffer (relative-position tables)
            m = sum(p.grad.numel() for p in rest)
            if self._flat is None or self._flat.numel() != m or self._flat.device != rest[0].grad.device:
                self._flat = torch.empty(m, dtype=torch.float32, device=rest[0].grad.device)
            fviews, off = [], 0
            for p in rest:
                fviews.append(self._flat[off:off + p.grad.numel()].view(p.grad.shape))
                off += p.grad.numel()
            torch._foreach_copy_(fviews, [p.grad for p in rest])
            self._all_reduce_mean(self._flat)
            torch._foreach_copy_([p.grad for p in rest], fviews)
            optimizer.step(only=rest, key=("dp", "rest"))
        return n

    def reduce(self) -> int:
        """Average the existing .grad tensors over the group in place; returns the number of elements exchanged."""
        grads = [p.grad for p in self.params if p.grad is not None]
        if not grads:
            return 0
        n = sum(g.numel() for g in grads)
        # gradients that share the most common buffer (the engine's flat unpack buffer) are reduced where they are; the
        # few that live elsewhere (relative-position tables, repeated parameters) are gathered into a small flat buffer
        by_storage = {}
        for g in grads:
            by_storage.setdefault(g.untyped_storage().data_ptr(), []).append(g)
        main = max(by_storage.values(), key=lambda gs: sum(g.numel() for g in gs))
        span = self._span_view(main) if len(main) > 1 else None
        if span is not None:
            self._all_reduce_mean(span)
            ids = {id(g) for g in main}
            grads = [g for g in grads if id(g) not in ids]
            if not grads:
                return n
        m = sum(g.numel() for g in grads)
        if self._flat is None or self._flat.numel() != m or self._flat.device != grads[0].device:
            self._flat = torch.empty(m, dtype=torch.float32, device=grads[0].device)
        views, off = [], 0
        for g in grads:
            views.append(self._flat[off:off + g.numel()].view(g.shape))
            off += g.numel()
        torch._foreach_copy_(views, grads)
        self._all_reduce_mean(self._flat)
        torch._foreach_copy_(grads, views)
        return n
