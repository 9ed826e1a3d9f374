"""Shared plumbing of the drop-in modules: engine construction, layout conversion at block boundaries and the
autograd bridge that makes every module trainable (trainer_CTUNet.py:87-109 calls loss.backward() through them)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

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


class Program:
    """What one module forward hands to the autograd bridge.

    in_acts : per module input, the engine-side tensor its gradient is read from (None: no gradient, e.g. the image)
    out_acts: engine-side output tensors (channels-last bf16 maps, or tensors returned as they are)
    outputs : what the caller receives, one per out_act (from_cl(out_act) when `converted`)
    """

    def __init__(self, in_acts: Sequence[Optional[torch.Tensor]], out_acts: Sequence[torch.Tensor], converted: bool = True):
        self.in_acts = list(in_acts)
        self.out_acts = list(out_acts)
        self.converted = converted
        # tensors handed to the caller are never the engine-side objects the tape closures hold: autograd attaches its
        # node to the returned objects, and node -> ctx -> tape -> closure -> tensor -> node would be a reference cycle
        # that keeps the whole step (and its AccumulateGrad nodes) alive until the garbage collector runs
        self.outputs = [from_cl(o) for o in out_acts] if converted else [o.detach() for o in out_acts]


class _EngineFn(torch.autograd.Function):
    """Runs a module's engine program with the tape on; backward replays the tape (Engine.backward)."""

    @staticmethod
    def forward(ctx, module, n_in, *tensors):
        eng = module._engine()
        eng.begin_training_forward()
        try:
            prog = module._program(eng, *[t.detach() for t in tensors[:n_in]])
        except Exception:
            eng.tape = None
            raise
        ctx.module, ctx.eng, ctx.tape, ctx.prog, ctx.n_in = module, eng, eng.tape, prog, n_in
        ctx.stats = eng.stats
        ctx.in_shapes = [t.shape for t in tensors[:n_in]]
        ctx.names = [n for n, _ in module._graph_parameters()]
        eng.tape = None  # a later inference call on the same module must not extend this tape
        outs = tuple(prog.outputs)
        prog.outputs = None
        return outs

    @staticmethod
    def backward(ctx, *grad_outputs):
        eng, prog = ctx.eng, ctx.prog
        eng.tape, eng.stats = ctx.tape, ctx.stats
        seeds = []
        for act, g in zip(prog.out_acts, grad_outputs):
            if g is not None and prog.converted:
                g = to_cl(g)
            seeds.append((act, g))
        pgrads, igrads = eng.backward(seeds, want=prog.in_acts)
        outs: List[Optional[torch.Tensor]] = [None, None]
        for act, g, shp, need in zip(prog.in_acts, igrads, ctx.in_shapes, ctx.needs_input_grad[2:2 + ctx.n_in]):
            if g is None or not need:
                outs.append(None)
            elif g.dim() == 5 and len(shp) == 5:  # channels-last activation gradient -> NCDHW fp32
                outs.append(from_cl(g))
            else:
                outs.append(g.float().reshape(shp))
        for name in ctx.names:
            outs.append(pgrads.get(name))
        ctx.tape = ctx.prog = None
        return tuple(outs)


class KernelModule(nn.Module):
    """nn.Module whose forward (and backward) runs on the sm_100a kernels.  There is no CPU / eager fallback:
    calling it without a CUDA tensor or without the built library raises."""

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

    def _graph_parameters(self):
        """(name, parameter) pairs that take part in the autograd graph: everything except the weights the forward never
        reads — `conv3` of a ResBlock with equal in/out channels (hybrid_CTUNet.py:88-91,100-102).  In the reference
        those are unreachable from the loss, so DistributedDataParallel(find_unused_parameters=True)
        (main_CTUNet.py:187-189) leaves their .grad None and AdamW never touches them; keeping them out of the graph
        here gives the same behaviour instead of a zero gradient (which would apply weight decay to them)."""
        dead = set()
        for mname, m in self.named_modules():
            if getattr(m, "downsample", None) is False and hasattr(m, "conv3"):
                dead.add((mname + "." if mname else "") + "conv3.conv.weight")
        return [(n, p) for n, p in self.named_parameters() if n not in dead]

    def invalidate_weight_cache(self):
        """Drop the packed bf16 copies of the parameters (they are rebuilt on the next call).  The cache notices
        parameter updates through Tensor._version (load_state_dict, copy_, non-fused optimizers); every training forward
        and every train()/eval() switch also drops it, because fused optimizers update parameters without bumping the
        version.  Call this after any other out-of-band in-place update of the weights between inference calls."""
        for m in self.modules():
            eng = getattr(m, "_eng", None)
            if eng is not None:
                eng.w.refresh_all()

    def train(self, mode: bool = True):
        eng = getattr(self, "_eng", None)
        if eng is not None:
            eng.w.refresh_all()
        return super().train(mode)

    def _input(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise CtuError("input must be a CUDA tensor (no CPU fallback)")
        return x.float().contiguous()

    def _program(self, eng: Engine, *inputs) -> Program:
        raise NotImplementedError

    def _call(self, *inputs: torch.Tensor) -> Tuple[torch.Tensor, ...]:
        """Run the module's program: through the autograd bridge when a gradient may be needed, plainly otherwise."""
        for x in inputs:
            if not x.is_cuda:
                raise CtuError("input must be a CUDA tensor (no CPU fallback)")
        needs = torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or
                                             any(x.requires_grad for x in inputs))
        if needs:
            return _EngineFn.apply(self, len(inputs), *inputs, *[p for _, p in self._graph_parameters()])
        eng = self._engine()
        eng.tape = None
        with torch.no_grad():
            return tuple(self._program(eng, *inputs).outputs)
