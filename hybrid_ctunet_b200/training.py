"""CUDA-graph replay of the training step's forward + loss + backward.

One CTUNet training step is ~1,900 kernel launches, many of them a few microseconds long (the ViT and the deep
ResNet stages); issued eagerly from Python they leave the GPU idle ~13 % of the step.  `GraphedTrainStep` captures
forward, loss and backward (including the per-step re-packing of the updated weights) once and replays them as a
single graph — the same mechanism `model.enable_cuda_graph()` offers for inference.  The optimizer (and, under data
parallelism, the gradient all-reduce) stay outside the graph: they read the static `.grad` tensors the replay fills.

    step = GraphedTrainStep(model, lambda logits, y: ctunet_loss(logits, y, loss_func), x, y)
    for x, y in loader:
        loss = step(x, y)          # copies into the static inputs, replays; returns the static loss tensor
        optimizer.step()           # never set .grad to None afterwards: the gradients are the graph's outputs
"""
from __future__ import annotations

import gc
from typing import Callable

import torch


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, loss_fn: Callable, example_x: torch.Tensor, example_y: torch.Tensor,
                 warmup: int = 2):
        self.model = model
        self.x = example_x.detach().clone()
        self.y = example_y.detach().clone()
        params = [p for p in model.parameters() if p.requires_grad]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):  # sets kernel attributes, sizes the arenas, primes the allocator
                for p in params:
                    p.grad = None
                loss_fn(model(self.x), self.y).backward()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for p in params:
            p.grad = None
        gc.collect()  # no autograd graph of an earlier eager step (with AccumulateGrad nodes on the default stream) survives
        model._engine().prepare_for_capture()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = loss_fn(model(self.x), self.y)
            self.loss.backward()

    def __call__(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        return self.loss
