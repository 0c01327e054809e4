"""Whole training step (forward, loss, backward, optimizer) captured in ONE CUDA graph.

The reference trainer (`src/trainer/deepfm.py:44-62`) launches ~70 small kernels per step from Python; at its
yaml batch (2048) a B200 finishes each of them in a few microseconds, so the step is launch-latency-bound.
Every kernel of this library takes its stream from torch, allocates through torch's caching allocator and keeps
no host-side state that changes between steps (the dropout stream position lives in device memory, see
`linalg.use_device_dropout_counter`), so the step can be captured once and replayed:

    step = GraphedTrainStep(model, optimizers, criterion, x0, y0)     # optimizers built with capturable=True
    for x, y in loader:                                              # host or device tensors
        loss = step(x, y)                                            # copies into static buffers, one graph launch

Constraints: fixed batch shape; dense gradients (torch's sparse optimizers are not capturable); optimizers that
support capture (`get_optimizers(..., {"capturable": True, "fused_adam": True})`).
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch

from . import linalg as LA


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, optimizers: Sequence[torch.optim.Optimizer], criterion: Callable,
                 example_inputs: torch.Tensor, example_labels: torch.Tensor, warmup: int = 3,
                 after_backward: Callable[[], None] = None):
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedTrainStep needs the model on a CUDA device")
        self.model, self.optimizers, self.criterion = model, list(optimizers), criterion
        self.after_backward = after_backward
        self.inputs = example_inputs.to(dev).clone()
        self.labels = example_labels.to(dev).clone()
        LA.use_device_dropout_counter(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):           # allocator warm-up + lazy optimizer state, outside the graph
                self._eager_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            LA.advance_dropout_counter()
            self.loss = self._eager_step()

    def _eager_step(self) -> torch.Tensor:
        logits = self.model(self.inputs)
        loss = self.criterion(logits, self.labels)
        for o in self.optimizers:
            o.zero_grad(set_to_none=True)
        loss.backward()
        if self.after_backward is not None:
            self.after_backward()
        for o in self.optimizers:
            o.step()
        return loss

    def __call__(self, inputs: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """Copy one batch (host or device) into the static buffers and replay; returns the static loss tensor
        (valid until the next call)."""
        self.inputs.copy_(inputs, non_blocking=True)
        self.labels.copy_(labels, non_blocking=True)
        self.graph.replay()
        return self.loss
