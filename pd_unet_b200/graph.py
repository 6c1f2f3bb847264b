"""CUDA-graph capture of an unrolled reconstruction pass.

One PD-UNet inference step is a few hundred short kernels (operators, fused updates, cuDNN
convolutions, PReLUs); launched eagerly, the gaps between them are a large part of the step at small
batch.  The operators enqueue on torch's current stream and allocate only through torch's caching
allocator, so the whole step can be captured once and replayed: the TMA descriptors, workspace
pointers and cuDNN algorithm choices are frozen into the graph.
"""
from __future__ import annotations

from typing import Callable

import torch


class GraphedInference:
    """Captures `fn(static_input)` under no_grad and replays it.

    fn:       callable on one CUDA tensor (e.g. a PrimalDualUNetCT in eval mode)
    example:  a tensor of the shape / dtype / device every later call will use
    """

    def __init__(self, fn: Callable[[torch.Tensor], torch.Tensor], example: torch.Tensor, warmup: int = 3):
        if not example.is_cuda:
            raise ValueError("GraphedInference needs a CUDA example input")
        self.fn = fn
        self.static_in = example.clone()
        side = torch.cuda.Stream(device=example.device)
        side.wait_stream(torch.cuda.current_stream(example.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):       # cuDNN autotuning, lazy plans, smem attributes: all before capture
                fn(self.static_in)
        torch.cuda.current_stream(example.device).wait_stream(side)
        torch.cuda.synchronize(example.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = fn(self.static_in)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """Copies x into the captured input (from host or device), replays, returns the captured output
        tensor (overwritten by the next call)."""
        if x.shape != self.static_in.shape or x.dtype != self.static_in.dtype:
            raise ValueError(f"captured for {tuple(self.static_in.shape)} {self.static_in.dtype}, "
                             f"got {tuple(x.shape)} {x.dtype}")
        if x.data_ptr() != self.static_in.data_ptr():
            self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        return self.static_out
