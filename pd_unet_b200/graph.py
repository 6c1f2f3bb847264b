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


class GraphedTrainingStep:
    """Captures one whole training step -- forward, loss, backward, optimizer.step() -- and replays it.

    An eager PD-UNet training step is ~1000 launches (cuDNN fprop / dgrad / wgrad, the operators and their
    adjoints, autograd glue); at a few slices per GPU the host cannot issue them as fast as the GPU retires
    them.  Replaying a captured graph removes that bound (and with it the size gate under which the fused
    differentiable epilogues of pd_unet_b200.updates do not pay).

    model:      the module to train (a DistributedDataParallel wrapper works: PyTorch's whole-network capture
                needs NCCL >= 2.9.6 and 11 eager warm-up steps, which this class runs)
    optimizer:  must be capturable, e.g. torch.optim.Adam(params, lr, capturable=True)
    loss_fn:    (output, target) -> scalar tensor
    inputs:     tuple of example input tensors (shapes / dtypes / devices fixed from now on); non-tensor entries
                are passed through unchanged
    The warm-up steps are REAL optimisation steps on the example batch.
    """

    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, loss_fn: Callable, inputs: tuple,
                 target: torch.Tensor, warmup: int = 3):
        if not target.is_cuda:
            raise ValueError("GraphedTrainingStep needs CUDA tensors")
        self.model, self.optimizer, self.loss_fn = model, optimizer, loss_fn
        self.static_inputs = tuple(t.clone() if isinstance(t, torch.Tensor) else t for t in inputs)
        self.static_target = target.clone()
        dev = target.device
        is_ddp = isinstance(model, torch.nn.parallel.DistributedDataParallel)
        n_warm = max(warmup, 11) if is_ddp else max(warmup, 1)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(n_warm):
                optimizer.zero_grad(set_to_none=True)
                loss = loss_fn(model(*self.static_inputs), self.static_target)
                loss.backward()
                optimizer.step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.static_loss = loss_fn(model(*self.static_inputs), self.static_target)
            self.static_loss.backward()
            optimizer.step()

    def __call__(self, inputs: tuple, target: torch.Tensor) -> torch.Tensor:
        """Copies the batch into the captured tensors, replays the step, returns the (captured) loss tensor."""
        for dst, src in zip(self.static_inputs, inputs):
            if isinstance(dst, torch.Tensor) and src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        if target.data_ptr() != self.static_target.data_ptr():
            self.static_target.copy_(target, non_blocking=True)
        self.graph.replay()
        return self.static_loss
