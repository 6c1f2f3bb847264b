"""Multi-GPU plumbing: the path shards by slice / sinogram batch.

Inference: every rank reconstructs its own contiguous share of the batch -- no data-path collective.
Training: stock DistributedDataParallel; the operators own no parameters, so the only collective is
DDP's bucketed gradient all-reduce over NCCL (NVLink 5 / NVSwitch).  One process per GPU, launched
by torchrun; everything also runs on the gloo backend on CPU for the host-side tests.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def init_distributed(backend: str | None = None, graph_capture: bool = False) -> Tuple[int, int, int]:
    """Initialise torch.distributed from the torchrun environment.  -> (rank, world_size, local_rank).
    A plain `python script.py` run (no RANK in the environment) is world_size 1 and touches nothing.
    graph_capture: the training step will be captured into a CUDA graph together with DDP's all-reduce
    (pd_unet_b200.graph.GraphedTrainingStep): NCCL's asynchronous error handling (a watchdog that polls the
    communicator from another thread) has to be off before the process group exists."""
    if graph_capture:
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
    if "RANK" not in os.environ or int(os.environ.get("WORLD_SIZE", "1")) == 1:
        return 0, 1, int(os.environ.get("LOCAL_RANK", "0"))
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [start, stop) share of n_items for `rank`; the first n_items % world_size ranks
    take one extra item, so shares differ by at most one and cover every item exactly once."""
    if n_items < 0 or world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad shard request")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(t: torch.Tensor, rank: int, world_size: int) -> torch.Tensor:
    a, b = shard_range(t.shape[0], rank, world_size)
    return t[a:b]


def max_over_ranks(value: float, device=None) -> float:
    """Max of a host scalar over all ranks (the timing rule: a step takes as long as its slowest rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier() -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def gather_batch(local: torch.Tensor, n_items: int) -> torch.Tensor | None:
    """Collect every rank's share on rank 0 (evaluation only; not part of the timed path)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_items, r, world) for r in range(world)]
    pad = max(b - a for a, b in sizes)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[:local.shape[0]] = local
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    if rank != 0:
        return None
    return torch.cat([o[:b - a] for o, (a, b) in zip(outs, sizes)], dim=0)


def wrap_ddp(model: torch.nn.Module, local_rank: int, graph_capture: bool = False) -> torch.nn.Module:
    """DistributedDataParallel with bucket views (no extra gradient copy); identity at world_size 1.
    graph_capture: construct the wrapper on a side stream, as whole-step CUDA-graph capture of a DDP model needs
    (PyTorch CUDA-graphs notes; GraphedTrainingStep then runs the 11 eager warm-up steps DDP asks for)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return model
    from torch.nn.parallel import DistributedDataParallel as DDP
    if next(model.parameters()).is_cuda:
        if graph_capture:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                ddp = DDP(model, device_ids=[local_rank], gradient_as_bucket_view=True)
            torch.cuda.current_stream().wait_stream(side)
            return ddp
        return DDP(model, device_ids=[local_rank], gradient_as_bucket_view=True)
    return DDP(model, gradient_as_bucket_view=True)
