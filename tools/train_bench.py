"""Training-step timings (BASELINE.json configs[1] and configs[3] shapes) under DDP.
  python tools/train_bench.py                       # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_bench.py
  ... --graph: the whole step (DDP all-reduce included) captured once and replayed (GraphedTrainingStep)
One step = forward + MSE loss + backward + (DDP all-reduce) + Adam update; CUDA-event time, max over ranks."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pd_unet_b200 as pdu
from pd_unet_b200 import data, parallel
from pd_unet_b200.model import PrimalDualUNetCT, PrimalDualUNetMRI

rank, world, local = parallel.init_distributed("nccl", graph_capture="--graph" in sys.argv)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
torch.backends.cudnn.benchmark = True
KW = dict(n_iter=4, n_primal=4, n_dual=4, unet_base=32, unet_depth=3, dual_features=32)


GRAPH = "--graph" in sys.argv      # replay one captured training step (pd_unet_b200.graph.GraphedTrainingStep)


def run(name, model, step_inputs, target, per_rank, steps=8, warm=3):
    ddp = parallel.wrap_ddp(model, local, graph_capture=GRAPH)
    opt = torch.optim.Adam(ddp.parameters(), 1e-4, capturable=GRAPH)
    times, losses = [], []
    if GRAPH:
        from pd_unet_b200.graph import GraphedTrainingStep
        step = GraphedTrainingStep(ddp, opt, lambda o, t: (o - t).abs().pow(2).mean(), step_inputs, target, warmup=warm)
        for it in range(steps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            parallel.barrier(); torch.cuda.synchronize()
            a.record()
            loss = step(step_inputs, target)
            b.record(); torch.cuda.synchronize()
            times.append(a.elapsed_time(b))
            losses.append(float(loss.detach()))
        ms = parallel.max_over_ranks(statistics.median(times), dev)
        mem = torch.cuda.max_memory_allocated(dev) / 2 ** 30
        if rank == 0:
            print(f"| {name}, CUDA graph | {world} | {per_rank} | {ms:.1f} | {per_rank * world / ms * 1e3:.0f} | {mem:.1f} | {losses[0]:.4g} -> {losses[-1]:.4g} |", flush=True)
        return
    for it in range(warm + steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        parallel.barrier(); torch.cuda.synchronize()
        a.record()
        opt.zero_grad(set_to_none=True)
        out = ddp(*step_inputs)
        loss = (out - target).abs().pow(2).mean()
        loss.backward()
        opt.step()
        b.record(); torch.cuda.synchronize()
        if it >= warm:
            times.append(a.elapsed_time(b))
        losses.append(float(loss.detach()))
    ms = parallel.max_over_ranks(statistics.median(times), dev)
    mem = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    if rank == 0:
        print(f"| {name} | {world} | {per_rank} | {ms:.1f} | {per_rank * world / ms * 1e3:.0f} | {mem:.1f} | {losses[0]:.4g} -> {losses[-1]:.4g} |", flush=True)


if rank == 0:
    print("| workload | GPUs | slices / GPU | ms / step | slices/s | peak GiB | loss |\n|---|---:|---:|---:|---:|---:|---|", flush=True)
# CT, configs[1] shape: 256^2, 64 -> 512 views
radon = pdu.Radon(256, np.linspace(0, np.pi, 512, endpoint=False))
torch.manual_seed(0)
ct = data.make_ct_batch(radon, 8, 8, seed=rank, device=dev)
run("CT PD-UNet 256^2, 64->512 views (train)", PrimalDualUNetCT(radon, upsample=8, **KW).to(dev), (ct["sino_sparse"],), ct["image"], 8)
del ct
torch.cuda.empty_cache()
# MRI, configs[3] shape: 320^2, 8 coils, 48 spokes
torch.manual_seed(0)
mri = data.make_mri_batch((320, 320), 48, 8, 2, seed=rank, device=dev)
m = PrimalDualUNetMRI((320, 320), 48, 640, coils=8, n_iter=4, n_primal=4, n_dual=16, unet_base=32, unet_depth=3, dual_features=32).to(dev)
run("MRI PD-UNet 320^2, 8 coils, 48 spokes (train)", m, (mri["kdata"], mri["omega"], mri["smaps"], mri["dcf"]), mri["image"], 2)
if world > 1:
    torch.distributed.destroy_process_group()
