"""Multi-GPU training smoke: PD-UNet CT under DistributedDataParallel over NCCL.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_smoke.py
Each rank trains on its own shard of a synthetic batch; the only collective is DDP's gradient all-reduce.
Checks: parameters stay identical across ranks, the loss falls, and the step is timed (max over ranks)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import pd_unet_b200 as pdu
from pd_unet_b200 import parallel
from pd_unet_b200.model import PrimalDualUNetCT
from pd_unet_b200.phantoms import phantom_batch

rank, world, local = parallel.init_distributed("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
n, A, up, B = 128, 128, 8, 8                      # per-rank batch 8 (weak scaling)
radon = pdu.Radon(n, np.linspace(0, np.pi, A, endpoint=False))
torch.manual_seed(0)
model = PrimalDualUNetCT(radon, upsample=up, n_iter=2, n_primal=4, n_dual=4, unet_base=16, unet_depth=2, dual_features=16).to(dev)
ddp = parallel.wrap_ddp(model, local)
opt = torch.optim.Adam(ddp.parameters(), 1e-3)
x_all = phantom_batch(B * world, n, seed=7)
x = parallel.shard_batch(x_all, rank, world).to(dev)
sparse = radon.forward(x)[:, None, ::up].contiguous()
losses, times = [], []
for it in range(8):
    torch.cuda.synchronize(); parallel.barrier(); t0 = time.perf_counter()
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.mse_loss(ddp(sparse)[:, 0], x)
    loss.backward()
    opt.step()
    torch.cuda.synchronize(); times.append(time.perf_counter() - t0)
    losses.append(parallel.sum_over_ranks(float(loss.detach()), dev) / world)
flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
ref = flat.clone()
if world > 1:
    dist.broadcast(ref, src=0)
same = bool(torch.equal(flat, ref))
step_ms = parallel.max_over_ranks(sorted(times[2:])[len(times[2:]) // 2] * 1e3, dev)
if rank == 0:
    print(f"ddp_smoke world={world} loss {losses[0]:.5f} -> {losses[-1]:.5f}  params_identical={same}  "
          f"step {step_ms:.1f} ms  ({B * world / step_ms * 1e3:.0f} slices/s training, 128^2, 2 iterations)", flush=True)
assert same and losses[-1] < losses[0]
if world > 1:
    dist.destroy_process_group()
