"""Experiment: CT PD-UNet step time per slice as a function of the batch per pass (L2 residency of the UNet activations).
   python tools/exp_batch_chunks.py"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pd_unet_b200.graph import GraphedInference

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
torch.backends.cudnn.benchmark = True
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
radon, model = bench.build_model(dev)
for B in (16, 8, 4, 2):
    sparse = bench.synthetic_sparse_sinograms(radon, dev, B, seed=100)
    with torch.no_grad():
        for _ in range(3):
            model(sparse)
    g = GraphedInference(model, sparse, warmup=2)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = statistics.median(ts)
    print(f"batch {B:3d}: {ms:8.3f} ms / pass   {ms / B * 1e3:8.1f} us / slice   {B / ms * 1e3:8.1f} slices/s", flush=True)
