"""Experiment: the CT PD-UNet pass over 16 slices as ONE graph of S independent sub-batches on S streams, so that the
issue-bound projector kernels of one sub-batch overlap the HBM-bound epilogues / convolutions of another (and every
kernel's tail wave is filled by the other stream's work).
   python tools/exp_two_streams.py"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
torch.backends.cudnn.benchmark = True
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
radon, model = bench.build_model(dev)
B = 16
sparse = bench.synthetic_sparse_sinograms(radon, dev, B, seed=100)


def capture(n_streams):
    parts = [p.contiguous() for p in sparse.chunk(n_streams, dim=0)]
    with torch.no_grad():
        for _ in range(3):
            for p in parts:
                model(p)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams - 1)]
    g = torch.cuda.CUDAGraph()
    outs = [None] * n_streams
    with torch.no_grad(), torch.cuda.graph(g):
        main = torch.cuda.current_stream(dev)
        for s in streams:
            s.wait_stream(main)
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                outs[i + 1] = model(parts[i + 1])
        outs[0] = model(parts[0])
        for s in streams:
            main.wait_stream(s)
    return g, outs


ref = None
for S in (1, 2, 4):
    g, outs = capture(S)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    out = torch.cat(outs, 0)
    if ref is None:
        ref = out.clone()
    err = float((out - ref).norm() / ref.norm())
    ts = []
    for _ in range(12):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = statistics.median(ts)
    print(f"streams {S}: {ms:8.3f} ms / pass of {B}   {B / ms * 1e3:8.1f} slices/s   rel diff vs 1 stream {err:.2e}", flush=True)
