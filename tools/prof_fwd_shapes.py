"""Forward projector variants on the sparse-view shapes of the sweep (heuristic tuning)."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pd_unet_b200 as pdu
dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return statistics.median(ts) * 1e3
for N, A in ((128, 64), (128, 128), (256, 64), (256, 128), (256, 192), (256, 256), (512, 64), (512, 128), (512, 256), (512, 384), (512, 512), (1024, 64), (1024, 256), (1024, 512), (1024, 768), (1024, 1024)):
    op = pdu.Radon(N, np.linspace(0, np.pi, A, endpoint=False))
    x = torch.rand(8, N, N, device=dev)
    row = []
    for v in (-1, 1, 13, 9, 11):
        pdu.set_option("radon_fwd_variant", v)
        row.append(f"v{v}: {timed(lambda: op._project(x)):8.1f}")
    pdu.set_option("radon_fwd_variant", -1)
    drift = np.pi / A * 0.7072 * N
    print(f"N={N:5d} A={A:5d} drift={drift:5.1f}  " + "  ".join(row), flush=True)
