"""Adjoint NUFFT, generic (sorted gather + FFT passes) against fused (row-binned) path as a function of the plane count.
   python tools/prof_nufft_adj_policy.py"""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pd_unet_b200 as pdu
dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def traj(spokes, readout):
    phi = np.arange(spokes) * (111.246117975 * np.pi / 180.0)
    r = (np.arange(readout) - readout / 2) * (2 * np.pi / readout)
    return torch.from_numpy(np.stack([(r[None] * np.sin(phi)[:, None]).reshape(-1), (r[None] * np.cos(phi)[:, None]).reshape(-1)]).astype(np.float32)).to(dev)
def timed(fn, reps=int(os.environ.get("REPS", "7"))):
    [fn() for _ in range(5)]; torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)
cases = ((128, 32), (256, 48), (320, 48), (512, 96), (1024, 128))
plane_counts = (8, 16, 24, 32, 48, 64)
if len(sys.argv) > 1 and sys.argv[1] == "dense":          # 0.15 < rho <= 0.3, few planes: fused against the sorted gather
    cases, plane_counts = ((128, 64), (256, 128), (320, 160), (512, 256)), (8, 12, 16)
for n, spokes in cases:
    for planes in plane_counts:
        if n * n * planes > 512 * 512 * 64:
            continue
        om = traj(spokes, 2 * n)
        ad = pdu.KbNufftAdjoint((n, n))
        k = torch.randn(planes, 1, om.shape[1], 2, device=dev)
        k = torch.view_as_complex(k)
        row = []
        for mode in (False, True, "auto"):
            ad._plan.use_fused = mode
            row.append(timed(lambda: ad(k, om)))
        print(f"N {n:4d} spokes {spokes:4d} planes {planes:3d}: adj generic {row[0]:8.1f} fused {row[1]:8.1f} auto {row[2]:8.1f} us {'  <-- auto not best' if row[2] > 1.03 * min(row[:2]) else ''}", flush=True)
