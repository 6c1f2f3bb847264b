"""Adjoint NUFFT on the generic path: sorted (CSR) gather against the atomic scatter as a function of the plane count.
   python tools/prof_nufft_csr_policy.py"""
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
def timed(fn, reps=7):
    fn(); fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)
for n, spokes in ((128, 64), (256, 32), (256, 256), (320, 48), (512, 256), (1024, 512)):
    for planes in (1, 2, 4, 8, 12):
        if n * n * planes > 1024 * 1024 * 2:
            continue
        om = traj(spokes, 2 * n)
        ad = pdu.KbNufftAdjoint((n, n))
        ad._plan.use_fused = False
        k = torch.view_as_complex(torch.randn(planes, 1, om.shape[1], 2, device=dev))
        row = []
        for mode in (False, True, "auto"):
            ad._plan.use_csr = mode
            row.append(timed(lambda: ad(k, om)))
        print(f"N {n:4d} spokes {spokes:4d} planes {planes:3d}: adj atomics {row[0]:8.1f} sorted gather {row[1]:8.1f} auto {row[2]:8.1f} us {'  <-- auto not best' if row[2] > 1.03 * min(row[:2]) else ''}", flush=True)
