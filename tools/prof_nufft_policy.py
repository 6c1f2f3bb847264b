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
def timed(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)
for planes in (8, 16, 64):
    for n in (128, 256, 320, 512):
        for spokes in (48, 64, 256, 1024):
            om = traj(spokes, 2 * n)
            fw, ad = pdu.KbNufft((n, n)), pdu.KbNufftAdjoint((n, n))
            img = torch.randn(planes, 1, n, n, dtype=torch.complex64, device=dev)
            k = fw(img, om)
            row = []
            for mode in (False, True, "auto"):
                fw._plan.use_fused = mode; ad._plan.use_fused = mode
                row.append((timed(lambda: fw(img, om)), timed(lambda: ad(k, om))))
            fw._plan.use_fused = "auto"; ad._plan.use_fused = "auto"
            print(f"planes {planes} N {n:4d} spokes {spokes:5d}: fwd generic {row[0][0]:8.1f} fused {row[1][0]:8.1f} | adj generic {row[0][1]:8.1f} fused {row[1][1]:8.1f} | auto {row[2][0]:8.1f} {row[2][1]:8.1f}", flush=True)
