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
n, planes = 320, 16
g = torch.Generator().manual_seed(0)
cases = {"radial 48 spokes": traj(48, 640),
         "uniform random": ((torch.rand(2, 30720, generator=g) * 2 - 1) * np.pi).to(dev),
         "random in |k| < pi/4": ((torch.rand(2, 30720, generator=g) * 2 - 1) * np.pi / 4).to(dev),
         "random in |k| < pi/16": ((torch.rand(2, 30720, generator=g) * 2 - 1) * np.pi / 16).to(dev),
         "cartesian 48 lines": torch.stack([torch.linspace(-np.pi, np.pi * (1 - 2 / 48), 48).repeat_interleave(640), torch.linspace(-np.pi, np.pi * (1 - 2 / 640), 640).repeat(48)]).to(dev),
         "cartesian 48 columns": torch.stack([torch.linspace(-np.pi, np.pi * (1 - 2 / 640), 640).repeat(48), torch.linspace(-np.pi, np.pi * (1 - 2 / 48), 48).repeat_interleave(640)]).to(dev)}
for name, om in cases.items():
    om = om.contiguous().float()
    fw, ad = pdu.KbNufft((n, n)), pdu.KbNufftAdjoint((n, n))
    img = torch.randn(planes, 1, n, n, dtype=torch.complex64, device=dev)
    k = fw(img, om)
    row = []
    for mode in (False, True, "auto"):
        fw._plan.use_fused = mode; ad._plan.use_fused = mode
        row.append((timed(lambda: fw(img, om)), timed(lambda: ad(k, om))))
    ent = fw._plan._entry(om)
    print(f"{name:24s} max_row {ent.get('max_row')}: fwd generic {row[0][0]:8.1f} fused {row[1][0]:8.1f} auto {row[2][0]:8.1f} | adj generic {row[0][1]:8.1f} fused {row[1][1]:8.1f} auto {row[2][1]:8.1f}", flush=True)
