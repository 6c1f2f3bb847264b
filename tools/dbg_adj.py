import sys, os, statistics
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import pd_unet_b200 as pdu
from pd_unet_b200.phantoms import coil_maps
dev="cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def traj(spokes, readout):
    phi = np.arange(spokes) * (111.246117975 * np.pi / 180.0)
    r = (np.arange(readout) - readout / 2) * (2 * np.pi / readout)
    return torch.from_numpy(np.stack([(r[None] * np.sin(phi)[:, None]).reshape(-1), (r[None] * np.cos(phi)[:, None]).reshape(-1)]).astype(np.float32)).to(dev)
n, coils, B, spokes = 320, 8, 8, 48
om = traj(spokes, 2*n); M = om.shape[1]
ad = pdu.KbNufftAdjoint((n,n)); sm = coil_maps(coils, n)[None].to(dev)
k = torch.randn(B, coils, M, dtype=torch.complex64, device=dev)
def timed(name, fn):
    fn(); torch.cuda.synchronize(); ts=[]
    for _ in range(5):
        flush.zero_(); a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print(f"{name:40s} {statistics.median(ts)*1e3:8.1f} us", flush=True)
for dbg, label in ((0,"full"),(2,"no gather taps"),(4,"no staging"),(6,"no staging, no taps"),(8,"no FFT"),(14,"nothing but tables + stores")):
    pdu.set_option("debug_fault", dbg if dbg else -1)
    timed("adj " + label, lambda: ad(k, om, smaps=sm))
pdu.set_option("debug_fault", -1)
