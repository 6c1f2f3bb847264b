"""Runs each CT operator a few times at BASELINE configs[1] sizes (for ncu captures and quick timings)."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pd_unet_b200 as pdu

N, A, B = 256, 512, 16
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = "cuda:0"
op = pdu.Radon(N, np.linspace(0, np.pi, A, endpoint=False))
x = torch.rand(B, N, N, device=dev)
s = torch.rand(B, A, N, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(name, fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"{name:28s} median {statistics.median(ts)*1e3:9.1f} us  min {min(ts)*1e3:9.1f} us", flush=True)


for v in (1, 13):
    pdu.set_option("radon_fwd_variant", v)
    timed(f"radon_fwd variant {v}", lambda: op._project(x))
pdu.set_option("radon_fwd_variant", -1)
for v in (3,):
    pdu.set_option("radon_adj_variant", v)
    timed(f"radon_adj variant {v}", lambda: op._backproject(s))
pdu.set_option("radon_adj_variant", -1)
for v in (0, 1):
    pdu.set_option("filter_variant", v)
    timed(f"filter variant {v}", lambda: op._filter(s, "ramp"))
pdu.set_option("filter_variant", -1)
# cfg3 share: fan 512^2, 1024 views, batch 8
fan = pdu.RadonFanbeam(512, np.linspace(0, 2 * np.pi, 1024, endpoint=False), 1024.0)
xf = torch.rand(8, 512, 512, device=dev)
sf = torch.rand(8, 1024, 512, device=dev)
for v in (1, 13):
    pdu.set_option("radon_fwd_variant", v)
    timed(f"fan512 fwd variant {v}", lambda: fan._project(xf))
pdu.set_option("radon_fwd_variant", -1)
timed("fan512 fwd", lambda: fan._project(xf))
for v in (3,):
    pdu.set_option("radon_adj_variant", v)
    timed(f"fan512 adj variant {v}", lambda: fan._backproject(sf))
pdu.set_option("radon_adj_variant", -1)
for v in (0, 1):
    pdu.set_option("filter_variant", v)
    timed(f"fan512 filter variant {v}", lambda: fan._filter(sf, "ramp"))
pdu.set_option("filter_variant", -1)
# MRI cfg4 share: 320^2, 8 coils, 48 spokes, batch 2
from pd_unet_b200.phantoms import coil_maps
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
im = (320, 320)
phi = np.arange(48) * (111.246117975 * np.pi / 180.0)
r = (np.arange(640) - 320) * (2 * np.pi / 640)
om = torch.from_numpy(np.stack([(r[None] * np.sin(phi)[:, None]).reshape(-1), (r[None] * np.cos(phi)[:, None]).reshape(-1)]).astype(np.float32)).to(dev)
sm = coil_maps(8, 320)[None].to(dev)
img = torch.randn(2, 1, 320, 320, dtype=torch.complex64, device=dev)
fw, ad = pdu.KbNufft(im), pdu.KbNufftAdjoint(im)
k = fw(img, om, smaps=sm)
timed("nufft fwd 320 c8 b2", lambda: fw(img, om, smaps=sm))
timed("nufft adj 320 c8 b2", lambda: ad(k, om, smaps=sm))
