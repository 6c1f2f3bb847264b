import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pd_unet_b200 as pdu
from pd_unet_b200 import _lib
dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)
N, A, B = 256, 512, 16
x = torch.rand(B, N, N, device=dev)
for sp in (2.0, 1.0, 0.75, 0.5, 0.4, 0.25):
    D = int(round(N / sp / 32)) * 32
    op = pdu.Radon(N, np.linspace(0, np.pi, A, endpoint=False), det_count=D, det_spacing=sp)
    s = op._project(x)
    tf = timed(lambda: op._project(x)); kf = _lib.last_kernel("radon_fwd")[:60]
    ta = timed(lambda: op._backproject(s)); ka = _lib.last_kernel("radon_adj")[:60]
    print(f"spacing {sp:5.2f} det {D:5d}: fwd {tf:8.1f} us ({1e-6*B*A*D*N/tf:6.3f} T samp/s) {kf} | adj {ta:8.1f} us ({1e-6*B*A*N*N/ta:6.3f} T taps/s) {ka}", flush=True)
