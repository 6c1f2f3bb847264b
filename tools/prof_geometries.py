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
print("filter, 8192 rows")
for D in (256, 320, 363, 384, 500, 512, 736, 768):
    op = pdu.Radon(N, np.linspace(0, np.pi, A, endpoint=False), det_count=D)
    s = torch.rand(B, A, D, device=dev)
    t = timed(lambda: op._filter(s, "ramp")); k = _lib.last_kernel("filter")[:50]
    print(f"  D {D:4d}: {t:8.1f} us  {k}", flush=True)
print("fan beam 256^2 x 512 views x 16, det_count 256: source distance sweep")
x = torch.rand(B, N, N, device=dev)
for sd in (0.75, 1.0, 1.5, 2.0, 4.0):
    fan = pdu.RadonFanbeam(N, np.linspace(0, 2 * np.pi, A, endpoint=False), sd * N)
    s = fan._project(x)
    tf = timed(lambda: fan._project(x)); kf = _lib.last_kernel("radon_fwd")[38:100]
    ta = timed(lambda: fan._backproject(s)); ka = _lib.last_kernel("radon_adj")[:45]
    print(f"  s_dist {sd:4.2f} N (det_spacing {fan.det_spacing:5.2f}): fwd {tf:8.1f} us {kf} | adj {ta:8.1f} us {ka}", flush=True)
print("parallel, det_count sweep (256^2 x 512 x 16)")
for D, clip in ((256, False), (363, False), (384, False), (256, True)):
    op = pdu.Radon(N, np.linspace(0, np.pi, A, endpoint=False), det_count=D, clip_to_circle=clip)
    s = op._project(x)
    tf = timed(lambda: op._project(x)); ta = timed(lambda: op._backproject(s))
    print(f"  D {D} clip {clip}: fwd {tf:8.1f} us ({_lib.last_kernel('radon_fwd')[38:80]}) adj {ta:8.1f} us", flush=True)
