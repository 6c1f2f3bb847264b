"""NUFFT forward / adjoint a few times at the cfg1 / cfg4 / sweep shapes (for ncu launch lists and timings)."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pd_unet_b200 as pdu
from pd_unet_b200 import _lib
from pd_unet_b200.phantoms import coil_maps

dev = "cuda:0"
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
which = sys.argv[2] if len(sys.argv) > 2 else "all"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def traj(spokes, readout):
    phi = np.arange(spokes) * (111.246117975 * np.pi / 180.0)
    r = (np.arange(readout) - readout / 2) * (2 * np.pi / readout)
    return torch.from_numpy(np.stack([(r[None] * np.sin(phi)[:, None]).reshape(-1),
                                      (r[None] * np.cos(phi)[:, None]).reshape(-1)]).astype(np.float32)).to(dev)


def timed(name, fn, nbytes):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = statistics.median(ts)
    print(f"{name:46s} median {t*1e3:8.1f} us   {nbytes / t / 1e6:8.1f} GB/s algorithmic", flush=True)


CASES = (("cfg1 256^2 c1 b1 32sp", 256, 1, 1, 32), ("cfg1 256^2 c1 b1 256sp", 256, 1, 1, 256),
         ("cfg4 320^2 c8 b2 48sp", 320, 8, 2, 48), ("cfg4 320^2 c8 b8 48sp", 320, 8, 8, 48),
         ("sweep 128^2 c1 b16 64sp", 128, 1, 16, 64), ("sweep 512^2 c1 b8 256sp", 512, 1, 8, 256),
         ("sweep 1024^2 c1 b2 512sp", 1024, 1, 2, 512))
for name, n, coils, B, spokes in CASES:
    if which != "all" and not name.startswith(which):
        continue
    im = (n, n)
    om = traj(spokes, 2 * n)
    M = om.shape[1]
    fw, ad = pdu.KbNufft(im), pdu.KbNufftAdjoint(im)
    img = torch.randn(B, 1, n, n, dtype=torch.complex64, device=dev)
    if coils > 1:
        sm = coil_maps(coils, n)[None].to(dev)
        nb = 8.0 * B * coils * (n * n + M) + 8.0 * M + 8.0 * coils * n * n
    else:
        sm = None
        nb = 8.0 * B * coils * (n * n + M) + 8.0 * M
    k = fw(img, om, smaps=sm)
    timed(name + " fwd (default)", lambda: fw(img, om, smaps=sm), nb)
    print("    ", _lib.last_kernel("nufft_fwd")[:150])
    timed(name + " adj (default)", lambda: ad(k, om, smaps=sm), nb)
    print("    ", _lib.last_kernel("nufft_adj")[:150])
    if "--pg" in sys.argv:
        for pg in (2, 8):
            pdu.set_option("nufft_fwd_variant", pg); pdu.set_option("nufft_adj_variant", pg)
            timed(name + f" fwd (fused, {pg} planes / CTA)", lambda: fw(img, om, smaps=sm), nb)
            timed(name + f" adj (fused, {pg} planes / CTA)", lambda: ad(k, om, smaps=sm), nb)
        pdu.set_option("nufft_fwd_variant", -1); pdu.set_option("nufft_adj_variant", -1)
    if "--variants" in sys.argv:
        fw._plan.use_fused = ad._plan.use_fused = False
        for v, label in ((0, "pad + cuFFT"), (1, "generic pruned FFT"), (2, "r01 register FFT + gathers")):
            pdu.set_option("nufft_fwd_variant", v); pdu.set_option("nufft_adj_variant", v)
            timed(name + f" fwd ({label})", lambda: fw(img, om, smaps=sm), nb)
            timed(name + f" adj ({label}, auto interp)", lambda: ad(k, om, smaps=sm), nb)
        pdu.set_option("nufft_fwd_variant", -1); pdu.set_option("nufft_adj_variant", -1)
        ad._plan.use_csr = False
        timed(name + " adj (r01, atomic scatter)", lambda: ad(k, om, smaps=sm), nb)
        ad._plan.use_csr = "auto"
        fw._plan.use_fused = ad._plan.use_fused = True
