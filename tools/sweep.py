"""BASELINE.json configs[4]: operator microbench sweep -- Radon fwd + adjoint and NUFFT fwd + adjoint over
128^2..1024^2 images and 64..2048 views / spokes.  Writes a markdown table (default profiles/r01_sweep.md).
Each entry: median of `reps` CUDA-event timings with L2 flushed between, G samples/s and the fraction of
the measured HBM roof the algorithmic bytes amount to."""
import json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import pd_unet_b200 as pdu

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sweep.md")
reps = 5
dev = "cuda:0"
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


lines = ["# r02 operator sweep (BASELINE.json configs[4]) on one B200\n",
         f"median of {reps} CUDA-event timings, 256 MiB L2 flush between; HBM roof {peak:.0f} GB/s (measured).\n",
         "\n## Radon, parallel beam, batch 8, det_count = N\n",
         "| N | views | fwd us | fwd GSamples/s | fwd HBM frac | adj us | adj GSamples/s | adj HBM frac |",
         "|---:|---:|---:|---:|---:|---:|---:|---:|"]
B = 8
for n in (128, 256, 512, 1024):
    for A in (64, 256, 1024, 2048):
        op = pdu.Radon(n, np.linspace(0, np.pi, A, endpoint=False))
        x = torch.rand(B, n, n, device=dev)
        s = torch.rand(B, A, n, device=dev)
        tf, ta = timed(lambda: op._project(x)), timed(lambda: op._backproject(s))
        nb = 4.0 * B * (n * n + A * n)
        smp = B * A * n * n
        lines.append(f"| {n} | {A} | {tf*1e3:.1f} | {smp/tf/1e6:.0f} | {nb/tf/1e6/peak:.4f} | {ta*1e3:.1f} | {smp/ta/1e6:.0f} | {nb/ta/1e6/peak:.4f} |")
        print(lines[-1], flush=True)
lines += ["\n## NUFFT, golden-angle radial, readout 2N, 1 coil, batch 4 (samples = B * M * 36 taps)\n",
          "| N | spokes | fwd us | fwd GTaps/s | fwd HBM frac | adj us | adj GTaps/s | adj HBM frac |",
          "|---:|---:|---:|---:|---:|---:|---:|---:|"]
Bm = 4
for n in (128, 256, 512, 1024):
    for sp in (64, 256, 1024, 2048):
        phi = np.arange(sp) * (111.246117975 * np.pi / 180.0)
        r = (np.arange(2 * n) - n) * (2 * np.pi / (2 * n))
        om = torch.from_numpy(np.stack([(r[None] * np.sin(phi)[:, None]).reshape(-1),
                                        (r[None] * np.cos(phi)[:, None]).reshape(-1)]).astype(np.float32)).to(dev)
        M = om.shape[1]
        fw, ad = pdu.KbNufft((n, n)), pdu.KbNufftAdjoint((n, n))
        img = torch.randn(Bm, 1, n, n, dtype=torch.complex64, device=dev)
        k = fw(img, om)
        tf, ta = timed(lambda: fw(img, om)), timed(lambda: ad(k, om))
        nb = 8.0 * Bm * (n * n + M) + 8.0 * M
        taps = Bm * M * 36
        lines.append(f"| {n} | {sp} | {tf*1e3:.1f} | {taps/tf/1e6:.0f} | {nb/tf/1e6/peak:.4f} | {ta*1e3:.1f} | {taps/ta/1e6:.0f} | {nb/ta/1e6/peak:.4f} |")
        print(lines[-1], flush=True)
        del fw, ad, img, k
        torch.cuda.empty_cache()
os.makedirs(os.path.dirname(out_path), exist_ok=True)
open(out_path, "w").write("\n".join(lines) + "\n")
