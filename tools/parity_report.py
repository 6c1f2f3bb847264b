"""Parity table: rel-L2 of every hot-path operator (through the C ABI) against the float64 oracle at the
BASELINE.json shapes, written as markdown (default profiles/r02_parity.md).  Runs on the GPU box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import oracle
from oracle import c_port, updates as ou
from oracle.radon import FAN
import pd_unet_b200 as pdu
from pd_unet_b200.phantoms import phantom_batch, coil_maps

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_parity.md")
dev = "cuda:0"


def rel(a, b):
    a = a.detach().cpu()
    dt = torch.complex128 if a.is_complex() else torch.float64
    a, b = a.to(dt), torch.as_tensor(b).to(dt)
    return float((a - b).norm() / b.norm())


def seeded(shape, seed, cplx=False):
    g = torch.Generator().manual_seed(seed)
    if cplx:
        return torch.complex(torch.randn(shape, generator=g), torch.randn(shape, generator=g)).to(torch.complex64)
    return torch.randn(shape, generator=g)


rows = []
for name, n, A, B, fan in (("cfg2 parallel 256^2 x 512 views, batch 16", 256, 512, 16, False),
                           ("cfg2 sparse set 256^2 x 64 views, batch 16", 256, 64, 16, False),
                           ("cfg3 fan 512^2 x 1024 views, batch 8", 512, 1024, 8, True)):
    if fan:
        ang = np.linspace(0, 2 * np.pi, A, endpoint=False)
        op = pdu.RadonFanbeam(n, ang, 2.0 * n)
        g = oracle.RadonGeom(n=n, n_angles=A, det_count=n, det_spacing=2.0, geom=FAN, s_dist=2.0 * n, d_dist=2.0 * n)
    else:
        ang = np.linspace(0, np.pi, A, endpoint=False)
        op = pdu.Radon(n, ang)
        g = oracle.RadonGeom(n=n, n_angles=A, det_count=n)
    trig = oracle.trig_table(-ang)
    x = phantom_batch(B, n, seed=1)
    y = op.forward(x.to(dev))
    rows.append((name, "radon forward (phantoms + 1% noise)", rel(y, c_port.radon_forward(x, trig, g))))
    xn = seeded((2, n, n), 3)
    rows.append((name, "radon forward (white noise, 2 slices)", rel(op.forward(xn.to(dev)), c_port.radon_forward(xn, trig, g))))
    q = op.filter_sinogram(y)
    rows.append((name, "ramp filter of the object sinogram (tcgen05)", rel(q, c_port.filter_sinogram(y.cpu()))))
    rows.append((name, "backprojection of the filtered sinogram (FBP)", rel(op.backprojection(q), c_port.radon_backprojection(q.cpu(), trig, g))))
    sn = seeded((2, A, n), 5)
    rows.append((name, "backprojection (white noise, 2 slices)", rel(op.backprojection(sn.to(dev)), c_port.radon_backprojection(sn, trig, g))))
for name, n, coils, B, spokes in (("cfg1 MRI 256^2, 32 spokes, 1 coil", 256, 1, 1, 32), ("cfg4 MRI 320^2, 48 spokes, 8 coils, batch 2", 320, 8, 2, 48)):
    spec = oracle.NufftSpec((n, n))
    om = oracle.radial_trajectory(spokes, 2 * n)
    omd = torch.from_numpy(om).to(dev)
    sm = coil_maps(coils, n)[None] if coils > 1 else None
    img = seeded((B, 1, n, n), 7, cplx=True)
    k = pdu.KbNufft((n, n))(img.to(dev), omd, smaps=sm.to(dev) if sm is not None else None, norm="ortho")
    rows.append((name, "NUFFT forward", rel(k, oracle.nufft_forward(img, om, spec, smaps=sm, norm="ortho"))))
    kd = seeded((B, coils, om.shape[1]), 9, cplx=True)
    xa = pdu.KbNufftAdjoint((n, n))(kd.to(dev), omd, smaps=sm.to(dev) if sm is not None else None, norm="ortho")
    rows.append((name, "NUFFT adjoint", rel(xa, oracle.nufft_adjoint(kd, om, spec, smaps=sm, norm="ortho"))))
    w = pdu.calc_density_compensation_function(omd, (n, n))
    rows.append((name, "density compensation (10 iterations)", rel(w.real.reshape(-1), oracle.calc_dcf(om, spec, 10))))
h, d = seeded((4, 4, 512, 256), 11), seeded((4, 4, 512, 256), 12)
o, s = pdu.updates.residual_slice(h.to(dev), d.to(dev), 0)
rows.append(("updates", "residual + slice", max(rel(o, ou.dual_update(h, d, 0)[0]), rel(s[:, 0], ou.dual_update(h, d, 0)[1]))))
sp = seeded((4, 64, 256), 13)
rows.append(("updates", "angular upsample 64 -> 512 views", rel(pdu.updates.angular_upsample(sp.to(dev), 8, "flip"), ou.angular_upsample(sp, 8, "flip"))))
lines = ["# r02 parity: CUDA path (through the C ABI) vs the float64 oracle\n",
         "rel-L2 = ||cuda - oracle|| / ||oracle||; budget 1e-5 (BASELINE.json north_star). The oracle is this repo's own",
         "restatement (parity unpinned, DESIGN.md section 0).\n", "| configuration | operator | rel-L2 |", "|---|---|---:|"]
lines += [f"| {a} | {b} | {c:.2e} |" for a, b, c in rows]
open(out_path, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
