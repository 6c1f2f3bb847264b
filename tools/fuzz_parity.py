"""Randomised parity sweep: the CUDA operators (through the C ABI) against the CPU oracle on random geometries,
trajectories and batch shapes -- the shapes nobody wrote a named test for.  Prints every case and a summary; exit
status 1 when a case exceeds the tolerance.
   python tools/fuzz_parity.py [n_ct_cases] [n_mri_cases] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from oracle import c_port
from oracle.radon import FAN
import pd_unet_b200 as pdu
from pd_unet_b200 import _lib

TOL = 1e-5
# Fan beams that need the backprojector's 192-entry segments (source within ~1.3 n of the centre, or detector bins well
# below half a pixel): float32 evaluates the projective detector coordinate num / den at the magnitude |c| den, c < 96
# bins, which on WHITE-NOISE sinograms of a few views gives 6e-6 (median) .. 1.7e-5 (worst of ~600 random geometries)
# against the float64 oracle.  set_option("radon_adj_variant", 0) is the float64 path for whoever needs more there.
TOL_FAN_WIDE = 2e-5
dev = "cuda:0"
FOCUS = os.environ.get("FUZZ_FOCUS", "")        # "fanwide": fan beams that need the backprojector's 192-entry segments


def rel(a, b):
    a = a.detach().cpu()
    dt = torch.complex128 if a.is_complex() else torch.float64
    a, b = a.to(dt), torch.as_tensor(b).to(dt)
    return float((a - b).norm() / max(float(b.norm()), 1e-300))


def ct_case(rng):
    """One random CT geometry: (description, operator, oracle geometry, internal angles)."""
    n = int(rng.choice([rng.integers(16, 64), rng.integers(64, 200), 4 * rng.integers(8, 80), 256, 128, 320]))
    A = int(rng.choice([1, 2, 3, rng.integers(4, 40), rng.integers(40, 200)]))
    fan = rng.random() < 0.4 or FOCUS == "fanwide"
    span = 2 * np.pi if (fan or rng.random() < 0.2) else np.pi
    kind = rng.integers(0, 4)
    if kind == 0:
        ang = np.linspace(0, span, A, endpoint=False)
    elif kind == 1:
        ang = np.linspace(0, span, A, endpoint=False) + rng.uniform(-3, 3)
    elif kind == 2:
        ang = np.sort(rng.uniform(0, span, A))
    else:
        ang = rng.uniform(-2 * np.pi, 2 * np.pi, A)               # unsorted, repeated quadrants
    D = int(rng.choice([n, n, rng.integers(max(8, n // 2), 2 * n + 1), 4 * rng.integers(max(2, n // 8), n // 2 + 2)]))
    circle = bool(rng.random() < 0.3)
    if fan:
        sd = float(rng.uniform(0.75, 4.0) * n)
        dd = float(rng.choice([sd, rng.uniform(0.0, 3.0) * n]))
        sp = float(rng.choice([(sd + dd) / sd, rng.uniform(0.3, 2.5)]))
        if FOCUS == "fanwide":                                      # strong magnification or fine detectors: wide segments
            sd = float(rng.uniform(0.75, 1.6) * n)
            dd = float(rng.choice([sd, rng.uniform(0.5, 3.0) * n]))
            sp = float(rng.choice([(sd + dd) / sd, rng.uniform(0.3, 1.0)]))
        op = pdu.RadonFanbeam(n, ang, sd, det_distance=dd, det_count=D, det_spacing=sp, clip_to_circle=circle)
        g = oracle.RadonGeom(n=n, n_angles=A, det_count=D, det_spacing=sp, geom=FAN, s_dist=sd, d_dist=dd, clip_to_circle=circle)
        desc = f"fan n{n} A{A} D{D} sp{sp:.3f} sd{sd / n:.2f}n dd{dd / n:.2f}n circ{int(circle)} ang{kind}"
    else:
        sp = float(rng.choice([1.0, 1.0, rng.uniform(0.3, 2.5)]))
        op = pdu.Radon(n, ang, det_count=D, det_spacing=sp, clip_to_circle=circle)
        g = oracle.RadonGeom(n=n, n_angles=A, det_count=D, det_spacing=sp, clip_to_circle=circle)
        desc = f"par n{n} A{A} D{D} sp{sp:.3f} circ{int(circle)} ang{kind} span{span / np.pi:.0f}pi"
    return desc, op, g, -ang


def run_ct(rng, i):
    desc, op, g, internal = ct_case(rng)
    B = int(rng.integers(1, 5))
    trig = oracle.trig_table(internal)
    gen = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
    x = torch.rand(B, g.n, g.n, generator=gen) + 0.1 * torch.randn(B, g.n, g.n, generator=gen)
    s = torch.randn(B, g.n_angles, g.det_count, generator=gen)
    e = {}
    e["fwd"] = rel(op.forward(x.to(dev)), c_port.radon_forward(x, trig, g))
    kf = _lib.last_kernel("radon_fwd")[:34]
    e["adj"] = rel(op.backprojection(s.to(dev)), c_port.radon_backprojection(s, trig, g))
    ka = _lib.last_kernel("radon_adj")[:40]
    filt = str(rng.choice(["ramp", "hann", "shepp-logan", "cosine", "hamming"]))
    e["filt"] = rel(op.filter_sinogram(s.to(dev), filt), c_port.filter_sinogram(s, filt))
    assert _lib.device_error() == 0
    wide = ",192,fan" in ka
    ok = e["fwd"] <= TOL and e["filt"] <= TOL and e["adj"] <= (TOL_FAN_WIDE if wide else TOL)
    flag = "" if ok else "  <-- FAIL"
    print(f"ct {i:3d} B{B} {desc:70s} fwd {e['fwd']:.1e} adj {e['adj']:.1e} {filt} {e['filt']:.1e} | {kf} | {ka}{flag}", flush=True)
    return dict(e, ok=ok, wide=wide, desc=desc)


def run_mri(rng, i):
    n0 = int(rng.choice([rng.integers(8, 48), 2 * rng.integers(8, 80), 64, 128, 160]))
    n1 = int(rng.choice([n0, n0, rng.integers(8, 100)]))
    J = int(rng.choice([6, 6, 6, 4, 5, 8]))
    grid = None if rng.random() < 0.75 else (int(n0 * rng.choice([1.5, 2, 2.5]) // 2 * 2), int(n1 * rng.choice([1.5, 2]) // 2 * 2))
    B, C = int(rng.integers(1, 4)), int(rng.choice([1, 1, 2, 5, 8, 12]))
    kind = int(rng.integers(0, 4))
    if kind == 0:
        om = oracle.radial_trajectory(int(rng.integers(1, 40)), int(rng.choice([2 * n0, n0, 37])))
    elif kind == 1:
        om = rng.uniform(-np.pi, np.pi, (2, int(rng.integers(1, 3000))))
    elif kind == 2:                                                # edge values: +-pi, 0, just inside
        base = np.array([-np.pi, np.pi, 0.0, np.pi - 1e-6, -np.pi + 1e-6, np.pi / 2, 1e-7])
        om = np.stack([rng.choice(base, 200), rng.choice(base, 200)])
    else:                                                          # concentrated cluster
        om = np.clip(rng.normal(0, 0.2, (2, int(rng.integers(10, 4000)))), -np.pi, np.pi)
    om = om.astype(np.float32).astype(np.float64)
    use_smaps = C > 1 and rng.random() < 0.5
    norm = "ortho" if rng.random() < 0.5 else None
    spec = oracle.NufftSpec((n0, n1), grid_size=grid, numpoints=J)
    gen = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
    cr = lambda *s: torch.complex(torch.randn(*s, generator=gen), torch.randn(*s, generator=gen)).to(torch.complex64)
    omd = torch.from_numpy(om).to(dev).float()
    fwd, adj = pdu.KbNufft((n0, n1), grid_size=grid, numpoints=J), pdu.KbNufftAdjoint((n0, n1), grid_size=grid, numpoints=J)
    M = om.shape[1]
    if use_smaps:
        sm = cr(1, C, n0, n1)
        img, kd = cr(B, 1, n0, n1), cr(B, C, M)
        y = fwd(img.to(dev), omd, smaps=sm.to(dev), norm=norm)
        kf = _lib.last_kernel("nufft_fwd")[:28]
        xa = adj(kd.to(dev), omd, smaps=sm.to(dev), norm=norm)
        ka = _lib.last_kernel("nufft_adj")[:28]
        wy, wx = oracle.nufft_forward(img, om, spec, smaps=sm, norm=norm), oracle.nufft_adjoint(kd, om, spec, smaps=sm, norm=norm)
    else:
        img, kd = cr(B, C, n0, n1), cr(B, C, M)
        y = fwd(img.to(dev), omd, norm=norm)
        kf = _lib.last_kernel("nufft_fwd")[:28]
        xa = adj(kd.to(dev), omd, norm=norm)
        ka = _lib.last_kernel("nufft_adj")[:28]
        wy, wx = oracle.nufft_forward(img, om, spec, norm=norm), oracle.nufft_adjoint(kd, om, spec, norm=norm)
    assert _lib.device_error() == 0
    ef, ea = rel(y, wy), rel(xa, wx)
    ok = max(ef, ea) <= TOL
    flag = "" if ok else "  <-- FAIL"
    desc = f"im{n0}x{n1} grid{grid} J{J} B{B} C{C} smaps{int(use_smaps)} {norm} traj{kind} M{M}"
    print(f"mri {i:3d} {desc}  fwd {ef:.1e} adj {ea:.1e} | {kf} | {ka}{flag}", flush=True)
    return dict(fwd=ef, adj=ea, ok=ok, desc=desc)


def main():
    n_ct = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    n_mri = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rng = np.random.default_rng(seed)
    ct = [run_ct(rng, i) for i in range(n_ct)]
    mri = [run_mri(rng, i) for i in range(n_mri)]
    mx = lambda rows, k, sel=lambda r: True: max([r[k] for r in rows if sel(r)], default=0.0)
    print(f"seed {seed}: {n_ct} CT cases: forward {mx(ct, 'fwd'):.2e}, filter {mx(ct, 'filt'):.2e}, backprojection "
          f"{mx(ct, 'adj', lambda r: not r['wide']):.2e} (fan beams through the wide segment: {mx(ct, 'adj', lambda r: r['wide']):.2e}, "
          f"{sum(r['wide'] for r in ct)} cases); {n_mri} MRI cases: forward {mx(mri, 'fwd'):.2e}, adjoint {mx(mri, 'adj'):.2e}; "
          f"tolerance {TOL:.0e} ({TOL_FAN_WIDE:.0e} for the wide fan segment)")
    sys.exit(0 if all(r["ok"] for r in ct + mri) else 1)


if __name__ == "__main__":
    main()
