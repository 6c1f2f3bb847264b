"""Pins the MRI oracle: gridded NUFFT against the exact non-uniform DFT, exact
adjointness, density compensation of a radial trajectory.  CPU only."""
import numpy as np
import torch

import oracle
from oracle import NufftSpec


def _rel(a, b):
    return float((a - b).norm() / b.norm())


def _rand_c(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.complex(torch.randn(*shape, generator=g, dtype=torch.float64),
                         torch.randn(*shape, generator=g, dtype=torch.float64))


def test_delta_is_a_plane_wave():
    spec = NufftSpec((32, 32))
    om = oracle.radial_trajectory(5, 64)
    x = torch.zeros(1, 1, 32, 32, dtype=torch.complex128)
    x[0, 0, 20, 9] = 1.0
    y = oracle.ndft_forward(x, om, spec)[0, 0]
    o = torch.from_numpy(om.astype(np.float64))
    assert _rel(y, torch.exp(-1j * (o[0] * (20 - 16) + o[1] * (9 - 16)))) < 1e-13


def test_gridded_forward_matches_exact_ndft():
    for n in [(32, 32), (48, 40)]:
        spec = NufftSpec(n)
        om = oracle.radial_trajectory(12, 2 * n[0])
        x = _rand_c(2, 1, *n)
        y = oracle.nufft_forward(x, om, spec)
        # J = 6, 2x grid, NEAREST lookup in a 1024x table: 6.6e-4 (7e-6 with a 2^18 table)
        assert _rel(y, oracle.ndft_forward(x, om, spec)) < 1e-3


def test_gridded_adjoint_matches_exact_ndft_adjoint():
    spec = NufftSpec((32, 32))
    om = oracle.radial_trajectory(12, 64)
    y = _rand_c(1, 2, om.shape[1])
    assert _rel(oracle.nufft_adjoint(y, om, spec), oracle.ndft_adjoint(y, om, spec)) < 1e-3


def test_adjoint_is_the_exact_conjugate_transpose():
    spec = NufftSpec((24, 40))
    om = oracle.radial_trajectory(7, 48)
    x = _rand_c(2, 3, 24, 40, seed=1)
    y = _rand_c(2, 3, om.shape[1], seed=2)
    for norm in (None, "ortho"):
        lhs = (oracle.nufft_forward(x, om, spec, norm=norm).conj() * y).sum()
        rhs = (x.conj() * oracle.nufft_adjoint(y, om, spec, norm=norm)).sum()
        assert abs(lhs - rhs) / abs(lhs) < 1e-12


def test_smaps_forward_and_adjoint():
    from pd_unet_b200.phantoms import coil_maps
    spec = NufftSpec((32, 32))
    om = oracle.radial_trajectory(9, 64)
    smaps = coil_maps(4, 32)[None].to(torch.complex128)
    x = _rand_c(2, 1, 32, 32, seed=3)
    y = _rand_c(2, 4, om.shape[1], seed=4)
    fwd = oracle.nufft_forward(x, om, spec, smaps=smaps)
    assert fwd.shape == (2, 4, om.shape[1])
    assert _rel(fwd, oracle.nufft_forward(x * smaps, om, spec)) < 1e-14
    adj = oracle.nufft_adjoint(y, om, spec, smaps=smaps)
    assert adj.shape == (2, 1, 32, 32)
    lhs = (fwd.conj() * y).sum()
    rhs = (x.conj() * adj).sum()
    assert abs(lhs - rhs) / abs(lhs) < 1e-12


def test_table_and_scaling_shapes_and_symmetry():
    spec = NufftSpec((64, 64))
    t = oracle.kb_table(spec, 0)
    assert t.shape == (6 * 1024 + 1,) and t[0] == 0 and t[-1] == 0
    assert abs(t[3 * 1024] - 1.0) < 1e-14                      # kb(0) = 1, phase(0) = 1
    assert np.allclose(np.abs(t), np.abs(t[::-1]))             # |kb| even in u
    s = oracle.scaling_coef(spec, 0)
    assert s.shape == (64,) and np.allclose(s, s[::-1]) and (s > 0).all()


def test_dcf_of_a_radial_trajectory_is_a_ramp():
    spec = NufftSpec((32, 32))
    n_sp, n_ro = 16, 64
    om = oracle.radial_trajectory(n_sp, n_ro, golden=False)
    w = oracle.calc_dcf(om, spec, 10).reshape(n_sp, n_ro).numpy()
    r = np.abs(np.arange(n_ro) - n_ro / 2)
    prof = w.mean(0)
    assert (w > 0).all()
    # grows with radius away from the centre, smallest at the centre
    assert prof.argmin() in (n_ro // 2 - 1, n_ro // 2, n_ro // 2 + 1)
    assert np.corrcoef(prof[20:45], r[20:45])[0, 1] > 0.97


def test_updates_upsample_adjoint():
    g = torch.Generator().manual_seed(0)
    for mode in ("flip", "periodic", "clamp"):
        s = torch.randn(2, 5, 7, generator=g, dtype=torch.float64)
        y = torch.randn(2, 20, 7, generator=g, dtype=torch.float64)
        lhs = (oracle.angular_upsample(s, 4, mode) * y).sum()
        rhs = (s * oracle.angular_upsample_adjoint(y, 4, mode)).sum()
        assert abs(lhs - rhs) < 1e-10
    up = oracle.angular_upsample(torch.ones(1, 4, 3, dtype=torch.float64), 8)
    assert torch.allclose(up, torch.ones_like(up))
    h, sl = oracle.dual_update(torch.ones(1, 3, 2, 2), 2 * torch.ones(1, 3, 2, 2), k=1)
    assert float(h.sum()) == 36 and sl.shape == (1, 2, 2)


def test_finer_table_converges_to_the_exact_ndft():
    # the structure (scaling centre, table phase, n_shift phase) is right iff the only
    # error left is the table quantisation: it must fall with the oversampling factor
    x = _rand_c(1, 1, 32, 32)
    om = oracle.radial_trajectory(12, 64)
    exact = oracle.ndft_forward(x, om, NufftSpec((32, 32)))
    errs = [_rel(oracle.nufft_forward(x, om, NufftSpec((32, 32), table_oversamp=L)), exact)
            for L in (1 << 10, 1 << 14, 1 << 18)]
    assert errs[0] > 8 * errs[1] and errs[2] < 2e-5
