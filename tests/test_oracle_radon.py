"""Pins the CT oracle against first-principles known answers (the reference mount
holds no golden vectors: SURVEY.md section 8c).  CPU only."""
import math

import numpy as np
import pytest
import torch

import oracle
from oracle import RadonGeom
from oracle.radon import FAN, PARALLEL
from pd_unet_b200.phantoms import disc, shepp_logan


def _rel(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).norm() / b.norm())


def test_parallel_disc_matches_analytic():
    n, A, r = 128, 45, 40.0
    g = RadonGeom(n=n, n_angles=A, det_count=n)
    trig = oracle.trig_table(-np.linspace(0, np.pi, A, endpoint=False))
    sino = oracle.radon_forward(disc(n, r)[None], trig, g)[0]
    s = np.arange(n) - n / 2.0 + 0.5
    exact = 2.0 * np.sqrt(np.clip(r * r - s * s, 0, None))
    for a in range(A):
        assert _rel(sino[a], exact) < 6e-3
    # every view carries the mass of the object
    assert np.allclose(sino.sum(-1).numpy(), math.pi * r * r, rtol=2e-3)


def test_parallel_offcentre_disc_moves_with_cos_sin():
    n, r, cx, cy = 96, 10.0, 17.0, -9.0
    theta = np.array([0.0, 0.3, 1.2, 2.5])
    g = RadonGeom(n=n, n_angles=len(theta), det_count=n)
    internal = -theta                      # the host wrapper negates ([RECALL] torch_radon)
    sino = oracle.radon_forward(disc(n, r, cx, cy)[None], oracle.trig_table(internal), g)[0]
    s = np.arange(n) - n / 2.0 + 0.5
    for i, a in enumerate(internal):
        centre = cx * math.cos(a) + cy * math.sin(a)
        exact = 2.0 * np.sqrt(np.clip(r * r - (s - centre) ** 2, 0, None))
        assert _rel(sino[i], exact) < 5e-2


def test_fan_disc_matches_analytic():
    n, A, r = 128, 16, 30.0
    sd, dd, sp = 2.0 * n, 2.0 * n, 2.0
    g = RadonGeom(n=n, n_angles=A, det_count=n, det_spacing=sp, geom=FAN, s_dist=sd, d_dist=dd)
    trig = oracle.trig_table(-np.linspace(0, 2 * np.pi, A, endpoint=False))
    sino = oracle.radon_forward(disc(n, r)[None], trig, g)[0]
    u = (np.arange(n) - n / 2.0 + 0.5) * sp
    dist = np.abs(u) * sd / np.sqrt(u * u + (sd + dd) ** 2)
    exact = 2.0 * np.sqrt(np.clip(r * r - dist * dist, 0, None))
    for a in range(A):
        assert _rel(sino[a], exact) < 1.2e-2


def test_axis_aligned_ray_is_a_column_sum():
    n = 32
    rng = np.random.default_rng(0)
    img = rng.random((n, n))
    g = RadonGeom(n=n, n_angles=1, det_count=n)
    sino = oracle.radon_forward(img[None], oracle.trig_table([0.0]), g)[0, 0]
    # theta = 0: ray d runs along y at x = d + .5 (pixel centre column d).  n steps of
    # length 1 starting ON the border: the two end samples see half a pixel each.
    col = img.sum(0)
    mid = 0.5 * (img[:-1] + img[1:]).sum(0) + 0.5 * (img[0] + img[-1])
    assert np.allclose(sino.numpy(), mid, atol=1e-9)
    assert np.allclose(sino.numpy(), col, atol=1e-9)


def test_rays_missing_the_volume_are_zero():
    n = 32
    g = RadonGeom(n=n, n_angles=3, det_count=96)
    rs = oracle.ray_setup_f32(g, oracle.trig_table([0.0, 0.7, 1.5]))
    assert (rs["n_steps"][0, :30] == -1).all() and (rs["n_steps"][0, -30:] == -1).all()
    sino = oracle.radon_forward(np.ones((1, n, n)), oracle.trig_table([0.0, 0.7, 1.5]), g)
    assert float(sino[0, 0, :30].abs().max()) == 0.0


def test_clip_to_circle_only_sees_the_circle():
    n = 64
    g_sq = RadonGeom(n=n, n_angles=8, det_count=n)
    g_ci = RadonGeom(n=n, n_angles=8, det_count=n, clip_to_circle=True)
    trig = oracle.trig_table(np.linspace(0, np.pi, 8, endpoint=False))
    inside = disc(n, 20.0)
    a = oracle.radon_forward(inside[None], trig, g_sq)
    b = oracle.radon_forward(inside[None], trig, g_ci)
    assert _rel(b, a) < 2e-2
    corner = np.zeros((n, n))
    corner[:6, :6] = 1.0
    assert float(oracle.radon_forward(corner[None], trig, g_ci).abs().max()) < 1e-12


@pytest.mark.parametrize("geom", [PARALLEL, FAN])
def test_backprojection_is_the_approximate_transpose(geom):
    # The pair is unmatched (ray-driven / pixel-driven) like the library it restates,
    # so <Ax, y> and <x, A^T y> agree to discretisation error on smooth inputs, not to 1e-12.
    n, A = 64, 48
    if geom == PARALLEL:
        g = RadonGeom(n=n, n_angles=A, det_count=n)
        ang = np.linspace(0, np.pi, A, endpoint=False)
    else:
        g = RadonGeom(n=n, n_angles=A, det_count=n, det_spacing=2.0, geom=FAN, s_dist=2.0 * n, d_dist=2.0 * n)
        ang = np.linspace(0, 2 * np.pi, A, endpoint=False)
    trig = oracle.trig_table(-ang)
    c = np.arange(n) - n / 2 + 0.5
    xx, yy = np.meshgrid(c, c)
    x = np.exp(-((xx - 5) ** 2 + (yy + 3) ** 2) / (2 * 8.0 ** 2))
    y = np.exp(-((c[None, :] * g.det_spacing / (2.0 if geom == FAN else 1.0)) ** 2) / (2 * 10.0 ** 2)) \
        * (1 + 0.3 * np.cos(ang)[:, None])
    lhs = float((oracle.radon_forward(x[None], trig, g)[0] * torch.from_numpy(y)).sum())
    rhs = float((torch.from_numpy(x) * oracle.radon_backprojection(y[None], trig, g)[0]).sum())
    assert abs(lhs - rhs) / abs(lhs) < 5e-3


@pytest.mark.parametrize("name", ["ramp", "shepp-logan", "cosine", "hamming", "hann"])
@pytest.mark.parametrize("D", [32, 100, 256])
def test_filter_matrix_equals_fft_route(name, D):
    A = 12
    rng = np.random.default_rng(1)
    s = rng.standard_normal((2, A, D))
    from oracle.radon import filter_sinogram_fft
    via_fft = filter_sinogram_fft(s, name)          # [RECALL] torch_radon's route: pad, FFT, multiply, inverse, crop
    via_mat = torch.from_numpy(s) @ torch.from_numpy(oracle.filter_matrix(D, A, name))
    assert _rel(via_mat, via_fft) < 1e-12
    assert _rel(oracle.filter_sinogram(s, name), via_fft) < 1e-12


def test_ramp_taps_are_the_band_limited_ramp():
    t = oracle.filter_taps(64, "ramp")
    mid = 63
    assert abs(t[mid] - 0.5) < 1e-12
    assert abs(t[mid + 1] + 2.0 / math.pi ** 2) < 1e-12 and abs(t[mid - 3] + 2.0 / (3 * math.pi) ** 2) < 1e-12
    assert np.abs(t[mid + 2::2]).max() < 1e-12


def test_fbp_reconstructs_the_phantom():
    n, A = 128, 180
    g = RadonGeom(n=n, n_angles=A, det_count=n)
    trig = oracle.trig_table(-np.linspace(0, np.pi, A, endpoint=False))
    img = shepp_logan(n)
    rec = oracle.fbp(oracle.radon_forward(img[None], trig, g), trig, g)[0].numpy()
    mse = np.mean((rec - img) ** 2)
    psnr = 10 * np.log10(img.max() ** 2 / mse)
    assert psnr > 24.0, psnr
    c = np.arange(n) - n / 2 + 0.5
    inner = (c[None, :] ** 2 + c[:, None] ** 2) < (0.2 * n) ** 2
    assert abs(rec[inner].mean() - img[inner].mean()) < 0.02


def test_fan_beam_fbp_reconstructs_the_phantom():
    """fan-beam FBP = cosine pre-weight + ramp filter + 1 / U^2-weighted backprojection
    (Kak & Slaney section 3.4.2).  With a source one image width away the weights matter: the weighted form recovers the
    object's mean to 0.2 %, the unweighted composition is 8 % low."""
    n, A = 128, 360
    ang = np.linspace(0, 2 * np.pi, A, endpoint=False)
    g = RadonGeom(n=n, n_angles=A, det_count=320, det_spacing=2.0, geom=FAN, s_dist=1.0 * n, d_dist=1.0 * n)
    trig = oracle.trig_table(-ang)
    img = shepp_logan(n)
    sino = oracle.radon_forward(img[None], trig, g)
    psnr = lambda rec: 10 * np.log10(img.max() ** 2 / np.mean((rec - img) ** 2))
    rec = oracle.fbp(sino, trig, g)[0].numpy()
    plain = oracle.fbp(sino, trig, g, fan_weights=False)[0].numpy()
    assert psnr(rec) > 25.5, psnr(rec)
    assert psnr(rec) > psnr(plain) + 0.5, (psnr(rec), psnr(plain))
    c = np.arange(n) - n / 2 + 0.5
    inner = (c[None, :] ** 2 + c[:, None] ** 2) < (0.2 * n) ** 2
    m = img[inner].mean()
    assert abs(rec[inner].mean() - m) < 0.005 * m
    assert abs(plain[inner].mean() - m) > 0.05 * m
