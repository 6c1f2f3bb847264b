"""Parity of the CUDA CT operators (through the C ABI, via pd_unet_b200.radon) with the CPU oracle."""
import numpy as np
import pytest
import torch

import oracle
import pd_unet_b200 as pdu
from oracle.radon import FAN, PARALLEL
from pd_unet_b200.phantoms import phantom_batch
from util import TOL, rel_l2, seeded, user_angles

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# name -> (operator kwargs, oracle geometry kwargs)
def _case(name):
    if name == "par64":
        n, A = 64, 24
        return pdu.Radon(n, user_angles(A)), oracle.RadonGeom(n=n, n_angles=A, det_count=n), -user_angles(A)
    if name == "par96_wide_det":
        n, A, D = 96, 40, 140
        return (pdu.Radon(n, user_angles(A), det_count=D), oracle.RadonGeom(n=n, n_angles=A, det_count=D), -user_angles(A))
    if name == "par100_spacing":       # n % 4 == 0 but odd-ish sizes, det_spacing != 1
        n, A, D = 100, 33, 77
        return (pdu.Radon(n, user_angles(A), det_count=D, det_spacing=1.5),
                oracle.RadonGeom(n=n, n_angles=A, det_count=D, det_spacing=1.5), -user_angles(A))
    if name == "par63_no_tma":         # n % 4 != 0: the forward projector must take its gather path
        n, A = 63, 17
        return pdu.Radon(n, user_angles(A)), oracle.RadonGeom(n=n, n_angles=A, det_count=n), -user_angles(A)
    if name == "par128_circle":
        n, A = 128, 48
        return (pdu.Radon(n, user_angles(A), clip_to_circle=True),
                oracle.RadonGeom(n=n, n_angles=A, det_count=n, clip_to_circle=True), -user_angles(A))
    if name == "par256_sparse":        # BASELINE configs[1] sparse view set
        n, A = 256, 64
        return pdu.Radon(n, user_angles(A)), oracle.RadonGeom(n=n, n_angles=A, det_count=n), -user_angles(A)
    if name == "fan96":
        n, A = 96, 36
        ang = user_angles(A, 2 * np.pi)
        return (pdu.RadonFanbeam(n, ang, 2.0 * n),
                oracle.RadonGeom(n=n, n_angles=A, det_count=n, det_spacing=2.0, geom=FAN, s_dist=2.0 * n, d_dist=2.0 * n),
                -ang)
    if name == "fan128_short":
        n, A, D = 128, 30, 200
        ang = user_angles(A, 2 * np.pi)
        return (pdu.RadonFanbeam(n, ang, 1.2 * n, det_distance=0.8 * n, det_count=D, det_spacing=1.3, clip_to_circle=True),
                oracle.RadonGeom(n=n, n_angles=A, det_count=D, det_spacing=1.3, geom=FAN, s_dist=1.2 * n, d_dist=0.8 * n,
                                 clip_to_circle=True), -ang)
    raise KeyError(name)


CASES = ["par64", "par96_wide_det", "par100_spacing", "par63_no_tma", "par128_circle", "par256_sparse", "fan96",
         "fan128_short"]


@pytest.fixture(autouse=True)
def _reset_variants():
    yield
    for k in ("radon_fwd_variant", "radon_adj_variant"):
        pdu.set_option(k, -1)


@pytest.mark.parametrize("variant", [-1, 0, 1, 9, 11, 13])     # default, L1 gather, float tiles, three cell-tile shapes
@pytest.mark.parametrize("name", CASES)
def test_forward_matches_oracle(name, variant):
    op, g, internal = _case(name)
    pdu.set_option("radon_fwd_variant", variant)
    x = phantom_batch(2, g.n, seed=3) + 0.05 * seeded((2, g.n, g.n), 5)
    got = op.forward(x.to(DEV))
    want = oracle.radon_forward(x, oracle.trig_table(internal), g)
    assert got.shape == want.shape
    assert rel_l2(got, want) <= TOL


@pytest.mark.parametrize("variant", [-1, 0])                    # default (line-form tile kernel), float64 gather
@pytest.mark.parametrize("name", CASES)
def test_backprojection_matches_oracle(name, variant):
    op, g, internal = _case(name)
    pdu.set_option("radon_adj_variant", variant)
    s = seeded((2, g.n_angles, g.det_count), 7)
    got = op.backprojection(s.to(DEV))
    want = oracle.radon_backprojection(s, oracle.trig_table(internal), g)
    assert rel_l2(got, want) <= TOL


@pytest.mark.parametrize("name", ["par64", "par100_spacing", "fan96"])
@pytest.mark.parametrize("filt", ["ramp", "hann", "shepp-logan", "cosine", "hamming"])
def test_filter_matches_oracle(name, filt):
    op, g, _ = _case(name)
    s = seeded((3, g.n_angles, g.det_count), 11)
    got = op.filter_sinogram(s.to(DEV), filt)
    want = oracle.filter_sinogram(s, filt)
    assert rel_l2(got, want) <= TOL


def test_fbp_reconstructs_phantom_like_the_oracle():
    n, A = 128, 180
    op = pdu.Radon(n, user_angles(A), clip_to_circle=True)
    g = oracle.RadonGeom(n=n, n_angles=A, det_count=n, clip_to_circle=True)
    trig = oracle.trig_table(-user_angles(A))
    x = phantom_batch(1, n, noise=0.0)
    sino = op.forward(x.to(DEV))
    rec = op.fbp(sino)
    want = oracle.fbp(oracle.radon_forward(x, trig, g), trig, g)
    assert rel_l2(rec, want) <= TOL
    mse_got = float(((rec.cpu().double() - x.double()) ** 2).mean())
    mse_want = float(((want - x.double()) ** 2).mean())
    psnr = lambda m: 10 * np.log10(float(x.max()) ** 2 / m)
    assert abs(psnr(mse_got) - psnr(mse_want)) < 0.01          # north_star: PSNR within 0.01 dB


def test_leading_dims_empty_and_errors():
    op, g, _ = _case("par64")
    x = seeded((2, 3, 64, 64), 1).to(DEV)
    y = op.forward(x)
    assert y.shape == (2, 3, g.n_angles, g.det_count)
    assert torch.equal(y[1, 2], op.forward(x[1, 2]))
    assert op.forward(x[:0]).shape == (0, 3, g.n_angles, g.det_count)
    assert op.backprojection(y[:0]).shape == (0, 3, 64, 64)
    with pytest.raises(pdu.PduError):
        op.forward(x.cpu())                       # no CPU fallback
    with pytest.raises(TypeError):
        op.forward(x.double())
    with pytest.raises(ValueError):
        op.forward(x[..., :32])


def test_autograd_pairs_forward_with_backprojection():
    op, g, _ = _case("par64")
    x = seeded((2, 64, 64), 2).to(DEV).requires_grad_()
    w = seeded((2, g.n_angles, g.det_count), 3).to(DEV)
    (op.forward(x) * w).sum().backward()
    assert torch.equal(x.grad, op.backprojection(w))
    s = seeded((2, g.n_angles, g.det_count), 4).to(DEV).requires_grad_()
    v = seeded((2, 64, 64), 5).to(DEV)
    (op.backprojection(s) * v).sum().backward()
    assert torch.equal(s.grad, op.forward(v))
    s2 = seeded((2, g.n_angles, g.det_count), 6).to(DEV).requires_grad_()
    (op.filter_sinogram(s2) * w).sum().backward()
    assert rel_l2(s2.grad, op.filter_sinogram(w)) <= 1e-6


# ---------------------------------------------------------------- BASELINE.json full sizes
def _full_cfg(name):
    if name == "cfg2":      # parallel 256^2, 512 views, batch 16
        n, A, B = 256, 512, 16
        return pdu.Radon(n, user_angles(A)), oracle.RadonGeom(n=n, n_angles=A, det_count=n), -user_angles(A), B
    n, A, B = 512, 1024, 8   # cfg3 per-GPU share: fan 512^2, 1024 views, batch 8
    ang = user_angles(A, 2 * np.pi)
    return (pdu.RadonFanbeam(n, ang, 2.0 * n),
            oracle.RadonGeom(n=n, n_angles=A, det_count=n, det_spacing=2.0, geom=FAN, s_dist=2.0 * n, d_dist=2.0 * n),
            -ang, B)


# White-noise sinograms are the worst case for a float32 backprojector (the interpolated value changes by
# O(1) per unit of detector coordinate).  The tile-relative float64 set-up of radon_adj_tile_kernel keeps
# the coordinate error at the magnitude of the tile, which holds the 1e-5 budget even there
# (measured 3.4e-6 at 512 bins x 1024 views, profiles/r01_parity.md).
TOL_NOISE_FULL = TOL


@pytest.mark.parametrize("name", ["cfg2", "cfg3"])
def test_full_size_properties(name):
    from oracle import c_port                      # OpenMP C restatement: whole batches in seconds
    op, g, internal, B = _full_cfg(name)
    trig = oracle.trig_table(internal)
    x = (phantom_batch(B, g.n, seed=1)).to(DEV)
    z = seeded((B, g.n, g.n), 9).to(DEV)
    y = op.forward(x)
    assert rel_l2(y, c_port.radon_forward(x.cpu(), trig, g)) <= TOL          # every slice, every view
    # linearity and batch independence
    assert rel_l2(op.forward(2.0 * x - 0.5 * z), 2.0 * y - 0.5 * op.forward(z)) <= TOL
    assert torch.equal(op.forward(x[3:5]), y[3:5])
    # the variants agree
    pdu.set_option("radon_fwd_variant", 0)
    assert rel_l2(op.forward(x), y) <= TOL
    # backprojection of what FBP feeds it: the ramp-filtered sinogram of the phantoms
    s_real = op.filter_sinogram(y)
    assert rel_l2(s_real, c_port.filter_sinogram(y.cpu())) <= TOL
    img = op.backprojection(s_real)
    assert rel_l2(img, c_port.radon_backprojection(s_real.cpu(), trig, g)) <= TOL
    # ... and of white noise (worst case, see TOL_NOISE_FULL)
    s = seeded((B, g.n_angles, g.det_count), 13).to(DEV)
    img_n = op.backprojection(s)
    assert rel_l2(img_n[:2], c_port.radon_backprojection(s[:2].cpu(), trig, g)) <= TOL_NOISE_FULL
    pdu.set_option("radon_adj_variant", 0)
    assert rel_l2(op.backprojection(s_real), img) <= TOL
    # the parallel-beam pair is close to adjoint (ray- vs pixel-driven discretisations), as in torch_radon
    if g.geom == PARALLEL:
        s2 = op.forward(phantom_batch(B, g.n, seed=21).to(DEV))
        lhs = float((y.double() * s2.double()).sum())
        rhs = float((x.double() * op.backprojection(s2).double()).sum())
        assert abs(lhs - rhs) / abs(lhs) < 2e-2


@pytest.mark.parametrize("D,A,B", [(128, 24, 2), (256, 64, 2), (512, 50, 3), (384, 7, 1), (320, 40, 2), (500, 33, 1), (132, 9, 2)])
def test_filter_tensor_core_variant(D, A, B):
    """The tcgen05 split-TF32 Toeplitz GEMM (filter_variant 1, the default when D % 4 == 0 and D >= 128; widths that
    are not a multiple of the 128-column tile are zero-padded by TMA and by the prepared filter matrix) against the
    float64 oracle and the CUDA-core kernel.  rows = B*A is deliberately not always a multiple of the
    128-row tile.  The object sinogram is the hard input: its ramp-filtered output is ~50x smaller than
    sum |x||h|, which is what the rounding error scales with."""
    from pd_unet_b200 import _lib
    op = pdu.Radon(D, user_angles(A), det_count=D)
    noise = seeded((B, A, D), 17)
    obj = op.forward(phantom_batch(B, D, seed=5).to(DEV)).cpu()
    for s in (noise, obj, obj + 0.01 * noise):
        want = oracle.filter_sinogram(s)
        for variant in (1, 0):
            try:
                pdu.set_option("filter_variant", variant)
                got = op.filter_sinogram(s.to(DEV))
                torch.cuda.synchronize()
            finally:
                pdu.set_option("filter_variant", -1)
            assert ("filter_tc_kernel" in _lib.last_kernel("filter")) == (variant == 1)
            assert rel_l2(got, want) <= TOL, f"variant {variant}"
    try:
        pdu.set_option("filter_variant", 1)
        assert rel_l2(op.filter_sinogram(obj.to(DEV), "hann"), oracle.filter_sinogram(obj, "hann")) <= TOL
    finally:
        pdu.set_option("filter_variant", -1)


def test_unsorted_angles_and_odd_geometries_take_the_fallback_paths_correctly():
    """Views of a CTA that are NOT neighbours (random angles) blow the strip box past its width, and a
    coarse detector makes the backprojector's interval overflow its shared segment: both kernels must
    then serve those strips / chunks from global memory with identical results."""
    rng = np.random.default_rng(3)
    n = 96
    for ang, D, sp in ((rng.uniform(0, 2 * np.pi, 37), 96, 1.0), (rng.uniform(0, np.pi, 16), 300, 0.25),
                       (np.array([0.0, np.pi / 2, np.pi / 4, 3.0]), 40, 4.0)):
        op = pdu.Radon(n, ang, det_count=D, det_spacing=sp)
        g = oracle.RadonGeom(n=n, n_angles=len(ang), det_count=D, det_spacing=sp)
        trig = oracle.trig_table(-ang)
        x = phantom_batch(2, n, seed=4) + 0.05 * seeded((2, n, n), 6)
        assert rel_l2(op.forward(x.to(DEV)), oracle.radon_forward(x, trig, g)) <= TOL
        s = seeded((2, len(ang), D), 8)
        assert rel_l2(op.backprojection(s.to(DEV)), oracle.radon_backprojection(s, trig, g)) <= TOL
    fan = pdu.RadonFanbeam(n, rng.uniform(0, 2 * np.pi, 11), 70.0, det_distance=30.0, det_count=180, det_spacing=0.7)
    gf = oracle.RadonGeom(n=n, n_angles=11, det_count=180, det_spacing=0.7, geom=FAN, s_dist=70.0, d_dist=30.0)
    trig = oracle.trig_table(-fan.angles)
    x = phantom_batch(1, n, seed=9)
    assert rel_l2(fan.forward(x.to(DEV)), oracle.radon_forward(x, trig, gf)) <= TOL
    s = seeded((1, 11, 180), 10)
    assert rel_l2(fan.backprojection(s.to(DEV)), oracle.radon_backprojection(s, trig, gf)) <= TOL


def test_fine_detectors_stay_on_the_shared_memory_taps():
    """det_spacing < 0.5 (and strong fan magnification): a 32 x 32 tile projects onto more than 96 bins.  r02 gives those
    geometries a 192-entry segment with a centred coordinate instead of the float64 global-load path (4 - 6 x slower):
    same tolerance, for object-like and white-noise sinograms, parallel and fan beam, plain and FBP-weighted."""
    from pd_unet_b200 import _lib
    n = 128
    ang = np.linspace(0, np.pi, 48, endpoint=False)
    for sp, D in ((0.4, 352), (0.25, 544)):
        op = pdu.Radon(n, ang, det_count=D, det_spacing=sp)
        g = oracle.RadonGeom(n=n, n_angles=len(ang), det_count=D, det_spacing=sp)
        trig = oracle.trig_table(-ang)
        x = phantom_batch(2, n, seed=11)
        sino = oracle.radon_forward(x, trig, g).float()
        assert rel_l2(op.backprojection(sino.to(DEV)), oracle.radon_backprojection(sino, trig, g)) <= TOL
        assert ",192," in _lib.last_kernel("radon_adj")
        s = seeded((2, len(ang), D), 12)
        assert rel_l2(op.backprojection(s.to(DEV)), oracle.radon_backprojection(s, trig, g)) <= TOL
    angf = np.linspace(0, 2 * np.pi, 40, endpoint=False)
    fan = pdu.RadonFanbeam(n, angf, 160.0, det_distance=160.0, det_count=640, det_spacing=0.5)
    gf = oracle.RadonGeom(n=n, n_angles=len(angf), det_count=640, det_spacing=0.5, geom=FAN, s_dist=160.0, d_dist=160.0)
    trig = oracle.trig_table(-fan.angles)
    s = seeded((1, len(angf), 640), 13)
    assert rel_l2(fan.backprojection(s.to(DEV)), oracle.radon_backprojection(s, trig, gf)) <= TOL
    assert ",192," in _lib.last_kernel("radon_adj")
    assert rel_l2(fan._backproject(s.to(DEV), True), oracle.radon_backprojection(s, trig, gf, fbp_weight=True)) <= TOL


def test_a_pipeline_timeout_is_reported_not_silently_wrong():
    """a timed-out mbarrier wait must not yield a quietly wrong sinogram.  `debug_fault` makes the TMA
    producers skip their loads; the kernels must give up within their (shortened) time-out, raise the device error
    word, and every later library call must fail with PDU_ECUDA until the word is cleared."""
    from pd_unet_b200 import _lib
    L = _lib.lib()
    n, A = 256, 64
    dense = pdu.Radon(n, user_angles(512))           # cell ("quad") kernel
    sparse = pdu.Radon(n, user_angles(A))            # float-tile kernel
    x = phantom_batch(1, n).to(DEV)
    good = dense.forward(x)
    s = sparse.forward(x)
    good_f = sparse.filter_sinogram(s)
    torch.cuda.synchronize()
    assert L.pdu_device_error(0) == 0
    for name, call, code in (("cell projector", lambda: dense._project(x), 1), ("tile projector", lambda: sparse._project(x), 1),
                             ("tensor-core filter", lambda: sparse._filter(s, "ramp"), 2)):
        try:
            pdu.set_option("debug_fault", 1)
            call()                                   # the launch itself succeeds ...
            torch.cuda.synchronize()
        finally:
            pdu.set_option("debug_fault", -1)
        assert L.pdu_device_error(0) == code, name   # ... the kernel reports
        with pytest.raises(_lib.PduError, match="device-side failure"):
            dense.forward(x)                         # and the next call refuses to run
        assert L.pdu_device_error(1) == code         # read and clear
        assert L.pdu_device_error(0) == 0
    # the library works again and gives the same numbers as before
    assert torch.equal(dense.forward(x), good)
    assert torch.equal(sparse.filter_sinogram(s), good_f)


def test_last_kernel_names_the_dispatch():
    from pd_unet_b200 import _lib
    n = 256
    x = phantom_batch(1, n).to(DEV)
    dense, sparse = pdu.Radon(n, user_angles(512)), pdu.Radon(n, user_angles(64))
    dense.forward(x)
    assert "radon_fwd_quad_kernel<32,8,16,92,2,4>" in _lib.last_kernel("radon_fwd")
    s = sparse.forward(x)
    assert "radon_fwd_strip_kernel" in _lib.last_kernel("radon_fwd")
    sparse.filter_sinogram(s)
    assert "filter_tc_kernel" in _lib.last_kernel("filter")
    sparse.backprojection(s)
    assert "radon_adj_tile_kernel<32,32,8,32,64,parallel>" in _lib.last_kernel("radon_adj")


def test_fan_beam_fbp_with_its_weights_matches_the_oracle():
    """the cosine pre-weight rides in the filter kernel (tensor-core path: D % 128 == 0, CUDA-core
    path otherwise), the distance weight in the backprojector; both against the float64 oracle, PSNR within 0.01 dB."""
    from pd_unet_b200.phantoms import shepp_logan
    for n, A, D, sp, s_, d_ in ((128, 180, 256, 2.0, 1.0, 1.0), (96, 120, 200, 1.7, 1.3, 0.6)):
        ang = user_angles(A, 2 * np.pi)
        op = pdu.RadonFanbeam(n, ang, s_ * n, det_distance=d_ * n, det_count=D, det_spacing=sp)
        g = oracle.RadonGeom(n=n, n_angles=A, det_count=D, det_spacing=sp, geom=FAN, s_dist=s_ * n, d_dist=d_ * n)
        trig = oracle.trig_table(-ang)
        x = torch.from_numpy(shepp_logan(n)).float()[None]
        sino = op.forward(x.to(DEV))
        rec = op.fbp(sino)
        want = oracle.fbp(sino.cpu(), trig, g)
        assert rel_l2(rec, want) <= TOL
        assert rel_l2(op.fbp(sino, fan_weights=False), oracle.fbp(sino.cpu(), trig, g, fan_weights=False)) <= TOL
        mse = lambda a: float(((a.double().cpu() - x.double()) ** 2).mean())
        assert abs(10 * np.log10(mse(rec) / mse(want))) < 0.01
        # the weights remove the bias of the plain composition (8 % low for a source one image width away)
        c = np.arange(n) - n / 2 + 0.5
        inner = torch.from_numpy((c[None, :] ** 2 + c[:, None] ** 2) < (0.2 * n) ** 2)
        m = float(x[0][inner].mean())
        assert abs(float(rec[0].cpu()[inner].mean()) - m) < 0.01 * m
        assert abs(float(op.fbp(sino, fan_weights=False)[0].cpu()[inner].mean()) - m) > 0.03 * m
        with pytest.raises(NotImplementedError):
            s2 = sino.clone().requires_grad_()
            op.fbp(s2).sum().backward()


@pytest.mark.parametrize("name", ["par64", "par256_sparse", "fan96", "par63_no_tma"])
def test_texture_weight_emulation_switch(name):
    """[RECALL] torch_radon interpolates through the texture unit (8-bit weights).  The option
    "tex_weights" makes every projector round its interpolation fractions the same way; the oracle has the twin.  The
    two conventions differ by 2e-4 on a phantom -- far more than the 1e-5 budget -- which is why the switch exists.
    Tolerances: a fraction within float32 rounding of a 2^-9 tie rounds differently in the kernels (strip- / tile-local
    float32 coordinates) and in the oracle (float64, or float32 at image magnitude); each such sample is off by
    2^-8 x the local difference of its neighbours -- 1e-5-level on images and object sinograms, 1e-4-level on white noise."""
    op, g, internal = _case(name)
    trig = oracle.trig_table(internal)
    x = phantom_batch(2, g.n, seed=3)
    s_obj = oracle.radon_forward(x, trig, g).float()
    s_noise = seeded((2, g.n_angles, g.det_count), 7)
    exact_f = oracle.radon_forward(x, trig, g)
    try:
        pdu.set_option("tex_weights", 1)
        got_f = op.forward(x.to(DEV))
        got_b = op.backprojection(s_obj.to(DEV))
        got_bn = op.backprojection(s_noise.to(DEV))
        dense = pdu.Radon(g.n, user_angles(8 * g.n_angles), det_count=g.det_count) if name == "par64" else None
        got_d = dense.forward(x.to(DEV)) if dense is not None else None       # the cell-tile kernel
    finally:
        pdu.set_option("tex_weights", -1)
    want_f = oracle.radon_forward(x, trig, g, tex_weights=True)
    assert rel_l2(got_f, want_f) <= 3e-5
    assert rel_l2(got_b, oracle.radon_backprojection(s_obj, trig, g, tex_weights=True)) <= 3e-5
    assert rel_l2(got_bn, oracle.radon_backprojection(s_noise, trig, g, tex_weights=True)) <= 5e-4
    assert rel_l2(want_f, exact_f) > 3e-5                              # the conventions really differ (6e-5 .. 2e-4)
    assert rel_l2(got_f, exact_f) > 3e-5
    if dense is not None:
        gd = oracle.RadonGeom(n=g.n, n_angles=8 * g.n_angles, det_count=g.det_count)
        assert rel_l2(got_d, oracle.radon_forward(x, oracle.trig_table(-user_angles(8 * g.n_angles)), gd, tex_weights=True)) <= 3e-5
    assert rel_l2(op.forward(x.to(DEV)), exact_f) <= TOL               # and the switch is off again
