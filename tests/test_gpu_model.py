"""End-to-end: the PD-UNet assembly on the CUDA operators against the same weights run on the CPU
oracle in float64 (north_star: final reconstruction PSNR within 0.01 dB)."""
import copy

import numpy as np
import pytest
import torch

import oracle
from oracle import updates as ou
import pd_unet_b200 as pdu
from pd_unet_b200.model import PrimalDualUNetCT, PrimalDualUNetMRI
from pd_unet_b200.phantoms import coil_maps, phantom_batch
from util import rel_l2, seeded, user_angles

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _reference_forward(model, g, image_shape, op_forward, op_adjoint):
    """PrimalDualUNet.forward restated with torch.cat / slicing on CPU float64."""
    m = copy.deepcopy(model).cpu().double()
    B = g.shape[0]
    h = g.new_zeros((B, m.n_dual) + tuple(g.shape[2:]))
    f = g.new_zeros((B, m.n_primal) + tuple(image_shape))
    inv = 1.0 / m.op_scale
    for i in range(m.n_iter):
        kf = op_forward(f[:, :m.kc]) * inv
        h = h + m.dual[i](torch.cat([h, kf, g], 1))
        kth = op_adjoint(h[:, :m.kd]) * inv
        f = f + m.primal[i](torch.cat([f, kth], 1))
    return f[:, :m.kc]


def test_ct_model_matches_cpu_oracle_model():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    n, A, up = 64, 32, 4
    ang = user_angles(A)
    radon = pdu.Radon(n, ang)
    torch.manual_seed(0)
    model = PrimalDualUNetCT(radon, upsample=up, n_iter=2, n_primal=3, n_dual=3, unet_base=8, unet_depth=2,
                             dual_features=8).to(DEV).eval()
    assert all("radon" not in k for k in model.state_dict())          # operators add nothing to checkpoints
    x = phantom_batch(2, n)
    g = oracle.RadonGeom(n=n, n_angles=A, det_count=n)
    trig = oracle.trig_table(-ang)
    full = oracle.radon_forward(x, trig, g).float()
    sparse = full[:, None, ::up].contiguous()
    with torch.no_grad():
        got = model(sparse.to(DEV))
        gg = ou.angular_upsample(sparse[:, 0].double(), up, "flip")[:, None] / model.op_scale
        want = _reference_forward(
            model, gg, (n, n),
            lambda im: oracle.radon_forward(im[:, 0], trig, g)[:, None],
            lambda s: oracle.fbp(s[:, 0], trig, g)[:, None])
    assert got.shape == (2, 1, n, n)
    assert rel_l2(got, want) < 1e-4                                    # cuDNN fp32 convolutions vs float64
    mse = lambda a: float(((a.double().cpu() - x[:, None].double()) ** 2).mean())
    assert abs(10 * np.log10(mse(got) / mse(want))) < 0.01


def test_ct_model_trains():
    n, A, up = 64, 32, 4
    radon = pdu.Radon(n, user_angles(A))
    torch.manual_seed(1)
    model = PrimalDualUNetCT(radon, upsample=up, n_iter=2, n_primal=2, n_dual=2, unet_base=8, unet_depth=2,
                             dual_features=8).to(DEV)
    opt = torch.optim.Adam(model.parameters(), 1e-3)
    x = phantom_batch(2, n).to(DEV)
    sparse = radon.forward(x)[:, None, ::up].contiguous()
    losses = []
    for _ in range(5):
        opt.zero_grad()
        loss = torch.nn.functional.mse_loss(model(sparse)[:, 0], x)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
    assert losses[-1] < losses[0]


def test_graphed_training_step_matches_eager():
    """One captured training step (forward + loss + backward + Adam) replayed N times lands on the same parameters
    as N eager steps (the fused epilogues are used at every size while capturing, so allow rounding differences)."""
    from pd_unet_b200.graph import GraphedTrainingStep
    n, A, up, warm, steps = 64, 32, 4, 2, 4
    radon = pdu.Radon(n, user_angles(A))
    x = phantom_batch(2, n).to(DEV)
    sparse = radon.forward(x)[:, None, ::up].contiguous()
    loss_fn = lambda out, tgt: torch.nn.functional.mse_loss(out[:, 0], tgt)

    def make():
        torch.manual_seed(3)
        m = PrimalDualUNetCT(radon, upsample=up, n_iter=2, n_primal=2, n_dual=2, unet_base=8, unet_depth=2,
                             dual_features=8).to(DEV)
        return m, torch.optim.Adam(m.parameters(), 1e-3, capturable=True)

    m_e, opt_e = make()
    eager_losses = []
    for _ in range(warm + steps):
        opt_e.zero_grad(set_to_none=True)
        loss = loss_fn(m_e(sparse), x)
        loss.backward()
        opt_e.step()
        eager_losses.append(float(loss))
    m_g, opt_g = make()
    step = GraphedTrainingStep(m_g, opt_g, loss_fn, (sparse,), x, warmup=warm)
    graph_losses = [float(step((sparse,), x)) for _ in range(steps)]
    assert graph_losses[-1] < graph_losses[0]
    assert abs(graph_losses[-1] - eager_losses[-1]) <= 1e-3 * abs(eager_losses[-1])
    for pe, pg in zip(m_e.parameters(), m_g.parameters()):
        assert rel_l2(pg, pe) <= 2e-3
    # a different batch goes through the captured tensors
    x2 = phantom_batch(2, n, seed=9).to(DEV)
    s2 = radon.forward(x2)[:, None, ::up].contiguous()
    assert torch.isfinite(step((s2,), x2))


def test_mri_model_matches_cpu_oracle_model():
    torch.backends.cudnn.allow_tf32 = False
    im, spokes, readout, coils = (32, 32), 8, 64, 2
    spec = oracle.NufftSpec(im)
    om = oracle.radial_trajectory(spokes, readout)
    smaps = coil_maps(coils, 32)[None]
    torch.manual_seed(2)
    model = PrimalDualUNetMRI(im, spokes, readout, coils=coils, n_iter=2, n_primal=4, n_dual=2 * coils, unet_base=8,
                              unet_depth=2, dual_features=8).to(DEV).eval()
    assert len([k for k in model.state_dict() if "nufft" in k]) == 0
    x = seeded((1, 1) + im, 3, complex_=True)
    kdata = oracle.nufft_forward(x, om, spec, smaps=smaps, norm="ortho").to(torch.complex64)
    omd = torch.from_numpy(om).to(DEV)
    dcf = pdu.calc_density_compensation_function(omd, im)
    with torch.no_grad():
        got = model(kdata.to(DEV), omd, smaps.to(DEV), dcf)
    dcf_cpu = dcf.cpu().to(torch.complex128)

    def to_real(y):      # [B, C, M] complex -> [B, 2C, spokes, readout]
        return torch.view_as_real(y).permute(0, 1, 3, 2).reshape(y.shape[0], 2 * coils, spokes, readout)

    def opf(xr):         # [B, 2, N, N] real -> data layout
        z = torch.view_as_complex(xr.permute(0, 2, 3, 1).contiguous())[:, None]
        return to_real(oracle.nufft_forward(z, om, spec, smaps=smaps, norm="ortho"))

    def opa(yr):
        z = torch.view_as_complex(yr.reshape(yr.shape[0], coils, 2, spokes * readout).permute(0, 1, 3, 2).contiguous())
        xx = oracle.nufft_adjoint(z * dcf_cpu, om, spec, smaps=smaps, norm="ortho")
        return torch.view_as_real(xx[:, 0]).permute(0, 3, 1, 2)

    with torch.no_grad():
        want = _reference_forward(model, to_real(kdata.to(torch.complex128)) / model.op_scale, im, opf, opa)
    want = torch.view_as_complex(want.permute(0, 2, 3, 1).contiguous())[:, None]
    assert got.shape == (1, 1) + im
    assert rel_l2(got, want) < 1e-4


def test_on_gpu_data_generation_feeds_both_models():
    from pd_unet_b200 import data
    radon = pdu.Radon(64, user_angles(32))
    ct = data.make_ct_batch(radon, 2, 4, seed=1, device=DEV)
    assert ct["sino_sparse"].shape == (2, 1, 8, 64) and torch.equal(ct["sino_sparse"], ct["sino_full"][:, :, ::4])
    torch.manual_seed(0)
    net = PrimalDualUNetCT(radon, upsample=4, n_iter=1, n_primal=4, n_dual=4, unet_base=8, unet_depth=2, dual_features=8).to(DEV).eval()
    with torch.no_grad():
        assert net(ct["sino_sparse"]).shape == ct["image"].shape
    mri = data.make_mri_batch((32, 32), 8, 2, 1, device=DEV)
    assert mri["kdata"].shape == (1, 2, 8 * 64) and mri["dcf"].shape == (1, 1, 8 * 64)
    spec = oracle.NufftSpec((32, 32))
    want = oracle.nufft_forward(mri["image"].cpu(), mri["omega"].cpu().numpy(), spec, smaps=mri["smaps"].cpu(), norm="ortho")
    assert rel_l2(mri["kdata"], want) <= 1e-5
    m = PrimalDualUNetMRI((32, 32), 8, 64, coils=2, n_iter=1, n_primal=4, n_dual=4, unet_base=8, unet_depth=2, dual_features=8).to(DEV).eval()
    with torch.no_grad():
        assert m(mri["kdata"], mri["omega"], mri["smaps"], mri["dcf"]).shape == (1, 1, 32, 32)


def test_fan_beam_ct_model_matches_cpu_oracle_model():
    """BASELINE configs[2] flavour (fan beam, views over 2 pi, periodic angular upsampling) at a small size."""
    from oracle.radon import FAN
    torch.backends.cudnn.allow_tf32 = False
    n, A, up = 64, 32, 4
    ang = user_angles(A, 2 * np.pi)
    radon = pdu.RadonFanbeam(n, ang, 2.0 * n)
    torch.manual_seed(3)
    model = PrimalDualUNetCT(radon, upsample=up, n_iter=2, n_primal=4, n_dual=4, unet_base=8, unet_depth=2,
                             dual_features=8).to(DEV).eval()
    assert model.wrap == "periodic"
    g = oracle.RadonGeom(n=n, n_angles=A, det_count=n, det_spacing=2.0, geom=FAN, s_dist=2.0 * n, d_dist=2.0 * n)
    trig = oracle.trig_table(-ang)
    x = phantom_batch(2, n)
    sparse = oracle.radon_forward(x, trig, g).float()[:, None, ::up].contiguous()
    with torch.no_grad():
        got = model(sparse.to(DEV))
        gg = ou.angular_upsample(sparse[:, 0].double(), up, "periodic")[:, None] / model.op_scale
        want = _reference_forward(model, gg, (n, n),
                                  lambda im: oracle.radon_forward(im[:, 0], trig, g)[:, None],
                                  lambda s: oracle.fbp(s[:, 0], trig, g)[:, None])
    assert rel_l2(got, want) < 1e-4


@pytest.mark.parametrize("tf32", [False, True])
def test_ct_model_at_the_benched_size_matches_cpu_oracle_model(tf32):
    """PSNR-within-0.01-dB proven at the size and model bench.py times (configs[1]: 256^2,
    64 -> 512 views, n_iter 4, UNet base 32 depth 3), on 2 slices, against the same weights on the CPU in float64 with
    the oracle's operators (OpenMP C restatement).  tf32=False: cuDNN in plain fp32 -- the north_star tolerance.
    tf32=True: PyTorch's default for convolutions (what bench.py and the reference run with); the operators stay
    fp32, the convolutions round their inputs to 10 mantissa bits -- a looser, measured bound."""
    from oracle import c_port
    import bench
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = tf32
    try:
        n, up, A = bench.N, bench.UP, bench.A_FULL
        ang = user_angles(A)
        radon = pdu.Radon(n, ang)
        torch.manual_seed(1234)
        model = PrimalDualUNetCT(radon, upsample=up, adjoint="fbp", **bench.MODEL_KW).to(DEV).eval()
        x = phantom_batch(2, n, seed=100)
        g = oracle.RadonGeom(n=n, n_angles=A, det_count=n)
        trig = oracle.trig_table(-ang)
        sparse = c_port.radon_forward(x, trig, g).float()[:, None, ::up].contiguous()
        with torch.no_grad():
            got = model(sparse.to(DEV))
            gg = ou.angular_upsample(sparse[:, 0].double(), up, "flip")[:, None] / model.op_scale
            want = _reference_forward(model, gg, (n, n),
                                      lambda im: c_port.radon_forward(im[:, 0], trig, g)[:, None],
                                      lambda s: c_port.fbp(s[:, 0], trig, g)[:, None])
        err = rel_l2(got, want)
        mse = lambda a: float(((a.double().cpu() - x[:, None].double()) ** 2).mean())
        dpsnr = abs(10 * np.log10(mse(got) / mse(want)))
        print(f"benched-size model, cudnn tf32={tf32}: rel-L2 {err:.2e}, |dPSNR| {dpsnr:.5f} dB")
        assert err < (5e-3 if tf32 else 1e-4)
        assert dpsnr < 0.01
    finally:
        torch.backends.cudnn.allow_tf32 = old
