"""Parity of the fused primal-dual update kernels with the CPU oracle (bit-exact: these are single
rounded float32 operations or pure data movement)."""
import pytest
import torch

import oracle
from oracle import updates as ou
import pd_unet_b200 as pdu
from pd_unet_b200 import updates
from util import rel_l2, seeded

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("plane", [(8, 16), (5, 7), (64, 64)])     # vectorised and scalar paths
def test_concat_is_torch_cat(plane):
    a, b, c = (seeded((3, ch) + plane, i) for i, ch in enumerate((4, 1, 2)))
    assert torch.equal(updates.concat(a.to(DEV), b.to(DEV), c.to(DEV)).cpu(), torch.cat([a, b, c], 1))
    assert torch.equal(updates.concat(a.to(DEV), b.to(DEV)).cpu(), torch.cat([a, b], 1))


@pytest.mark.parametrize("plane", [(8, 16), (5, 7)])
@pytest.mark.parametrize("k,kn", [(0, 1), (2, 1), (1, 2)])
def test_residual_slice(plane, k, kn):
    h, d = seeded((2, 4) + plane, 1), seeded((2, 4) + plane, 2)
    out, sl = updates.residual_slice(h.to(DEV), d.to(DEV), k, kn)
    assert torch.equal(out.cpu(), h + d)
    assert torch.equal(sl.cpu(), (h + d)[:, k:k + kn])
    ref, _ = ou.dual_update(h, d, k)
    assert rel_l2(out, ref) < 1e-7


def test_axpby_and_autograd():
    x, y = seeded((1000,), 3), seeded((1000,), 4)
    got = updates.axpby(0.75, x.to(DEV), -1.5, y.to(DEV))
    assert rel_l2(got, ou.axpby(0.75, x, -1.5, y)) < 1e-7
    a = seeded((2, 3, 4, 8), 5).to(DEV).requires_grad_()
    b = seeded((2, 3, 4, 8), 6).to(DEV).requires_grad_()
    out, sl = updates.residual_slice(a, b, 1)
    (out.sum() + 2 * sl.sum()).backward()
    want = torch.ones_like(a)
    want[:, 1] += 2
    assert torch.equal(a.grad, want) and torch.equal(b.grad, want)
    p, q = (seeded((2, c, 4, 4), 7 + c).to(DEV).requires_grad_() for c in (2, 3))
    w = seeded((2, 5, 4, 4), 11).to(DEV)
    (updates.concat(p, q) * w).sum().backward()
    assert torch.equal(p.grad, w[:, :2]) and torch.equal(q.grad, w[:, 2:])


@pytest.mark.parametrize("mode", ["flip", "periodic", "clamp"])
@pytest.mark.parametrize("shape,factor", [((2, 8, 16), 4), ((1, 64, 256), 8), ((3, 5, 7), 3), ((2, 6, 10), 1)])
def test_angular_upsample_and_its_transpose(mode, shape, factor):
    s = seeded(shape, 1)
    up = updates.angular_upsample(s.to(DEV), factor, mode)
    assert rel_l2(up, ou.angular_upsample(s, factor, mode)) < 1e-7
    g = seeded(tuple(up.shape), 2)
    down = updates.angular_upsample_adjoint(g.to(DEV), factor, mode)
    assert rel_l2(down, ou.angular_upsample_adjoint(g, factor, mode)) < 1e-6
    sr = s.to(DEV).requires_grad_()
    (updates.angular_upsample(sr, factor, mode) * g.to(DEV)).sum().backward()
    assert torch.equal(sr.grad, down)


def test_errors():
    a = seeded((2, 2, 4, 4), 1)
    with pytest.raises(pdu.PduError):
        updates.concat(a, a)
    with pytest.raises(ValueError):
        updates.residual_slice(a.to(DEV), a.to(DEV), 3)
    with pytest.raises(ValueError):
        updates.angular_upsample(a.to(DEV), 2, "nearest")


@pytest.mark.parametrize("plane", [(8, 16), (5, 7)])
def test_channels_last_layout_and_scaled_concat(plane):
    a, b, c = (seeded((3, ch) + plane, i) for i, ch in enumerate((4, 1, 2)))
    al = a.to(DEV).contiguous(memory_format=torch.channels_last)
    cl = c.to(DEV).contiguous(memory_format=torch.channels_last)
    out = updates.concat(al, b.to(DEV), cl, scale_b=0.25)
    assert out.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(out.cpu(), torch.cat([a, 0.25 * b, c], 1))
    assert torch.equal(updates.concat(a.to(DEV), b.to(DEV), c.to(DEV), scale_b=0.25).cpu(), torch.cat([a, 0.25 * b, c], 1))
    h, d = seeded((2, 4) + plane, 5), seeded((2, 4) + plane, 6)
    hl = h.to(DEV).contiguous(memory_format=torch.channels_last)
    dl = d.to(DEV).contiguous(memory_format=torch.channels_last)
    res, sl = updates.residual_slice(hl, dl, 1, 2)
    assert res.is_contiguous(memory_format=torch.channels_last) and sl.is_contiguous()
    assert torch.equal(res.cpu(), h + d) and torch.equal(sl.cpu(), (h + d)[:, 1:3])
    # mixed: planar state, channels-last delta (what the first unrolled iteration may see)
    res2, sl2 = updates.residual_slice(h.to(DEV), dl, 0, 1)
    assert torch.equal(res2.cpu(), h + d) and torch.equal(sl2.cpu(), (h + d)[:, :1])
    # gradients flow through the scaled input
    br = b.to(DEV).requires_grad_()
    (updates.concat(al, br, cl, scale_b=0.25) * 2.0).sum().backward()
    assert torch.equal(br.grad, torch.full_like(br, 0.5))


def test_graphed_inference_matches_eager():
    import numpy as np
    from pd_unet_b200.graph import GraphedInference
    from pd_unet_b200.model import PrimalDualUNetCT
    n, A, up = 64, 32, 4
    radon = pdu.Radon(n, np.linspace(0, np.pi, A, endpoint=False))
    torch.manual_seed(0)
    model = PrimalDualUNetCT(radon, upsample=up, n_iter=2, n_primal=2, n_dual=2, unet_base=8, unet_depth=2,
                             dual_features=8).to(DEV).eval()
    x = seeded((2, 1, A // up, n), 3).to(DEV)
    with torch.no_grad():
        want = model(x).clone()
    g = GraphedInference(model, x)
    assert torch.allclose(g(x), want, atol=1e-5, rtol=1e-5)
    x2 = seeded((2, 1, A // up, n), 4)
    with torch.no_grad():
        want2 = model(x2.to(DEV)).clone()
    assert torch.allclose(g(x2.pin_memory()), want2, atol=1e-5, rtol=1e-5)     # host input: H2D inside the call


@pytest.mark.parametrize("shape", [(2, 8, 6, 10), (3, 5, 7, 9), (1, 32, 16, 16)])
@pytest.mark.parametrize("channels_last", [True, False])
@pytest.mark.parametrize("slope_kind", ["per_channel", "single", "none"])
def test_bias_prelu_epilogue(shape, channels_last, slope_kind):
    y = seeded(shape, 1)
    bias = seeded((shape[1],), 2)
    slope = {"per_channel": seeded((shape[1],), 3).abs(), "single": torch.tensor([0.25]), "none": None}[slope_kind]
    want = y + bias.view(1, -1, 1, 1)
    if slope is not None:
        want = torch.nn.functional.prelu(want, slope)
    yd = y.to(DEV)
    if channels_last:
        yd = yd.contiguous(memory_format=torch.channels_last)
    out = updates.bias_prelu_(yd, bias.to(DEV), slope.to(DEV) if slope is not None else None)
    assert out.data_ptr() == yd.data_ptr()
    assert torch.equal(out.cpu(), want)


def test_fused_conv_modules_match_stock_modules():
    from pd_unet_b200.model import ConvAct, UpConv
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    for mod, x in ((ConvAct(6, 8).to(DEV), seeded((2, 6, 12, 20), 1)), (ConvAct(8, 4, act=False, kernel=1).to(DEV), seeded((2, 8, 9, 9), 2)),
                   (UpConv(8, 4).to(DEV), seeded((2, 8, 6, 6), 3))):
        mod = mod.to(memory_format=torch.channels_last)
        xd = x.to(DEV).contiguous(memory_format=torch.channels_last)
        with torch.enable_grad():
            want = mod(xd.clone().requires_grad_()).detach()      # stock path
        with torch.no_grad():
            got = mod(xd)                                         # fused epilogue
        assert torch.allclose(got, want, atol=1e-5, rtol=1e-5)


def test_concat_fast_path_4_1_1_to_8():
    a, b, c = seeded((2, 4, 6, 10), 1), seeded((2, 1, 6, 10), 2), seeded((2, 1, 6, 10), 3)
    al = a.to(DEV).contiguous(memory_format=torch.channels_last)
    out = updates.concat(al, b.to(DEV), c.to(DEV), scale_b=0.5, pad_to=8)
    assert torch.equal(out.cpu(), torch.cat([a, 0.5 * b, c, torch.zeros(2, 2, 6, 10)], 1))
    out2 = updates.concat(al, b.to(DEV), scale_b=2.0, pad_to=8)
    assert torch.equal(out2.cpu(), torch.cat([a, 2.0 * b, torch.zeros(2, 3, 6, 10)], 1))


def test_concat_zero_pads_channels_for_the_tensor_cores():
    a, b, c = (seeded((2, ch, 6, 10), i) for i, ch in enumerate((4, 1, 1)))
    want = torch.cat([a, 0.5 * b, c, torch.zeros(2, 2, 6, 10)], 1)
    al = a.to(DEV).contiguous(memory_format=torch.channels_last)
    out = updates.concat(al, b.to(DEV), c.to(DEV), scale_b=0.5, pad_to=8)
    assert out.shape == (2, 8, 6, 10) and torch.equal(out.cpu(), want)
    assert torch.equal(updates.concat(a.to(DEV), b.to(DEV), c.to(DEV), scale_b=0.5, pad_to=8).cpu(), want)
    # a convolution fed the padded tensor (fused path pads its weights) equals the stock one on 6 channels
    from pd_unet_b200.model import ConvAct
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    conv = ConvAct(6, 8).to(DEV).to(memory_format=torch.channels_last)
    with torch.enable_grad():
        ref = conv(want[:, :6].to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_()).detach()
    with torch.no_grad():
        got = conv(out)
    assert torch.allclose(got, ref, atol=1e-5, rtol=1e-5)


def test_fused_unet_path_matches_the_plain_modules():
    """UNet inference path with the one-pass encoder epilogue (bias + PReLU + skip placement + 2x2 max) against
    the stock-module path on the same weights."""
    from pd_unet_b200.model import UNet
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    net = UNet(5, 4, base=8, depth=2).to(DEV).to(memory_format=torch.channels_last).eval()
    x = seeded((2, 5, 32, 24), 1).to(DEV).contiguous(memory_format=torch.channels_last)
    with torch.enable_grad():
        want = net(x.clone().requires_grad_()).detach()       # stock modules (autograd path)
    with torch.no_grad():
        assert net._fused_ok(x)
        got = net(x)
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-5)
    # sizes the fused path cannot pool evenly fall back to the stock path
    x2 = seeded((1, 5, 18, 22), 2).to(DEV).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        assert not net._fused_ok(x2) and net(x2).shape == (1, 4, 18, 22)
    # the kernel itself, against ATen
    y = seeded((2, 8, 6, 10), 3).to(DEV).contiguous(memory_format=torch.channels_last)
    bias, slope = seeded((8,), 4).to(DEV), seeded((8,), 5).abs().to(DEV)
    cat = torch.zeros((2, 16, 6, 10), device=DEV).contiguous(memory_format=torch.channels_last)
    pooled = torch.empty((2, 8, 3, 5), device=DEV).contiguous(memory_format=torch.channels_last)
    updates.bias_prelu_place_(y, bias, slope, cat[:, 8:], pooled)
    ref = torch.nn.functional.prelu(y + bias.view(1, -1, 1, 1), slope)
    assert torch.equal(cat[:, 8:], ref) and torch.equal(cat[:, :8], torch.zeros_like(ref))
    assert torch.equal(pooled, torch.nn.functional.max_pool2d(ref, 2))


@pytest.mark.parametrize("C,n_slope", [(32, 32), (64, 1), (8, 8), (256, 256)])
def test_bias_prelu_training_matches_aten(C, n_slope):
    """The differentiable epilogue (pdu_bias_prelu_fwd_f32 / _bwd_f32) against conv2d(bias) + PReLU in ATen:
    output, input gradient, bias gradient and slope gradient; and the reduction is bit-reproducible."""
    torch.manual_seed(7)
    B, H, W = 3, 24, 40
    x = torch.randn(B, 5, H, W, device=DEV).contiguous(memory_format=torch.channels_last)
    conv = torch.nn.Conv2d(5, C, 3, padding=1).to(DEV).to(memory_format=torch.channels_last)
    act = torch.nn.PReLU(n_slope).to(DEV)
    monkey = pytest.MonkeyPatch()
    monkey.setattr(updates, "FUSED_TRAIN_MIN_ELEMS", 0)      # the test tensors are small: force the fused kernels
    with torch.no_grad():
        act.weight.uniform_(-0.2, 0.6)        # negative and zero-crossing slopes included
    w = torch.randn(B, C, H, W, device=DEV)

    def run(fused):
        for p in list(conv.parameters()) + list(act.parameters()):
            p.grad = None
        xr = x.clone().requires_grad_()
        if fused:
            y = torch.nn.functional.conv2d(xr, conv.weight, None, padding=1)
            out = pdu.updates.bias_prelu(y, conv.bias, act.weight)
        else:
            out = act(conv(xr))
        (out * w).sum().backward()
        return out.detach(), xr.grad, conv.bias.grad.clone(), act.weight.grad.clone(), conv.weight.grad.clone()

    ref = run(False)
    got = run(True)
    again = run(True)
    for name, a, b in zip(("out", "grad_x", "grad_bias", "grad_slope", "grad_weight"), got, ref):
        assert rel_l2(a, b) <= 2e-5, name          # TF32 convolutions on both sides; the epilogue itself is exact fp32
    assert torch.equal(got[2], again[2]) and torch.equal(got[3], again[3])
    # odd layouts fall back to the ATen ops on the GPU (same numbers)
    yp = torch.randn(2, 6, 9, 9, device=DEV, requires_grad=True)
    o = pdu.updates.bias_prelu(yp, torch.randn(6, device=DEV), torch.full((1,), 0.25, device=DEV))
    assert o.shape == yp.shape
    monkey.undo()


@pytest.mark.parametrize("C", [4, 32, 128])
def test_bias_add_training_matches_aten(C):
    """updates.bias_add (in-place add + pdu_channel_sum_f32 in the backward) against conv2d(bias=True)."""
    torch.manual_seed(11)
    x = torch.randn(2, 6, 20, 28, device=DEV).contiguous(memory_format=torch.channels_last)
    conv = torch.nn.Conv2d(6, C, 3, padding=1).to(DEV).to(memory_format=torch.channels_last)
    w = torch.randn(2, C, 20, 28, device=DEV)
    monkey = pytest.MonkeyPatch()
    monkey.setattr(updates, "FUSED_TRAIN_MIN_ELEMS", 0)

    def run(fused):
        conv.zero_grad(set_to_none=True)
        xr = x.clone().requires_grad_()
        out = updates.bias_add(torch.nn.functional.conv2d(xr, conv.weight, None, padding=1), conv.bias) if fused else conv(xr)
        (out * w).sum().backward()
        return out.detach(), xr.grad, conv.bias.grad.clone(), conv.weight.grad.clone()

    ref, got, again = run(False), run(True), run(True)
    monkey.undo()
    for name, a, b in zip(("out", "grad_x", "grad_bias", "grad_weight"), got, ref):
        assert rel_l2(a, b) <= 2e-5, name
    assert torch.equal(got[2], again[2])


@pytest.mark.parametrize("mode", ["flip", "periodic", "clamp"])
@pytest.mark.parametrize("ca,shape,factor,pad_to,nhwc", [(4, (2, 8, 16), 8, 8, True), (4, (3, 5, 7), 3, 0, True),
                                                         (3, (2, 6, 12), 4, 0, False), (4, (1, 64, 256), 8, 8, False)])
def test_concat_with_the_upsampling_fused(mode, ca, shape, factor, pad_to, nhwc):
    """the dual update's concatenation interpolates the sparse-view sinogram on the fly; the
    result is torch.cat([a, scale_b b, scale_c upsample(sparse)]) to the last bit (the same single FMA per element)."""
    B, As, D = shape
    sparse = seeded((B, 1, As, D), 1)
    a = seeded((B, ca, As * factor, D), 2)
    b = seeded((B, 1, As * factor, D), 3)
    ad = a.to(DEV).contiguous(memory_format=torch.channels_last) if nhwc else a.to(DEV)
    got = updates.concat_upsampled(ad, b.to(DEV), sparse.to(DEV), factor, mode, scale_b=0.5, scale_c=0.25, pad_to=pad_to)
    up = updates.angular_upsample(sparse.to(DEV), factor, mode)
    want = torch.cat([a.to(DEV), 0.5 * b.to(DEV), 0.25 * up], 1)
    assert got.shape[1] == (8 if pad_to else ca + 2)
    assert torch.equal(got[:, :ca + 2], want)
    assert not got[:, ca + 2:].any()
    assert rel_l2(got[:, ca + 1], 0.25 * ou.angular_upsample(sparse[:, 0], factor, mode)) < 1e-7
    # gradients reach a and b (the measured data takes none)
    ar, br = ad.clone().requires_grad_(), b.to(DEV).requires_grad_()
    w = seeded(tuple(got.shape), 4).to(DEV)
    (updates.concat_upsampled(ar, br, sparse.to(DEV), factor, mode, scale_b=0.5, scale_c=0.25, pad_to=pad_to) * w).sum().backward()
    assert torch.equal(ar.grad, w[:, :ca]) and torch.equal(br.grad, 0.5 * w[:, ca:ca + 1])


@pytest.mark.parametrize("ca,cb,cc,pad_to", [(4, 16, 16, 0), (4, 2, 0, 8), (3, 5, 1, 0)])
def test_concat_takes_planar_operands_into_a_channels_last_state(ca, cb, cc, pad_to):
    """The MRI data-space update concatenates a channels-last state with the NUFFT's planar multi-channel output: the
    layout change happens inside the concatenation (pdu_concat_mixed_f32), with torch.cat's values to the last bit."""
    B, plane = 2, (12, 20)
    a = seeded((B, ca) + plane, 1).to(DEV).contiguous(memory_format=torch.channels_last)
    b = seeded((B, cb) + plane, 2).to(DEV)                      # planar
    c = seeded((B, cc) + plane, 3).to(DEV) if cc else None
    got = updates.concat(a, b, c, scale_b=0.25, pad_to=pad_to)
    want = torch.cat([a, 0.25 * b] + ([c] if cc else []), 1)
    assert got.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(got[:, :ca + cb + cc], want)
    assert not got[:, ca + cb + cc:].any()
