"""CPU check of the PD-UNet assembly's plumbing (shapes, layouts, channel bookkeeping) with the
CUDA-only fused ops and operators swapped for torch / oracle stand-ins.  The product path itself is
covered by tests/test_gpu_model.py; this only keeps host-side mistakes from reaching the GPU box."""
import numpy as np
import pytest
import torch

import oracle
from oracle import updates as ou
from pd_unet_b200 import model as M, updates


@pytest.fixture
def cpu_ops(monkeypatch):
    monkeypatch.setattr(updates, "concat", lambda a, b, c=None, scale_b=1.0, pad_to=0:
                        torch.cat([a, scale_b * b] + ([c] if c is not None else []), 1))

    def residual_slice(state, delta, k=0, kn=1):
        out = state + delta
        return out, out[:, k:k + kn].contiguous()
    monkeypatch.setattr(updates, "residual_slice", residual_slice)
    monkeypatch.setattr(updates, "angular_upsample",
                        lambda s, f, mode="flip": ou.angular_upsample(s.reshape(-1, *s.shape[-2:]), f, mode)
                        .reshape(*s.shape[:-2], s.shape[-2] * f, s.shape[-1]).float())


class _FakeRadon(M._BaseRadon):
    def __init__(self, n, angles):
        super().__init__(n, angles, n, 1.0, False, 0)
        self._g = oracle.RadonGeom(n=n, n_angles=len(angles), det_count=n)
        self._t = oracle.trig_table(-np.asarray(angles))

    def forward(self, x):
        return oracle.radon_forward(x.reshape(-1, *x.shape[-2:]), self._t, self._g).float().reshape(*x.shape[:-2], -1, x.shape[-1])

    def fbp(self, s, filter_name="ramp"):
        return oracle.fbp(s.reshape(-1, *s.shape[-2:]), self._t, self._g).float().reshape(*s.shape[:-2], s.shape[-1], s.shape[-1])


@pytest.mark.parametrize("channels_last", [True, False])
def test_ct_model_runs_and_keeps_shapes(cpu_ops, channels_last):
    n, A, up = 16, 8, 2
    radon = _FakeRadon(n, np.linspace(0, np.pi, A, endpoint=False))
    torch.manual_seed(0)
    net = M.PrimalDualUNetCT(radon, upsample=up, n_iter=2, n_primal=3, n_dual=2, unet_base=4, unet_depth=2,
                             dual_features=4, channels_last=channels_last).eval()
    assert all("radon" not in k for k in net.state_dict())
    with torch.no_grad():
        out = net(torch.randn(2, 1, A // up, n))
    assert out.shape == (2, 1, n, n) and torch.isfinite(out).all()


def test_unet_handles_odd_sizes():
    net = M.UNet(3, 2, base=4, depth=2).eval()
    with torch.no_grad():
        assert net(torch.randn(1, 3, 20, 28)).shape == (1, 2, 20, 28)
        assert net(torch.randn(1, 3, 18, 22)).shape == (1, 2, 18, 22)
