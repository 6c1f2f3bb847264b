"""Helpers shared by the parity tests."""
import numpy as np
import torch

TOL = 1e-5          # BASELINE.json north_star: rel-L2 <= 1e-5 in fp32


def rel_l2(got, want) -> float:
    got = got.detach().cpu() if isinstance(got, torch.Tensor) else torch.as_tensor(got)
    want = want.detach().cpu() if isinstance(want, torch.Tensor) else torch.as_tensor(want)
    dt = torch.complex128 if (got.is_complex() or want.is_complex()) else torch.float64
    got, want = got.to(dt), want.to(dt)
    den = float(want.norm())
    return float((got - want).norm()) / (den if den > 0 else 1.0)


def seeded(shape, seed=0, complex_=False):
    g = torch.Generator().manual_seed(seed)
    if complex_:
        return torch.complex(torch.randn(shape, generator=g), torch.randn(shape, generator=g)).to(torch.complex64)
    return torch.randn(shape, generator=g, dtype=torch.float32)


def user_angles(n_angles, span=np.pi):
    return np.linspace(0.0, span, n_angles, endpoint=False)
