"""Float64 restatement of the elementwise steps around the operators
(TEST INFRASTRUCTURE, parity unpinned -- see oracle/__init__.py).

The unrolled primal-dual iteration (Adler & Oktem's Learned Primal-Dual, which
the paper cited at /root/reference/README.md:3 extends) is

    h <- h + Gamma(cat(h, K f[:, k], g))        dual (sinogram / k-space side)
    f <- f + Lambda(cat(f, K* h[:, k]))         primal (image side)

The pieces that are not convolutions are: the channel concatenation feeding
each network, the residual add, the channel slice handed to the next operator
call, and (PD-UNet's sinogram upsampling) the linear interpolation of a sparse
set of views onto the full angular grid.  These are what the fused CUDA
kernels in pd_unet_b200/csrc/updates.cu compute.
"""
from __future__ import annotations

import torch


def _f64(x):
    return torch.as_tensor(x, dtype=torch.float64)


def concat(*parts) -> torch.Tensor:
    """cat along the channel axis of [B, c_i, ...] tensors."""
    return torch.cat([_f64(p) for p in parts], dim=1)


def dual_update(h, dh, k: int = 0):
    """-> (h + dh, contiguous copy of channel k of the result)."""
    out = _f64(h) + _f64(dh)
    return out, out[:, k].clone()


def primal_update(f, df, k: int = 0):
    return dual_update(f, df, k)


def axpby(a: float, x, b: float, y) -> torch.Tensor:
    return a * _f64(x) + b * _f64(y)


def _wrap_row(s: torch.Tensor, mode: str) -> torch.Tensor:
    if mode == "flip":        # parallel beam over [0, pi): p(theta + pi, u) = p(theta, -u)
        return s[:, :1].flip(-1)
    if mode == "periodic":    # fan beam over [0, 2 pi)
        return s[:, :1]
    if mode == "clamp":
        return s[:, -1:]
    raise ValueError(mode)


def angular_upsample(sino, factor: int, mode: str = "flip") -> torch.Tensor:
    """[B, As, D] -> [B, As*factor, D]; view i*factor + r = (1 - r/factor) s[i] + (r/factor) s[i+1]."""
    s = _f64(sino)
    B, As, D = s.shape
    ext = torch.cat([s, _wrap_row(s, mode)], dim=1)                 # [B, As+1, D]
    r = torch.arange(factor, dtype=torch.float64) / factor          # [f]
    lo = ext[:, :-1, None, :]
    hi = ext[:, 1:, None, :]
    out = lo * (1.0 - r)[None, None, :, None] + hi * r[None, None, :, None]
    return out.reshape(B, As * factor, D)


def angular_upsample_adjoint(full, factor: int, mode: str = "flip") -> torch.Tensor:
    """Exact transpose of angular_upsample: [B, As*factor, D] -> [B, As, D]."""
    g = _f64(full)
    B, A, D = g.shape
    As = A // factor
    g = g.reshape(B, As, factor, D)
    r = torch.arange(factor, dtype=torch.float64) / factor
    lo = (g * (1.0 - r)[None, None, :, None]).sum(2)                # weight on s[i]
    hi = (g * r[None, None, :, None]).sum(2)                        # weight on s[i+1]
    out = lo.clone()
    out[:, 1:] += hi[:, :-1]
    last = hi[:, -1]
    if mode == "flip":
        out[:, 0] += last.flip(-1)
    elif mode == "periodic":
        out[:, 0] += last
    elif mode == "clamp":
        out[:, -1] += last
    else:
        raise ValueError(mode)
    return out
