"""Float64 CPU restatement of the radial-MRI operators (TEST INFRASTRUCTURE, parity unpinned).

The reference names torchkbnufft as its MRI operator library (BASELINE.json
north_star; /root/reference/README.md:3-5 only points at unmounted branches).
The library is not in the image, so the structure below is [RECALL] of
torchkbnufft >= 1.0 (`KbNufft`, `KbNufftAdjoint`, `kb_table_interp`,
`calc_density_compensation_function`), itself a port of Fessler's NUFFT:

    forward :  x  --(* scaling_coef)--> zero-pad to grid --> FFT
                  --> Kaiser-Bessel table interpolation at tm = omega / (2 pi / K)
                  --> * exp(+i omega . n_shift)
    adjoint :  exact conjugate transpose of the above (scatter, un-normalised
               inverse FFT, crop, * conj(scaling_coef))

* J = numpoints taps per axis at grid offsets floor(tm - J/2) + 1 + j;
* the table holds kb(u) * exp(-i (2 pi / K) ((N-1)/2) u) at u = q / L - J/2,
  q = 0 .. J L, and is read with NEAREST lookup, q = rint((tm - g) L) + J L / 2;
* scaling_coef[n] = 1 / FT{kb}((n - (N-1)/2) / K), real;
* kb(u) = I0(alpha sqrt(1 - (2u/J)^2)) / I0(alpha), alpha = kbwidth * J.

The mathematical object being approximated is
    y_k = sum_n x_n exp(-i omega_k . (n - n_shift)),
which `ndft_forward` evaluates exactly for small sizes.

The grid offset and the table index are computed in IEEE float32 op for op
what pd_unet_b200/csrc/nufft.cu does (`_tap_indices_f32`), so oracle and
kernel read the same table entries; sums are complex128.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch
from scipy import special


@dataclass(frozen=True)
class NufftSpec:
    im_size: Tuple[int, int]
    grid_size: Optional[Tuple[int, int]] = None
    numpoints: int = 6
    n_shift: Optional[Tuple[int, int]] = None
    table_oversamp: int = 1024
    kbwidth: float = 2.34
    order: float = 0.0

    def resolved(self) -> "NufftSpec":
        gs = self.grid_size or tuple(2 * n for n in self.im_size)
        ns = self.n_shift or tuple(n // 2 for n in self.im_size)
        if self.order != 0.0:
            raise NotImplementedError("only order 0 Kaiser-Bessel is restated")
        return NufftSpec(tuple(self.im_size), tuple(gs), self.numpoints, tuple(ns),
                         self.table_oversamp, self.kbwidth, self.order)

    @property
    def alpha(self) -> float:
        return self.kbwidth * self.numpoints


# kb_table / scaling_coef are written independently of pd_unet_b200/nufft.py on purpose: the
# product evaluates np.i0 on a vector grid and the closed-form transform sinh(w)/w; here the kernel goes through the
# exponentially scaled Bessel function entry by entry, and the apodisation is the kernel's Fourier transform by
# Gauss-Legendre quadrature -- which also checks the closed form itself.
def _kb(u: float, J: int, alpha: float) -> float:
    """Kaiser-Bessel kernel of order 0 and width J at offset u (grid units): I0(alpha sqrt(1 - (2u/J)^2)) / I0(alpha)."""
    r = 2.0 * u / J
    if abs(r) >= 1.0:
        return 0.0
    s = math.sqrt(1.0 - r * r)
    # I0(x) = i0e(x) e^x: the ratio stays in range for any alpha
    return float(special.i0e(alpha * s) / special.i0e(alpha) * math.exp(alpha * (s - 1.0)))


def kb_table(spec: NufftSpec, dim: int) -> np.ndarray:
    """complex128 [J*L + 1]; entry q is the coefficient at u = (q - J L / 2) / L: the kernel times the linear phase
    exp(-i (2 pi / K) ((N - 1) / 2) u) that centres the image ([RECALL] Fessler's / torchkbnufft's table)."""
    s = spec.resolved()
    J, L = s.numpoints, s.table_oversamp
    N, K = s.im_size[dim], s.grid_size[dim]
    half = (J * L) // 2
    out = np.empty(J * L + 1, dtype=np.complex128)
    slope = (2.0 * math.pi / K) * ((N - 1) / 2.0)
    for q in range(J * L + 1):
        u = (q - half) / L
        out[q] = _kb(u, J, s.alpha) * complex(math.cos(slope * u), -math.sin(slope * u))
    return out


def scaling_coef(spec: NufftSpec, dim: int, nodes: int = 96) -> np.ndarray:
    """float64 [N]: 1 / FT{kb}((n - (N-1)/2) / K), the transform taken numerically:
    FT(f) = int_{-J/2}^{J/2} kb(u) cos(2 pi f u) du by Gauss-Legendre quadrature (the integrand is smooth inside the
    support and the kernel is even; 96 nodes reach 1e-15 relative for J = 6, alpha = 14)."""
    s = spec.resolved()
    J = s.numpoints
    N, K = s.im_size[dim], s.grid_size[dim]
    x, w = np.polynomial.legendre.leggauss(nodes)
    u = 0.5 * J * x                                              # [-J/2, J/2]
    kb = np.array([_kb(float(v), J, s.alpha) for v in u])
    f = (np.arange(N, dtype=np.float64) - (N - 1) / 2.0) / K
    ft = (np.cos(2.0 * np.pi * f[:, None] * u[None, :]) * (kb * w)[None, :]).sum(1) * (0.5 * J)
    return 1.0 / ft


def _tap_indices_f32(omega_d: np.ndarray, K: int, J: int, L: int):
    """float32 restatement of the per-axis tap selection.
    Returns grid index [M, J] (already wrapped mod K) and table index [M, J]."""
    f = np.float32
    gam = f(2.0 * np.pi / K)
    tm = omega_d.astype(f) / gam
    koff = np.floor(tm - f(J / 2.0)).astype(np.int64)
    g = koff[:, None] + 1 + np.arange(J, dtype=np.int64)[None, :]
    dist = (tm[:, None] - g.astype(f)) * f(L)
    q = np.rint(dist).astype(np.int64) + (J * L) // 2
    q = np.clip(q, 0, J * L)
    return np.mod(g, K), q


def _taps(omega: np.ndarray, spec: NufftSpec, table_c64: bool = True):
    s = spec.resolved()
    J, L = s.numpoints, s.table_oversamp
    K0, K1 = s.grid_size
    t0, t1 = kb_table(s, 0), kb_table(s, 1)
    if table_c64:                      # the kernels hold the tables as complex64
        t0 = t0.astype(np.complex64).astype(np.complex128)
        t1 = t1.astype(np.complex64).astype(np.complex128)
    g0, q0 = _tap_indices_f32(omega[0], K0, J, L)
    g1, q1 = _tap_indices_f32(omega[1], K1, J, L)
    flat = (g0[:, :, None] * K1 + g1[:, None, :]).reshape(omega.shape[1], J * J)
    coef = (t0[q0][:, :, None] * t1[q1][:, None, :]).reshape(omega.shape[1], J * J)
    om64 = omega.astype(np.float64)
    phase = np.exp(1j * (om64[0] * s.n_shift[0] + om64[1] * s.n_shift[1]))
    return torch.from_numpy(flat), torch.from_numpy(coef), torch.from_numpy(phase)


def _as_omega(omega) -> np.ndarray:
    om = omega.detach().cpu().numpy() if isinstance(omega, torch.Tensor) else np.asarray(omega)
    om = om.astype(np.float32)
    assert om.ndim == 2 and om.shape[0] == 2, "omega must be [2, M] radians"
    return om


def interp_forward(grid, omega, spec: NufftSpec, chunk: int = 1 << 15) -> torch.Tensor:
    """grid [B, C, K0, K1] complex -> samples [B, C, M] (table interpolation only)."""
    s = spec.resolved()
    grid = torch.as_tensor(grid).to(torch.complex128)
    B, C = grid.shape[:2]
    om = _as_omega(omega)
    flat, coef, phase = _taps(om, s)
    gf = grid.reshape(B, C, -1)
    out = torch.empty(B, C, om.shape[1], dtype=torch.complex128)
    for m0 in range(0, om.shape[1], chunk):
        m1 = min(om.shape[1], m0 + chunk)
        v = gf[:, :, flat[m0:m1].reshape(-1)].reshape(B, C, m1 - m0, -1)
        out[:, :, m0:m1] = (v * coef[m0:m1]).sum(-1) * phase[m0:m1]
    return out


def interp_adjoint(kdata, omega, spec: NufftSpec) -> torch.Tensor:
    """samples [B, C, M] -> grid [B, C, K0, K1]; exact conjugate transpose of interp_forward."""
    s = spec.resolved()
    kdata = torch.as_tensor(kdata).to(torch.complex128)
    B, C, M = kdata.shape
    K0, K1 = s.grid_size
    om = _as_omega(omega)
    flat, coef, phase = _taps(om, s)
    vals = (kdata * phase.conj())[..., None] * coef.conj()          # [B, C, M, J*J]
    grid = torch.zeros(B, C, K0 * K1, dtype=torch.complex128)
    grid.index_add_(2, flat.reshape(-1), vals.reshape(B, C, -1))
    return grid.reshape(B, C, K0, K1)


def _scal2d(s: NufftSpec) -> torch.Tensor:
    return torch.from_numpy(np.outer(scaling_coef(s, 0), scaling_coef(s, 1)))


def nufft_forward(image, omega, spec: NufftSpec, smaps=None, norm: Optional[str] = None) -> torch.Tensor:
    """image [B, C, N0, N1] complex (C == 1 when smaps [.., Cc, N0, N1] is given) -> [B, C, M]."""
    s = spec.resolved()
    x = torch.as_tensor(image).to(torch.complex128)
    if smaps is not None:
        x = x * torch.as_tensor(smaps).to(torch.complex128)
    x = x * _scal2d(s)
    Z = torch.fft.fft2(x, s=s.grid_size)
    if norm == "ortho":
        Z = Z / math.sqrt(s.grid_size[0] * s.grid_size[1])
    elif norm is not None:
        raise ValueError("norm must be None or 'ortho'")
    return interp_forward(Z, omega, s)


def nufft_adjoint(kdata, omega, spec: NufftSpec, smaps=None, norm: Optional[str] = None) -> torch.Tensor:
    """samples [B, C, M] -> image [B, C, N0, N1] (or [B, 1, N0, N1] with smaps)."""
    s = spec.resolved()
    K0, K1 = s.grid_size
    N0, N1 = s.im_size
    grid = interp_adjoint(kdata, omega, s)
    x = torch.fft.ifft2(grid) * (K0 * K1)
    if norm == "ortho":
        x = x / math.sqrt(K0 * K1)
    elif norm is not None:
        raise ValueError("norm must be None or 'ortho'")
    x = x[..., :N0, :N1] * _scal2d(s)
    if smaps is not None:
        x = (x * torch.as_tensor(smaps).to(torch.complex128).conj()).sum(1, keepdim=True)
    return x


def ndft_forward(image, omega, spec: NufftSpec) -> torch.Tensor:
    """Exact y_k = sum_n x_n exp(-i omega_k . (n - n_shift)); O(N M), small sizes only."""
    s = spec.resolved()
    x = torch.as_tensor(image).to(torch.complex128)
    om = torch.from_numpy(_as_omega(omega).astype(np.float64))
    n0 = torch.arange(s.im_size[0], dtype=torch.float64) - s.n_shift[0]
    n1 = torch.arange(s.im_size[1], dtype=torch.float64) - s.n_shift[1]
    e0 = torch.exp(-1j * om[0][:, None] * n0[None, :])       # [M, N0]
    e1 = torch.exp(-1j * om[1][:, None] * n1[None, :])       # [M, N1]
    return torch.einsum("ma,bcad,md->bcm", e0, x, e1)


def ndft_adjoint(kdata, omega, spec: NufftSpec) -> torch.Tensor:
    s = spec.resolved()
    y = torch.as_tensor(kdata).to(torch.complex128)
    om = torch.from_numpy(_as_omega(omega).astype(np.float64))
    n0 = torch.arange(s.im_size[0], dtype=torch.float64) - s.n_shift[0]
    n1 = torch.arange(s.im_size[1], dtype=torch.float64) - s.n_shift[1]
    e0 = torch.exp(1j * om[0][:, None] * n0[None, :])
    e1 = torch.exp(1j * om[1][:, None] * n1[None, :])
    return torch.einsum("ma,bcm,md->bcad", e0, y, e1)


def calc_dcf(omega, spec: NufftSpec, num_iterations: int = 10) -> torch.Tensor:
    """[RECALL] Pipe-style iteration w <- w / |G G^H w| with the interpolator only.  -> [M] float64."""
    om = _as_omega(omega)
    w = torch.ones(1, 1, om.shape[1], dtype=torch.complex128)
    for _ in range(num_iterations):
        new = interp_forward(interp_adjoint(w, om, spec), om, spec)
        w = w / new.abs()
    return w.real.reshape(-1)


def radial_trajectory(n_spokes: int, n_readout: int, golden: bool = True) -> np.ndarray:
    """[2, n_spokes*n_readout] float32; row 0 pairs with image axis 0.  omega in [-pi, pi)."""
    if golden:
        phi = np.arange(n_spokes, dtype=np.float64) * (111.246117975 * np.pi / 180.0)
    else:
        phi = np.arange(n_spokes, dtype=np.float64) * (np.pi / n_spokes)
    r = (np.arange(n_readout, dtype=np.float64) - n_readout / 2.0) * (2.0 * np.pi / n_readout)
    om0 = (r[None, :] * np.sin(phi)[:, None]).reshape(-1)
    om1 = (r[None, :] * np.cos(phi)[:, None]).reshape(-1)
    return np.stack([om0, om1]).astype(np.float32)
