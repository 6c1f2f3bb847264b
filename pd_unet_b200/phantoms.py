"""Synthetic inputs shared by the tests and bench.py.
No datasets exist offline; everything is generated from fixed tables and seeds."""
from __future__ import annotations

import numpy as np
import torch

# (value, a, b, x0, y0, phi_deg) -- modified Shepp-Logan ellipse table, unit square [-1, 1]^2
_ELLIPSES = [
    (1.0, .69, .92, 0, 0, 0), (-.8, .6624, .874, 0, -.0184, 0),
    (-.2, .11, .31, .22, 0, -18), (-.2, .16, .41, -.22, 0, 18),
    (.1, .21, .25, 0, .35, 0), (.1, .046, .046, 0, .1, 0), (.1, .046, .046, 0, -.1, 0),
    (.1, .046, .023, -.08, -.605, 0), (.1, .023, .023, 0, -.606, 0), (.1, .023, .046, .06, -.605, 0),
]


def shepp_logan(n: int, scale: float = 0.95) -> np.ndarray:
    c = (np.arange(n) + 0.5 - n / 2.0) / (n / 2.0) / scale
    x, y = np.meshgrid(c, -c)
    img = np.zeros((n, n))
    for val, a, b, x0, y0, phi in _ELLIPSES:
        p = np.deg2rad(phi)
        xr = (x - x0) * np.cos(p) + (y - y0) * np.sin(p)
        yr = -(x - x0) * np.sin(p) + (y - y0) * np.cos(p)
        img[(xr / a) ** 2 + (yr / b) ** 2 <= 1.0] += val
    return img


def phantom_batch(batch: int, n: int, seed: int = 0, noise: float = 0.01) -> torch.Tensor:
    """[batch, n, n] float32: Shepp-Logan (slightly different scale per slice) + U[0, noise)."""
    g = torch.Generator().manual_seed(seed)
    out = torch.empty(batch, n, n, dtype=torch.float32)
    for b in range(batch):
        base = torch.from_numpy(shepp_logan(n, 0.95 - 0.02 * (b % 8))).float()
        out[b] = base + noise * torch.rand(n, n, generator=g)
    return out


def disc(n: int, radius: float, cx: float = 0.0, cy: float = 0.0, supersample: int = 8) -> np.ndarray:
    """Area-sampled disc (world units = pixels, origin at the image centre)."""
    ss = supersample
    c = (np.arange(n * ss) + 0.5) / ss - n / 2.0
    x, y = np.meshgrid(c, c)
    m = ((x - cx) ** 2 + (y - cy) ** 2 <= radius ** 2).astype(np.float64)
    return m.reshape(n, ss, n, ss).mean(axis=(1, 3))


def coil_maps(n_coils: int, n: int) -> torch.Tensor:
    """[n_coils, n, n] complex64, Gaussian lobes on a ring, sum |S|^2 = 1."""
    c = (np.arange(n) + 0.5 - n / 2.0) / (n / 2.0)
    x, y = np.meshgrid(c, c)
    maps = []
    for k in range(n_coils):
        a = 2 * np.pi * k / n_coils
        mag = np.exp(-((x - 1.1 * np.cos(a)) ** 2 + (y - 1.1 * np.sin(a)) ** 2) / (2 * 0.8 ** 2))
        maps.append(mag * np.exp(1j * (a + 0.5 * x * np.cos(a) + 0.5 * y * np.sin(a))))
    s = np.stack(maps)
    s = s / np.sqrt((np.abs(s) ** 2).sum(0, keepdims=True))
    return torch.from_numpy(s.astype(np.complex64))
