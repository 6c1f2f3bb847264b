"""On-GPU synthetic data for PD-UNet: the undersampled measurements the model
consumes are produced by the operators themselves, so no CPU NUFFT / Radon preprocessing is needed
(BASELINE.json configs[0] prepares its k-space with torchkbnufft on the CPU).  Phantoms are Shepp-Logan-style
ellipse sets with per-slice variation (pd_unet_b200.phantoms); there is no dataset offline."""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from .nufft import KbNufft, calc_density_compensation_function
from .phantoms import coil_maps, phantom_batch
from .radon import _BaseRadon


def radial_trajectory(n_spokes: int, n_readout: int, golden: bool = True, device=None) -> torch.Tensor:
    """[2, n_spokes * n_readout] float32 radians in [-pi, pi); row 0 pairs with image axis 0."""
    if golden:
        phi = np.arange(n_spokes, dtype=np.float64) * (111.246117975 * np.pi / 180.0)
    else:
        phi = np.arange(n_spokes, dtype=np.float64) * (np.pi / n_spokes)
    r = (np.arange(n_readout, dtype=np.float64) - n_readout / 2.0) * (2.0 * np.pi / n_readout)
    om = np.stack([(r[None, :] * np.sin(phi)[:, None]).reshape(-1), (r[None, :] * np.cos(phi)[:, None]).reshape(-1)])
    return torch.from_numpy(om.astype(np.float32)).to(device) if device is not None else torch.from_numpy(om.astype(np.float32))


def make_ct_batch(radon_full: _BaseRadon, batch: int, upsample: int, seed: int = 0, device="cuda") -> Dict[str, torch.Tensor]:
    """image [B, 1, N, N], full sinogram [B, 1, A, D] and its every-`upsample`-th-view subset [B, 1, A/up, D]."""
    n = radon_full.resolution
    image = phantom_batch(batch, n, seed=seed).to(device)[:, None]
    full = radon_full.forward(image)
    return {"image": image, "sino_full": full, "sino_sparse": full[:, :, ::upsample].contiguous()}


def make_mri_batch(im_size, n_spokes: int, coils: int, batch: int, seed: int = 0, device="cuda") -> Dict[str, torch.Tensor]:
    """Complex phantom images, golden-angle radial k-space (readout = 2 N) through `coils` sensitivity maps,
    the trajectory and its density compensation -- everything computed on the GPU."""
    n = im_size[0]
    g = torch.Generator().manual_seed(seed)
    mag = phantom_batch(batch, n, seed=seed)
    phase = 0.3 * torch.rand(batch, 1, 1, generator=g) * torch.linspace(-1, 1, n)[None, None, :]
    image = (mag * torch.exp(1j * phase)).to(torch.complex64).to(device)[:, None]
    omega = radial_trajectory(n_spokes, 2 * n, device=device)
    smaps = coil_maps(coils, n)[None].to(device) if coils > 1 else None
    kdata = KbNufft(im_size)(image, omega, smaps=smaps, norm="ortho")
    dcf = calc_density_compensation_function(omega, im_size)
    return {"image": image, "kdata": kdata, "omega": omega, "smaps": smaps, "dcf": dcf}
