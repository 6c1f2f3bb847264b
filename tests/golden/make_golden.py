"""Generates tests/golden/golden_v1.npz from the CPU oracle (python tests/golden/make_golden.py).

The reference mount holds no golden vectors and the libraries it uses cannot be
imported here, so these fixtures freeze the oracle's own outputs on small seeded inputs: they guard
the oracle against drift (tests/test_golden.py, CPU) and give the CUDA path a committed target that
does not depend on the oracle code at test time (tests/test_gpu_golden.py).  If torch_radon /
torchkbnufft ever become available, add vectors dumped from them here and compare all three."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import updates as ou  # noqa: E402
from oracle.radon import FAN  # noqa: E402


def seeded(shape, seed, complex_=False):
    g = torch.Generator().manual_seed(seed)
    if complex_:
        return torch.complex(torch.randn(shape, generator=g), torch.randn(shape, generator=g)).to(torch.complex64)
    return torch.randn(shape, generator=g, dtype=torch.float32)


def main():
    out = {}
    # CT, parallel beam: 32^2, 10 views over pi, 40 bins
    n, A, D = 32, 10, 40
    ang = np.linspace(0, np.pi, A, endpoint=False)
    g = oracle.RadonGeom(n=n, n_angles=A, det_count=D)
    trig = oracle.trig_table(-ang)
    x = seeded((2, n, n), 1)
    s = seeded((2, A, D), 2)
    out.update(par_angles=ang, par_x=x.numpy(), par_s=s.numpy(),
               par_fwd=oracle.radon_forward(x, trig, g).numpy(),
               par_adj=oracle.radon_backprojection(s, trig, g).numpy(),
               par_filt=oracle.filter_sinogram(s).numpy(),
               par_filt_hann=oracle.filter_sinogram(s, "hann").numpy())
    # CT, fan beam: 32^2, 12 views over 2 pi, source = detector distance = 64, circle clip
    A2 = 12
    ang2 = np.linspace(0, 2 * np.pi, A2, endpoint=False)
    gf = oracle.RadonGeom(n=n, n_angles=A2, det_count=n, det_spacing=2.0, geom=FAN, s_dist=64.0, d_dist=64.0,
                          clip_to_circle=True)
    trig2 = oracle.trig_table(-ang2)
    s2 = seeded((2, A2, n), 3)
    out.update(fan_angles=ang2, fan_s=s2.numpy(), fan_fwd=oracle.radon_forward(x, trig2, gf).numpy(),
               fan_adj=oracle.radon_backprojection(s2, trig2, gf).numpy())
    # MRI: 16^2 image, 5 golden-angle spokes of 32 samples, 2 coils
    im = (16, 16)
    spec = oracle.NufftSpec(im)
    om = oracle.radial_trajectory(5, 32)
    img = seeded((1, 2) + im, 4, complex_=True)
    k = seeded((1, 2, om.shape[1]), 5, complex_=True)
    out.update(mri_omega=om, mri_img=img.numpy(), mri_k=k.numpy(),
               mri_fwd=oracle.nufft_forward(img, om, spec).numpy(),
               mri_adj=oracle.nufft_adjoint(k, om, spec).numpy(),
               mri_fwd_ortho=oracle.nufft_forward(img, om, spec, norm="ortho").numpy(),
               mri_dcf=oracle.calc_dcf(om, spec, 5).numpy())
    # updates
    sp = seeded((2, 4, 6), 6)
    out.update(up_in=sp.numpy(), up_flip=ou.angular_upsample(sp, 3, "flip").numpy(),
               up_periodic=ou.angular_upsample(sp, 3, "periodic").numpy(),
               up_flip_adj=ou.angular_upsample_adjoint(seeded((2, 12, 6), 7), 3, "flip").numpy())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
