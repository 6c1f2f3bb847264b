"""CPU oracle for the PD-UNet measurement operators.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  The mounted reference (/root/reference) is the `main` branch
stub of phernst/pd-unet: README.md:1-5 (a title, the paper link, a pointer to
unmounted branches), LICENSE and a stock .gitignore.  It holds no operator
code, no tests and no golden vectors, and the two third-party libraries that
carry the arithmetic -- torch_radon (matteo-ronchetti/torch-radon) and
torchkbnufft (mmuckley/torchkbnufft), pinned versions unknown -- are absent
from the image and cannot be fetched.  This package therefore restates the
*published algorithms* of those two libraries as the author recalls them
(every convention is a parameter, every recollection is tagged [RECALL] in
the docstrings) and pins itself against first-principles known answers
instead (tests/test_oracle_*.py): the analytic Radon transform of a disc,
<Ax,y> = <x,A^H y> in float64, an O(NK) exact non-uniform DFT, FBP(Radon(disc))
~ disc.  Golden vectors generated from this oracle are committed under
tests/golden/ with the script that made them (tests/golden/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package.  Nothing under pd_unet_b200/
imports it; the product path has no CPU fallback.

Arithmetic: float64 / complex128 everywhere, EXCEPT the discrete decisions
(ray clipping and step count for the Radon forward projector; grid offset and
table index for the Kaiser-Bessel interpolator), which are restated in IEEE
float32 op-for-op so that the CUDA kernels and the oracle select the same
samples.  Without that, two correct implementations differ by 1e-4..1e-3 on
the handful of rays / taps whose decision sits on a rounding boundary.
"""
from .radon import (RadonGeom, trig_table, ray_setup_f32, radon_forward,
                    radon_backprojection, filter_taps, filter_matrix,
                    filter_sinogram, fbp)
from .nufft import (NufftSpec, kb_table, scaling_coef, nufft_forward,
                    nufft_adjoint, ndft_forward, ndft_adjoint, interp_forward,
                    interp_adjoint, calc_dcf, radial_trajectory)
from .updates import (dual_update, primal_update, axpby, angular_upsample,
                      angular_upsample_adjoint)
