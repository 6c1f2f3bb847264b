"""The oracle (Python and C restatements) reproduces the committed golden vectors.  CPU only."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import c_port, updates as ou
from oracle.radon import FAN
from util import rel_l2

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz"))


def _par():
    return oracle.RadonGeom(n=32, n_angles=10, det_count=40), oracle.trig_table(-G["par_angles"])


def _fan():
    return (oracle.RadonGeom(n=32, n_angles=12, det_count=32, det_spacing=2.0, geom=FAN, s_dist=64.0, d_dist=64.0,
                             clip_to_circle=True), oracle.trig_table(-G["fan_angles"]))


@pytest.mark.parametrize("impl", [oracle, c_port], ids=["python", "c"])
def test_ct_oracles_reproduce_golden(impl):
    g, trig = _par()
    x, s = torch.from_numpy(G["par_x"]), torch.from_numpy(G["par_s"])
    assert rel_l2(impl.radon_forward(x, trig, g), G["par_fwd"]) < 1e-12
    assert rel_l2(impl.radon_backprojection(s, trig, g), G["par_adj"]) < 1e-12
    assert rel_l2(impl.filter_sinogram(s), G["par_filt"]) < 1e-11
    assert rel_l2(impl.filter_sinogram(s, "hann"), G["par_filt_hann"]) < 1e-11
    gf, trig2 = _fan()
    assert rel_l2(impl.radon_forward(x, trig2, gf), G["fan_fwd"]) < 1e-12
    assert rel_l2(impl.radon_backprojection(torch.from_numpy(G["fan_s"]), trig2, gf), G["fan_adj"]) < 1e-12


def test_mri_and_update_oracles_reproduce_golden():
    spec = oracle.NufftSpec((16, 16))
    om = G["mri_omega"]
    assert rel_l2(oracle.nufft_forward(torch.from_numpy(G["mri_img"]), om, spec), G["mri_fwd"]) < 1e-12
    assert rel_l2(oracle.nufft_adjoint(torch.from_numpy(G["mri_k"]), om, spec), G["mri_adj"]) < 1e-12
    assert rel_l2(oracle.nufft_forward(torch.from_numpy(G["mri_img"]), om, spec, norm="ortho"), G["mri_fwd_ortho"]) < 1e-12
    assert rel_l2(oracle.calc_dcf(om, spec, 5), G["mri_dcf"]) < 1e-12
    sp = torch.from_numpy(G["up_in"])
    assert rel_l2(ou.angular_upsample(sp, 3, "flip"), G["up_flip"]) < 1e-15
    assert rel_l2(ou.angular_upsample(sp, 3, "periodic"), G["up_periodic"]) < 1e-15
