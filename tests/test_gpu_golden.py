"""The CUDA path against the committed golden vectors (no oracle code runs in these tests)."""
import os

import numpy as np
import pytest
import torch

import pd_unet_b200 as pdu
from util import TOL, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz"))


def _t(name):
    return torch.from_numpy(G[name]).to(DEV)


def test_ct_parallel_and_fan():
    op = pdu.Radon(32, G["par_angles"], det_count=40)
    assert rel_l2(op.forward(_t("par_x")), G["par_fwd"]) <= TOL
    assert rel_l2(op.backprojection(_t("par_s")), G["par_adj"]) <= TOL
    assert rel_l2(op.filter_sinogram(_t("par_s")), G["par_filt"]) <= TOL
    assert rel_l2(op.filter_sinogram(_t("par_s"), "hann"), G["par_filt_hann"]) <= TOL
    fan = pdu.RadonFanbeam(32, G["fan_angles"], 64.0, clip_to_circle=True)
    assert rel_l2(fan.forward(_t("par_x")), G["fan_fwd"]) <= TOL
    assert rel_l2(fan.backprojection(_t("fan_s")), G["fan_adj"]) <= TOL


def test_mri_and_updates():
    om = _t("mri_omega")
    assert rel_l2(pdu.KbNufft((16, 16))(_t("mri_img"), om), G["mri_fwd"]) <= TOL
    assert rel_l2(pdu.KbNufftAdjoint((16, 16))(_t("mri_k"), om), G["mri_adj"]) <= TOL
    assert rel_l2(pdu.KbNufft((16, 16))(_t("mri_img"), om, norm="ortho"), G["mri_fwd_ortho"]) <= TOL
    w = pdu.calc_density_compensation_function(om, (16, 16), num_iterations=5)
    assert rel_l2(w.real.reshape(-1), G["mri_dcf"]) <= 5e-5
    assert rel_l2(pdu.updates.angular_upsample(_t("up_in"), 3, "flip"), G["up_flip"]) < 1e-7
    assert rel_l2(pdu.updates.angular_upsample(_t("up_in"), 3, "periodic"), G["up_periodic"]) < 1e-7
