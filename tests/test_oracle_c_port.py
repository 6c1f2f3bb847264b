"""The C restatement (oracle/radon_c.c) equals the numpy/torch oracle it restates.  CPU only."""
import numpy as np
import pytest
import torch

import oracle
from oracle import c_port
from oracle.radon import FAN
from util import rel_l2, seeded

GEOMS = {
    "par": oracle.RadonGeom(n=48, n_angles=19, det_count=60),
    "par_circle_spacing": oracle.RadonGeom(n=40, n_angles=13, det_count=33, det_spacing=1.5, clip_to_circle=True),
    "fan": oracle.RadonGeom(n=48, n_angles=17, det_count=48, det_spacing=2.0, geom=FAN, s_dist=96.0, d_dist=96.0),
    "fan_short_circle": oracle.RadonGeom(n=44, n_angles=11, det_count=70, det_spacing=1.3, geom=FAN, s_dist=52.8,
                                         d_dist=35.2, clip_to_circle=True),
}


def _trig(g):
    span = np.pi if g.geom == 0 else 2 * np.pi
    return oracle.trig_table(-np.linspace(0, span, g.n_angles, endpoint=False))


@pytest.mark.parametrize("name", GEOMS)
def test_ray_setup_is_bit_identical(name):
    g = GEOMS[name]
    a, b = oracle.ray_setup_f32(g, _trig(g)), c_port.ray_setup(g, _trig(g))
    for k in a:
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("name", GEOMS)
def test_operators_equal_the_python_oracle(name):
    g = GEOMS[name]
    trig = _trig(g)
    x = seeded((2, g.n, g.n), 1).double()
    s = seeded((2, g.n_angles, g.det_count), 2).double()
    assert rel_l2(c_port.radon_forward(x, trig, g), oracle.radon_forward(x, trig, g)) < 1e-13
    assert rel_l2(c_port.radon_backprojection(s, trig, g), oracle.radon_backprojection(s, trig, g)) < 1e-13
    assert rel_l2(c_port.filter_sinogram(s), oracle.filter_sinogram(s)) < 1e-12
    assert rel_l2(c_port.filter_sinogram(s, "hann"), oracle.filter_sinogram(s, "hann")) < 1e-12
