"""Randomised parity: the operators against the CPU oracle on seeded random geometries, trajectories and batch shapes
(tools/fuzz_parity.py holds the generators; `python tools/fuzz_parity.py 150 100 <seed>` is the long form)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [101, 102])
def test_random_ct_geometries_match_the_oracle(seed):
    import fuzz_parity as fz
    rng = np.random.default_rng(seed)
    rows = [fz.run_ct(rng, i) for i in range(40)]
    bad = [r for r in rows if not r["ok"]]
    assert not bad, bad


@pytest.mark.parametrize("seed", [201, 202])
def test_random_nufft_shapes_and_trajectories_match_the_oracle(seed):
    import fuzz_parity as fz
    rng = np.random.default_rng(seed)
    rows = [fz.run_mri(rng, i) for i in range(30)]
    bad = [r for r in rows if not r["ok"]]
    assert not bad, bad
