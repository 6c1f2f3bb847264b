"""The C-ABI library builds, loads without a GPU, and exports every symbol include/pdu.h declares.
No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__
    __graft_entry__.build()
    from pd_unet_b200 import _lib
    return _lib


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pdu.h")).read()
    return sorted(set(re.findall(r"PDU_API\s+[\w\s\*]+?\b(pdu_\w+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(built):
    names = _declared_symbols()
    assert len(names) >= 25
    handle = ctypes.CDLL(built.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/pdu.h but not exported"
        assert n in built.SIGNATURES, f"{n} has no ctypes signature in pd_unet_b200/_lib.py"
    assert sorted(built.SIGNATURES) == names


def test_library_reports_errors_instead_of_crashing(built):
    L = built.lib()
    assert L.pdu_version() >= 100
    assert L.pdu_set_option(b"no_such_option", 1) == -1
    assert b"unknown key" in L.pdu_last_error()
    v = ctypes.c_int(7)
    assert L.pdu_get_option(b"radon_fwd_variant", ctypes.byref(v)) == 0 and v.value == -1
    # argument validation happens before any CUDA call, so it is testable without a device
    g = built.RadonGeomC(0, 0, 4, 4, 1.0, 0.0, 0.0, 0)
    assert L.pdu_radon_fwd_f32(None, None, None, 1, ctypes.byref(g), None, 0, None) == -1
    assert L.pdu_radon_adj_f32(None, None, None, 1, ctypes.byref(g), None, 0, None) == -1
    assert L.pdu_filter_sinogram_f32(None, None, None, None, 0, 4, 4, None) == -1
    assert L.pdu_concat_f32(None, None, None, None, 1, 1, 1, 0, 2, 4, 1.0, 0, None) == -1
    assert L.pdu_nufft_plan_create(None, 8, 8, 16, 16, 6, 1024, 4, 4, None, None, None, None) == -1
    assert L.pdu_nufft_plan_destroy(None) == 0
    assert L.pdu_radon_workspace_bytes(None, 1) == 0


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pd_unet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, re.M), f"{f} imports the oracle"


def test_operators_refuse_cpu_tensors(built):
    import numpy as np
    import torch
    import pd_unet_b200 as pdu
    op = pdu.Radon(16, np.linspace(0, np.pi, 4, endpoint=False))
    with pytest.raises(pdu.PduError):
        op.forward(torch.zeros(1, 16, 16))
    with pytest.raises(pdu.PduError):
        pdu.KbNufft((16, 16))(torch.zeros(1, 1, 16, 16, dtype=torch.complex64), torch.zeros(2, 8))
    with pytest.raises(pdu.PduError):
        pdu.updates.axpby(1.0, torch.zeros(4), 1.0, torch.zeros(4))
