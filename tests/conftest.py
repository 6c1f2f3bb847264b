import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _cuda_context_guard(request):
    """After every GPU test make sure the CUDA context is still alive.  A faulted context poisons
    every later test in the process, so leave at once: under `pytest -n 1` xdist reports the test
    as crashed and carries on in a fresh worker."""
    yield
    if "gpu" not in request.keywords:
        return
    import torch
    try:
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        sys.stderr.write(f"\nCUDA context lost after {request.node.nodeid}: {e}\n")
        sys.stderr.flush()
        os._exit(17)
