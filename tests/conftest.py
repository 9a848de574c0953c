import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # a fresh checkout has no built artefacts: build them once (nvcc cross-compiles without a GPU)
    lib = os.path.join(ROOT, "cascade_b200", "libcascade_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def oracle():
    from oracle import cbo
    cbo.lib()
    return cbo


@pytest.fixture(scope="session")
def sim_driver():
    """The device thread bodies compiled for the host (logic check without a GPU)."""
    from tests.host_sim.driver import SimDriver, build
    build()
    return SimDriver()


@pytest.fixture(scope="session")
def gpu_driver():
    """The product path: libcascade_b200.so through its C ABI, host buffers in and out."""
    from cascade_b200 import _lib, api
    n = _lib.lib().cb200_device_count()
    assert n > 0, f"no CUDA device (cb200_device_count() = {n}); cascade_b200 has no CPU fallback"
    return api.HostBufferDriver(0)
