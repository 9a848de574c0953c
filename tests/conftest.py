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


def _have_gpu() -> bool:
    try:
        from cascade_b200 import _lib
        return _lib.lib().cb200_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the gpu-marked tests are skipped (so that `pytest tests` on a CPU box shows real failures
    only); with one they run and cascade_b200 has no CPU fallback to hide behind."""
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (cb200_device_count() <= 0); cascade_b200 has no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
    if n <= 0:
        pytest.skip(f"no CUDA device (cb200_device_count() = {n}); cascade_b200 has no CPU fallback")
    return api.HostBufferDriver(0)
