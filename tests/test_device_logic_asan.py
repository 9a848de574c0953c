"""The device thread bodies (compiled for the host, tests/host_sim) under AddressSanitizer: every workspace and
scratch array of the simulation has exactly the size the device code is given, so an out-of-bounds access of a
kernel body shows up here as a heap-buffer-overflow (compute-sanitizer is not available on the GPU pool)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNNER = """
import sys
sys.path.insert(0, {root!r})
from tests.host_sim import driver
driver._SO = {so!r}
driver.build = lambda force=False: None
import pytest
sys.exit(pytest.main(['-x', '-q', '-p', 'no:cacheprovider', '-m', 'not gpu',
                      {root!r} + '/tests/test_device_logic_sim.py', {root!r} + '/tests/test_golden_vectors.py']))
"""


def test_device_logic_under_address_sanitizer(tmp_path):
    asan = subprocess.run(["/usr/bin/gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan is not installed")
    so = str(tmp_path / "libcbsim_asan.so")
    subprocess.run(["/usr/bin/g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-ffp-contract=off",
                    "-fsanitize=address", "-fno-omit-frame-pointer", "-o", so,
                    os.path.join(ROOT, "tests", "host_sim", "host_sim.cpp")], check=True)
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1")
    r = subprocess.run([sys.executable, "-c", RUNNER.format(root=ROOT, so=so)], env=env, capture_output=True, text=True,
                       timeout=1500)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert "passed" in r.stdout
