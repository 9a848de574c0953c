"""The reference's own API names (include/cascade_compat.h) on the GPU: the README program and a
Frobenius / continuation / evaluate program, compiled with gcc and linked to libcascade_b200.so."""
import math
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_and_run(name):
    exe = os.path.join(ROOT, "tests", "c", name)
    subprocess.run(["/usr/bin/gcc", "-O1", "-I", os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "tests", "c", name + ".c"), "-L", os.path.join(ROOT, "cascade_b200"),
                    "-lcascade_b200", "-Wl,-rpath," + os.path.join(ROOT, "cascade_b200")], check=True)
    return subprocess.run([exe], check=True, capture_output=True, text=True).stdout


def test_readme_program():
    """reference README.md:66-77: operator rows (20,0,0), (0,-2,0), (1,0,-1); solution midpoints."""
    out = build_and_run("readme_example")
    assert "Order: 2" in out and "Degree: 2" in out
    sol = out[out.index("["):]
    mids = [float(line.strip("[ ").split("+")[0].strip("( ")) for line in sol.strip().splitlines()]
    assert mids == [0.375, 0.0, -3.75, 0.0, 4.375]


def test_frobenius_continuation_evaluate():
    out = build_and_run("continuation_example")
    vals = {line.split()[0]: [float(v) for v in line.split()[1:]] for line in out.strip().splitlines()}
    # 2F1(1/2, 1/4; 3/2; 1/8), reference value from mpmath at 50 digits
    assert abs(vals["hyp2f1"][0] - 1.0109404757371801) < 1e-15 and vals["hyp2f1"][1] < 1e-30
    assert vals["hyp2f1_eval"][0] == vals["hyp2f1"][0]
    # J0(1.5), J0'(1.5) = -J1(1.5); inputs were doubles, so agreement to ~1e-15
    assert abs(vals["J0(1.5)"][0] - 0.5118276717359181) < 1e-13
    assert abs(vals["J0'(1.5)"][0] - (-0.5579365079100996)) < 1e-13
