"""The reference's own API names (include/cascade_compat.h) on the GPU: the README program and a
Frobenius / continuation / evaluate program, compiled with gcc and linked to libcascade_b200_compat.so (the reference's names) + libcascade_b200.so (the C ABI)."""
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
                    "-lcascade_b200_compat", "-lcascade_b200", "-lm", "-Wl,-rpath," + os.path.join(ROOT, "cascade_b200")], check=True)
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


def test_logarithmic_frobenius_and_solution_bookkeeping():
    """acb_ode_solve_frobenius with M = 2, indicial_polynomial, _acb_ode_solution_extend/_update/_normalize
    through the reference's names (reference tests/frobenius.c, tests/indicial_polynomial.c,
    tests/solution_extend.c, tests/solution_update.c)."""
    from fractions import Fraction as F
    out = build_and_run("frobenius_log_example")
    rows = [line.split() for line in out.strip().splitlines()]
    gens = {int(r[1]): (float(r[2]), float(r[3]), float(r[4])) for r in rows if r[0] == "gens"}
    a, b, n = F(3, 8), F(5, 16), 12
    g, dg = F(1), F(0)
    dP = sum(F(2, nu) for nu in range(1, n + 1))  # the reference's rescaling, see tests/test_oracle_golden.py
    for k in range(n + 1):
        assert abs(gens[k][0] - float(g)) <= 1e-15 * abs(float(g))
        want = float(dg + dP * g)
        assert abs(gens[k][1] - want) <= 1e-14 * max(1.0, abs(want)) and gens[k][2] < 1e-290
        ld = 1 / (k + a) + 1 / (k + b) - F(2, k + 1)
        fac = (k + a) * (k + b) / F((k + 1) ** 2)
        dg = dg * fac + g * fac * ld
        g = g * fac
    # acb_ode_solution_evaluate, M = 2, rho = 0: gens[0](x) log(x) + 2 gens[1](x) with the printed coefficients
    x = 0.125
    want = sum(gens[k][0] * x ** k for k in range(n + 1)) * math.log(x) + 2 * sum(gens[k][1] * x ** k for k in range(n + 1))
    ev = [float(v) for v in next(r for r in rows if r[0] == "eval2")[1:]]
    assert abs(ev[0] - want) < 1e-13 * abs(want) and ev[1] == 0.0 and ev[2] < 1e-290
    lg = [float(v) for v in next(r for r in rows if r[0] == "log")[1:]]
    assert abs(lg[0] - math.log(0.125)) < 1e-15 and abs(lg[1] - math.pi) < 1e-15 and lg[2] < 1e-290
    pw = [float(v) for v in next(r for r in rows if r[0] == "pow")[1:]]
    assert abs(pw[0]) < 1e-290 and abs(pw[1] - 1 / math.sqrt(8)) < 1e-16 and pw[2] < 1e-290
    el = [float(v) for v in next(r for r in rows if r[0] == "explog")[1:]]
    assert abs(el[0] + 0.125) < 1e-16 and abs(el[1]) < 1e-290
    vals = {r[0] + (r[1] if r[0] in ("indicial", "extend", "update") else ""): r for r in rows if r[0] not in ("gens", "eval2", "log", "explog", "pow")}
    assert vals["indicial_len"][1] == "3"
    assert [float(vals[f"indicial{k}"][2]) for k in range(3)] == [9.0, 6.0, 1.0]
    assert float(vals["indicial_at_2"][1]) == 25.0 and float(vals["indicial_at_2"][2]) == 25.0
    assert [float(vals[f"extend{i}"][2]) for i in range(3)] == [2.75, 5.0, 6.0]
    assert all(vals[f"extend{i}"][3] == "5" for i in range(3)) and vals["extend_consumed"][1] == "0"
    # Leibniz with f = rho at 1/2: [f g, f g' + f' g, f g'' + 2 f' g'] = [1.375, 5.25, 13]
    assert [float(vals[f"update{i}"][2]) for i in range(3)] == [1.375, 5.25, 13.0]
    assert vals["update_consumed"][1] == "0" and vals["normalize_finite"][1] == "0"


def test_monodromy_matrix():
    """find_monodromy_matrix (reference src/fuchs_solver.c:165-206) for 2F1(1/2, 1/4; 3/4): eigenvalues 1 and
    exp(2 pi i / 4) = i.  The reference's continuation adds no truncation bound, so the test compares midpoints
    (the truncation order leaves about 2^-1024) and checks that the balls stayed tight over 256 steps."""
    out = build_and_run("monodromy_example")
    v = {line.split()[0]: [float(x) for x in line.split()[1:]] for line in out.strip().splitlines()}
    assert abs(v["radius"][0] - 1.0) < 1e-8 and v["radius"][0] <= 1.0
    assert 200 <= v["truncation_order"][0] <= 260
    assert abs(v["trace"][0] - 1.0) < 1e-13 and abs(v["trace"][1] - 1.0) < 1e-13 and v["trace"][2] < 1e-200
    assert abs(v["det"][0]) < 1e-13 and abs(v["det"][1] - 1.0) < 1e-13 and v["det"][2] < 1e-200


def test_reference_fuchs_test_program():
    """reference tests/fuchs.c through the compat names: acb_ode_random operators, acb_ode_solve_fuchs and the
    reference's own acceptance test acb_ode_solves, all on the device."""
    out = build_and_run("fuchs_random_example")
    lines = out.strip().splitlines()
    assert lines[-1] == "failures 0"
    assert sum(1 for line in lines if line.startswith("iter") and line.endswith("solves 1")) >= 6


def test_radius_and_reduce_program():
    """reference tests/radius.c and tests/reduce.c through the compat names.  radius_of_convergence is the rigorous
    device version (Graeffe iterations + root bound in ball arithmetic): the ball must CONTAIN the smallest root modulus
    |1 + i| = sqrt(2) of the leading coefficient, from below at its midpoint, with the error bar of 20 iterations."""
    out = build_and_run("radius_example")
    v = {line.split()[0]: line.split()[1:] for line in out.strip().splitlines()}
    r, rad = float(v["radius"][0]), float(v["radius"][1])
    assert r <= math.sqrt(2) <= r + rad * (1 + 1e-9) and rad < 1e-4
    assert v["reduced"] == ["2", "degree", "3", "valuation", "-1"]
    assert abs(float(v["radius_after_reduce"][0]) - r) < 1e-12
