"""Pins the CPU oracle (oracle/): against the reference's literal known answers, against exact
rational evaluation of the same recurrences, against an independent exact model of the rounding
semantics, and against its own plain-C limb backend.  No GPU."""
from fractions import Fraction as F

import numpy as np
import pytest

from cascade_b200.format import FLAG_NAN, Balls, from_complex, random_balls
from tests.exact_model import C, fuchs_exact, trunc_to_prec


def test_readme_known_answer(oracle):
    """reference README.md:43-47,66-77."""
    L = 32
    ode = oracle.example("legendre", L, 4)
    rows = [[ode.mid(2 * (3 * i + j)) for j in range(3)] for i in range(3)]
    assert rows == [[20, 0, 0], [0, -2, 0], [1, 0, -1]]
    rc, res = oracle.fuchs_batch(2, 2, 5, ode, from_complex([F(3, 8), 0], L), 1, True)
    assert rc == 0
    assert [res.mid(2 * k) for k in range(6)] == [F(3, 8), 0, F(-15, 4), 0, F(35, 8), 0]
    assert all(res.rad(2 * k) == 0 for k in range(6))  # every operation here is exact


def test_constructor_known_answers(oracle):
    """reference tests/legendre.c:13-20, tests/bessel.c:13-23, tests/hypgeom.c:20-35."""
    L = 32
    leg = oracle.example("legendre", L, 5)
    assert leg.contains(2 * 0, 30) and leg.contains(2 * 3, 0) and leg.contains(2 * 8, -1)
    bes = oracle.example("bessel", L, 0, from_complex([F(1, 2)], L))
    for e in (8, 4, 2):
        assert bes.contains(2 * e, 1)
    assert bes.contains(0, F(-1, 4))
    hyp = oracle.example("hypgeom", L, 0, from_complex([1, 2, 3], L))
    want = {0: -2, 4: -4, 3: 3, 6: 0, 7: 1, 8: -1}
    for e, v in want.items():
        assert hyp.contains(2 * e, v), (e, v)


def test_valuation_guard(oracle):
    """Bessel at the origin has valuation 0: the reference loop is undefined there."""
    L = 32
    bes = oracle.example("bessel", L, 0, from_complex([F(1, 2)], L))
    rc, _ = oracle.fuchs_batch(2, 2, 5, bes, Balls.zeros(0, L), 1, True)
    assert rc == -1


@pytest.mark.parametrize("L", [4, 32])
def test_fuchs_contains_exact_rationals(oracle, L):
    """Dyadic complex operator and initial values: every oracle ball must contain the exact value
    of the recurrence of SURVEY.md App. A (independent Fraction implementation)."""
    rng = np.random.default_rng(5)
    order, degree, n = 2, 3, 14
    dy = lambda: F(int(rng.integers(-40, 40)), 8)
    P = [[C(dy(), dy()) for _ in range(degree + 1)] for _ in range(order + 1)]
    P[order][0] = C(F(5, 4), F(-3, 8))  # ordinary point: valuation -order
    init = [C(dy(), dy()) for _ in range(order)]
    ode = from_complex([(P[i][j].re, P[i][j].im) for i in range(order + 1) for j in range(degree + 1)], L)
    rc, res = oracle.fuchs_batch(order, degree, n, ode, from_complex([(c.re, c.im) for c in init], L), 1, True)
    assert rc == 0
    exact = fuchs_exact(P, init, n, -order)
    for k in range(n + 1):
        assert res.contains(2 * k, exact[k].re) and res.contains(2 * k + 1, exact[k].im), k
    # and the balls are tight: relative radius of the last coefficient ~ n ulps
    if L == 32:
        assert res.rad(2 * n) / abs(res.mid(2 * n)) < F(1, 2 ** 1000)


def test_hypgeom_closed_forms(oracle):
    """(a)_n (b)_n / ((c)_n n!) via acb_ode_solve_fuchs (valuation -1) and via Frobenius at
    rho = 0; and 2F1(a-c+1, b-c+1; 2-c) at rho = 1-c."""
    L = 32
    a, b, c = F(3, 8), F(5, 16), F(7, 4)
    ode = oracle.example("hypgeom", L, 0, from_complex([a, b, c], L))
    n = 25

    def coeffs(a, b, c):
        out = [F(1)]
        for k in range(n):
            out.append(out[-1] * (a + k) * (b + k) / ((c + k) * (k + 1)))
        return out

    rc, res = oracle.fuchs_batch(2, 2, n, ode, from_complex([1], L), 1, True)
    fro0 = oracle.frobenius1_batch(2, 2, n, ode, from_complex([0], L), 1, True)
    fro1 = oracle.frobenius1_batch(2, 2, n, ode, from_complex([1 - c], L), 1, True)
    e0, e1 = coeffs(a, b, c), coeffs(a - c + 1, b - c + 1, 2 - c)
    for k in range(n + 1):
        assert res.contains(2 * k, e0[k]) and fro0.contains(2 * k, e0[k]) and fro1.contains(2 * k, e1[k]), k


@pytest.mark.parametrize("L", [4, 32, 64])
def test_rounding_semantics_against_exact_model(oracle, L):
    """Midpoints of add/sub/mul/div are the exact result truncated toward zero to 32 L bits, the
    radius is zero exactly when the operation was exact, and the ball contains the exact result."""
    rng = np.random.default_rng(L)
    n = 60
    x = random_balls(rng, 2 * n, L, -40, 40, "zero", short_frac=0.5)
    y = random_balls(rng, 2 * n, L, -40, 40, "zero", short_frac=0.5)
    y.meta[: n, 0] = x.meta[: n, 0]  # equal exponents: cancellation
    prec = 32 * L
    for op, fn in ((0, lambda p, q: p + q), (1, lambda p, q: p - q), (2, lambda p, q: p * q), (3, lambda p, q: p / q)):
        z = oracle.ew_binary(op, x, y)
        for i in range(n):
            xv, yv = C(x.mid(2 * i), x.mid(2 * i + 1)), C(y.mid(2 * i), y.mid(2 * i + 1))
            ev = fn(xv, yv)
            if op in (0, 1, 2):  # one rounding per component
                for part, exact in ((0, ev.re), (1, ev.im)):
                    mid, inexact = trunc_to_prec(exact, prec)
                    assert z.mid(2 * i + part) == mid
                    assert (z.rad(2 * i + part) != 0) == inexact
            assert z.contains(2 * i, ev.re) and z.contains(2 * i + 1, ev.im)


def test_division_by_ball_containing_zero_is_indeterminate(oracle):
    L = 32
    x = from_complex([1], L)
    y = from_complex([0], L)
    y.set_exact(0, F(1, 1024), rad=F(1, 512))
    z = oracle.ew_binary(3, x, y)
    assert int(z.meta[0, 1]) & FLAG_NAN


def test_gmp_and_plain_backends_agree(oracle):
    assert oracle.backend() == "gmp-mpn" and oracle.backend(True) == "plain-c"
    rng = np.random.default_rng(1)
    for L in (4, 32, 128):
        x = random_balls(rng, 120, L, -3, 3, "ulp", short_frac=0.2)
        y = random_balls(rng, 120, L, -3, 3, "ulp", short_frac=0.2)
        for op in range(5):
            assert oracle.ew_binary(op, x, y).same_as(oracle.ew_binary(op, x, y, plain=True))
    from tests.cases import bessel_shifted
    ode = bessel_shifted(oracle, rng, 32)
    init = random_balls(rng, 8, 32, -1, 0, "zero")
    a = oracle.fuchs_batch(2, 2, 30, ode, init, 2, True)[1]
    b = oracle.fuchs_batch(2, 2, 30, ode, init, 2, True, plain=True)[1]
    assert a.same_as(b)
