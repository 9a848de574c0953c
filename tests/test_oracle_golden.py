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


@pytest.mark.parametrize("L", [4, 32])
def test_general_frobenius_contains_exact_rationals(oracle, L):
    """acb_ode_solve_frobenius with M = 2 on the operators of reference tests/frobenius.c:24-31
    (dyadic coefficients): every ball of gens[0], gens[1] must contain the exact value of the same
    method carried out in rational arithmetic (tests/exact_model.py), and gens[0] shifted by rho
    must pass the reference's acceptance test acb_ode_solves (tests/frobenius.c:37-40)."""
    from tests.cases import _frobenius_test_operator
    from tests.exact_model import frobenius_general_exact
    rng = np.random.default_rng(3)
    for trial in range(4):
        degree = 3 + int(rng.integers(0, 3))
        order = 2 + int(rng.integers(0, degree - 2))
        ode = _frobenius_test_operator(rng, L, order, degree, dyadic=True)
        P = [[C(ode.mid(2 * (i * (degree + 1) + j)), ode.mid(2 * (i * (degree + 1) + j) + 1))
              for j in range(degree + 1)] for i in range(order + 1)]
        s_rho, n = order - 2, 2 + int(rng.integers(0, 9))
        gens = oracle.frobenius_general_batch(order, degree, n, ode, from_complex([s_rho], L), 1, True, True, 2, 2)
        ex = frobenius_general_exact(P, 0, C(s_rho), 2, 2, n)
        for i in range(2):
            for k in range(n + 1):
                s = 2 * (i * (n + 1) + k)
                assert gens.contains(s, ex[i][k].re) and gens.contains(s + 1, ex[i][k].im), (trial, i, k)
        shifted = Balls.concat([Balls.zeros(2 * s_rho, L), gens[0:2 * (n + 1)]])
        assert oracle.ode_solves(order, degree, ode, shifted, n)


def test_general_frobenius_log_solution_of_hypgeom(oracle):
    """c = 1: rho = 0 is a double indicial root of the hypergeometric equation.  gens[0] must be
    2F1(a, b; 1; z).  The reference rescales by f_0(rho + nu) / f_0(rho_0 + nu) at every step
    (src/frobenius_solver.c:170-176, :188), i.e. it differentiates P(rho) g_k(rho) with
    P = prod_nu f_0(rho + nu) / f_0(nu), P(0) = 1, P'(0) = sum_{nu <= n} 2 / nu (f_0(rho) = rho^2 here):
    gens[1][k] = g_k'(0) + P'(0) g_k(0) -- the second solution plus a multiple of the first."""
    L = 32
    a, b = F(3, 8), F(5, 16)
    ode = oracle.example("hypgeom", L, 0, from_complex([a, b, 1], L))
    n = 12
    gens = oracle.frobenius_general_batch(2, 2, n, ode, from_complex([0], L), 1, True, True, 2, 2)
    # g_k(rho) = prod_{j<k} (rho + j + a)(rho + j + b) / (rho + j + 1)^2; gens[0][k] = g_k(0), gens[1][k] = g_k'(0)
    g, dg = F(1), F(0)
    dP = sum(F(2, nu) for nu in range(1, n + 1))
    for k in range(n + 1):
        assert gens.contains(2 * k, g), k
        assert gens.contains(2 * ((n + 1) + k), dg + dP * g), k
        # logarithmic derivative of the next factor at rho = 0
        ld = 1 / (k + a) + 1 / (k + b) - F(2, k + 1)
        fac = (k + a) * (k + b) / F((k + 1) ** 2)
        dg = dg * fac + g * fac * ld
        g = g * fac


def test_gmp_and_plain_backends_agree_general_frobenius(oracle):
    from tests.cases import _frobenius_test_operator
    rng = np.random.default_rng(11)
    ode = _frobenius_test_operator(rng, 32, 2, 4)
    rho = from_complex([0], 32)
    a = oracle.frobenius_general_batch(2, 4, 9, ode, rho, 1, True, True, 2, 3)
    b = oracle.frobenius_general_batch(2, 4, 9, ode, rho, 1, True, True, 2, 3, plain=True)
    assert a.same_as(b)


@pytest.mark.parametrize("L", [4, 32])
def test_exp_log_enclose_mpmath(oracle, L):
    """The oracle's acb_exp / acb_log (oracle/cbo_trans.c) are rigorous enclosures: the true value computed by
    mpmath at a much higher precision lies inside every ball, and the balls stay within ~2^-(prec-40) relative."""
    import mpmath as mp
    from tests.cases import trans_inputs
    mp.mp.prec = 32 * L + 300
    rng = np.random.default_rng(9)
    n = 30
    x = trans_inputs(rng, L, n)
    x.meta[:, 2] = 0  # exact inputs, so that mpmath sees the same numbers
    x.meta[:, 3] = 0
    to_mp = lambda q: mp.mpf(q.numerator) / q.denominator
    for name, fn in (("exp", mp.exp), ("log", mp.log)):
        z = oracle.ew_trans(name, x)
        for i in range(n):
            if (int(z.meta[2 * i, 1]) | int(z.meta[2 * i + 1, 1])) & FLAG_NAN:
                assert name == "log" and x.mid(2 * i) == 0 and x.mid(2 * i + 1) == 0 or name == "exp"
                continue
            tv = fn(mp.mpc(to_mp(x.mid(2 * i)), to_mp(x.mid(2 * i + 1))))
            for part, t in ((0, tv.real), (1, tv.imag)):
                s = 2 * i + part
                assert abs(to_mp(z.mid(s)) - t) <= to_mp(z.rad(s)), (name, i, part)
            mag = abs(tv)
            if mag != 0:
                assert to_mp(z.rad(2 * i)) <= mag * mp.mpf(2) ** (-(32 * L - 40)), (name, i)


@pytest.mark.parametrize("L", [4, 32])
def test_log_across_the_branch_cut(oracle, L):
    """A ball that straddles the negative real axis contains points whose principal logarithms have imaginary parts
    near +pi AND near -pi; acb_log must enclose both (Arb returns Im in [-pi, pi] there).  Points of the ball on both
    sides of the cut, on the cut, and at the corners are checked with mpmath."""
    import mpmath as mp
    from fractions import Fraction as F
    mp.mp.prec = 32 * L + 200
    to_mp = lambda q: mp.mpf(q.numerator) / q.denominator
    cases_ = [((F(-3, 2), F(1, 2**30)), (-20, -25)),    # im = 2^-30 +- 2^-26: straddles (log wants radius/|a| < 2^-8)
              ((F(-5, 7), F(-1, 2**50)), (-60, -45)),   # re has a radius too
              ((F(-1, 2**20), 0), (-200, -100))]        # im midpoint exactly zero, radius 2^-101
    for (re, im), (rre, rim) in cases_:
        x = from_complex([(re, im)], L)
        x.meta[0, 2], x.meta[0, 3] = 1 << 29, rre
        x.meta[1, 2], x.meta[1, 3] = 1 << 29, rim
        z = oracle.ew_trans("log", x)
        assert not (int(z.meta[0, 1]) | int(z.meta[1, 1])) & FLAG_NAN
        for fr in (-1, F(-1, 3), 0, F(1, 2), 1):
            for fi in (-1, F(-9, 10), F(1, 10**9), F(9, 10), 1):
                pr = x.mid(0) + fr * x.rad(0)
                pi_ = x.mid(1) + fi * x.rad(1)
                tv = mp.log(mp.mpc(to_mp(pr), to_mp(pi_)))
                assert abs(to_mp(z.mid(0)) - tv.real) <= to_mp(z.rad(0)), (re, im, fr, fi)
                assert abs(to_mp(z.mid(1)) - tv.imag) <= to_mp(z.rad(1)), (re, im, fr, fi)
        # the real part stays sharp: only the argument is lost
        assert z.rad(0) < F(1, 2**20)
    # a ball next to the cut that does not cross it keeps a tight argument
    x = from_complex([(F(-3, 2), F(1, 2**30))], L)
    x.meta[1, 2], x.meta[1, 3] = 1 << 29, -40
    z = oracle.ew_trans("log", x)
    assert z.rad(1) < F(1, 2**8) and z.mid(1) > 3


def test_solution_evaluate_encloses_mpmath(oracle):
    """acb_ode_solution_evaluate (reference src/acb_ode_solution.c:24-58; property of reference
    tests/solution_eval.c): a^rho * sum_i w_i gens[i](a) log(a)^(M-1-i) with the weights the reference codes,
    w_0 = 1, w_i = w_{i-1} (M - i + 1) / i, evaluated independently with mpmath on exact (dyadic) inputs."""
    import mpmath as mp
    from cascade_b200 import api
    L, M, cap = 32, 3, 6
    mp.mp.prec = 32 * L + 300
    rng = np.random.default_rng(21)
    dy = lambda: (F(int(rng.integers(-64, 64)), 32), F(int(rng.integers(-64, 64)), 32))
    gvals = [[dy() for _ in range(cap)] for _ in range(M)]
    gens = from_complex([v for row in gvals for v in row], L)
    pts_v = [(F(3, 4), F(1, 8)), (F(-5, 16), F(7, 8)), (F(1, 2), 0)]
    pts = from_complex(pts_v, L)
    to_mp = lambda q: mp.mpf(q.numerator) / q.denominator
    for rho_v in ((F(0), F(0)), (F(3), F(0)), (F(1, 4), F(-3, 8))):
        rho = from_complex([rho_v], L)
        mode, e = api.pow_mode_of(rho)
        out = oracle.solution_evaluate(gens, M, rho, mode, e, pts)
        for p, (ar, ai) in enumerate(pts_v):
            a = mp.mpc(to_mp(ar), to_mp(ai))
            la = mp.log(a)
            tot, w = mp.mpc(0), 1
            for i in range(M):
                if i:
                    w = w * (M - i + 1) // i
                gi = sum(mp.mpc(to_mp(c[0]), to_mp(c[1])) * a ** k for k, c in enumerate(gvals[i]))
                tot += w * gi * la ** (M - 1 - i)
            tot *= a ** mp.mpc(to_mp(rho_v[0]), to_mp(rho_v[1]))
            for part, t in ((0, tot.real), (1, tot.imag)):
                s = 2 * p + part
                assert abs(to_mp(out.mid(s)) - t) <= to_mp(out.rad(s)), (rho_v, p, part)
                assert to_mp(out.rad(s)) <= abs(tot) * mp.mpf(2) ** (-(32 * L - 60))
