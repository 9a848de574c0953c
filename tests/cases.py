"""Shared parity cases: each takes a driver (host-sim or GPU) and compares with the oracle.

The bar is BIT-EXACT agreement of every mantissa limb, exponent, flag and radius word.
"""
from __future__ import annotations

from fractions import Fraction as F

import numpy as np

from cascade_b200 import _lib, api
from cascade_b200.format import FLAG_NAN, FLAG_ZERO, Balls, from_complex, random_balls


def assert_same(a: Balls, b: Balls, what: str):
    bm = np.nonzero((a.mant != b.mant).any(axis=1))[0]
    bt = np.nonzero((a.meta != b.meta).any(axis=1))[0]
    assert len(bm) == 0 and len(bt) == 0, (
        f"{what}: {len(bm)} mantissa and {len(bt)} meta slots differ "
        f"(first: slot {(list(bm) + list(bt))[0]}: oracle meta {a.meta[(list(bm) + list(bt))[0]]} "
        f"vs {b.meta[(list(bm) + list(bt))[0]]})")


def adversarial_pair(rng, L: int, n: int):
    """2n slots each: limb-boundary exponent gaps, exact and near cancellation, zeros, all-ones
    mantissas, short (integer-like) mantissas, radii from 1 ulp to huge."""
    x = random_balls(rng, 2 * n, L, -2, 2, "zero", short_frac=0.3)
    y = random_balls(rng, 2 * n, L, -2, 2, "zero", short_frac=0.3)
    gaps = rng.choice([0, 1, 2, 30, 31, 32, 33, 34, 63, 64, 65, 32 * L - 1, 32 * L, 32 * L + 1, 32 * L + 31,
                       32 * L + 32, 32 * L + 33, 64 * L, 64 * L + 40, 64 * L + 64, 64 * L + 65, 64 * L + 100,
                       5000], size=2 * n)
    y.meta[:, 0] = x.meta[:, 0] - gaps * rng.choice([1, -1], size=2 * n)
    for s in range(0, 2 * n, 5):
        y.mant[s] = x.mant[s]
        y.meta[s, 0] = x.meta[s, 0]
        mode = rng.integers(0, 4)
        if mode == 1:
            y.mant[s, 0] ^= np.uint32(1)
        if mode == 2:
            y.mant[s, rng.integers(0, L)] ^= np.uint32(1 << int(rng.integers(0, 31)))
        if mode == 3:
            y.mant[s, : L // 2] = 0
    for s in range(3, 2 * n, 17):
        x.mant[s] = 0
        x.meta[s] = (0, FLAG_ZERO, 0, 0)
    for s in range(7, 2 * n, 23):
        y.mant[s] = 0
        y.meta[s] = (0, FLAG_ZERO, 0, 0)
    for s in range(11, 2 * n, 29):
        x.mant[s] = 0xFFFFFFFF
    r = rng.random(2 * n) < 0.5
    for b in (x, y):
        idx = np.nonzero(r & ((b.meta[:, 1] & FLAG_ZERO) == 0))[0]
        b.meta[idx, 2] = rng.integers(1 << 29, 1 << 30, size=len(idx))
        b.meta[idx, 3] = b.meta[idx, 0] - rng.integers(1, 40 * L, size=len(idx))
    return x, y


def case_elementwise(oracle, drv, L: int, n: int, seed: int):
    rng = np.random.default_rng(seed)
    x, y = adversarial_pair(rng, L, n)
    for op, name in ((0, "add"), (1, "sub"), (2, "mul"), (3, "div"), (4, "inv")):
        zo = oracle.ew_binary(op, x, y)
        zs = api.ew(op, x, None if op == 4 else y, driver=drv)
        assert_same(zo, zs, f"{name} L={L}")
    k = rng.integers(-2**62, 2**62, size=n)
    k[:10] = [0, 1, -1, 2, 3, 2**62, -2**62, 5, 7, 10]
    k[10:40] = rng.integers(-100, 100, size=min(30, n - 10))
    for op, name in ((16, "mul_si"), (17, "add_si")):
        assert_same(oracle.ew_scalar_si(op, x, k), api.ew(op, x, k=k, driver=drv), f"{name} L={L}")


def case_exponent_overflow(oracle, drv, L: int):
    """Midpoint exponents beyond +-2^28 make the ball indeterminate (device exponents are 32-bit);
    the oracle applies the same rule, and an indeterminate ball has an all-zero mantissa."""
    rng = np.random.default_rng(L)
    n = 40
    x = random_balls(rng, 2 * n, L, -2, 2, "ulp")
    y = random_balls(rng, 2 * n, L, -2, 2, "ulp")
    big = (1 << 27) + 5
    x.meta[:, 0] += np.where(np.arange(2 * n) % 3 == 0, big, 0).astype(np.int32)
    y.meta[:, 0] += np.where(np.arange(2 * n) % 3 == 0, big, 0).astype(np.int32)
    x.meta[:, 3] = x.meta[:, 0] - 32 * L
    y.meta[:, 3] = y.meta[:, 0] - 32 * L
    ys = y.copy()
    ys.meta[:, 0] -= np.where(np.arange(2 * n) % 3 == 0, 2 * big, 0).astype(np.int32)
    ys.meta[:, 3] = ys.meta[:, 0] - 32 * L
    for op, name, yy in ((2, "mul", y), (3, "div", ys), (0, "add", y)):
        zo = oracle.ew_binary(op, x, yy)
        zs = api.ew(op, x, yy, driver=drv)
        assert_same(zo, zs, f"overflow {name} L={L}")
        if op == 2:
            assert (zo.meta[:, 1] & FLAG_NAN).any()


def case_readme_kat(oracle, drv):
    """reference README.md:43-47,72-77: legendre(4), c0 = 0.375, 5 terms at 1024 bits."""
    L = 32
    ode = oracle.example("legendre", L, 4)
    init = from_complex([F(3, 8), 0], L)
    res = api.acb_ode_solve_fuchs_batch(ode, init, 2, 2, 5, 1, True, driver=drv)
    assert [res.mid(2 * i) for i in range(6)] == [F(3, 8), 0, F(-15, 4), 0, F(35, 8), 0]
    assert all(res.mid(2 * i + 1) == 0 for i in range(6))
    rc, ro = oracle.fuchs_batch(2, 2, 5, ode, init, 1, True)
    assert_same(ro, res, "README KAT")


def case_fuchs_random(oracle, drv, L: int, seed: int, trials: int = 4, batch: int = 3):
    """reference tests/fuchs.c:18-41: random operator (degree 3..9, order 2..degree-1), random
    initial values, n = order + randint(32); parity with the oracle AND the reference's own
    acceptance test acb_ode_solves (residual contains zero)."""
    rng = np.random.default_rng(seed)
    for _ in range(trials):
        degree = 3 + int(rng.integers(0, 7))
        order = 2 + int(rng.integers(0, degree - 2))
        ne = (order + 1) * (degree + 1)
        ode = random_balls(rng, batch * ne * 2, L, -4, 4, "ulp", short_frac=0.2)
        n = order + int(rng.integers(0, 32))
        init = random_balls(rng, batch * order * 2, L, -3, 3, "ulp")
        rc, ro = oracle.fuchs_batch(order, degree, n, ode, init, batch, False)
        assert rc == 0
        rs = api.acb_ode_solve_fuchs_batch(ode, init, order, degree, n, batch, False, driver=drv)
        assert_same(ro, rs, f"fuchs L={L} order={order} degree={degree} n={n}")
        for i in range(batch):
            assert oracle.ode_solves(order, degree, ode[i * ne * 2:(i + 1) * ne * 2],
                                     rs[i * (n + 1) * 2:(i + 1) * (n + 1) * 2], n - order)


def case_fuchs_random_shared(oracle, drv, L: int, seed: int, trials: int = 4, batch: int = 5):
    """As case_fuchs_random, but ONE random operator shared by the batch (the table + recurrence
    kernels); includes operators whose leading coefficient is real (division path) and valuations
    between -order and -1."""
    rng = np.random.default_rng(seed)
    for trial in range(trials):
        degree = 3 + int(rng.integers(0, 7))
        order = 2 + int(rng.integers(0, degree - 2))
        ne = (order + 1) * (degree + 1)
        ode = random_balls(rng, ne * 2, L, -4, 4, "ulp", short_frac=0.2)
        m = ode.mant.reshape(order + 1, degree + 1, 2, L)
        t = ode.meta.reshape(order + 1, degree + 1, 2, 4)
        if trial % 2 == 1:  # real leading coefficient -> real divisor
            m[order, 0, 1] = 0
            t[order, 0, 1] = (0, FLAG_ZERO, 0, 0)
        if trial % 3 == 2:  # kill the low columns of the top row: valuation > -order
            m[order, 0] = 0
            t[order, 0] = (0, FLAG_ZERO, 0, 0)
        v = api.acb_ode_valuation(ode, order, degree)
        assert v < 0
        n = -v + int(rng.integers(0, 32))
        init = random_balls(rng, batch * (-v) * 2, L, -3, 3, "ulp")
        rc, ro = oracle.fuchs_batch(order, degree, n, ode, init, batch, True)
        assert rc == 0
        rs = api.acb_ode_solve_fuchs_batch(ode, init, order, degree, n, batch, True, driver=drv)
        assert_same(ro, rs, f"fuchs shared L={L} order={order} degree={degree} v={v} n={n}")


def case_fuchs_shared_adversarial(oracle, drv, L: int, seed: int, trials: int = 6, batch: int = 40):
    """Shared operator with hostile data for the recurrence kernel's fast paths: real/imaginary
    parts hundreds of bits apart, exact zeros, all-ones and short (integer-like) mantissas, huge
    and tiny radii -- in the operator and in the initial values."""
    rng = np.random.default_rng(seed)
    for trial in range(trials):
        order, degree = 2, 2 + trial % 2
        ne = (order + 1) * (degree + 1)
        ode, _ = adversarial_pair(rng, L, ne)
        ode.meta[:, 0] = rng.integers(-3, 4, size=2 * ne) + rng.choice([0, 0, 0, 40, -70, 200], size=2 * ne)
        lead = random_balls(rng, 2, L, -1, 1, "ulp")  # leading coefficient must not contain zero
        if trial % 3 == 0:
            lead.mant[1] = 0
            lead.meta[1] = (0, FLAG_ZERO, 0, 0)
        ode.mant[2 * (order * (degree + 1)):2 * (order * (degree + 1)) + 2] = lead.mant
        ode.meta[2 * (order * (degree + 1)):2 * (order * (degree + 1)) + 2] = lead.meta
        fix = (ode.meta[:, 1] & FLAG_ZERO) != 0  # zero midpoints keep exponent 0
        ode.meta[fix, 0] = 0
        v = api.acb_ode_valuation(ode, order, degree)
        assert v == -order
        init, _ = adversarial_pair(rng, L, batch * order)
        init.meta[:, 0] += rng.choice([0, 0, 0, 33, -64, 150, -300], size=init.n_slots).astype(np.int32)
        init.meta[(init.meta[:, 1] & FLAG_ZERO) != 0, 0] = 0
        # non-canonical zeros (flag set, stale limbs) must behave as zeros, as they do in the oracle
        init.mant[5] = 0xFFFFFFFF
        init.meta[5] = (0, FLAG_ZERO, 0, 0)
        ode.mant[3] = 0x12345678
        ode.meta[3] = (0, FLAG_ZERO, 0, 0)
        n = order + 12
        rc, ro = oracle.fuchs_batch(order, degree, n, ode, init, batch, True)
        assert rc == 0
        rs = api.acb_ode_solve_fuchs_batch(ode, init, order, degree, n, batch, True, driver=drv)
        assert_same(ro, rs, f"fuchs shared adversarial L={L} trial={trial}")


def case_fuchs_integer_operators(oracle, drv, L: int, seed: int):
    """BASELINE configs[3]'s path: Legendre sweep (acb_ode_legendre(n), reference src/examples.c:3-11)
    and random integer operators with complex data, through the integer-operator kernel; must give
    the same balls as the oracle's general loop (and as the general kernel)."""
    rng = np.random.default_rng(seed)
    ns = list(range(0, 40)) + [10**3, 2**20 - 1, 12345, 999999]
    batch = len(ns)
    ode = api.acb_ode_legendre_batch(ns, L)
    assert api.integer_operator(ode, 2, 2, 60) is not None
    init = from_complex([1, 1] * batch, L)  # c0 = c1 = 1
    n = 60
    rc, ro = oracle.fuchs_batch(2, 2, n, ode, init, batch, False)
    rs = api.acb_ode_solve_fuchs_batch(ode, init, 2, 2, n, batch, False, driver=drv)
    assert_same(ro, rs, f"legendre sweep L={L}")
    rg = api.acb_ode_solve_fuchs_batch(ode, init, 2, 2, n, batch, False, driver=drv, use_intop=False)
    assert_same(ro, rg, f"legendre sweep, general kernel L={L}")
    # n = 4, c0 = c1 = 1: c2 = -(20 c0)/2 = -10, c3 = -(20 - 2) c1 / 6 = -3 exactly
    k4 = ns.index(4)
    assert rs.mid(2 * (k4 * (n + 1) + 2)) == -10 and rs.mid(2 * (k4 * (n + 1) + 3)) == -3
    # random integer operators, complex balls with radii as data
    for trial in range(4):
        degree = 2 + int(rng.integers(0, 3))
        order = 2 + int(rng.integers(0, degree - 1))
        ne = (order + 1) * (degree + 1)
        batch = 9
        vals = rng.integers(-50, 50, size=(batch, ne))
        vals[:, order * (degree + 1)] = rng.integers(1, 9, size=batch) * rng.choice([-1, 1], size=batch)  # leading != 0
        vals[rng.random(vals.shape) < 0.3] = 0
        vals[:, order * (degree + 1)] = np.where(vals[:, order * (degree + 1)] == 0, 3, vals[:, order * (degree + 1)])
        ode = from_complex([int(x) for x in vals.reshape(-1)], L)
        v = api.acb_ode_valuation(ode, order, degree)
        if v >= 0:
            continue
        # all instances must share the valuation for one launch
        if any(api.acb_ode_valuation(ode, order, degree, inst=i) != v for i in range(batch)):
            continue
        n = -v + 25
        init = random_balls(rng, batch * (-v) * 2, L, -3, 3, "ulp", short_frac=0.3)
        rc, ro = oracle.fuchs_batch(order, degree, n, ode, init, batch, False)
        rs = api.acb_ode_solve_fuchs_batch(ode, init, order, degree, n, batch, False, driver=drv)
        assert_same(ro, rs, f"integer operators L={L} order={order} degree={degree} v={v}")


def bessel_shifted(oracle, rng, L: int):
    """BASELINE config 2's operator: Bessel with complex nu, shifted to a complex ordinary point
    (acb_ode_bessel + acb_ode_shift, reference src/examples.c:13-22, src/acb_ode.c:124-135)."""
    nu = random_balls(rng, 2, L, 0, 0, "zero")
    ode0 = oracle.example("bessel", L, 0, nu)
    a = random_balls(rng, 2, L, 0, 1, "zero")
    return oracle.ode_shift(2, 2, ode0, a)


def case_fuchs_bessel_shared(oracle, drv, L: int, seed: int, n: int, batch: int):
    rng = np.random.default_rng(seed)
    ode = bessel_shifted(oracle, rng, L)
    assert api.acb_ode_valuation(ode, 2, 2) == -2
    init = random_balls(rng, batch * 2 * 2, L, -1, 0, "zero")
    rc, ro = oracle.fuchs_batch(2, 2, n, ode, init, batch, True)
    rs = api.acb_ode_solve_fuchs_batch(ode, init, 2, 2, n, batch, True, driver=drv)
    assert_same(ro, rs, f"bessel shared L={L}")


def case_fuchs_hypgeom(oracle, drv, L: int, seed: int):
    """valuation -1 (regular singular point at 0) with a shared operator; closed form
    (a)_n (b)_n / ((c)_n n!) must lie inside every ball."""
    a, b, c = F(1, 3), F(2, 5), F(7, 4)
    ode = oracle.example("hypgeom", L, 0, from_complex([a, b, c], L))
    # a, b, c are not dyadic: from_complex stored truncations with a radius, so use dyadic ones
    a, b, c = F(3, 8), F(5, 16), F(7, 4)
    ode = oracle.example("hypgeom", L, 0, from_complex([a, b, c], L))
    init = from_complex([1], L)
    n = 30
    rs = api.acb_ode_solve_fuchs_batch(ode, init, 2, 2, n, 1, True, driver=drv)
    rc, ro = oracle.fuchs_batch(2, 2, n, ode, init, 1, True)
    assert_same(ro, rs, "hypgeom fuchs")
    exact = [F(1)]
    for k in range(n):
        exact.append(exact[-1] * (a + k) * (b + k) / ((c + k) * (k + 1)))
    for k in range(n + 1):
        assert rs.contains(2 * k, exact[k]) and rs.contains(2 * k + 1, 0)


def case_frobenius(oracle, drv, L: int, seed: int, n: int = 20, batch: int = 3):
    """BASELINE config 3's path: hypgeom sweep, M = 1, rho = 0 and a generic rho."""
    rng = np.random.default_rng(seed)
    par = random_balls(rng, batch * 6, L, -1, 1, "zero")
    odes = Balls.concat([oracle.example("hypgeom", L, 0, par[i * 6:(i + 1) * 6]) for i in range(batch)])
    for rho in (Balls.zeros(batch * 2, L), random_balls(rng, batch * 2, L, -1, 1, "ulp")):
        ro = oracle.frobenius1_batch(2, 2, n, odes, rho, batch, False)
        rs = api.acb_ode_solve_frobenius1_batch(odes, rho, 2, 2, n, batch, False, False, driver=drv)
        assert_same(ro, rs, f"frobenius L={L}")


def case_frobenius_singleton(oracle, drv, L: int, seed: int, trials: int = 3):
    """reference tests/singleton_frobenius.c:18-46: zero P[i][j] for j <= i, P[order][order] = 1,
    rho = order - 1; the series shifted by rho must pass acb_ode_solves."""
    rng = np.random.default_rng(seed)
    for _ in range(trials):
        degree = 3 + int(rng.integers(0, 7))
        order = 2 + int(rng.integers(0, degree - 2))
        ne = (order + 1) * (degree + 1)
        ode = random_balls(rng, ne * 2, L, -4, 4, "ulp")
        m = ode.mant.reshape(order + 1, degree + 1, 2, L)
        t = ode.meta.reshape(order + 1, degree + 1, 2, 4)
        for i in range(order + 1):
            m[i, : i + 1] = 0
            t[i, : i + 1] = (0, FLAG_ZERO, 0, 0)
        one = from_complex([1], L)
        m[order, order] = one.mant
        t[order, order] = one.meta
        n = 2 + int(rng.integers(0, 30))
        rho = from_complex([order - 1], L)
        ro = oracle.frobenius1_batch(order, degree, n, ode, rho, 1, True)
        rs = api.acb_ode_solve_frobenius1_batch(ode, rho, order, degree, n, 1, True, True, driver=drv)
        assert_same(ro, rs, f"singleton frobenius L={L} order={order} degree={degree}")
        shifted = Balls.concat([Balls.zeros(2 * (order - 1), L), rs])
        assert oracle.ode_solves(order, degree, ode, shifted, n - 1)


def case_continuation(oracle, drv, L: int, seed: int, steps: int = 5, n: int = 30, batch: int = 3):
    """BASELINE configs[4]'s driver: analytic_continuation (reference src/fuchs_solver.c:147-163) of
    `batch` solutions of the hypergeometric operator along a polygon that avoids 0 and 1; includes
    an exact-zero corner (acb_ode_shift then skips the shift, src/acb_ode.c:128)."""
    rng = np.random.default_rng(seed)
    par = random_balls(rng, 6, L, -1, 1, "zero")
    ode = oracle.example("hypgeom", L, 0, par)
    th = np.linspace(0.3, 2.0, steps + 1)
    corners = [(F(float(0.5 + 0.25 * np.cos(t))).limit_denominator(2**20), F(float(0.3 * np.sin(t))).limit_denominator(2**20))
               for t in th]
    path = from_complex(corners, L)
    y0 = random_balls(rng, batch * 2 * 2, L, -1, 0, "zero")
    yo = oracle.continuation(2, 2, n, ode, path, y0, batch)
    ys = api.analytic_continuation_batch(ode, path, y0, 2, 2, n, batch, driver=drv)
    assert_same(yo, ys, f"analytic continuation L={L}")
    assert not (ys.meta[:, 1] & FLAG_NAN).any()
    # Legendre operator from the regular point 0 (an exact-zero corner) to 1/4
    leg = oracle.example("legendre", L, 3)
    path0 = from_complex([0, F(1, 8), F(1, 4)], L)
    yo = oracle.continuation(2, 2, n, leg, path0, y0, batch)
    ys = api.analytic_continuation_batch(leg, path0, y0, 2, 2, n, batch, driver=drv)
    assert_same(yo, ys, f"analytic continuation from the origin L={L}")


def case_ode_shift(oracle, drv, L: int, seed: int, batch: int = 5):
    """acb_ode_shift (reference src/acb_ode.c:124-135) as ONE launch for a batch: per-instance operators, one operator
    moved to several points, an exactly zero point (the copy), degree-0 rows."""
    rng = np.random.default_rng(seed)
    for order, degree in ((2, 2), (3, 4), (1, 0)):
        ne = (order + 1) * (degree + 1)
        odes = random_balls(rng, batch * ne * 2, L, -2, 2, "ulp")
        a = random_balls(rng, batch * 2, L, -1, 1, "ulp")
        a.mant[2:4] = 0
        a.meta[2:4] = (0, FLAG_ZERO, 0, 0)  # instance 1 is not moved
        got = api.acb_ode_shift_batch(odes, a, order, degree, driver=drv)
        want = Balls.concat([oracle.ode_shift(order, degree, odes[i * ne * 2:(i + 1) * ne * 2], a[2 * i:2 * i + 2])
                             for i in range(batch)])
        assert_same(want, got, f"acb_ode_shift L={L} order={order} degree={degree}")
        one = odes[: ne * 2]
        got = api.acb_ode_shift_batch(one, a, order, degree, driver=drv)
        want = Balls.concat([oracle.ode_shift(order, degree, one, a[2 * i:2 * i + 2]) for i in range(batch)])
        assert_same(want, got, f"acb_ode_shift of one operator to {batch} points L={L}")


def case_taylor_shift(oracle, drv, L: int, seed: int, length: int = 30, batch: int = 3):
    """acb_poly_taylor_shift (reference src/fuchs_solver.c:159), Horner form, one CTA per series; its first
    coefficients are bit for bit the Horner evaluation with derivatives that the continuation pipeline keeps."""
    rng = np.random.default_rng(seed)
    f = random_balls(rng, batch * length * 2, L, -3, 3, "ulp")
    a = random_balls(rng, batch * 2, L, -2, 0, "ulp")
    got = api.acb_poly_taylor_shift_batch(f, a, batch, driver=drv)
    assert_same(oracle.taylor_shift(f, a, batch), got, f"taylor shift L={L}")
    a1 = a[:2]
    got1 = api.acb_poly_taylor_shift_batch(f, a1, batch, driver=drv)
    assert_same(oracle.taylor_shift(f, a1, batch), got1, f"taylor shift by a shared point L={L}")
    for s in range(batch):
        h = oracle.horner_batch(f[s * length * 2:(s + 1) * length * 2], a1, 3)
        assert h.same_as(got1[s * length * 2: s * length * 2 + 6]), "Horner-with-derivatives == head of the Taylor shift"
    z = Balls.zeros(2, L)
    assert_same(f, api.acb_poly_taylor_shift_batch(f, z, batch, driver=drv), "taylor shift by zero")
    assert_same(oracle.taylor_shift(f[:2], a1, 1), api.acb_poly_taylor_shift_batch(f[:2], a1, 1, driver=drv), "length 1")


def case_continuation_series(oracle, drv, L: int, seed: int, steps: int = 3, n: int = 24, batch: int = 2):
    """analytic_continuation returning what the reference leaves in `res` (the complete re-expanded series), and
    corners where the shifted operator is SINGULAR: valuation > -order (only -v initial values are read, as in
    reference src/fuchs_solver.c:105-117) and valuation >= 0 (rejected)."""
    rng = np.random.default_rng(seed)
    par = random_balls(rng, 6, L, -1, 1, "zero")
    ode = oracle.example("hypgeom", L, 0, par)
    th = np.linspace(0.3, 1.4, steps + 1)
    path = from_complex([(F(float(0.5 + 0.25 * np.cos(t))).limit_denominator(2**20),
                          F(float(0.3 * np.sin(t))).limit_denominator(2**20)) for t in th], L)
    y0 = random_balls(rng, batch * 2 * 2, L, -1, 0, "zero")
    rc, yo, so = oracle.continuation_series(2, 2, n, ode, path, y0, batch)
    assert rc == 0
    ys, ss = api.analytic_continuation_series_batch(ode, path, y0, 2, 2, n, batch, driver=drv)
    assert_same(yo, ys, f"continuation (series variant) y L={L}")
    assert_same(so, ss, f"continuation: complete re-expanded series L={L}")
    assert_same(yo, api.analytic_continuation_batch(ode, path, y0, 2, 2, n, batch, driver=drv), "both variants agree on y")
    # the hypergeometric operator AT its singular point 0 (exact-zero corner: valuation -1), then on to 1/8
    path0 = from_complex([0, F(1, 8), F(3, 16)], L)
    rc, yo, so = oracle.continuation_series(2, 2, n, ode, path0, y0, batch)
    assert rc == 0
    ys, ss = api.analytic_continuation_series_batch(ode, path0, y0, 2, 2, n, batch, driver=drv)
    assert_same(yo, ys, f"continuation from a regular singular corner L={L}")
    assert_same(so, ss, f"continuation from a regular singular corner, series L={L}")
    assert not (ys.meta[:, 1] & FLAG_NAN).any()
    # Legendre moved to its singular point 1: (1 - z^2) -> -2z - z^2 exactly, valuation -1 after the shift
    leg = oracle.example("legendre", L, 3)
    path1 = from_complex([1, F(7, 8)], L)
    rc, yo, so = oracle.continuation_series(2, 2, n, leg, path1, y0, batch)
    assert rc == 0
    ys, ss = api.analytic_continuation_series_batch(leg, path1, y0, 2, 2, n, batch, driver=drv)
    assert_same(yo, ys, f"continuation from a corner that is singular after the shift L={L}")
    assert_same(so, ss, f"... series L={L}")
    # Bessel at the origin has valuation 0: undefined in the reference, rejected here
    bes = oracle.example("bessel", L, 0, random_balls(rng, 2, L, 0, 0, "zero"))
    rc, _, _ = oracle.continuation_series(2, 2, n, bes, path0, y0, batch)
    assert rc == -2
    try:
        api.analytic_continuation_batch(bes, path0, y0, 2, 2, n, batch, driver=drv)
        raise AssertionError("valuation >= 0 must be rejected")
    except _lib.CascadeB200Error:
        pass


def _leading_poly_operator(roots, L: int, order: int = 2, zero_roots: int = 0):
    """an operator whose leading coefficient is z^zero_roots * prod (z - root) (roots: exact (re, im) Fraction pairs),
    the other rows small integers; returns (ode, degree)"""
    from fractions import Fraction as Fr
    coef = [(Fr(1), Fr(0))]
    for rr, ri in roots:
        out = [(Fr(0), Fr(0)) for _ in range(len(coef) + 1)]
        for i, (a, b) in enumerate(coef):
            out[i + 1] = (out[i + 1][0] + a, out[i + 1][1] + b)                       # * z
            out[i] = (out[i][0] - (a * rr - b * ri), out[i][1] - (a * ri + b * rr))   # * (-root)
        coef = out
    coef = [(Fr(0), Fr(0))] * zero_roots + coef
    degree = len(coef) - 1
    rows = []
    for i in range(order):
        rows += [(Fr(i + 1), Fr(0))] + [(Fr(0), Fr(0))] * degree
    rows += coef
    return from_complex(rows, L), degree


def case_radius_and_truncation(oracle, drv, L: int, seed: int, batch: int = 6):
    """radius_of_convergence / truncation_order (reference src/fuchs_solver.c:3-94; property of reference tests/radius.c):
    rigorous on the device -- the balls are the oracle's bit for bit AND contain the true distance of the nearest nonzero
    root of the leading coefficient (exact rational roots here), also with a singularity at the origin factored out, for
    several Graeffe counts including one that runs into the exponent guard."""
    rng = np.random.default_rng(seed)
    from fractions import Fraction as Fr
    import math
    for trial in range(3):
        nroots = 2 + trial
        odes, trues = [], []
        for inst in range(batch):
            roots = [(Fr(int(rng.integers(-40, 41)), 8), Fr(int(rng.integers(-40, 41)), 8)) for _ in range(nroots)]
            roots = [r if r != (0, 0) else (Fr(3, 8), Fr(1, 8)) for r in roots]
            ode, degree = _leading_poly_operator(roots, L, zero_roots=inst % 2)
            # all instances of a launch share (order, degree): pad with a zero root count of the same parity
            odes.append((ode, degree, roots))
        by_deg = {}
        for ode, degree, roots in odes:
            by_deg.setdefault(degree, []).append((ode, roots))
        for degree, lst in by_deg.items():
            ode = Balls.concat([o for o, _ in lst])
            nb = len(lst)
            for n in (6, 20, 40):
                want, wd = oracle.radius_of_convergence(2, degree, ode, nb, n)
                got, gd = api.radius_of_convergence_batch(ode, 2, degree, nb, n, driver=drv)
                assert_same(want, got, f"radius_of_convergence L={L} degree={degree} n={n}")
                assert (wd == gd).all() and (gd >= 1).all() and (gd <= n).all()
                for i, (_, roots) in enumerate(lst):
                    true2 = min(r[0] * r[0] + r[1] * r[1] for r in roots)
                    lo, hi = got.mid(2 * i) - got.rad(2 * i), got.mid(2 * i) + got.rad(2 * i)
                    assert lo >= 0 and lo * lo <= true2 <= hi * hi, (L, degree, n, i, float(lo), float(hi), math.sqrt(true2))
                    assert got.rad(2 * i) <= got.mid(2 * i) * Fr(4 * degree * degree + 2, 2 ** min(int(gd[i]), 30)) + Fr(1, 2 ** 20)
    # no singularity away from zero: indeterminate, n_done = -1 (reference :17-22)
    ode, degree = _leading_poly_operator([], L, zero_roots=2)
    got, gd = api.radius_of_convergence_batch(ode, 2, degree, 1, 10, driver=drv)
    want, wd = oracle.radius_of_convergence(2, degree, ode, 1, 10)
    assert_same(want, got, "radius_of_convergence without singularities")
    assert gd[0] == -1 and wd[0] == -1 and (int(got.meta[0, 1]) & FLAG_NAN)
    # truncation_order
    eta = from_complex([Fr(1, 10), Fr(1, 3), Fr(2, 5), Fr(1, 2), Fr(1, 1000), Fr(3, 4)], L)
    alpha = from_complex([Fr(1, 2), Fr(1, 2), Fr(1, 2), Fr(1, 2), Fr(1, 2), Fr(1, 2)], L)
    for bits in (64, 32 * L):
        wn, wc = oracle.truncation_order(eta, alpha, bits)
        gn, gc = api.truncation_order_batch(eta, alpha, bits, driver=drv)
        assert_same(wn, gn, f"truncation_order L={L} bits={bits}")
        assert (wc == gc).all()
        for i, (e, a) in enumerate(((0.1, 0.5), (1 / 3, 0.5), (0.4, 0.5), (0.5, 0.5), (0.001, 0.5), (0.75, 0.5))):
            if e >= a:
                assert gc[i] == 0
            else:
                r = e / ((e + a) / 2)
                exact = (math.log(1 - r) - bits * math.log(2)) / math.log(r)
                assert abs(gc[i] - math.ceil(exact)) <= 1 and gc[i] >= exact - 1e-9, (i, gc[i], exact)


def case_horner(oracle, drv, L: int, seed: int, length: int = 40, npts: int = 9):
    rng = np.random.default_rng(seed)
    f = random_balls(rng, 2 * length, L, -3, 3, "ulp")
    pts = random_balls(rng, 2 * npts, L, -2, 0, "zero")
    for nder in (1, 2, 3):
        assert_same(oracle.horner_batch(f, pts, nder), api.acb_poly_evaluate_batch(f, pts, nder, driver=drv),
                    f"horner L={L} nder={nder}")


def case_horner_cancellation(oracle, drv, L: int, seed: int, length: int = 24, npts: int = 40):
    """Points with re == im (and nearly so) under series whose coefficients have re == im (and nearly so): the
    real part of every product o * a cancels completely or down to a few low bits, which defeats the guard-bit
    proof of the truncated products and sends the step through the exact full-width path (at 2048 bits and up:
    the part of the scratch columns that lives in global memory)."""
    rng = np.random.default_rng(seed)
    f = random_balls(rng, 2 * length, L, -3, 3, "ulp")
    pts = random_balls(rng, 2 * npts, L, -2, 0, "zero")
    for b, stride in ((f, 1), (pts, 1)):
        b.mant[1::2] = b.mant[0::2]
        b.meta[1::2] = b.meta[0::2]
        n = b.n_slots // 2
        for k in range(0, n, 3):  # every third value: im differs from re in one low bit
            limb = int(rng.integers(0, 3))
            b.mant[2 * k + 1, limb] ^= np.uint32(1 << int(rng.integers(0, 32)))
    for nder in (1, 2):
        assert_same(oracle.horner_batch(f, pts, nder), api.acb_poly_evaluate_batch(f, pts, nder, driver=drv),
                    f"horner cancellation L={L} nder={nder}")


def _frobenius_test_operator(rng, L: int, order: int, degree: int, dyadic: bool = False):
    """reference tests/frobenius.c:24-31: random operator with P[i][j] = 0 for j <= i and
    P[order][order] = P[order-1][order-1] = 1: valuation 0, rho = order - 2 a double indicial root."""
    ne = (order + 1) * (degree + 1)
    if dyadic:
        vals = [(F(int(rng.integers(-40, 40)), 8), F(int(rng.integers(-40, 40)), 8)) for _ in range(ne)]
        ode = from_complex(vals, L)
    else:
        ode = random_balls(rng, ne * 2, L, -4, 4, "ulp")
    m = ode.mant.reshape(order + 1, degree + 1, 2, L)
    t = ode.meta.reshape(order + 1, degree + 1, 2, 4)
    for i in range(order + 1):
        m[i, : i + 1] = 0
        t[i, : i + 1] = (0, FLAG_ZERO, 0, 0)
    one = from_complex([1], L)
    for i in (order, order - 1):
        m[i, i] = one.mant
        t[i, i] = one.meta
    return ode


def case_frobenius_general(oracle, drv, L: int, seed: int, trials: int = 3):
    """acb_ode_solve_frobenius with M = 2 (reference src/frobenius_solver.c:131-205) on the operators
    of reference tests/frobenius.c: bit-exact against the oracle, and gens[0] shifted by rho passes
    the reference's own acceptance test acb_ode_solves (tests/frobenius.c:37-40)."""
    rng = np.random.default_rng(seed)
    for _ in range(trials):
        degree = 3 + int(rng.integers(0, 4))
        order = 2 + int(rng.integers(0, degree - 2))
        ode = _frobenius_test_operator(rng, L, order, degree)
        n = 2 + int(rng.integers(0, 12))
        s_rho = order - 2
        rho = from_complex([s_rho], L)
        go = oracle.frobenius_general_batch(order, degree, n, ode, rho, 1, True, True, 2, 2)
        gs = api.acb_ode_solve_frobenius_batch(ode, rho, 2, 0, order, degree, n, 1, True, True, driver=drv)
        assert_same(go, gs, f"general frobenius L={L} order={order} degree={degree} n={n}")
        shifted = Balls.concat([Balls.zeros(2 * s_rho, L), gs[0:2 * (n + 1)]])
        assert oracle.ode_solves(order, degree, ode, shifted, n)


def case_frobenius_general_sweep(oracle, drv, L: int, seed: int, n: int = 8, batch: int = 4):
    """M = 2 and M = 3 (mul + alpha) on per-instance hypergeometric operators with per-instance complex
    rho (nothing special about rho: the recurrence and the rho-derivative bookkeeping are generic),
    plus the exact double root rho = 0 of c = 1."""
    rng = np.random.default_rng(seed)
    par = random_balls(rng, batch * 6, L, -1, 1, "zero")
    odes = Balls.concat([oracle.example("hypgeom", L, 0, par[i * 6:(i + 1) * 6]) for i in range(batch)])
    rho = random_balls(rng, batch * 2, L, -1, 1, "ulp")
    for mul, alpha in ((2, 0), (2, 1), (1, 1)):
        M = mul + alpha
        go = oracle.frobenius_general_batch(2, 2, n, odes, rho, batch, False, False, mul, M)
        gs = api.acb_ode_solve_frobenius_batch(odes, rho, mul, alpha, 2, 2, n, batch, False, False, driver=drv)
        assert_same(go, gs, f"general frobenius sweep L={L} mul={mul} alpha={alpha}")
    # c = 1: rho = 0 is a double root, gens[1] carries the coefficient of the logarithmic solution
    ode = oracle.example("hypgeom", L, 0, from_complex([F(3, 8), F(5, 16), 1], L))
    rho0 = from_complex([0], L)
    go = oracle.frobenius_general_batch(2, 2, n, ode, rho0, 1, True, True, 2, 2)
    gs = api.acb_ode_solve_frobenius_batch(ode, rho0, 2, 0, 2, 2, n, 1, True, True, driver=drv)
    assert_same(go, gs, f"general frobenius c=1 L={L}")
    a, b = F(3, 8), F(5, 16)
    ex = F(1)
    for k in range(n):  # gens[0] = 2F1(a, b; 1; z)
        assert gs.contains(2 * k, ex)
        ex = ex * (a + k) * (b + k) / ((k + 1) * (k + 1))


def case_solution_ops(oracle, drv, L: int, seed: int, batch: int = 3):
    """reference tests/solution_extend.c and tests/solution_update.c: _acb_ode_solution_extend,
    _update and _normalize on random polynomials (length <= 20) and random rho, bit-exact against
    the oracle; extend's defining property gens[j][i] == f^(j)(rho) is checked with the oracle's
    Horner evaluation (the reference test uses acb_equal against acb_poly_evaluate)."""
    rng = np.random.default_rng(seed)
    for trial in range(3):
        M = 1 + int(rng.integers(0, 5))
        cap = 2 + int(rng.integers(0, 8))
        rho = random_balls(rng, batch * 2, L, -2, 2, "ulp")
        gens = Balls.zeros(batch * M * cap * 2, L)
        for nu in range(cap):
            flen = 1 + int(rng.integers(0, 20))
            f = random_balls(rng, batch * flen * 2, L, -3, 3, "ulp", short_frac=0.2)
            go, fo = oracle.solution_op("extend", rho, gens, batch, M, M, f, nu)
            gs, fs = api.acb_ode_solution_op_batch(api.SOL_EXTEND, rho, gens, batch, M, M, f, nu, driver=drv)
            assert_same(go, gs, f"solution_extend L={L} M={M} nu={nu}")
            assert_same(fo, fs, "solution_extend consumes g_nu")
            assert (fs.meta[:, 1] & FLAG_ZERO).all()
            for s in range(batch):  # gens[0][nu] = f(rho)
                val = oracle.horner_batch(f[s * flen * 2:(s + 1) * flen * 2], rho[2 * s:2 * s + 2], 1)
                got = gs[2 * ((s * M + 0) * cap + nu):2 * ((s * M + 0) * cap + nu) + 2]
                assert_same(val, got, "extend: gens[0][nu] == f(rho)")
            gens = gs
        flen = 1 + int(rng.integers(0, 20))
        f = random_balls(rng, batch * flen * 2, L, -3, 3, "ulp", short_frac=0.2)
        go, fo = oracle.solution_op("update", rho, gens, batch, M, M, f)
        gs, fs = api.acb_ode_solution_op_batch(api.SOL_UPDATE, rho, gens, batch, M, M, f, driver=drv)
        assert_same(go, gs, f"solution_update L={L} M={M}")
        assert_same(fo, fs, "solution_update consumes f")
        mul = 1 + int(rng.integers(0, M))
        go, _ = oracle.solution_op("normalize", rho, gs, batch, mul, M)
        gn, _ = api.acb_ode_solution_op_batch(api.SOL_NORMALIZE, rho, gs, batch, mul, M, driver=drv)
        assert_same(go, gn, f"solution_normalize L={L} M={M} mul={mul}")


def case_indicial_polynomial(oracle, drv, L: int, seed: int, trials: int = 4):
    """reference tests/indicial_polynomial.c:18-96: the polynomial f_nu(rho + shift), bit-exact against
    the oracle, and consistent with indicial_polynomial_evaluate at an integer rho (overlap)."""
    rng = np.random.default_rng(seed)
    for trial in range(trials):
        degree = 3 + int(rng.integers(0, 5))
        order = 2 + int(rng.integers(0, degree - 2))
        ne = (order + 1) * (degree + 1)
        ode = random_balls(rng, ne * 2, L, -4, 4, "ulp", short_frac=0.2)
        m = ode.mant.reshape(order + 1, degree + 1, 2, L)
        t = ode.meta.reshape(order + 1, degree + 1, 2, 4)
        for i in range(order + 1):
            m[i, :i] = 0
            t[i, :i] = (0, FLAG_ZERO, 0, 0)
        v = api.acb_ode_valuation(ode, order, degree)
        r = 4 + int(rng.integers(0, 10))
        for nu in range(degree + 2):
            for shift in (0, r):
                po = oracle.indicial_polynomial(order, degree, ode, nu, shift)
                ps = api.indicial_polynomial_batch(ode, order, degree, nu, shift, 1, True, driver=drv)
                assert_same(po, ps, f"indicial_polynomial L={L} nu={nu} shift={shift}")
                # f_nu(r) both ways: Horner of the polynomial at rho = r - shift vs direct evaluation
                at_ = from_complex([r - shift], L)
                hv = oracle.horner_batch(ps, at_, 1)
                ev = oracle.indicial_polynomial_evaluate(order, degree, ode, nu, from_complex([r], L), 0)
                for part in (0, 1):
                    assert abs(hv.mid(part) - ev.mid(part)) <= hv.rad(part) + ev.rad(part)


def case_frobenius_rhs(oracle, drv, L: int, seed: int, n: int = 14, batch: int = 4):
    """_acb_ode_solve_frobenius with a right-hand side (reference src/frobenius_solver.c:86-96,100):
    per-instance rhs series of different effective lengths, one instance with a zero rhs (it must
    take the homogeneous branch g_0 = 1), hypergeometric operators, generic complex rho."""
    rng = np.random.default_rng(seed)
    par = random_balls(rng, batch * 6, L, -1, 1, "zero")
    odes = Balls.concat([oracle.example("hypgeom", L, 0, par[i * 6:(i + 1) * 6]) for i in range(batch)])
    rho = random_balls(rng, batch * 2, L, -1, 1, "ulp")
    hl = 6
    rhs = random_balls(rng, batch * hl * 2, L, -2, 2, "ulp")
    m = rhs.mant.reshape(batch, hl, 2, L)
    t = rhs.meta.reshape(batch, hl, 2, 4)
    m[1, 3:] = 0
    t[1, 3:] = (0, FLAG_ZERO, 0, 0)  # shorter right-hand side
    m[2] = 0
    t[2] = (0, FLAG_ZERO, 0, 0)      # zero right-hand side: homogeneous branch
    m[3, 0] = 0
    t[3, 0] = (0, FLAG_ZERO, 0, 0)   # rhs_0 = 0 but rhs != 0
    ro = oracle.frobenius1_rhs_batch(2, 2, n, odes, rho, rhs, batch, False)
    rs = api.acb_ode_solve_frobenius1_batch(odes, rho, 2, 2, n, batch, False, False, driver=drv, rhs=rhs)
    assert_same(ro, rs, f"frobenius with rhs L={L}")
    assert rs.mid(2 * (2 * (n + 1))) == 1  # instance 2: g_0 = 1


def trans_inputs(rng, L: int, n: int) -> Balls:
    """Complex balls for exp/log: generic, real, imaginary, tiny, large, negative real axis (branch cut side),
    wide radii, an exact zero and a ball containing zero."""
    x = random_balls(rng, 2 * n, L, -3, 3, "ulp")
    special = [(F(3, 2), F(1, 4)), (F(-5, 8), F(7, 3)), (F(1, 1024), 0), (-3, 0), (0, F(-9, 4)), (F(12345, 7), -100),
               (F(1, 3), F(1, 10**6)), (20, 0), (F(-1, 2**40), F(1, 2**60)), (1, 0), (0, 0), (F(1, 2**300), 0)]
    sp = from_complex(special, L)
    k = min(len(special), n)
    x.mant[: 2 * k] = sp.mant[: 2 * k]
    x.meta[: 2 * k] = sp.meta[: 2 * k]
    if n > k + 2:  # a wide ball, and a ball containing zero
        x.meta[2 * k, 2], x.meta[2 * k, 3] = 1 << 29, x.meta[2 * k, 0] - 20
        x.meta[2 * k + 2, 2], x.meta[2 * k + 2, 3] = 1 << 29, x.meta[2 * k + 2, 0] + 3
        x.meta[2 * k + 3, 2], x.meta[2 * k + 3, 3] = 1 << 29, x.meta[2 * k + 3, 0] + 3
    if n > k + 4:
        # a ball straddling the branch cut of log (re < 0, im = small +- larger): Im log = [0 +- pi]; and its neighbour,
        # re < 0 with an im ball that stays on one side of the cut
        for e, wide in ((k + 2, 2), (k + 3, -6)):
            x.meta[2 * e, 1] |= 1            # re negative
            x.meta[2 * e, 0] = 1
            x.meta[2 * e + 1, 0] = -40       # |im| ~ 2^-40
            x.meta[2 * e + 1, 2], x.meta[2 * e + 1, 3] = 1 << 29, -40 + wide
    return x


def case_exp_log(oracle, drv, L: int, seed: int, n: int = 24):
    """acb_exp / acb_log on the device against the oracle's algorithms (bit-exact: same ball operations in
    the same order, double-precision seed from IEEE +,-,*,/ only), plus exp(log(a)) overlapping a."""
    rng = np.random.default_rng(seed)
    x = trans_inputs(rng, L, n)
    le = api.acb_exp_batch(x, driver=drv)
    assert_same(oracle.ew_trans("exp", x), le, f"acb_exp L={L}")
    ll = api.acb_log_batch(x, driver=drv)
    assert_same(oracle.ew_trans("log", x), ll, f"acb_log L={L}")
    back = api.acb_exp_batch(ll, driver=drv)
    for i in range(n):
        if (int(ll.meta[2 * i, 1]) | int(ll.meta[2 * i + 1, 1])) & FLAG_NAN:
            continue
        for part in (0, 1):
            s = 2 * i + part
            assert abs(back.mid(s) - x.mid(s)) <= back.rad(s) + x.rad(s), (i, part)


def case_solution_evaluate(oracle, drv, L: int, seed: int, npts: int = 7, cap: int = 12):
    """acb_ode_solution_evaluate (reference src/acb_ode_solution.c:24-58; reference tests/solution_eval.c):
    M = 1..3, rho zero / small integer / generic complex, points including an exact zero."""
    rng = np.random.default_rng(seed)
    pts = random_balls(rng, 2 * npts, L, -2, 0, "zero")
    pts.mant[2:4] = 0
    pts.meta[2:4] = (0, FLAG_ZERO, 0, 0)  # point 1 is exactly zero: indeterminate result
    for M in (1, 2, 3):
        gens = random_balls(rng, 2 * M * cap, L, -2, 2, "ulp")
        g = gens.mant.reshape(M, cap, 2, L)
        t = gens.meta.reshape(M, cap, 2, 4)
        g[M - 1, cap - 3:] = 0
        t[M - 1, cap - 3:] = (0, FLAG_ZERO, 0, 0)  # a shorter series
        for rho in (from_complex([0], L), from_complex([5], L), from_complex([(F(1, 3), F(-1, 7))], L),
                    random_balls(rng, 2, L, -1, 1, "ulp")):
            mode, e = api.pow_mode_of(rho)
            want = oracle.solution_evaluate(gens, M, rho, mode, e, pts)
            got = api.acb_ode_solution_evaluate_batch(gens, M, rho, pts, driver=drv)
            assert_same(want, got, f"solution_evaluate L={L} M={M} mode={mode}")
            assert int(got.meta[2, 1]) & FLAG_NAN


def case_ode_apply(oracle, drv, L: int, seed: int, batch: int = 3):
    """acb_ode_apply / acb_ode_solves on the device (reference src/acb_ode.c:185-240): the residual of computed
    Fuchs series is bit-identical to the oracle's and passes the acceptance test; a perturbed series fails it."""
    rng = np.random.default_rng(seed)
    degree, order = 4, 2
    ne = (order + 1) * (degree + 1)
    ode = random_balls(rng, batch * ne * 2, L, -4, 4, "ulp", short_frac=0.2)
    n = order + 14
    init = random_balls(rng, batch * order * 2, L, -3, 3, "ulp")
    rs = api.acb_ode_solve_fuchs_batch(ode, init, order, degree, n, batch, False, driver=drv)
    out, solved = api.acb_ode_apply_batch(ode, rs, order, degree, batch, False, deg=n - order, driver=drv)
    for i in range(batch):
        want = oracle.ode_apply(order, degree, ode[i * ne * 2:(i + 1) * ne * 2], rs[i * (n + 1) * 2:(i + 1) * (n + 1) * 2])
        assert_same(want, out[i * (n + 1 + degree) * 2:(i + 1) * (n + 1 + degree) * 2], f"ode_apply L={L} instance {i}")
    assert solved.all()
    bad = rs.copy()
    bad.mant[2 * (n + 1) + 10, L - 1] ^= np.uint32(1 << 20)  # instance 1, coefficient 5: off by 2^-11 relative
    _, solved = api.acb_ode_apply_batch(ode, bad, order, degree, batch, False, deg=n - order, driver=drv)
    assert list(solved) == [True, False] + [True] * (batch - 2)


def case_frobenius_general_edge(oracle, drv, L: int):
    """Edge cases of the polynomial-in-rho recurrence where lengths shrink and the loop ends early
    (reference src/frobenius_solver.c:193-194): integer operators with polynomial solutions (Legendre at an
    ordinary point: the series terminates and g_new becomes the zero polynomial), an operator whose rescaling
    value f_0(rho + nu) contains zero (the rescaling is skipped, :171), M = 4, and a shared operator with
    per-instance rho."""
    for n_leg, rho_v, mul, alpha in ((4, 0, 1, 1), (3, 1, 1, 1), (6, 0, 2, 0)):
        ode = oracle.example("legendre", L, n_leg)
        rho = from_complex([rho_v], L)
        M = mul + alpha
        go = oracle.frobenius_general_batch(2, 2, 14, ode, rho, 1, True, True, mul, M)
        gs = api.acb_ode_solve_frobenius_batch(ode, rho, mul, alpha, 2, 2, 14, 1, True, True, driver=drv)
        assert_same(go, gs, f"general frobenius, legendre({n_leg}) rho={rho_v} L={L}")
    # f_0(rho + nu) = (rho + nu)(rho + nu - 1 + c) vanishes at nu = 2 for rho = -2: the division is skipped there
    ode = oracle.example("hypgeom", L, 0, from_complex([F(3, 8), F(5, 16), F(7, 4)], L))
    rho = from_complex([-2], L)
    go = oracle.frobenius_general_batch(2, 2, 6, ode, rho, 1, True, True, 2, 2)
    gs = api.acb_ode_solve_frobenius_batch(ode, rho, 2, 0, 2, 2, 6, 1, True, True, driver=drv)
    assert_same(go, gs, f"general frobenius through a zero of f_0 L={L}")
    # M = 4, shared operator, per-instance rho
    rng = np.random.default_rng(17 + L)
    rho = random_balls(rng, 3 * 2, L, -1, 1, "ulp")
    go = oracle.frobenius_general_batch(2, 2, 5, ode, rho, 3, True, False, 3, 4)
    gs = api.acb_ode_solve_frobenius_batch(ode, rho, 3, 1, 2, 2, 5, 3, True, False, driver=drv)
    assert_same(go, gs, f"general frobenius M=4 L={L}")
