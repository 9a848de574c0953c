"""Independent exact models used to pin the oracle (pure Python integers / Fractions):
  * the Fuchs recurrence of SURVEY.md App. A in exact rational arithmetic,
  * correctly rounded (toward zero) midpoints of +, -, *, / and the fused two-term dot product.
Nothing here shares code with oracle/ or with the CUDA sources."""
from __future__ import annotations

from fractions import Fraction as F


def trunc_to_prec(v: F, prec: int):
    """(midpoint as Fraction, inexact) of v rounded toward zero to prec bits from the leading bit."""
    if v == 0:
        return F(0), False
    s = -1 if v < 0 else 1
    a = abs(v)
    e = a.numerator.bit_length() - a.denominator.bit_length()
    while F(2) ** e <= a:
        e += 1
    while F(2) ** (e - 1) > a:
        e -= 1
    scaled = a * F(2) ** (prec - e)
    m = scaled.numerator // scaled.denominator
    return s * F(m) * F(2) ** (e - prec), m * scaled.denominator != scaled.numerator


def fuchs_exact(P, init, n, v):
    """Exact coefficients c_0..c_n of L y = 0.  P[i][j] = coefficient of z^j in p_i (Fractions,
    complex as (re, im) is not needed: callers pass real or complex Python numbers made of Fractions)."""
    order, degree = len(P) - 1, len(P[0]) - 1
    c = list(init) + [None] * (n + 1 - len(init))

    def ff(b, i):
        r = 1
        for k in range(i):
            r *= (b - k)
        return r

    def Q(k, b):
        tot = 0
        for i in range(order + 1):
            if 0 <= i + k <= degree:
                tot += P[i][i + k] * ff(b, i)
        return tot

    for bmax in range(-v, n + 1):
        e = bmax + v
        bmin = max(0, e - degree)
        acc = 0
        for b in range(bmin, bmax):
            acc += c[b] * Q(e - b, b)
        c[bmax] = -acc / Q(v, bmax)
    return c


class C:
    """Exact complex rational."""

    def __init__(self, re, im=0):
        self.re, self.im = F(re), F(im)

    def __add__(self, o):
        o = o if isinstance(o, C) else C(o)
        return C(self.re + o.re, self.im + o.im)

    __radd__ = __add__

    def __neg__(self):
        return C(-self.re, -self.im)

    def __sub__(self, o):
        o = o if isinstance(o, C) else C(o)
        return C(self.re - o.re, self.im - o.im)

    def __mul__(self, o):
        o = o if isinstance(o, C) else C(o)
        return C(self.re * o.re - self.im * o.im, self.re * o.im + self.im * o.re)

    __rmul__ = __mul__

    def __truediv__(self, o):
        o = o if isinstance(o, C) else C(o)
        d = o.re * o.re + o.im * o.im
        return C((self.re * o.re + self.im * o.im) / d, (self.im * o.re - self.re * o.im) / d)


# ---------------------------------------------------------------- general Frobenius, exact
# The method of reference src/frobenius_solver.c:131-205 in exact arithmetic: polynomials in rho
# are lists of C (lowest degree first).  Written from the mathematics (Frobenius' method with the
# rescaling by f_0(rho + nu)), not from oracle/: used to check that every oracle ball contains the
# exact value.
def _ptrim(p):
    while p and p[-1].re == 0 and p[-1].im == 0:
        p = p[:-1]
    return p


def _padd(p, q, sign=1):
    n = max(len(p), len(q))
    out = []
    for i in range(n):
        a = p[i] if i < len(p) else C(0)
        b = q[i] if i < len(q) else C(0)
        out.append(a + b if sign > 0 else a - b)
    return _ptrim(out)


def _pmul(p, q):
    if not p or not q:
        return []
    out = [C(0)] * (len(p) + len(q) - 1)
    for i, a in enumerate(p):
        for j, b in enumerate(q):
            out[i + j] = out[i + j] + a * b
    return _ptrim(out)


def _peval(p, x):
    r = C(0)
    for a in reversed(p):
        r = r * x + a
    return r


def _pder(p):
    return _ptrim([p[i] * i for i in range(1, len(p))])


def indicial_poly_exact(P, v, nu, shift):
    """f_nu(rho + shift) as a polynomial in rho: sum_lam P[lam][lam+nu+v] (rho+shift)(rho+shift-1)...(lam factors)."""
    order, degree = len(P) - 1, len(P[0]) - 1
    k = nu + v
    out = []
    for lam in range(order + 1):
        if not (0 <= lam + k <= degree):
            continue
        term = [P[lam][lam + k]]
        for t in range(lam):
            term = _pmul(term, [C(shift - t), C(1)])
        out = _padd(out, term)
    return out


def frobenius_general_exact(P, v, rho, mul, M, n):
    """gens[i][nu] (i < M, nu <= n) exactly as the reference's algorithm defines them."""
    order, degree = len(P) - 1, len(P[0]) - 1
    g_rho = [[] for _ in range(degree)]
    g_rho[0] = [C(1)]
    gens = [[C(0)] * (n + 1) for _ in range(M)]
    gens[0][0] = C(1)
    for nu in range(1, n + 1):
        g_new = []
        i = min(max(nu, 1), degree)
        while i > 0:
            g_new = _padd(g_new, _pmul(indicial_poly_exact(P, v, i, nu - i), g_rho[i - 1]), -1)
            i -= 1
        ind = indicial_poly_exact(P, v, 0, nu)
        temp = _peval(ind, rho)
        if not (temp.re == 0 and temp.im == 0):
            ind = [a / temp for a in ind]
            g_new = [a / temp for a in g_new]
        all_zero = not g_new
        for i in range(degree - 1, 0, -1):
            g_rho[i] = _pmul(g_rho[i - 1], ind)
            all_zero = all_zero and not g_rho[i]
        g_rho[0] = list(g_new)
        # update: gens[m] <- sum_k C(m,k) ind^(k)(rho) gens[m-k]   (Leibniz)
        derivs, f = [], list(ind)
        for k in range(M):
            derivs.append(_peval(f, rho))
            f = _pder(f)
        from math import comb
        new = []
        for m in range(M):
            row = [C(0)] * (n + 1)
            for k in range(m + 1):
                for j in range(n + 1):
                    row[j] = row[j] + derivs[k] * comb(m, k) * gens[m - k][j]
            new.append(row)
        gens = new
        # extend: gens[i][nu] = g_new^(i)(rho)
        f = list(g_new)
        for i in range(M):
            gens[i][nu] = _peval(f, rho)
            f = _pder(f)
        if all_zero:
            break
    t = gens[M - mul][0]
    return [[a / t for a in row] for row in gens]
