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
