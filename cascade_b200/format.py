"""Host-side ball format helpers (marshalling only -- no ball ARITHMETIC lives here).

A *slot* is one real ball: ``L`` little-endian 32-bit mantissa limbs plus a 4-word meta record
``(exp, flags, rman, rexp)`` -- see ``include/cb_format.h``.  A complex ball is two consecutive
slots (re, im).  This is the device replacement for Arb's ``acb_struct`` that the reference
stores everywhere (reference ``src/acb_ode.h:9-15,45-50``).

Host layout ("AoS"):   mant[slot, limb] uint32,  meta[slot, 4] int32.
Device layout ("SoA"): mant[limb, slot] uint32 (limb-major: a warp reading limb k of 32
consecutive slots touches one 128-byte line), meta[slot, 4] int32.  For a series batch the slot
axis is ``coefficient * (2*B) + 2*instance + part``.
"""
from __future__ import annotations

from fractions import Fraction

import numpy as np

FLAG_SIGN = 1
FLAG_ZERO = 2
FLAG_NAN = 4
MAG_BITS = 30
MAG_INF = 0xFFFFFFFF


class Balls:
    """A flat array of real-ball slots in host (AoS) layout."""

    __slots__ = ("L", "mant", "meta")

    def __init__(self, L: int, mant: np.ndarray, meta: np.ndarray):
        assert L % 2 == 0 and 2 <= L <= 128
        assert mant.dtype == np.uint32 and meta.dtype == np.int32
        assert mant.ndim == 2 and mant.shape[1] == L and meta.shape == (mant.shape[0], 4)
        self.L = L
        self.mant = np.ascontiguousarray(mant)
        self.meta = np.ascontiguousarray(meta)

    # ---- construction
    @classmethod
    def zeros(cls, n_slots: int, L: int) -> "Balls":
        meta = np.zeros((n_slots, 4), np.int32)
        meta[:, 1] = FLAG_ZERO
        return cls(L, np.zeros((n_slots, L), np.uint32), meta)

    @property
    def n_slots(self) -> int:
        return self.mant.shape[0]

    @property
    def prec(self) -> int:
        return 32 * self.L

    def copy(self) -> "Balls":
        return Balls(self.L, self.mant.copy(), self.meta.copy())

    def __getitem__(self, idx) -> "Balls":
        m, t = self.mant[idx], self.meta[idx]
        if m.ndim == 1:
            m, t = m[None, :], t[None, :]
        return Balls(self.L, m.copy(), t.copy())

    def __setitem__(self, idx, other: "Balls"):
        self.mant[idx] = other.mant if not isinstance(idx, (int, np.integer)) else other.mant[0]
        self.meta[idx] = other.meta if not isinstance(idx, (int, np.integer)) else other.meta[0]

    @classmethod
    def concat(cls, parts) -> "Balls":
        parts = list(parts)
        return cls(parts[0].L, np.concatenate([p.mant for p in parts]),
                   np.concatenate([p.meta for p in parts]))

    def same_as(self, other: "Balls") -> bool:
        return (self.L == other.L and np.array_equal(self.mant, other.mant)
                and np.array_equal(self.meta, other.meta))

    # ---- exact scalar access
    def set_exact(self, slot: int, value, rad=0) -> None:
        """Store ``value`` (int / Fraction with power-of-two denominator / float).  The midpoint
        is truncated toward zero to prec bits; one ulp is added to ``rad`` when that is inexact
        (the arb_set_round convention)."""
        fr = Fraction(value)
        L, prec = self.L, self.prec
        self.mant[slot] = 0
        self.meta[slot] = (0, FLAG_ZERO, 0, 0)
        radf = Fraction(rad)
        if fr != 0:
            sign = fr < 0
            fr = abs(fr)
            num, den = fr.numerator, fr.denominator
            # exponent e with fr in [2^(e-1), 2^e)
            e = num.bit_length() - den.bit_length()
            if Fraction(2) ** e <= fr:
                e += 1
            while Fraction(2) ** (e - 1) > fr:
                e -= 1
            scaled = fr * Fraction(2) ** (prec - e)
            m = scaled.numerator // scaled.denominator
            if m * scaled.denominator != scaled.numerator:
                radf += Fraction(2) ** (e - prec)
            assert m >> (prec - 1) == 1
            self.mant[slot] = np.frombuffer(m.to_bytes(4 * L, "little"), dtype=np.uint32)
            self.meta[slot, 0] = e
            self.meta[slot, 1] = FLAG_SIGN if sign else 0
        if radf != 0:
            rman, rexp = mag_from_fraction_upper(radf)
            self.meta[slot, 2] = np.int32(rman)
            self.meta[slot, 3] = rexp

    def mid(self, slot: int) -> Fraction:
        exp, flags, _, _ = (int(v) for v in self.meta[slot])
        if flags & FLAG_NAN:
            raise ValueError("indeterminate ball")
        if flags & FLAG_ZERO:
            return Fraction(0)
        m = int.from_bytes(self.mant[slot].tobytes(), "little")
        v = Fraction(m) * Fraction(2) ** (exp - self.prec)
        return -v if flags & FLAG_SIGN else v

    def rad(self, slot: int) -> Fraction:
        _, flags, rman, rexp = (int(v) for v in self.meta[slot])
        if flags & FLAG_NAN:
            raise ValueError("indeterminate ball")
        rman &= 0xFFFFFFFF
        if rman == 0:
            return Fraction(0)
        return Fraction(rman) * Fraction(2) ** (rexp - MAG_BITS)

    def is_nan(self, slot: int) -> bool:
        return bool(int(self.meta[slot, 1]) & FLAG_NAN)

    def contains(self, slot: int, value) -> bool:
        if self.is_nan(slot):
            return True
        return abs(Fraction(value) - self.mid(slot)) <= self.rad(slot)

    def complex_mid(self, i: int):
        return self.mid(2 * i), self.mid(2 * i + 1)

    # ---- device layout
    def to_soa(self) -> tuple[np.ndarray, np.ndarray]:
        """(mant[L, n_slots], meta[n_slots, 4]) contiguous arrays in device layout."""
        return np.ascontiguousarray(self.mant.T), self.meta.copy()

    @classmethod
    def from_soa(cls, L: int, mant_soa: np.ndarray, meta: np.ndarray) -> "Balls":
        return cls(L, np.ascontiguousarray(mant_soa.reshape(L, -1).T),
                   np.ascontiguousarray(meta.reshape(-1, 4)))


def mag_from_fraction_upper(r: Fraction) -> tuple[int, int]:
    """Smallest (man, exp) with man in [2^29, 2^30) and man * 2^(exp-30) >= r."""
    assert r > 0
    e = r.numerator.bit_length() - r.denominator.bit_length()
    while Fraction(2) ** e <= r:
        e += 1
    while Fraction(2) ** (e - 1) > r:
        e -= 1
    scaled = r * Fraction(2) ** (MAG_BITS - e)
    man = -((-scaled.numerator) // scaled.denominator)  # ceil
    if man >> MAG_BITS:
        man = 1 << (MAG_BITS - 1)
        e += 1
    return man, e


def from_values(values, L: int, rads=None) -> Balls:
    """One slot per entry of ``values`` (real numbers)."""
    b = Balls.zeros(len(values), L)
    for i, v in enumerate(values):
        b.set_exact(i, v, 0 if rads is None else rads[i])
    return b


def from_complex(values, L: int) -> Balls:
    """Two slots per entry; entries are ``(re, im)`` pairs or plain reals."""
    b = Balls.zeros(2 * len(values), L)
    for i, v in enumerate(values):
        re, im = v if isinstance(v, tuple) else (v, 0)
        b.set_exact(2 * i, re)
        b.set_exact(2 * i + 1, im)
    return b


def random_balls(rng: np.random.Generator, n_slots: int, L: int, exp_lo: int = -1, exp_hi: int = 1,
                 rad: str = "zero", signed: bool = True, short_frac: float = 0.0) -> Balls:
    """Dense random slots: full-width mantissas (top bit set), exponents uniform in
    [exp_lo, exp_hi], random signs.  rad: "zero" | "ulp" (radius 2^(exp-prec)*k, k<=8) |
    "wide" (radius ~2^-40 relative).  ``short_frac`` of the slots get mantissas with only a few
    significant limbs (integers / dyadic rationals), which exercises exact results."""
    mant = rng.integers(0, 2**32, size=(n_slots, L), dtype=np.uint64).astype(np.uint32)
    mant[:, L - 1] |= np.uint32(0x80000000)
    if short_frac > 0:
        short = rng.random(n_slots) < short_frac
        keep = rng.integers(1, 3, size=n_slots)
        for s in np.nonzero(short)[0]:
            mant[s, : L - keep[s]] = 0
    meta = np.zeros((n_slots, 4), np.int32)
    meta[:, 0] = rng.integers(exp_lo, exp_hi + 1, size=n_slots)
    if signed:
        meta[:, 1] = rng.integers(0, 2, size=n_slots) * FLAG_SIGN
    if rad == "ulp":
        meta[:, 2] = rng.integers(1 << 29, 1 << 30, size=n_slots)
        meta[:, 3] = meta[:, 0] - 32 * L + rng.integers(1, 4, size=n_slots)
    elif rad == "wide":
        meta[:, 2] = rng.integers(1 << 29, 1 << 30, size=n_slots)
        meta[:, 3] = meta[:, 0] - 40
    elif rad != "zero":
        raise ValueError(rad)
    return Balls(L, mant, meta)


# ---------------------------------------------------------------- device array layout
def pack_entries(b: Balls, entries: int, S: int) -> tuple[np.ndarray, np.ndarray]:
    """Balls with slot index ``e*S + s``  ->  (mant[E, L, S] uint32, meta[E, S, 4] int32)."""
    assert b.n_slots == entries * S
    mant = np.ascontiguousarray(b.mant.reshape(entries, S, b.L).transpose(0, 2, 1))
    meta = np.ascontiguousarray(b.meta.reshape(entries, S, 4))
    return mant, meta


def unpack_entries(L: int, mant: np.ndarray, meta: np.ndarray) -> Balls:
    """Inverse of :func:`pack_entries`."""
    E, L_, S = mant.shape
    assert L_ == L
    return Balls(L, np.ascontiguousarray(mant.transpose(0, 2, 1)).reshape(E * S, L),
                 np.ascontiguousarray(meta).reshape(E * S, 4))


def instance_major_to_entry_major(b: Balls, batch: int, per_inst: int) -> Balls:
    """[batch][per_inst] complex balls -> [per_inst][batch] complex balls."""
    m = b.mant.reshape(batch, per_inst, 2, b.L).transpose(1, 0, 2, 3).reshape(-1, b.L)
    t = b.meta.reshape(batch, per_inst, 2, 4).transpose(1, 0, 2, 3).reshape(-1, 4)
    return Balls(b.L, np.ascontiguousarray(m), np.ascontiguousarray(t))


def entry_major_to_instance_major(b: Balls, batch: int, per_inst: int) -> Balls:
    m = b.mant.reshape(per_inst, batch, 2, b.L).transpose(1, 0, 2, 3).reshape(-1, b.L)
    t = b.meta.reshape(per_inst, batch, 2, 4).transpose(1, 0, 2, 3).reshape(-1, 4)
    return Balls(b.L, np.ascontiguousarray(m), np.ascontiguousarray(t))
