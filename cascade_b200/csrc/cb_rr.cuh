// cb_rr.cuh -- the fused two-term dot product of the 1024-bit recurrence in REDUCED RADIX.
//
// Why: a 32x32-limb product on 32-bit limbs is a carry chain (IMAD.WIDE.U32.X), a form that
// issues at half the IMAD.WIDE rate and serialises on the carry predicate (measured on B200,
// profiles/imad_peak_r01.json: 8.55 T/s).  With 28-bit digits a digit product is < 2^56 and a
// 64-bit accumulator holds a whole column of BOTH products of x1*y1 +- x2*y2, so the work is
// plain signed IMAD.WIDE accumulations, one row of the x operand at a time (the x digit is the
// same register for the whole row: operand-reuse cache), with no carry chain at all and 37
// independent accumulators per row (tools/microbench_radix29.cu: 16.3 T/s in this form).
//
// What is computed (L = 32).  Term B is the one with the larger exponent, term A the other,
// d = eB - eA >= 0.  With X, Y the 1024-bit mantissas as integers:
//     XA' = floor(XA * 2^84 / 2^d)          (40 digits; 84 = 3 digits of guard below XB's frame)
//     XB' = XB * 2^84                       (the same formula with d = 0)
//     S'  = XB' * YB +- XA' * YA            (units of 2^-84 of the B product's frame)
// restricted to the digit products whose column is >= C0 = 36 (28-bit columns of S').  The exact
// value differs from the computed one by
//     dropped columns:  < (34 + 37) * 2^(56 + 28*35) * (1 + 2^-27)  < 2^1043   (units of S')
//     floor in XA', and the digitwise complement standing in for the negation: < 2 YA < 2^1025
//     two's complement negation of a negative total: <= 2^1008
// i.e. by less than 2^960 in the B frame (bits of XB*YB).  The result is accepted only if the top
// word of the sum is word 63 or 64 of that frame (at most 31 bits of cancellation), so the
// rounding position is at bit >= 993, and if the 24 bits just below the rounding position are
// neither all zero nor all one: then an error below 2^969 cannot carry or borrow into the kept
// bits, the truncated mantissa equals the exactly rounded one and the result is certainly
// inexact.  Everything else (zero or indeterminate operands, short operands whose products are
// exact, deep cancellation, the 2^-23 unlucky guard patterns) returns *ok = false and the caller
// falls back to the carry-chain / exact paths, so results stay bit-identical to the oracle.
#pragma once

namespace cb {

constexpr int RR_BITS = 28;
constexpr uint32_t RR_MASK = (1u << RR_BITS) - 1u;

// digit i (28 bits) of the little-endian word array w[0..NW)
template <int NW>
CB_HD uint32_t rr_digit(const uint32_t (&w)[NW], int i) {
    const int bit = RR_BITS * i, q = bit >> 5, r = bit & 31;
    uint32_t lo = q < NW ? w[q] : 0u, hi = (q + 1) < NW ? w[q + 1] : 0u;
    return (r + RR_BITS <= 32 ? (lo >> r) : shr_pair(lo, hi, r)) & RR_MASK;
}

template <int L, class XS = RRef, class YS = RRef, class DST = WRef>
CB_NOINLINE Meta arb_fdot2_rr(DST dst, OpdT<XS> x1, OpdT<YS> y1, OpdT<XS> x2, OpdT<YS> y2, bool neg2, bool* ok) {
    static_assert(L == 32, "reduced-radix fused dot product: 1024-bit mantissas");
    constexpr int YD = 37;        // digits of a y operand (1024 bits)
    constexpr int XD = 40;        // digits of a shifted x operand (1024 + 84 bits)
    constexpr int XW = 36;        // 32-bit words of a shifted x operand
    constexpr int C0 = 36;        // first kept column of S'
    constexpr int NC = 42;        // kept columns C0 .. C0 + NC - 1 (products reach column 75)
    constexpr int KW = 29;        // first B-frame word fully covered by the kept columns
    bool nan = is_nan(x1.t) || is_nan(y1.t) || is_nan(x2.t) || is_nan(y2.t);
    bool z1 = is_zero_mid(x1.t) || is_zero_mid(y1.t), z2 = is_zero_mid(x2.t) || is_zero_mid(y2.t);
    Mag rad = mul_rad(mag_zero(), x1.t, top_limb<L>(x1), y1.t, top_limb<L>(y1));
    rad = mul_rad(rad, x2.t, top_limb<L>(x2), y2.t, top_limb<L>(y2));
    const bool usable = !(nan || z1 || z2);
    bool n1 = ((x1.t.flags ^ y1.t.flags) & F_SIGN) != 0;
    bool n2 = (((x2.t.flags ^ y2.t.flags) & F_SIGN) != 0) != neg2;
    const int64_t e1 = (int64_t)x1.t.exp + y1.t.exp, e2 = (int64_t)x2.t.exp + y2.t.exp;
    const bool a_is_1 = e1 <= e2;
    const bool nA = a_is_1 ? n1 : n2, nB = a_is_1 ? n2 : n1;
    const int64_t eA = a_is_1 ? e1 : e2, eB = a_is_1 ? e2 : e1;
    const XS xA = a_is_1 ? x1.r : x2.r, xB = a_is_1 ? x2.r : x1.r;
    const YS yA = a_is_1 ? y1.r : y2.r, yB = a_is_1 ? y2.r : y1.r;
    const bool sub = nA != nB;
    const int64_t d = usable ? eB - eA : 0;

    // unsigned column sums modulo 2^64 (read as two's complement at the end)
    uint64_t acc[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) acc[c] = 0;
#pragma unroll 1
    for (int term = 0; term < 2; term++) {  // a loop, so that the row code exists once
        const XS xr = term ? xB : xA;
        const YS yr = term ? yB : yA;
        // XW words of floor(X * 2^96 / 2^D), D = d + 12 (term A) or 12 (term B)
        const int64_t D = (term ? 0 : d) + 12;
        const int mw = D > 32 * 40 ? 40 : (int)(D >> 5);
        const int s = (int)(D & 31);
        // a subtracted term enters as the digitwise complement 2^1120 - 1 - XA' (all digits stay
        // non-negative, so every product is a plain unsigned IMAD.WIDE); the surplus 2^1120 * YA is
        // taken out of columns 40.. below, the missing YA is part of the error bound
        const uint32_t cmask = (term == 0 && sub) ? RR_MASK : 0u;
        uint32_t yd[YD];
        {
            uint32_t w[L];
            load_limbs<L>(w, yr);
#pragma unroll
            for (int j = 0; j < YD; j++) yd[j] = rr_digit<L>(w, j);
        }
        uint32_t xw[XW];
        {
            auto ldx = [&](int k) -> uint32_t {
                int idx = k + mw - 3;
                return (unsigned)idx < (unsigned)L ? xr.ld(idx) : 0u;
            };
            uint32_t prev = ldx(0);
#pragma unroll
            for (int k = 0; k < XW; k++) {
                uint32_t nxt = ldx(k + 1);
                xw[k] = shr_pair(prev, nxt, s);
                prev = nxt;
            }
        }
#pragma unroll
        for (int i = 0; i < XD; i++) {
            const uint32_t a = rr_digit<XW>(xw, i) ^ cmask;
#pragma unroll
            for (int j = 0; j < YD; j++) {
                if (i + j < C0) continue;
                acc[i + j - C0] += (uint64_t)a * yd[j];
            }
        }
        if (cmask) {
#pragma unroll
            for (int j = 0; j < YD; j++)
                if (XD + j - C0 < NC) acc[XD + j - C0] -= yd[j];
        }
    }
    // carry-normalise the signed column sums into 28-bit digits (two's complement over the window)
    uint32_t dg[NC + 1];
    int64_t carry = 0;
#pragma unroll
    for (int c = 0; c < NC; c++) {
        int64_t t = (int64_t)acc[c] + carry;
        dg[c] = (uint32_t)t & RR_MASK;
        carry = t >> RR_BITS;  // arithmetic
    }
    dg[NC] = 0u;
    const bool negtot = carry < 0;  // |A| > |B| in a subtraction: negate
    bool good = usable && (carry == 0 || carry == -1);
    {
        const uint32_t mask = negtot ? RR_MASK : 0u;
        uint32_t cin = negtot ? 1u : 0u;
#pragma unroll
        for (int c = 0; c < NC; c++) {
            uint32_t v = (dg[c] ^ mask) + cin;
            dg[c] = v & RR_MASK;
            cin = v >> RR_BITS;
        }
    }
    // B-frame words KW .. 64 of the sum: word k = bits [32 k - 924, +32) of the kept digit string
    constexpr int NS = 66;
    uint32_t sw[NS - KW + 1];
#define SW_(k) sw[(k) - KW]
#pragma unroll
    for (int k = KW; k <= NS; k++) {
        const int p = 32 * k + 84 - RR_BITS * C0;
        const int q = p / RR_BITS, r = p % RR_BITS;
        uint64_t v = 0;
        if (q < NC) v |= (uint64_t)dg[q] >> r;
        if (q + 1 < NC) v |= (uint64_t)dg[q + 1] << (RR_BITS - r);
        if (q + 2 < NC && 2 * RR_BITS - r < 32) v |= (uint64_t)dg[q + 2] << (2 * RR_BITS - r);
        SW_(k) = (uint32_t)v;
    }
    int hi = 0;
    uint32_t top = 0u;
    if (SW_(64)) {
        hi = 64;
        top = SW_(64);
    } else if (SW_(63)) {
        hi = 63;
        top = SW_(63);
    } else {
        good = false;  // a limb or more cancelled
        hi = 63;
        top = 1u;
    }
    if (SW_(65) | SW_(66)) good = false;
    const int bl = 32 * hi + 32 - clz32(top);  // bit length of the sum in the B frame
    const int off = bl - 32 * L;                // rounding position: in [993, 1025]
    const int q = off >> 5, r = off & 31;       // q in {31, 32}
    {
        uint32_t lo_w = (q == 31) ? SW_(30) : SW_(31), hi_w = (q == 31) ? SW_(31) : SW_(32);
        uint32_t g = shr_pair(lo_w, hi_w, r) >> 8;  // bits [off - 24, off)
        if (g == 0u || g == 0xFFFFFFu) good = false;
    }
    if (good) {
#pragma unroll
        for (int i = 31; i < 31 + L + 1; i++) {
            uint32_t v = shr_pair(SW_(i), SW_(i + 1), r);
            int j = i - q;
            if ((unsigned)j < (unsigned)L) dst.st(j, v);
        }
    }
#undef SW_
    *ok = good;
    Rounded rr;
    int64_t e = eB - 64 * (int64_t)L + bl;
    rr.overflow = (e > EXP_LIMIT || e < -EXP_LIMIT);
    rr.exp = (int32_t)e;
    rr.zero = false;
    rr.inexact = true;
    const bool neg = sub ? (negtot ? nA : nB) : nB;
    if (!good) return meta_zero();
    return finish_to<L>(dst, rr, neg, rad);
}

}  // namespace cb
