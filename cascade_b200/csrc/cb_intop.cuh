// cb_intop.cuh -- register-resident fast path of the integer-operator Fuchs recurrence (BASELINE configs[3]:
// Legendre sweeps, reference src/fuchs_solver.c:96-145 on the operators of src/examples.c:3-11).
//
// fuchs_intop_thread (cb_solvers.cuh) forms every coefficient through shared-memory columns with dynamically indexed
// normalisation loops: about 4 500 instructions per coefficient for ~200 limb products, which kept the sweep at 13 %
// of the HBM write bandwidth that bounds it.  Here one coefficient of one instance is formed entirely in registers
// with statically indexed code:
//   p   = round(x * k)      x * (k normalised to 64 bits): 2 L IMAD.WIDE in carry chains; the top bit of the product
//                           is in one of two known places, so rounding is one funnel shift + select per limb
//   acc = round(acc - p)    fast_add_round on register operands (operands within 31 bits, < 1 limb cancelled)
//   c   = round(acc / d)    quotient digits by Moeller-Granlund 3-by-2 steps in registers; the quotient of two
//                           normalised operands needs a shift of 0 or 1 bit
// Radii follow arb_mul_si / arb_addsub / arb_div_u64 step by step; results are BIT-IDENTICAL to
// fuchs_intop_thread.  Whatever the fast path does not cover -- indeterminate or zero-midpoint balls with a radius,
// exponent gaps beyond 31 bits, cancellation of a limb or more, exponent overflow -- makes it return false before
// anything is stored, and the caller runs that one coefficient through fuchs_intop_thread.
#pragma once

namespace cb {

// P[0 .. L+2) = x[0 .. L) * k (k != 0 any 64-bit value), exact
template <int L>
CB_HD void mul_by_u64(uint32_t (&P)[L + 2], const uint32_t (&x)[L], uint64_t k) {
    const uint32_t k0 = (uint32_t)k, k1 = (uint32_t)(k >> 32);
#if defined(__CUDA_ARCH__)
    static_assert(L % 2 == 0, "even limb count");
    // even/odd column accumulators as in mul_full: row k0 then row k1, each two carry chains
    uint32_t ev[L + 4], od[L + 4];
#pragma unroll
    for (int i = 0; i < L + 4; i++) { ev[i] = 0u; od[i] = 0u; }
    mul_chain<L, 0, 0, true>(ev, x, k0);   // x[even] * k0 -> columns (j, j+1), j even
    mul_chain<L, 0, 1, true>(od, x, k0);   // x[odd]  * k0 -> columns (j+1, j+2): od[i] is column i+1
    mul_chain<L, 0, 0, true>(od, x, k1);   // x[even] * k1 -> columns (j+1, j+2)
    mul_chain<L, 2, 1, true>(ev, x, k1);   // x[odd]  * k1 -> columns (j+2, j+3)
    P[0] = ev[0];
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(P[1]) : "r"(ev[1]), "r"(od[0]));
#pragma unroll
    for (int i = 2; i < L + 2; i++) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(P[i]) : "r"(ev[i]), "r"(od[i - 1]));
#else
    uint64_t c = 0;
    uint32_t lo[L + 1];
    for (int j = 0; j < L; j++) {
        uint64_t t = (uint64_t)x[j] * k0 + c;
        lo[j] = (uint32_t)t;
        c = t >> 32;
    }
    lo[L] = (uint32_t)c;
    P[0] = lo[0];
    c = 0;
    for (int j = 0; j < L; j++) {
        uint64_t t = (uint64_t)x[j] * k1 + lo[j + 1] + c;
        P[j + 1] = (uint32_t)t;
        c = t >> 32;
    }
    P[L + 1] = (uint32_t)c;
#endif
}

// out = round(x * ka), x finite with a nonzero midpoint, ka != 0 (arb_mul_si_impl + extract + finish, restated for
// registers).  Returns false (nothing defined) when the exponent leaves the supported range.
template <int L>
CB_HD bool reg_mul_u64(uint32_t (&out)[L], Meta& to, const uint32_t (&x)[L], const Meta& tx, uint64_t ka, bool kneg) {
    const int sh = clz64(ka);
    const uint64_t kn = ka << sh;  // top bit at 63
    Meta kt{64 - sh, 0u, 0u, 0};
    const Mag rad = mag_mul(mag_mid_upper(kt, (uint32_t)(kn >> 32)), mag_of(tx));
    uint32_t P[L + 2];
    mul_by_u64<L>(P, x, kn);
    // x >= 2^(32L-1), kn >= 2^63: the product has 32(L+2) or 32(L+2)-1 bits
    const bool full = (P[L + 1] >> 31) != 0u;
    const uint32_t low = P[0] | (full ? P[1] : (P[1] & 0x7FFFFFFFu));
#pragma unroll
    for (int j = 0; j < L; j++) {
        uint32_t hi = (j + 2 <= L + 1) ? P[j + 2] : 0u;
        uint32_t a = shr_pair(P[j + 1], hi, 31);
        out[j] = full ? hi : a;
    }
    const int64_t e = (int64_t)tx.exp + 64 - sh - (full ? 0 : 1);
    if (e > EXP_LIMIT || e < -EXP_LIMIT) return false;
    Rounded rr;
    rr.exp = (int32_t)e;
    rr.zero = false;
    rr.inexact = low != 0u;
    rr.overflow = false;
    to = finish<L>(rr, ((tx.flags & F_SIGN) != 0) != kneg, rad);
    return true;
}

// a register destination for fast_add_round: `st_static<I>(c, v)` is called with the STATIC loop index I (relative to
// QMIN) and the runtime offset class c in [0, 4): limb I - c receives v -- four predicated moves, no dynamic register
// indexing (see the REG_DST branch of fast_add_round's store loop)
template <int L>
struct RegDst {
    uint32_t* out;
    static constexpr bool REG_DST = true;
    template <int I>
    CB_HD void st_static(int c, uint32_t v) const {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            constexpr int dummy = 0;
            (void)dummy;
            const int j = I - k;
            if (j >= 0 && j < L) out[j] = (c == k) ? v : out[j];
        }
    }
    CB_HD void st(int, uint32_t) const {}
};

// acc = round(acc -+ p) on register operands (the fast path of arb_addsub / arb_addsub_cols); false: not the easy case
template <int L>
CB_HD bool reg_addsub(uint32_t (&acc)[L], Meta& ta, const uint32_t (&p)[L], const Meta& tp, bool negp) {
    const bool na = (ta.flags & F_SIGN) != 0, nb = ((tp.flags & F_SIGN) != 0) != negp;
    const int64_t d = (int64_t)ta.exp - tp.exp;
    if (d > 31 || d < -31) return false;
    const int64_t E = ta.exp > tp.exp ? ta.exp : tp.exp;
    const int s1 = (int)(E - ta.exp), s2 = (int)(E - tp.exp);
    const bool sub = na != nb;
    uint32_t res[L];
#pragma unroll
    for (int j = 0; j < L; j++) res[j] = 0u;
    RegDst<L> dst{res};
    Rounded rr;
    bool neg;
    auto w1 = [&](int k) -> uint32_t { return acc[k]; };
    auto w2 = [&](int k) -> uint32_t { return p[k]; };
    if (!fast_add_round<L, L>(dst, w1, w2, s1, s2, sub, na, nb, E, &rr, &neg)) return false;
    if (rr.overflow) return false;
    ta = finish<L>(rr, neg, mag_add(mag_of(ta), mag_of(tp)));
#pragma unroll
    for (int j = 0; j < L; j++) acc[j] = res[j];
    return true;
}
template <int L>
CB_HD bool reg_sub(uint32_t (&acc)[L], Meta& ta, const uint32_t (&p)[L], const Meta& tp) {
    return reg_addsub<L>(acc, ta, p, tp, true);
}

// what arb_div_u64 derives from the divisor alone: formed once per coefficient and shared by both parts
struct DivSetup {
    uint64_t D;      // d << sh, top bit set
    uint32_t d1;     // its top limb
    uint32_t recip;  // 2-by-1 reciprocal of d1 (D's low limb zero) or 3-by-2 reciprocal of D
    Meta yt;         // exponent and sign of the divisor as a ball
    Mag yl, num_y;   // lower bound of |d| and upper bound of |d| (30-bit magnitudes)
    bool one_limb;
};
CB_HD DivSetup div_setup(uint64_t d, bool dneg) {
    DivSetup s;
    const int sh = clz64(d);
    s.D = d << sh;
    s.d1 = (uint32_t)(s.D >> 32);
    s.yt = Meta{64 - sh, dneg ? F_SIGN : 0u, 0u, 0};
    s.yl = mag_mid_lower(s.yt, s.d1);
    s.num_y = mag_mid_upper(s.yt, s.d1);
    s.one_limb = (uint32_t)s.D == 0u;
    // one-limb divisor (d < 2^32, e.g. the b(b-1) of a Legendre sweep): floor(x * 2^64 / D) = floor(x * 2^32 / d1),
    // digit for digit the same quotient, by Moeller-Granlund 2-by-1 steps (half the work of a 3-by-2 step)
    s.recip = s.one_limb ? (uint32_t)(0xFFFFFFFFFFFFFFFFull / s.d1 - 0x100000000ull) : recip_3by2(s.d1, (uint32_t)s.D);
    return s;
}

// out = round(x / d), x finite with a nonzero midpoint, d != 0 (arb_div_u64 restated for registers)
template <int L>
CB_HD bool reg_div_u64(uint32_t (&out)[L], Meta& to, const uint32_t (&x)[L], const Meta& tx, const DivSetup& ds) {
    Mag rad = mag_zero();
    if (tx.rman != 0) rad = mag_div(mag_mul(ds.num_y, mag_of(tx)), mag_mul_lower(ds.yl, ds.yl));
    uint32_t q[L + 1];
    bool rem_nz;
    if (ds.one_limb) {
        const uint32_t dinv = ds.recip, d1 = ds.d1;
        uint32_t r = 0u;
#pragma unroll
        for (int j = L; j >= 0; j--) {
            const uint32_t u = (j >= 1) ? x[j - 1] : 0u;
            const uint64_t qq = (uint64_t)dinv * r + (((uint64_t)r << 32) | u);
            uint32_t q1 = (uint32_t)(qq >> 32) + 1u;
            const uint32_t q0 = (uint32_t)qq;
            uint32_t rr2 = u - q1 * d1;
            if (rr2 > q0) {
                q1--;
                rr2 += d1;
            }
            if (rr2 >= d1) {
                q1++;
                rr2 -= d1;
            }
            q[j] = q1;
            r = rr2;
        }
        rem_nz = r != 0u;
    } else {
        uint64_t rem = 0;
        (void)div_3by2(rem, x[L - 1], ds.D, ds.recip);
#pragma unroll
        for (int j = L; j >= 0; j--) q[j] = div_3by2(rem, (j >= 2) ? x[j - 2] : 0u, ds.D, ds.recip);
        rem_nz = rem != 0;
    }
    // x in [2^(32L-1), 2^(32L)), D in [2^63, 2^64): the quotient has 32L or 32L+1 bits
    const bool one = q[L] != 0u;
    uint32_t low = rem_nz ? 1u : 0u;
    if (one) low |= q[0] & 1u;
#pragma unroll
    for (int j = 0; j < L; j++) out[j] = one ? shr_pair(q[j], q[j + 1], 1) : q[j];
    const int64_t e = (int64_t)tx.exp - ds.yt.exp + (one ? 1 : 0);
    if (e > EXP_LIMIT || e < -EXP_LIMIT) return false;
    Rounded rr;
    rr.exp = (int32_t)e;
    rr.zero = false;
    rr.inexact = low != 0u;
    rr.overflow = false;
    to = finish<L>(rr, ((tx.flags ^ ds.yt.flags) & F_SIGN) != 0, rad);
    return true;
}

// multipliers of one coefficient for order = degree = 2 when e - bmin = degree (e >= 2): cur[k], k = 0 .. cnt, with
//   imin = max(k - 2, 0), imax = min(k, 2), column index i + 2 - k -- all compile-time constants (src/fuchs_solver.c:121-135)
CB_HD void intop_multipliers_22(int64_t* cur, const int64_t* P, int64_t bmin, int cnt) {
    const int64_t p02 = P[2], p01 = P[1], p00 = P[0], p12 = P[5], p11 = P[4], p10 = P[3], p22 = P[8], p21 = P[7], p20 = P[6];
    cur[0] = p02;
    if (cnt >= 1) {  // k = 1: i in [0, 1], columns 1, 2
        const int64_t b = bmin + 1;
        cur[1] = p01 + p12 * b;
    }
    if (cnt >= 2) {  // k = 2: i in [0, 2], columns 0, 1, 2
        const int64_t b = bmin + 2;
        cur[2] = p00 + p11 * b + p22 * (b * (b - 1));
    }
    if (cnt >= 3) {  // k = 3: imin = 1: i in [1, 2], columns 0, 1; common factor b
        const int64_t b = bmin + 3;
        cur[3] = (p10 + p21 * (b - 1)) * b;
    }
    if (cnt >= 4) {  // k = 4: imin = 2: i = 2, column 0; common factor (b - 1) b
        const int64_t b = bmin + 4;
        cur[4] = p20 * ((b - 1) * b);
    }
}

// One coefficient of one instance in registers.  Returns false when a step is outside the fast path; the caller then
// forms this coefficient (both parts) with fuchs_intop_thread.
constexpr int INTOP_MAXWIN = 12;  // multipliers per coefficient: degree - valuation + 1 <= 12 on the fast path
template <int L>
CB_HD bool fuchs_intop_fast_coeff(const FuchsIntParams& p, int64_t inst, const int64_t* P, int64_t bmax) {
    const size_t slot = 2 * (size_t)inst;
    const int v = p.valuation, D1 = p.degree + 1;
    const size_t S = p.res.S;
    const int64_t e = bmax + v;
    const int64_t bmin = clampi(e - p.degree, 0, bmax);
    const int cnt = (int)(bmax - bmin);
    if (cnt + 1 > INTOP_MAXWIN) return false;
    // the multipliers of res[bmin .. bmax-1] and the divisor, exactly as the reference forms temp2 (all integers;
    // reference src/fuchs_solver.c:114,121-135)
    int64_t cur[INTOP_MAXWIN];
    if (p.order == 2 && p.degree == 2 && e >= 2 && cnt <= 4) {
        // the shipped example operators (reference src/examples.c) in the steady state (e - bmin = degree): every index
        // below is a compile-time constant, so the nine operator words are loaded up front and the multipliers are a
        // few integer multiply-adds (same integers as the general loop in the else branch)
        intop_multipliers_22(cur, P, bmin, cnt);
    } else {
        cur[0] = P[0 * D1 + (e - bmin)];
        for (int k = 1; k <= cnt; k++) {
            const int64_t b = bmin + k;
            int64_t t2 = 0, fac = 1;
            const int64_t imin = clampi(b - e, 0, -v);
            const int64_t imax = clampi(b - bmin, 0, p.order);
            for (int64_t i = imin; i <= imax; i++) {
                t2 += P[i * D1 + (i + e - b)] * fac;
                fac *= (b - i);
            }
            fac = 1;
            for (int64_t q = 0; q < imin; q++) fac *= (b - imin + 1 + q);
            cur[k] = t2 * fac;
        }
    }
    const int64_t dv = cur[cnt];
    if (dv == 0) return false;  // division by zero: indeterminate, the general path writes it
    uint32_t mask = 0u;         // terms with a nonzero multiplier (x * 0 is an exact zero and acc - 0 leaves acc untouched)
    for (int k = 0; k < cnt; k++) mask |= (cur[k] != 0) ? (1u << k) : 0u;
    // the loads of the first term (both meta records and the limbs of its real part) are issued BEFORE the divisor
    // is prepared (a 64-bit division), so that their latency is covered
    const int k0 = mask ? (int)(31 - clz32(mask & (0u - mask))) : 0;
    Meta t0re, t0im;
    uint32_t x[L];
    {
        const size_t b0 = (size_t)(bmin + k0);
        t0re = ld_meta_user(p.res.t + b0 * S + slot);
        t0im = ld_meta_user(p.res.t + b0 * S + slot + 1);
        const uint32_t* q = p.res.m + b0 * L * S + slot;
#pragma unroll
        for (int j = 0; j < L; j++) {
            x[j] = *q;
            q += S;
#if defined(__CUDA_ARCH__)
            asm volatile("" : "+l"(q));
#endif
        }
    }
    const DivSetup ds = div_setup(dv < 0 ? (uint64_t)0 - (uint64_t)dv : (uint64_t)dv, dv < 0);
    // a real sweep (Legendre with real initial values): one term whose imaginary part is an exact zero -- the
    // imaginary part of the coefficient is an exact zero too and goes out with the real part, a 64-bit (re, im) pair
    // per limb (re and im words of a limb are adjacent: slot = 2 * instance + part)
    const bool im_zero = (mask & (mask - 1u)) == 0u && is_exact_zero(t0im);
#pragma unroll 1
    for (int part = 0; part < (im_zero ? 1 : 2); part++) {
        uint32_t acc[L];
        Meta tacc = meta_zero();
        bool have = false;
        uint32_t todo = mask;
#pragma unroll 1
        while (todo) {
            const int k = (int)(31 - clz32(todo & (0u - todo)));
            todo &= todo - 1u;
            const int64_t c = cur[k];
            const int64_t b = bmin + k;
            const Meta tx = (k == k0) ? (part ? t0im : t0re) : ld_meta_user(p.res.t + (size_t)b * S + slot + part);
            if (is_exact_zero(tx)) continue;
            if (is_nan(tx) || is_zero_mid(tx)) return false;
            if (!(k == k0 && part == 0)) {  // the first term of the real part is already in x
                const uint32_t* q = p.res.m + (size_t)b * L * S + slot + part;
#pragma unroll
                for (int j = 0; j < L; j++) {
                    x[j] = *q;
                    q += S;
#if defined(__CUDA_ARCH__)
                    asm volatile("" : "+l"(q));
#endif
                }
            }
            uint32_t pr[L];
            Meta tp;
            const uint64_t ka = c < 0 ? (uint64_t)0 - (uint64_t)c : (uint64_t)c;
            if (!reg_mul_u64<L>(pr, tp, x, tx, ka, c < 0)) return false;
            if (!have) {  // 0 - p = -p exactly
#pragma unroll
                for (int j = 0; j < L; j++) acc[j] = pr[j];
                tacc = tp;
                tacc.flags ^= F_SIGN;
                have = true;
            } else if (!reg_sub<L>(acc, tacc, pr, tp) || is_zero_mid(tacc)) {
                return false;
            }
        }
        // this part of the coefficient: c = acc / divisor.  (A part stored here stays valid if the other part falls
        // back: the general path recomputes both from earlier coefficients only and writes the same words.)
        uint32_t o[L];
        Meta to;
        if (!have) {
#pragma unroll
            for (int j = 0; j < L; j++) o[j] = 0u;
            to = meta_zero();
        } else if (!reg_div_u64<L>(o, to, acc, tacc, ds)) {
            return false;
        }
        uint32_t* om = p.res.m + (size_t)bmax * L * S + slot + part;
        if (im_zero) {
#pragma unroll
            for (int j = 0; j < L; j++) {
#if defined(__CUDA_ARCH__)
                *reinterpret_cast<uint2*>(om) = make_uint2(o[j], 0u);
                asm volatile("" : "+l"(om));
#else
                om[0] = o[j];
                om[1] = 0u;
#endif
                om += S;
            }
            st_meta(p.res.t + (size_t)bmax * S + slot, to);
            st_meta(p.res.t + (size_t)bmax * S + slot + 1, meta_zero());
        } else {
#pragma unroll
            for (int j = 0; j < L; j++) {
                *om = o[j];
                om += S;
#if defined(__CUDA_ARCH__)
                asm volatile("" : "+l"(om));
#endif
            }
            st_meta(p.res.t + (size_t)bmax * S + slot + part, to);
        }
    }
    return true;
}

// The recurrence of one instance over coefficients [b_first, b_last]: the fast path per coefficient, the general
// shared-memory path (fuchs_intop_thread on that single coefficient: it resumes from res) where it does not apply.
template <int L, class S0T, class SCR>
CB_HD void fuchs_intop_fast_thread(const FuchsIntParams& p, int64_t inst, S0T S0, SCR A0, SCR A1, SCR A2,
                                   int64_t n_from = -1, int64_t n_to = -1) {
    const int ne = (p.order + 1) * (p.degree + 1);
    int64_t P[INTOP_MAXENT];
    uint64_t pmax = 0;
    for (int e = 0; e < ne; e++) {
        P[e] = p.ode[(size_t)e * (p.ostride ? p.ostride : 1) + (p.ostride ? (size_t)inst : 0)];
        uint64_t a = P[e] < 0 ? (uint64_t)0 - (uint64_t)P[e] : (uint64_t)P[e];
        pmax = a > pmax ? a : pmax;
    }
    const int64_t b_first = n_from >= 0 ? n_from : -p.valuation, b_last = n_from >= 0 ? n_to : p.n;
    double bound = (double)pmax * (double)(p.order + 1);
    for (int i = 0; i < p.order; i++) bound *= (double)(p.n + 1);
    if (!(bound < 4.0e18)) {  // the general body writes the indeterminate balls of an operator beyond the int64 bound
        fuchs_intop_thread<L>(p, inst, S0, A0, A1, A2, b_first, b_last);
        return;
    }
#pragma unroll 1
    for (int64_t bmax = b_first; bmax <= b_last; bmax++)
        if (!fuchs_intop_fast_coeff<L>(p, inst, P, bmax)) fuchs_intop_thread<L>(p, inst, S0, A0, A1, A2, bmax, bmax);
}

// ================================================================ per-instance multiplier tables in registers
// fuchs_table_thread (cb_solvers.cuh) forms the multipliers  sum_i P[i][i+e-b] * fac_i  of one table row through
// acb_mul_si / acb_add on temporaries in global memory (reference src/fuchs_solver.c:121-135): every step waits for
// the stores of the one before it (ncu: long_scoreboard 7.3 cycles per issue, profiles/
// ncu_fuchs_table32_pi_r02_before_summary.txt).  The real and the imaginary part of that sum never mix, so here each
// part of each multiplier is a short chain of register operations -- reg_mul_u64 (ball x integer), reg_addsub -- and
// goes to the table once, with its trailing-zero hint.  Same operations in the same order: bit-identical.  A step
// outside the fast path (an indeterminate or zero-midpoint-with-radius operand, an exponent gap beyond 31 bits,
// cancellation of a limb or more) returns false and the caller recomputes the row with fuchs_table_thread.
template <int L>
CB_HD bool fuchs_table_fast_multipliers(const FuchsTableParams& p, int64_t row, size_t oslot, size_t qslot, int64_t base,
                                        int* cnt_out) {
    const int v = p.valuation, D1 = p.degree + 1;
    const int64_t bmax = row - v, e = bmax + v;
    const int64_t bmin = clampi(e - p.degree, 0, bmax);
    const int cnt = (int)(bmax - bmin);
    *cnt_out = cnt;
    const size_t So = p.ode.S, St = p.tbl.S;
#pragma unroll 1
    for (int k = 0; k <= cnt; k++) {
        const int64_t b = bmin + k;
        const int64_t imin = k == 0 ? 0 : clampi(b - e, 0, -v), imax = k == 0 ? 0 : clampi(b - bmin, 0, p.order);
        int64_t fac2 = 1;
        for (int64_t q = 0; q < imin; q++) fac2 *= (b - imin + 1 + q);
#pragma unroll 1
        for (int part = 0; part < 2; part++) {
            uint32_t acc[L];
            Meta tacc = meta_zero();
            bool have = false;
            int64_t fac = 1;
#pragma unroll 1
            for (int64_t i = imin; i <= imax; i++) {
                const int64_t ent = i * D1 + (k == 0 ? (e - bmin) : (i + e - b));
                const int64_t f = fac;
                fac *= (b - i);
                const Meta tx = ld_meta_user(p.ode.t + (size_t)ent * So + oslot + part);
                if (is_exact_zero(tx) || f == 0) continue;  // an exact zero term: t2 + 0 leaves t2 untouched
                if (is_nan(tx) || is_zero_mid(tx)) return false;
                uint32_t x[L];
                {
                    const uint32_t* q = p.ode.m + (size_t)ent * L * So + oslot + part;
#pragma unroll
                    for (int j = 0; j < L; j++) {
                        x[j] = *q;
                        q += So;
#if defined(__CUDA_ARCH__)
                        asm volatile("" : "+l"(q));
#endif
                    }
                }
                if (k == 0) {  // slot 0 is a plain copy of P[0][e - bmin]
#pragma unroll
                    for (int j = 0; j < L; j++) acc[j] = x[j];
                    tacc = tx;
                    have = true;
                    continue;
                }
                uint32_t pr[L];
                Meta tp;
                const uint64_t ka = f < 0 ? (uint64_t)0 - (uint64_t)f : (uint64_t)f;
                if (!reg_mul_u64<L>(pr, tp, x, tx, ka, f < 0)) return false;
                if (!have) {  // 0 + t1 = t1 exactly
#pragma unroll
                    for (int j = 0; j < L; j++) acc[j] = pr[j];
                    tacc = tp;
                    have = true;
                } else if (!reg_addsub<L>(acc, tacc, pr, tp, false) || is_zero_mid(tacc)) {
                    return false;
                }
            }
            if (have && k > 0) {  // t2 *= fac2 (acb_mul_si runs even for a factor 1: the radius bound is |k| rounded up)
                uint32_t pr[L];
                Meta tp;
                if (!reg_mul_u64<L>(pr, tp, acc, tacc, (uint64_t)fac2, false)) return false;
#pragma unroll
                for (int j = 0; j < L; j++) acc[j] = pr[j];
                tacc = tp;
            }
            // store this part of slot k with its trailing-zero hint (acb_mark_tz)
            uint32_t* om = p.tbl.m + (size_t)(base + k) * L * St + qslot + part;
            if (!have) {
#pragma unroll
                for (int j = 0; j < L; j++) {
                    *om = 0u;
                    om += St;
#if defined(__CUDA_ARCH__)
                    asm volatile("" : "+l"(om));
#endif
                }
                st_meta(p.tbl.t + (size_t)(base + k) * St + qslot + part, meta_zero());
            } else {
                int tz = L - 1;
#pragma unroll
                for (int j = L - 2; j >= 0; j--) tz = acc[j] ? j : tz;
#pragma unroll
                for (int j = 0; j < L; j++) {
                    *om = acc[j];
                    om += St;
#if defined(__CUDA_ARCH__)
                    asm volatile("" : "+l"(om));
#endif
                }
                tacc.flags = (tacc.flags & 0xFFu) | ((uint32_t)tz << F_TZ_SHIFT);
                st_meta(p.tbl.t + (size_t)(base + k) * St + qslot + part, tacc);
            }
        }
    }
    return true;
}

}  // namespace cb
