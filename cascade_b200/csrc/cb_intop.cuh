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
// re and im words of a limb are adjacent in memory (slot = 2 * instance + part), so limbs are loaded and stored as
// 64-bit pairs.  Radii follow arb_mul_si / arb_addsub / arb_div_u64 step by step; results are BIT-IDENTICAL to
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

// acc = round(acc - p) on register operands (arb_addsub_cols's fast path); false: not the easy case
template <int L>
CB_HD bool reg_sub(uint32_t (&acc)[L], Meta& ta, const uint32_t (&p)[L], const Meta& tp) {
    const bool na = (ta.flags & F_SIGN) != 0, nb = (tp.flags & F_SIGN) == 0;  // acc + (-p)
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

// out = round(x / d), x finite with a nonzero midpoint, d != 0 (arb_div_u64 restated for registers)
template <int L>
CB_HD bool reg_div_u64(uint32_t (&out)[L], Meta& to, const uint32_t (&x)[L], const Meta& tx, uint64_t d, bool dneg) {
    const int sh = clz64(d);
    const uint64_t D = d << sh;
    const uint32_t d1 = (uint32_t)(D >> 32);
    Meta yt{64 - sh, dneg ? F_SIGN : 0u, 0u, 0};
    const Mag yl = mag_mid_lower(yt, d1);
    Mag rad = mag_zero();
    if (tx.rman != 0) rad = mag_div(mag_mul(mag_mid_upper(yt, d1), mag_of(tx)), mag_mul_lower(yl, yl));
    const uint32_t v = recip_3by2(d1, (uint32_t)D);
    uint64_t rem = 0;
    (void)div_3by2(rem, x[L - 1], D, v);
    uint32_t q[L + 1];
#pragma unroll
    for (int j = L; j >= 0; j--) q[j] = div_3by2(rem, (j >= 2) ? x[j - 2] : 0u, D, v);
    // x in [2^(32L-1), 2^(32L)), D in [2^63, 2^64): the quotient has 32L or 32L+1 bits
    const bool one = q[L] != 0u;
    uint32_t low = (rem != 0) ? 1u : 0u;
    if (one) low |= q[0] & 1u;
#pragma unroll
    for (int j = 0; j < L; j++) out[j] = one ? shr_pair(q[j], q[j + 1], 1) : q[j];
    const int64_t e = (int64_t)tx.exp - yt.exp + (one ? 1 : 0);
    if (e > EXP_LIMIT || e < -EXP_LIMIT) return false;
    Rounded rr;
    rr.exp = (int32_t)e;
    rr.zero = false;
    rr.inexact = low != 0u;
    rr.overflow = false;
    to = finish<L>(rr, ((tx.flags & F_SIGN) != 0) != dneg, rad);
    return true;
}

// One coefficient of one instance in registers.  Returns false -- with NOTHING stored -- when a step is outside the
// fast path; the caller then forms this coefficient with fuchs_intop_thread.
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
    const int64_t dv = cur[cnt];
    if (dv == 0) return false;  // division by zero: indeterminate, the general path writes it
    const uint64_t dmag = dv < 0 ? (uint64_t)0 - (uint64_t)dv : (uint64_t)dv;
    uint32_t o[2][L];  // only o[part] of the part in flight is live at a time, see the stores below
    Meta to[2];
    bool okay = true;
#pragma unroll 1
    for (int part = 0; part < 2 && okay; part++) {
        uint32_t acc[L];
        Meta tacc = meta_zero();
        bool have = false;
#pragma unroll 1
        for (int k = 0; k < cnt && okay; k++) {
            const int64_t c = cur[k];
            if (c == 0) continue;  // x * 0 is an exact zero and acc - 0 leaves acc untouched
            const int64_t b = bmin + k;
            const Meta tx = ld_meta_user(p.res.t + (size_t)b * S + slot + part);
            if (is_exact_zero(tx)) continue;
            if (is_nan(tx) || is_zero_mid(tx)) {
                okay = false;
                break;
            }
            uint32_t x[L], pr[L];
            {
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
            Meta tp;
            const uint64_t ka = c < 0 ? (uint64_t)0 - (uint64_t)c : (uint64_t)c;
            if (!reg_mul_u64<L>(pr, tp, x, tx, ka, c < 0)) {
                okay = false;
                break;
            }
            if (!have) {  // 0 - p = -p exactly
#pragma unroll
                for (int j = 0; j < L; j++) acc[j] = pr[j];
                tacc = tp;
                tacc.flags ^= F_SIGN;
                have = true;
            } else if (!reg_sub<L>(acc, tacc, pr, tp) || is_zero_mid(tacc)) {
                okay = false;
                break;
            }
        }
        if (!okay) break;
        if (!have) {
#pragma unroll
            for (int j = 0; j < L; j++) o[part][j] = 0u;
            to[part] = meta_zero();
        } else if (!reg_div_u64<L>(o[part], to[part], acc, tacc, dmag, dv < 0)) {
            okay = false;
        }
    }
    if (!okay) return false;
    // re and im words of a limb are adjacent (slot = 2 * instance + part): one 64-bit store per limb
    uint32_t* om = p.res.m + (size_t)bmax * L * S + slot;
#pragma unroll
    for (int j = 0; j < L; j++) {
#if defined(__CUDA_ARCH__)
        *reinterpret_cast<uint2*>(om) = make_uint2(o[0][j], o[1][j]);
        om += S;
        asm volatile("" : "+l"(om));
#else
        om[0] = o[0][j];
        om[1] = o[1][j];
        om += S;
#endif
    }
    Meta* ot = p.res.t + (size_t)bmax * S + slot;
    st_meta(ot, to[0]);
    st_meta(ot + 1, to[1]);
    return true;
}

// The recurrence of one instance over coefficients [b_first, b_last]: the fast path per coefficient, the general
// shared-memory path (fuchs_intop_thread on that single coefficient: it resumes from res) where it does not apply.
template <int L, class SCR>
CB_HD void fuchs_intop_fast_thread(const FuchsIntParams& p, int64_t inst, SCR S0, SCR A0, SCR A1, SCR A2,
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

}  // namespace cb
