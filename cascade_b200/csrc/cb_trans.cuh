// cb_trans.cuh -- the transcendental part of acb_ode_solution_evaluate on the device
// (reference src/acb_ode_solution.c:24-58): acb_log(a), acb_pow(a, rho) and the Horner-in-log(a)
// combination of the M series of a solution, one thread per evaluation point.
//
// Arb's acb_log / acb_exp are not correctly rounded, so there is no unique ball to reproduce; the
// algorithms are the ones stated in oracle/cbo_trans.c, built from the ball primitives of
// cb_arith.cuh in the SAME order, so device and oracle agree bit for bit:
//   exp(z): w = z / 2^k, |w| < 2^-16; Taylor polynomial by Horner; remainder into the radius; k squarings
//   log(a): y0 = double-precision seed from IEEE +,-,*,/ only (round-to-nearest intrinsics, so no FMA
//           contraction: identical bits on CPU and GPU); t = a exp(-y0); u = t - 1; v = u / (u + 2);
//           log(a) = y0 + 2 atanh(v) by its Taylor series
//   pow:    rho = 0 -> 1; exact integer 0 <= rho < 2^31 -> binary exponentiation; else exp(rho log a)
#pragma once
#include <string.h>

namespace cb {

#if defined(__CUDA_ARCH__)
#define CB_DMUL(a, b) __dmul_rn((a), (b))
#define CB_DADD(a, b) __dadd_rn((a), (b))
#define CB_DSUB(a, b) __dsub_rn((a), (b))
#define CB_DDIV(a, b) __ddiv_rn((a), (b))
#else  // host build (tests/host_sim): compiled with -ffp-contract=off
#define CB_DMUL(a, b) ((a) * (b))
#define CB_DADD(a, b) ((a) + (b))
#define CB_DSUB(a, b) ((a) - (b))
#define CB_DDIV(a, b) ((a) / (b))
#endif

constexpr int64_t TR_BIG = ((int64_t)1) << 40;
constexpr bool EVAL_SYNC = true;  // phase barriers in evaluate_thread's Horner loops (uniform trip counts)
constexpr int TRANS_NTMP = 16;  // complex temporaries of acb_exp (4) + acb_log (10 + the negated argument) + evaluate (1)

template <int L>
CB_HD void acb_set_nan(const CRef& z) {
    store_nan<L>(wre(z));
    store_nan<L>(wim(z));
    st_meta(z.t, meta_nan());
    st_meta(z.t + 1, meta_nan());
}
// E with |z| < 2^E; TR_BIG for an indeterminate ball, -TR_BIG for the exact zero
CB_HD int64_t acb_bound_exp(const CRef& z) {
    int64_t e = -TR_BIG;
    for (int part = 0; part < 2; part++) {
        Meta t = ld_meta(z.t + part);
        if (is_nan(t)) return TR_BIG;
        if (!is_zero_mid(t) && t.exp > e) e = t.exp;
        if (t.rman && t.rexp > e) e = t.rexp;
    }
    return e == -TR_BIG ? -TR_BIG : e + 2;
}
CB_HD void acb_mul_2exp(const CRef& z, int64_t k) {
    for (int part = 0; part < 2; part++) {
        Meta t = ld_meta(z.t + part);
        if (is_nan(t)) continue;
        if (!is_zero_mid(t)) t.exp = (int32_t)(t.exp + k);
        if (t.rman) t.rexp = (int32_t)(t.rexp + k);
        st_meta(z.t + part, t);
    }
}
CB_HD void acb_add_error_2exp(const CRef& z, int64_t e) {
    for (int part = 0; part < 2; part++) {
        Meta t = ld_meta(z.t + part);
        if (is_nan(t)) continue;
        Mag r = mag_add(mag_of(t), Mag{1u << 29, (int32_t)(e + 1)});
        t.rman = r.man;
        t.rexp = r.man ? r.exp : 0;
        st_meta(z.t + part, t);
    }
}
CB_HD int tr_floor_log2(int64_t n) {
    int r = 0;
    while (n > 1) {
        n >>= 1;
        r++;
    }
    return r;
}

// res = exp(z).  T[0..3]: w, s, t, kball.  res may alias z.
template <int L, class SCR>
CB_HD void acb_exp(const CRef& res, const CRef& z, const CRef* T, SCR S0, SCR S1) {
    int64_t E = acb_bound_exp(z);
    if (E > 36) {
        acb_set_nan<L>(res);
        return;
    }
    if (E == -TR_BIG) {
        acb_one<L>(res);
        return;
    }
    const int r = 16;
    int64_t k = E + r;
    if (k < 0) k = 0;
    CRef w = T[0], s = T[1], t = T[2], kb = T[3];
    acb_set<L>(w, z);
    acb_mul_2exp(w, -k);
    int64_t q = k > 0 ? r : -E;
    if (q > 8192) q = 8192;
    int64_t target = 32 * (int64_t)L + 8, bits = 0, N = 0;
    while (bits < target) {
        N++;
        bits += q + tr_floor_log2(N);
    }
    acb_one<L>(s);
#pragma unroll 1
    for (int64_t n = N - 1; n >= 1; n--) {
        acb_mul<L>(t, s, w, S0, S1);
        acb_div_si<L>(s, t, n, S0);
        acb_add_si<L>(s, s, 1, kb, S0, S1);
    }
    acb_add_error_2exp(s, -(32 * (int64_t)L + 7));
#pragma unroll 1
    for (int64_t i = 0; i < k; i++) {
        acb_mul<L>(t, s, s, S0, S1);
        CRef sw = s;
        s = t;
        t = sw;
    }
    acb_set<L>(res, s);
}

CB_HD double tr_pow2(int64_t k) {  // 2^k, -1022 <= k <= 1023
    uint64_t b = (uint64_t)(k + 1023) << 52;
    double d;
    memcpy(&d, &b, 8);
    return d;
}
template <int L>
CB_HD double tr_top53(const uint32_t* m, size_t stride, const Meta& t) {  // mid / 2^exp: +-[1/2, 1) or 0
    if (t.flags & (F_ZERO | F_NAN)) return 0.0;
    uint64_t mm = ((uint64_t)m[(size_t)(L - 1) * stride] << 32) | m[(size_t)(L - 2) * stride];
    double d = CB_DMUL((double)(int64_t)(mm >> 11), tr_pow2(-53));
    return (t.flags & F_SIGN) ? -d : d;
}
constexpr int TR_SEED_TERMS = 24;
template <int L>
CB_HD void tr_seed_log(double* yr, double* yi, const CRef& a) {
    const double PI = 3.141592653589793, HALF_PI = 1.5707963267948966, LN2 = 0.6931471805599453;
    Meta tr = ld_meta(a.t), ti = ld_meta(a.t + 1);
    bool zr0 = is_zero_mid(tr), zi0 = is_zero_mid(ti);
    int64_t er = tr.exp, ei = ti.exp;
    int64_t E = zr0 ? ei : (zi0 ? er : (er > ei ? er : ei));
    double xr = (zr0 || er - E < -200) ? 0.0 : CB_DMUL(tr_top53<L>(a.m, a.stride, tr), tr_pow2(er - E));
    double xi = (zi0 || ei - E < -200) ? 0.0 : CB_DMUL(tr_top53<L>(a.m + 1, a.stride, ti), tr_pow2(ei - E));
    double ar = xr < 0 ? -xr : xr, ai = xi < 0 ? -xi : xi;
    double pr, pi, rot;
    if (ar >= ai) {
        if (xr > 0) {
            pr = xr; pi = xi; rot = 0.0;
        } else {
            pr = -xr; pi = -xi; rot = xi >= 0 ? PI : -PI;
        }
    } else if (xi > 0) {
        pr = xi; pi = -xr; rot = HALF_PI;
    } else {
        pr = -xi; pi = xr; rot = -HALF_PI;
    }
    double nr = CB_DSUB(pr, 1.0), ni = pi, dr = CB_DADD(pr, 1.0), di = pi;
    double den = CB_DADD(CB_DMUL(dr, dr), CB_DMUL(di, di));
    double wr = CB_DDIV(CB_DADD(CB_DMUL(nr, dr), CB_DMUL(ni, di)), den);
    double wi = CB_DDIV(CB_DSUB(CB_DMUL(ni, dr), CB_DMUL(nr, di)), den);
    double w2r = CB_DSUB(CB_DMUL(wr, wr), CB_DMUL(wi, wi)), w2i = CB_DMUL(2.0, CB_DMUL(wr, wi));
    double sr = CB_DDIV(1.0, (double)(2 * TR_SEED_TERMS - 1)), si = 0.0;
#pragma unroll 1
    for (int j = TR_SEED_TERMS - 2; j >= 0; j--) {
        double t_r = CB_DSUB(CB_DMUL(sr, w2r), CB_DMUL(si, w2i)), t_i = CB_DADD(CB_DMUL(sr, w2i), CB_DMUL(si, w2r));
        sr = CB_DADD(t_r, CB_DDIV(1.0, (double)(2 * j + 1)));
        si = t_i;
    }
    double lr = CB_DSUB(CB_DMUL(sr, wr), CB_DMUL(si, wi)), li = CB_DADD(CB_DMUL(sr, wi), CB_DMUL(si, wr));
    *yr = CB_DADD(CB_DMUL(2.0, lr), CB_DMUL((double)E, LN2));
    *yi = CB_DADD(CB_DMUL(2.0, li), rot);
}
// exact ball of a double (subnormals flush to zero); part 0 re, 1 im
template <int L>
CB_HD void arb_set_d(const CRef& z, int part, double d) {
    uint64_t b;
    memcpy(&b, &d, 8);
    int e11 = (int)((b >> 52) & 0x7FF);
#pragma unroll 8
    for (int k = 0; k < L; k++) z.m[(size_t)k * z.stride + part] = 0u;
    if (e11 == 0 || e11 == 0x7FF) {
        st_meta(z.t + part, meta_zero());
        return;
    }
    uint64_t m64 = ((b & 0xFFFFFFFFFFFFFull) | (1ull << 52)) << 11;
    z.m[(size_t)(L - 1) * z.stride + part] = (uint32_t)(m64 >> 32);
    z.m[(size_t)(L - 2) * z.stride + part] = (uint32_t)m64;
    st_meta(z.t + part, Meta{e11 - 1022, (b >> 63) ? F_SIGN : 0u, 0u, 0});
}

// res = log(a) for a ball that does not straddle the branch cut.  T[0..13]: exp's four, then y0, e, t, u, d, v, v2,
// s, c, one.  res must not alias a.
template <int L, class SCR>
CB_HD void acb_log_main(const CRef& res, const CRef& a, const CRef* T, SCR S0, SCR S1);
// res = log(a).  A ball that straddles the branch cut (re < 0, im a non-exact ball containing 0) holds points with
// arg near +pi and near -pi; like Arb's acb_log the result is then Re = log|a| (from log(-a): -a lies around the
// positive real axis) and Im = [0 +- pi] (oracle/cbo_trans.c, cbo_acb_log).  T[14] holds -a.
constexpr uint32_t PI_MAG_MAN = 843314857u;  // ceil(pi * 2^28): pi <= man * 2^(2-30)
constexpr int32_t PI_MAG_EXP = 2;
template <int L, class SCR>
CB_HD void acb_log(const CRef& res, const CRef& a, const CRef* T, SCR S0, SCR S1) {
    Meta tr = ld_meta(a.t), ti = ld_meta(a.t + 1);
    if (!is_nan(tr) && !is_nan(ti) && !is_zero_mid(tr) && (tr.flags & F_SIGN) && !is_exact_zero(ti) &&
        arb_contains_zero<L>(a.m + 1, a.stride, ti)) {
        CRef na = T[14];
        acb_set<L>(na, a);
        acb_neg_inplace(na);
        acb_log_main<L>(res, na, T, S0, S1);
        if (acb_bound_exp(res) != TR_BIG) {
            store_nan<L>(wim(res));  // zeros
            st_meta(res.t + 1, Meta{0, F_ZERO, PI_MAG_MAN, PI_MAG_EXP});
        }
        return;
    }
    acb_log_main<L>(res, a, T, S0, S1);
}
template <int L, class SCR>
CB_HD void acb_log_main(const CRef& res, const CRef& a, const CRef* T, SCR S0, SCR S1) {
    if (acb_bound_exp(a) == TR_BIG || acb_contains_zero<L>(a)) {
        acb_set_nan<L>(res);
        return;
    }
    double yr, yi;
    tr_seed_log<L>(&yr, &yi, a);
    CRef kb = T[3], y0 = T[4], e = T[5], t = T[6], u = T[7], d = T[8], v = T[9], v2 = T[10], s = T[11], c = T[12],
         one = T[13];
    arb_set_d<L>(y0, 0, yr);
    arb_set_d<L>(y0, 1, yi);
    acb_set<L>(t, y0);
    acb_neg_inplace(t);
    acb_exp<L>(e, t, T, S0, S1);
    acb_mul<L>(t, a, e, S0, S1);
    acb_add_si<L>(u, t, -1, kb, S0, S1);
    int64_t Eu = acb_bound_exp(u);
    if (Eu > -8) {
        acb_set_nan<L>(res);
        return;
    }
    if (Eu == -TR_BIG) {
        acb_set<L>(res, y0);
        return;
    }
    acb_add_si<L>(d, u, 2, kb, S0, S1);
    acb_div<L>(v, u, d, e, S0, S1);  // e is free again: the inverse of a complex divisor goes there
    int64_t q = -Eu;
    if (q > 8192) q = 8192;
    int64_t target = 32 * (int64_t)L + 10, K = 1;
    while (q * (2 * K + 1) < target) K++;
    acb_mul<L>(v2, v, v, S0, S1);
    acb_one<L>(one);
    acb_div_si<L>(s, one, 2 * K - 1, S0);
#pragma unroll 1
    for (int64_t j = K - 2; j >= 0; j--) {
        acb_mul<L>(t, s, v2, S0, S1);
        acb_div_si<L>(c, one, 2 * j + 1, S0);
        acb_addsub<L>(s, t, c, false, S0, S1);
    }
    acb_mul<L>(t, s, v, S0, S1);
    acb_mul_2exp(t, 1);
    acb_add_error_2exp(t, -(32 * (int64_t)L + 8));
    acb_addsub<L>(res, y0, t, false, S0, S1);
}

// acb_ode_solution_evaluate for one solution at many points
struct EvalParams {
    int M, pow_mode;     // pow_mode: 0 rho == 0; 1 exact integer pow_e; 2 general
    uint64_t pow_e;
    int64_t cap, npts;
    CArr gens;  // M * cap entries, S = 2 (one solution for all points)
    CArr rho;   // 1 entry, S = 2
    CArr pts;   // 1 entry, S = 2*npts
    CArr out;   // 1 entry, S = 2*npts
    CArr wp;    // TRANS_NTMP + 4 entries, S = 2*npts
};
template <int L, class SCR>
CB_HD void evaluate_thread(const EvalParams& p, int64_t pt, SCR S0, SCR S1) {
    const size_t slot = 2 * (size_t)pt;
    CRef T[TRANS_NTMP];
    for (int i = 0; i < TRANS_NTMP; i++) T[i] = at<L>(p.wp, i, slot);
    CRef lg = at<L>(p.wp, TRANS_NTMP, slot), res = at<L>(p.wp, TRANS_NTMP + 1, slot),
         qv = at<L>(p.wp, TRANS_NTMP + 2, slot), pw = at<L>(p.wp, TRANS_NTMP + 3, slot);
    CRef a = at<L>(p.pts, 0, slot), out = at<L>(p.out, 0, slot), rho = at<L>(p.rho, 0, 0);
    if (acb_is_exact_zero(a)) {  // reference :29-33
        acb_set_nan<L>(out);
        return;
    }
    const bool have_log = p.M > 1 || p.pow_mode == 2;
    if (have_log) acb_log<L>(lg, a, T, S0, S1);
    // the series live in the shared array: a polynomial context on it (slot 0), temporaries per point
    PolyCtx<L, SCR> pc{p.gens, 0, S0, S1, T[0], T[1], T[2], T[3], T[4]};
    auto series_eval = [&](const CRef& y, int i) {
        int64_t e0 = (int64_t)i * p.cap;
        // the series are shared by all points, so every thread runs the same Horner loop: phase barriers -- unless this
        // point is an exact zero of the series variable (poly_eval then returns the constant term without looping)
        poly_eval<L, EVAL_SYNC>(pc, y, e0, pc.norm(e0, p.cap), a);
    };
    series_eval(res, 0);
    int64_t binom = 1;
#pragma unroll 1
    for (int i = 1; i < p.M; i++) {
        acb_mul<L>(T[5], res, lg, S0, S1);
        series_eval(qv, i);
        binom = (binom * (p.M - i + 1)) / i;
        acb_mul_si<L>(qv, qv, binom, S0);
        acb_addsub<L>(res, T[5], qv, false, S0, S1);
    }
    // a^rho
    if (p.pow_mode == 0 || (p.pow_mode == 1 && p.pow_e == 0)) {
        acb_one<L>(pw);
    } else if (p.pow_mode == 1) {
        CRef pp = pw, tt = T[5];
        acb_set<L>(pp, a);
        int top = 63;
        while (!((p.pow_e >> top) & 1)) top--;
#pragma unroll 1
        for (int bit = top - 1; bit >= 0; bit--) {
            acb_mul<L>(tt, pp, pp, S0, S1);
            CRef sw = pp; pp = tt; tt = sw;
            if ((p.pow_e >> bit) & 1) {
                acb_mul<L>(tt, pp, a, S0, S1);
                sw = pp; pp = tt; tt = sw;
            }
        }
        if (pp.m != pw.m) acb_set<L>(pw, pp);
    } else {
        acb_mul<L>(T[5], lg, rho, S0, S1);
        acb_exp<L>(pw, T[5], T, S0, S1);
    }
    acb_mul<L>(out, res, pw, S0, S1);
}

// elementwise acb_exp / acb_log (unit-test surface and the compat layer's acb_log / acb_exp)
struct TransParams {
    int which;  // 0 exp, 1 log
    int64_t n;
    CArr z, x;  // 1 entry each, S = 2*n
    CArr wp;    // TRANS_NTMP entries
};
template <int L, class SCR>
CB_HD void trans_thread(const TransParams& p, int64_t i, SCR S0, SCR S1) {
    const size_t slot = 2 * (size_t)i;
    CRef T[TRANS_NTMP];
    for (int k = 0; k < TRANS_NTMP; k++) T[k] = at<L>(p.wp, k, slot);
    CRef z = at<L>(p.z, 0, slot), x = at<L>(p.x, 0, slot);
    if (p.which == 0) acb_exp<L>(z, x, T, S0, S1);
    else acb_log<L>(z, x, T, S0, S1);
}

// ---------------------------------------------------------------- radius_of_convergence / truncation_order
// reference src/fuchs_solver.c:3-94, for a batch, on the device ball type: a RIGOROUS enclosure (every step a ball
// operation), operation for operation the algorithm stated in oracle/cbo_trans.c (cbo_radius_of_convergence,
// cbo_truncation_order) -- Graeffe iterations on the reversed leading polynomial with an exact power-of-two rescaling
// after every step (as many of the n requested as the 32-bit exponent range allows: n_done), the sum form of
// Fujiwara's root bound, its 2^n_done-th root inverted, and the sharpness (2 m^2 + 1)^(1/2^n_done) - 1 of the bound as
// the error bar.  One thread per operator.
constexpr int RADIUS_MAXLEN = 40;
constexpr int32_t RADIUS_EXP_GUARD = 1 << 24;
constexpr int RADIUS_NTMP = 14;  // t, tc, w, U, l, v, S, R, X, one, E, inv, kb + 1 spare
struct RadiusParams {
    int order, degree;
    int64_t batch, n;
    CArr ode;      // S = 2*batch or 2
    CArr out;      // 1 entry, S = 2*batch
    CArr wp;       // 2 * (degree + 1) + RADIUS_NTMP + TRANS_NTMP entries, S = 2*batch
    int32_t* n_done;  // [batch]
};
// z.re = the exact ball of a 30-bit upper bound of |x.re|; z.im = 0
template <int L>
CB_HD void acb_upper_ball_re(const CRef& z, const CRef& x) {
    Meta t = ld_meta(x.t);
    acb_zero<L>(z);
    if (is_nan(t)) {
        acb_set_nan<L>(z);
        return;
    }
    Mag u = mag_add(mag_mid_upper(t, x.m[(size_t)(L - 1) * x.stride]), mag_of(t));
    if (u.man == 0) return;
    z.m[(size_t)(L - 1) * z.stride] = u.man << 2;
    st_meta(z.t, Meta{u.exp, 0u, 0u, 0});
}
template <int L>
CB_HD void acb_set_si_exact(const CRef& z, int64_t k) {
    acb_zero<L>(z);
    if (k == 0) return;
    uint64_t ka = k < 0 ? (uint64_t)0 - (uint64_t)k : (uint64_t)k;
    int sh = clz64(ka);
    uint64_t kn = ka << sh;
    z.m[(size_t)(L - 1) * z.stride] = (uint32_t)(kn >> 32);
    z.m[(size_t)(L - 2) * z.stride] = (uint32_t)kn;
    st_meta(z.t, Meta{64 - sh, k < 0 ? F_SIGN : 0u, 0u, 0});
}
CB_HD bool radius_part_ok(const Meta& t) {
    if (is_nan(t)) return false;
    if (!is_zero_mid(t) && (t.exp > RADIUS_EXP_GUARD || t.exp < -RADIUS_EXP_GUARD)) return false;
    if (t.rman && (t.rexp > RADIUS_EXP_GUARD || t.rexp < -2 * RADIUS_EXP_GUARD)) return false;
    return true;
}
template <int L, class SCR>
CB_HD void radius_thread(const RadiusParams& p, int64_t inst, SCR S0, SCR S1) {
    const size_t slot = 2 * (size_t)inst, oslot = (p.ode.S == 2) ? 0 : slot;
    const int D1 = p.degree + 1;
    CRef out = at<L>(p.out, 0, slot);
    int low = -1, top = -1;
    for (int i = 0; i <= p.degree; i++)
        if (!acb_is_exact_zero(at<L>(p.ode, p.order * D1 + i, oslot))) {
            if (low < 0) low = i;
            top = i;
        }
    if (low < 0 || top - low < 1 || top - low + 1 > RADIUS_MAXLEN) {
        acb_set_nan<L>(out);
        p.n_done[inst] = -1;
        return;
    }
    const int m = top - low;
    const int64_t tb = 2 * (int64_t)D1;
    auto A = [&](int i) { return at<L>(p.wp, i, slot); };
    auto G = [&](int i) { return at<L>(p.wp, D1 + i, slot); };
    CRef t = at<L>(p.wp, tb + 0, slot), tc = at<L>(p.wp, tb + 1, slot), w = at<L>(p.wp, tb + 2, slot),
         U = at<L>(p.wp, tb + 3, slot), l = at<L>(p.wp, tb + 4, slot), v = at<L>(p.wp, tb + 5, slot),
         S = at<L>(p.wp, tb + 6, slot), R = at<L>(p.wp, tb + 7, slot), X = at<L>(p.wp, tb + 8, slot),
         one = at<L>(p.wp, tb + 9, slot), E = at<L>(p.wp, tb + 10, slot), inv = at<L>(p.wp, tb + 11, slot),
         kb = at<L>(p.wp, tb + 12, slot);
    CRef T[TRANS_NTMP];
    for (int i = 0; i < TRANS_NTMP; i++) T[i] = at<L>(p.wp, tb + RADIUS_NTMP + i, slot);
    for (int i = 0; i <= m; i++) acb_set<L>(A(i), at<L>(p.ode, p.order * D1 + (top - i), oslot));
    int64_t n_done = 0;
#pragma unroll 1
    for (int64_t it = 0; it < p.n; it++) {
        bool ok = true;
        for (int i = 0; i <= m && ok; i++) ok = radius_part_ok(ld_meta(A(i).t)) && radius_part_ok(ld_meta(A(i).t + 1));
        if (!ok) break;
        const int ne = m / 2 + 1, no = (m + 1) / 2;
        for (int j = 0; j <= m; j++) acb_zero<L>(G(j));
#pragma unroll 1
        for (int i = 0; i < ne; i++)
#pragma unroll 1
            for (int k = 0; k < ne; k++) {
                acb_mul<L>(t, A(2 * i), A(2 * k), S0, S1);
                acb_addsub<L>(G(i + k), G(i + k), t, false, S0, S1);
            }
#pragma unroll 1
        for (int i = 0; i < no; i++)
#pragma unroll 1
            for (int k = 0; k < no; k++) {
                acb_mul<L>(t, A(2 * i + 1), A(2 * k + 1), S0, S1);
                acb_addsub<L>(G(i + k + 1), G(i + k + 1), t, true, S0, S1);
            }
        CRef gm = G(m);
        if (acb_bound_exp(gm) == TR_BIG || acb_contains_zero<L>(gm)) break;
        Meta gr = ld_meta(gm.t), gi = ld_meta(gm.t + 1);
        int64_t e0 = -TR_BIG;
        if (!is_zero_mid(gr)) e0 = gr.exp;
        if (!is_zero_mid(gi) && gi.exp > e0) e0 = gi.exp;
        for (int j = 0; j <= m; j++) {
            acb_set<L>(A(j), G(j));
            acb_mul_2exp(A(j), -e0);
        }
        n_done++;
    }
    p.n_done[inst] = -1;
    if (acb_bound_exp(A(m)) == TR_BIG || acb_contains_zero<L>(A(m))) {
        acb_set_nan<L>(out);
        return;
    }
    acb_zero<L>(S);
#pragma unroll 1
    for (int k = 1; k <= m; k++) {
        if (acb_is_exact_zero(A(m - k))) continue;
        acb_div<L>(t, A(m - k), A(m), inv, S0, S1);
        if (k == m) acb_mul_2exp(t, -1);
        acb_set<L>(tc, t);
        {   // conj(t)
            Meta ti = ld_meta(tc.t + 1);
            if (!(ti.flags & (F_ZERO | F_NAN))) {
                ti.flags ^= F_SIGN;
                st_meta(tc.t + 1, ti);
            }
        }
        acb_mul<L>(w, t, tc, S0, S1);
        acb_upper_ball_re<L>(U, w);
        if (is_zero_mid(ld_meta(U.t)) && !is_nan(ld_meta(U.t))) continue;
        acb_log<L>(l, U, T, S0, S1);
        acb_div_si<L>(l, l, 2 * (int64_t)k, S0);
        acb_exp<L>(v, l, T, S0, S1);
        acb_addsub<L>(S, S, v, false, S0, S1);
    }
    acb_mul_2exp(S, 1);
    acb_upper_ball_re<L>(U, S);
    if (acb_bound_exp(U) == TR_BIG || is_zero_mid(ld_meta(U.t))) {
        acb_set_nan<L>(out);
        return;
    }
    acb_log<L>(l, U, T, S0, S1);
    acb_mul_2exp(l, -n_done);
    acb_exp<L>(R, l, T, S0, S1);
    acb_one<L>(one);
    acb_div<L>(X, one, R, inv, S0, S1);
    acb_set_si_exact<L>(w, 2 * (int64_t)m * m + 1);
    acb_log<L>(l, w, T, S0, S1);
    acb_mul_2exp(l, -n_done);
    acb_exp<L>(E, l, T, S0, S1);
    acb_add_si<L>(E, E, -1, kb, S0, S1);
    acb_mul<L>(w, X, E, S0, S1);
    if (acb_bound_exp(X) == TR_BIG || acb_bound_exp(w) == TR_BIG) {
        acb_set_nan<L>(out);
        return;
    }
    acb_set<L>(out, X);
    {   // arb_add_error: the radius grows by a 30-bit upper bound of |w.re|; the imaginary part is an exact zero
        Meta to = ld_meta(out.t), tw = ld_meta(w.t);
        Mag r = mag_add(mag_of(to), mag_add(mag_mid_upper(tw, w.m[(size_t)(L - 1) * w.stride]), mag_of(tw)));
        to.rman = r.man;
        to.rexp = r.man ? r.exp : 0;
        st_meta(out.t, to);
        store_nan<L>(wim(out));
        st_meta(out.t + 1, meta_zero());
    }
    p.n_done[inst] = (int32_t)n_done;
}

struct TruncParams {
    int64_t batch, bits;
    CArr eta, alpha;  // 1 entry each, S = 2*batch (real balls: the imaginary parts are ignored as in the reference)
    CArr nball;       // 1 entry, S = 2*batch: the ball N
    CArr wp;          // 8 + TRANS_NTMP entries
    int64_t* n_out;   // [batch]
};
template <int L, class SCR>
CB_HD void truncation_thread(const TruncParams& p, int64_t inst, SCR S0, SCR S1) {
    const size_t slot = 2 * (size_t)inst;
    CRef eta = at<L>(p.eta, 0, slot), alpha = at<L>(p.alpha, 0, slot), N = at<L>(p.nball, 0, slot);
    CRef r = at<L>(p.wp, 0, slot), tmp = at<L>(p.wp, 1, slot), one = at<L>(p.wp, 2, slot), two = at<L>(p.wp, 3, slot),
         inv = at<L>(p.wp, 4, slot), U = at<L>(p.wp, 5, slot), q = at<L>(p.wp, 6, slot), lg = at<L>(p.wp, 7, slot);
    CRef T[TRANS_NTMP];
    for (int i = 0; i < TRANS_NTMP; i++) T[i] = at<L>(p.wp, 8 + i, slot);
    acb_set_nan<L>(N);
    if (acb_bound_exp(eta) == TR_BIG || acb_bound_exp(alpha) == TR_BIG) {
        p.n_out[inst] = 2 * p.bits;
        return;
    }
    acb_addsub<L>(tmp, alpha, eta, true, S0, S1);
    {
        Meta t = ld_meta(tmp.t);
        if ((t.flags & (F_ZERO | F_SIGN)) || arb_contains_zero<L>(tmp.m, tmp.stride, t)) {
            p.n_out[inst] = 0;
            return;
        }
    }
    acb_addsub<L>(r, eta, alpha, false, S0, S1);
    acb_mul_2exp(r, -1);
    acb_div<L>(q, eta, r, inv, S0, S1);  // r = eta / r
    acb_one<L>(one);
    acb_addsub<L>(tmp, one, q, true, S0, S1);
    acb_log<L>(N, tmp, T, S0, S1);
    acb_set_si_exact<L>(two, 2);
    acb_log<L>(tmp, two, T, S0, S1);
    acb_mul_si<L>(lg, tmp, p.bits, S0);
    acb_addsub<L>(N, N, lg, true, S0, S1);
    acb_log<L>(tmp, q, T, S0, S1);
    acb_div<L>(r, N, tmp, inv, S0, S1);
    acb_set<L>(N, r);
    if (acb_bound_exp(N) == TR_BIG) {
        p.n_out[inst] = 2 * p.bits;
        return;
    }
    acb_upper_ball_re<L>(U, N);
    Meta tu = ld_meta(U.t), tn = ld_meta(N.t);
    if (is_zero_mid(tu) || (tn.flags & F_SIGN)) {
        p.n_out[inst] = 0;
        return;
    }
    if (tu.exp > 62) {
        p.n_out[inst] = 2 * p.bits;
        return;
    }
    if (tu.exp <= 0) {
        p.n_out[inst] = 1;
        return;
    }
    uint64_t topw = ((uint64_t)U.m[(size_t)(L - 1) * U.stride] << 32) | U.m[(size_t)(L - 2) * U.stride];
    uint64_t ip = topw >> (64 - tu.exp);
    bool frac = (topw << tu.exp) != 0;  // (the upper ball has a 30-bit mantissa: nothing below the top limb)
    p.n_out[inst] = (int64_t)ip + (frac ? 1 : 0);
}

// ---------------------------------------------------------------- residual check on the device
// acb_ode_apply / acb_ode_solves (reference src/acb_ode.c:185-240), the reference's own acceptance test
// for both solvers, for a batch: out = sum_i p_i * in^(i) truncated to out_len coefficients, and
// solved[instance] = every one of the first `deg` residual coefficients is finite and contains zero.
// Same operation order as the oracle's cbo_ode_apply (classical products, row by row).
struct ApplyParams {
    int order, degree;
    int64_t batch, len_in, out_len, deg;
    CArr ode;  // S = 2*batch or 2
    CArr in;   // len_in entries, S = 2*batch
    CArr out;  // out_len entries, S = 2*batch
    CArr wp;   // len_in + 1 entries, S = 2*batch: the running derivative and one product
    int32_t* solved;  // [batch] or nullptr
};
template <int L, class SCR>
CB_HD void apply_thread(const ApplyParams& p, int64_t inst, SCR S0, SCR S1) {
    const size_t slot = 2 * (size_t)inst;
    const size_t oslot = (p.ode.S == 2) ? 0 : slot;
    const int D1 = p.degree + 1;
    CRef t = at<L>(p.wp, p.len_in, slot);
#pragma unroll 1
    for (int64_t k = 0; k < p.len_in; k++) acb_set<L>(at<L>(p.wp, k, slot), at<L>(p.in, k, slot));
#pragma unroll 1
    for (int64_t k = 0; k < p.out_len; k++) acb_zero<L>(at<L>(p.out, k, slot));
    int64_t dl = p.len_in;
#pragma unroll 1
    for (int i = 0; i <= p.order; i++) {
#pragma unroll 1
        for (int j = 0; j <= p.degree; j++) {
            CRef c = at<L>(p.ode, i * D1 + j, oslot);
            if (acb_is_exact_zero(c)) continue;
#pragma unroll 1
            for (int64_t k = 0; k < dl && j + k < p.out_len; k++) {
                acb_mul<L>(t, c, at<L>(p.wp, k, slot), S0, S1);
                CRef o = at<L>(p.out, j + k, slot);
                acb_addsub<L>(o, o, t, false, S0, S1);
            }
        }
#pragma unroll 1
        for (int64_t k = 1; k < dl; k++) acb_mul_si<L>(at<L>(p.wp, k - 1, slot), at<L>(p.wp, k, slot), k, S0);
        if (dl > 0) dl--;
    }
    if (p.solved) {
        int ok = 1;
        for (int64_t k = 0; k < p.deg && k < p.out_len; k++) {
            CRef o = at<L>(p.out, k, slot);
            if (acb_bound_exp(o) == TR_BIG || !acb_contains_zero<L>(o)) {
                ok = 0;
                break;
            }
        }
        p.solved[inst] = ok;
    }
}

}  // namespace cb
