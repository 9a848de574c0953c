/* cbo_trans.c -- ORACLE restatement of the transcendental part of acb_ode_solution_evaluate
 * (reference src/acb_ode_solution.c:24-58): acb_log(a), acb_pow(a, rho) and the Horner-in-log(a)
 * combination.  TEST INFRASTRUCTURE ONLY (see cbo.h).
 *
 * Arb's acb_log / acb_exp are not correctly rounded functions (their balls depend on Arb's internal
 * algorithms and version), so unlike +, -, *, / there is no unique answer to restate: any rigorous
 * enclosure is "correct".  The algorithms below are chosen so that the device can run the SAME
 * sequence of ball operations (bit-identical results) using only the primitives it already has:
 *   exp(z):  w = z / 2^k with |w| < 2^-16; Taylor polynomial by Horner (s = s w / n + 1) with the
 *            remainder 2^-(prec+7) added to the radius; k squarings.
 *   log(a):  y0 = a double-precision approximation of log(a) formed with IEEE +,-,*,/ only
 *            (deterministic on CPU and GPU); t = a exp(-y0); u = t - 1; v = u / (u + 2);
 *            log(a) = y0 + 2 atanh(v) by its Taylor series, remainder added to the radius.
 *   pow(a, rho) = exp(rho log(a)); exact non-negative integer rho < 2^31: binary exponentiation;
 *            rho = 0: 1  (the cases Arb's acb_pow distinguishes as well).
 * A ball that contains zero (log) or an exponent beyond 2^36 (exp) gives an indeterminate result. */
#include "cbo.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

void cbo_arb_set_nan(cbo_arb *z);
void cbo_arb_add_error_2exp(cbo_arb *z, int64_t e);
void cbo_arb_mul_2exp(cbo_arb *z, int64_t k);

#define BIG (((int64_t)1) << 40)

static void acb_set_nan(cbo_acb *z) {
    cbo_arb_set_nan(&z->re);
    cbo_arb_set_nan(&z->im);
}
/* E with |z| < 2^E; BIG for an indeterminate ball, -BIG for the exact zero */
static int64_t acb_bound_exp(const cbo_acb *z) {
    if (!cbo_acb_is_finite(z)) return BIG;
    int64_t e = -BIG;
    const cbo_arb *p[2] = {&z->re, &z->im};
    for (int i = 0; i < 2; i++) {
        if (!(p[i]->flags & CBO_ZERO) && p[i]->exp > e) e = p[i]->exp;
        if (p[i]->rman && p[i]->rexp > e) e = p[i]->rexp;
    }
    return e == -BIG ? -BIG : e + 2;
}
static void acb_mul_2exp(cbo_acb *z, int64_t k) {
    cbo_arb_mul_2exp(&z->re, k);
    cbo_arb_mul_2exp(&z->im, k);
}
static void acb_add_error_2exp(cbo_acb *z, int64_t e) {
    cbo_arb_add_error_2exp(&z->re, e);
    cbo_arb_add_error_2exp(&z->im, e);
}
static void acb_div_si(cbo_acb *z, const cbo_acb *x, int64_t n, int L) {
    cbo_acb k;
    cbo_acb_set_si(&k, n, L);
    cbo_acb_div(z, x, &k, L);
}
static int floor_log2(int64_t n) {
    int r = 0;
    while (n > 1) {
        n >>= 1;
        r++;
    }
    return r;
}

/* ---- exp */
void cbo_acb_exp(cbo_acb *res, const cbo_acb *z, int L) {
    int64_t E = acb_bound_exp(z);
    if (E > 36) {
        acb_set_nan(res);
        return;
    }
    if (E == -BIG) {
        cbo_acb_one(res, L);
        return;
    }
    const int r = 16;
    int64_t k = E + r;
    if (k < 0) k = 0;
    cbo_acb w = *z, s, t;
    acb_mul_2exp(&w, -k);
    int64_t q = k > 0 ? r : -E; /* |w| < 2^-q */
    if (q > 8192) q = 8192;
    int64_t target = 32 * (int64_t)L + 8, bits = 0, N = 0;
    while (bits < target) {
        N++;
        bits += q + floor_log2(N);
    }
    cbo_acb_one(&s, L);
    for (int64_t n = N - 1; n >= 1; n--) {
        cbo_acb_mul(&t, &s, &w, L);
        acb_div_si(&s, &t, n, L);
        cbo_acb_add_si(&s, &s, 1, L);
    }
    acb_add_error_2exp(&s, -(32 * (int64_t)L + 7));
    for (int64_t i = 0; i < k; i++) {
        cbo_acb_mul(&t, &s, &s, L);
        s = t;
    }
    *res = s;
}

/* ---- the double-precision seed of log: IEEE +,-,*,/ only (no libm), so that the GPU computes the
 * same bits.  Built without FMA contraction (oracle/Makefile: -ffp-contract=off). */
static double d_pow2(int64_t k) { /* 2^k, -1022 <= k <= 1023 */
    uint64_t b = (uint64_t)(k + 1023) << 52;
    double d;
    memcpy(&d, &b, 8);
    return d;
}
static double top53(const cbo_arb *x, int L) { /* mid / 2^exp: +-[1/2, 1) or 0 */
    if (x->flags & (CBO_ZERO | CBO_NAN)) return 0.0;
    uint64_t m = ((uint64_t)x->m.w[L - 1] << 32) | x->m.w[L - 2];
    double d = (double)(int64_t)(m >> 11) * d_pow2(-53);
    return (x->flags & CBO_SIGN) ? -d : d;
}
#define SEED_TERMS 24
static void seed_log(double *yr, double *yi, const cbo_acb *a, int L) {
    const double PI = 3.141592653589793, HALF_PI = 1.5707963267948966, LN2 = 0.6931471805599453;
    int zr0 = (a->re.flags & CBO_ZERO) != 0, zi0 = (a->im.flags & CBO_ZERO) != 0;
    int64_t er = a->re.exp, ei = a->im.exp;
    int64_t E = zr0 ? ei : (zi0 ? er : (er > ei ? er : ei));
    double xr = zr0 || er - E < -200 ? 0.0 : top53(&a->re, L) * d_pow2(er - E);
    double xi = zi0 || ei - E < -200 ? 0.0 : top53(&a->im, L) * d_pow2(ei - E);
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
    /* w = (z' - 1) / (z' + 1) */
    double nr = pr - 1.0, ni = pi, dr = pr + 1.0, di = pi;
    double den = dr * dr + di * di;
    double wr = (nr * dr + ni * di) / den, wi = (ni * dr - nr * di) / den;
    double w2r = wr * wr - wi * wi, w2i = 2.0 * (wr * wi);
    double sr = 1.0 / (double)(2 * SEED_TERMS - 1), si = 0.0;
    for (int j = SEED_TERMS - 2; j >= 0; j--) {
        double tr = sr * w2r - si * w2i, ti = sr * w2i + si * w2r;
        sr = tr + 1.0 / (double)(2 * j + 1);
        si = ti;
    }
    double lr = sr * wr - si * wi, li = sr * wi + si * wr;
    *yr = 2.0 * lr + (double)E * LN2;
    *yi = 2.0 * li + rot;
}
static void arb_set_d(cbo_arb *z, double d, int L) { /* exact; subnormals flush to zero */
    uint64_t b;
    memcpy(&b, &d, 8);
    int e11 = (int)((b >> 52) & 0x7FF);
    cbo_arb_zero(z);
    if (e11 == 0 || e11 == 0x7FF) return;
    uint64_t m64 = ((b & 0xFFFFFFFFFFFFFull) | (1ull << 52)) << 11;
    z->m.w[L - 1] = (uint32_t)(m64 >> 32);
    z->m.w[L - 2] = (uint32_t)m64;
    z->exp = e11 - 1022;
    z->flags = (b >> 63) ? CBO_SIGN : 0;
}

/* ---- log */
static void acb_log_main(cbo_acb *res, const cbo_acb *a, int L);
/* A ball that straddles the branch cut (re < 0, im a non-exact ball containing 0) holds points with arg near +pi
 * and points with arg near -pi: the principal log is discontinuous on it.  Like Arb's acb_log (acb_arg returns
 * [-pi, pi] there) the result is Re = log|a| (taken from log(-a): -a lies around the positive real axis, away from
 * the cut, and |-a| = |a|) and Im = [0 +- pi], pi rounded up to a 30-bit magnitude.  A point ON the cut (im an
 * exact zero) keeps the principal value +i pi. */
#define CBO_PI_MAG_MAN 843314857u /* ceil(pi * 2^28): pi <= man * 2^(2-30) */
#define CBO_PI_MAG_EXP 2
void cbo_acb_log(cbo_acb *res, const cbo_acb *a, int L) {
    int im_exact_zero = (a->im.flags & CBO_ZERO) && a->im.rman == 0;
    if (cbo_acb_is_finite(a) && !(a->re.flags & CBO_ZERO) && (a->re.flags & CBO_SIGN) && !im_exact_zero &&
        cbo_arb_contains_zero(&a->im)) {
        cbo_acb na;
        cbo_acb_neg(&na, a);
        acb_log_main(res, &na, L);
        if (cbo_acb_is_finite(res)) {
            cbo_arb_zero(&res->im);
            res->im.rman = CBO_PI_MAG_MAN;
            res->im.rexp = CBO_PI_MAG_EXP;
        }
        return;
    }
    acb_log_main(res, a, L);
}
static void acb_log_main(cbo_acb *res, const cbo_acb *a, int L) {
    if (!cbo_acb_is_finite(a) || cbo_acb_contains_zero(a)) {
        acb_set_nan(res);
        return;
    }
    double yr, yi;
    seed_log(&yr, &yi, a, L);
    cbo_acb y0, ny, e, t, u, d, v, v2, s, c, one;
    arb_set_d(&y0.re, yr, L);
    arb_set_d(&y0.im, yi, L);
    cbo_acb_neg(&ny, &y0);
    cbo_acb_exp(&e, &ny, L);
    cbo_acb_mul(&t, a, &e, L);
    cbo_acb_add_si(&u, &t, -1, L);
    int64_t Eu = acb_bound_exp(&u);
    if (Eu > -8) { /* the seed is not close enough (wide ball, or indeterminate exp) */
        acb_set_nan(res);
        return;
    }
    if (Eu == -BIG) {
        *res = y0;
        return;
    }
    cbo_acb_add_si(&d, &u, 2, L);
    cbo_acb_div(&v, &u, &d, L);
    int64_t q = -Eu; /* |v| < 2^-q */
    if (q > 8192) q = 8192;
    int64_t target = 32 * (int64_t)L + 10, K = 1;
    while (q * (2 * K + 1) < target) K++;
    cbo_acb_mul(&v2, &v, &v, L);
    cbo_acb_one(&one, L);
    acb_div_si(&s, &one, 2 * K - 1, L);
    for (int64_t j = K - 2; j >= 0; j--) {
        cbo_acb_mul(&t, &s, &v2, L);
        acb_div_si(&c, &one, 2 * j + 1, L);
        cbo_acb_add(&s, &t, &c, L);
    }
    cbo_acb_mul(&t, &s, &v, L);
    acb_mul_2exp(&t, 1);
    acb_add_error_2exp(&t, -(32 * (int64_t)L + 8));
    cbo_acb_add(res, &y0, &t, L);
}

/* ---- pow: mode 0: rho == 0 -> 1; mode 1: exact integer e >= 0 -> binary exponentiation;
 * mode 2: exp(rho log a).  lg: log(a) if the caller already has it, else NULL */
void cbo_acb_pow_mode(cbo_acb *res, const cbo_acb *a, const cbo_acb *rho, int mode, uint64_t e,
                      const cbo_acb *lg, int L) {
    cbo_acb p, t;
    if (mode == 0 || (mode == 1 && e == 0)) {
        cbo_acb_one(res, L);
        return;
    }
    if (mode == 1) {
        p = *a;
        int top = 63;
        while (!((e >> top) & 1)) top--;
        for (int bit = top - 1; bit >= 0; bit--) {
            cbo_acb_mul(&t, &p, &p, L);
            p = t;
            if ((e >> bit) & 1) {
                cbo_acb_mul(&t, &p, a, L);
                p = t;
            }
        }
        *res = p;
        return;
    }
    cbo_acb l;
    if (lg) l = *lg;
    else cbo_acb_log(&l, a, L);
    cbo_acb_mul(&t, &l, rho, L);
    cbo_acb_exp(res, &t, L);
}

/* ---- acb_ode_solution_evaluate, reference src/acb_ode_solution.c:24-58.
 * gens: M series of `cap` coefficients each (length implied by trailing exact zeros), shared by all
 * points; out[p] for every point a[p]. */
static void series_eval(cbo_acb *y, const cbo_acb *f, int64_t cap, const cbo_acb *x, int L) {
    int64_t len = cap;
    while (len > 0 && cbo_acb_is_zero(f + len - 1)) len--;
    if (len == 0) cbo_acb_zero(y);
    else if (len == 1) *y = f[0];
    else cbo_poly_evaluate_horner(y, f, len, x, 1, L);
}
int cbo_solution_evaluate_flat(int L, int64_t M, int64_t cap, int64_t npts, const uint32_t *gens_m,
                               const int32_t *gens_t, const uint32_t *rho_m, const int32_t *rho_t,
                               int pow_mode, uint64_t pow_e, const uint32_t *a_m, const int32_t *a_t,
                               uint32_t *out_m, int32_t *out_t) {
    cbo_acb *g = (cbo_acb *)malloc(sizeof(cbo_acb) * (M * cap > 0 ? M * cap : 1));
    cbo_acb rho;
    for (int64_t i = 0; i < M * cap; i++) {
        cbo_load(&g[i].re, gens_m + (2 * i) * L, gens_t + (2 * i) * 4, L);
        cbo_load(&g[i].im, gens_m + (2 * i + 1) * L, gens_t + (2 * i + 1) * 4, L);
    }
    cbo_load(&rho.re, rho_m, rho_t, L);
    cbo_load(&rho.im, rho_m + L, rho_t + 4, L);
    for (int64_t p = 0; p < npts; p++) {
        cbo_acb a, l, res, q, pw, out;
        cbo_load(&a.re, a_m + (2 * p) * L, a_t + (2 * p) * 4, L);
        cbo_load(&a.im, a_m + (2 * p + 1) * L, a_t + (2 * p + 1) * 4, L);
        if (cbo_acb_is_zero(&a)) { /* :29-33 */
            acb_set_nan(&out);
        } else {
            int have_log = M > 1 || pow_mode == 2;
            if (have_log) cbo_acb_log(&l, &a, L); /* :39 */
            series_eval(&res, g, cap, &a, L);     /* :40 */
            int64_t binom = 1;
            for (int64_t i = 1; i < M; i++) {
                cbo_acb_mul(&res, &res, &l, L);           /* :43 */
                series_eval(&q, g + i * cap, cap, &a, L); /* :45 */
                binom = (binom * (M - i + 1)) / i;
                cbo_acb_mul_si(&q, &q, binom, L);         /* :47 */
                cbo_acb_add(&res, &res, &q, L);           /* :49 */
            }
            cbo_acb_pow_mode(&pw, &a, &rho, pow_mode, pow_e, have_log ? &l : NULL, L); /* :52 */
            cbo_acb_mul(&out, &res, &pw, L);                                             /* :53 */
        }
        cbo_store(out_m + (2 * p) * L, out_t + (2 * p) * 4, &out.re, L);
        cbo_store(out_m + (2 * p + 1) * L, out_t + (2 * p + 1) * 4, &out.im, L);
    }
    free(g);
    return 0;
}

/* ---------------------------------------------------------------- radius_of_convergence / truncation_order
 * reference src/fuchs_solver.c:3-54: Graeffe iterations on the reversed leading polynomial, a root bound, its
 * 2^n-th root inverted, and the sharpness of the bound as the error bar.  Restated on this ball type with two stated
 * differences (Arb's versions of the same steps are unpinned anyway): (1) after every Graeffe step the polynomial
 * is rescaled by a power of two (exact; roots unchanged) and the iteration stops early when an exponent would
 * leave the 32-bit range of the ball type -- n_done <= n iterations are performed and used in the formulas; (2) the
 * root bound is the SUM form of Fujiwara's bound, B = 2 sum_k |a_{m-k}/a_m|^(1/k) (the last term halved inside the
 * root), which satisfies M <= B <= 2 m^2 M for the largest root modulus M, so that the true radius lies in
 * [X, X (2 m^2 + 1)^(1/2^n_done)] with X = B^(-1/2^n_done).  Every step is a ball operation: the result is a rigorous
 * enclosure.  Returns n_done, or -1 (out indeterminate) when there is no singularity away from zero (:17-22). */
void cbo_arb_upper_ball(cbo_arb *z, const cbo_arb *x, int L);
void cbo_arb_add_error_arb(cbo_arb *z, const cbo_arb *err, int L);
#define RADIUS_MAXLEN 40
#define RADIUS_EXP_GUARD (1 << 24)
static int radius_exps_ok(const cbo_acb *a, int m) {
    for (int i = 0; i <= m; i++) {
        const cbo_arb *p[2] = {&a[i].re, &a[i].im};
        for (int k = 0; k < 2; k++) {
            if (p[k]->flags & CBO_NAN) return 0;
            if (!(p[k]->flags & CBO_ZERO) && (p[k]->exp > RADIUS_EXP_GUARD || p[k]->exp < -RADIUS_EXP_GUARD)) return 0;
            if (p[k]->rman && (p[k]->rexp > RADIUS_EXP_GUARD || p[k]->rexp < -2 * RADIUS_EXP_GUARD)) return 0;
        }
    }
    return 1;
}
int cbo_radius_of_convergence(cbo_acb *out, cbo_ode *o, int64_t n, int L) {
    int low = -1, top = -1;
    for (int i = 0; i <= o->degree; i++)
        if (!cbo_acb_is_zero(&o->polys[o->order * (o->degree + 1) + i])) {
            if (low < 0) low = i;
            top = i;
        }
    if (low < 0 || top - low < 1 || top - low + 1 > RADIUS_MAXLEN) {
        acb_set_nan(out);
        return -1;
    }
    const int m = top - low;
    cbo_acb a[RADIUS_MAXLEN], g[RADIUS_MAXLEN], t;
    for (int i = 0; i <= m; i++) a[i] = o->polys[o->order * (o->degree + 1) + (top - i)];
    int64_t n_done = 0;
    for (int64_t it = 0; it < n; it++) {
        if (!radius_exps_ok(a, m)) break;
        const int ne = m / 2 + 1, no = (m + 1) / 2;
        for (int j = 0; j <= m; j++) cbo_acb_zero(&g[j]);
        for (int i = 0; i < ne; i++)
            for (int l = 0; l < ne; l++) {
                cbo_acb_mul(&t, &a[2 * i], &a[2 * l], L);
                cbo_acb_add(&g[i + l], &g[i + l], &t, L);
            }
        for (int i = 0; i < no; i++)
            for (int l = 0; l < no; l++) {
                cbo_acb_mul(&t, &a[2 * i + 1], &a[2 * l + 1], L);
                cbo_acb_sub(&g[i + l + 1], &g[i + l + 1], &t, L);
            }
        if (!cbo_acb_is_finite(&g[m]) || cbo_acb_contains_zero(&g[m])) break;
        int64_t e0 = -BIG;
        if (!(g[m].re.flags & CBO_ZERO)) e0 = g[m].re.exp;
        if (!(g[m].im.flags & CBO_ZERO) && g[m].im.exp > e0) e0 = g[m].im.exp;
        for (int j = 0; j <= m; j++) {
            a[j] = g[j];
            acb_mul_2exp(&a[j], -e0);
        }
        n_done++;
    }
    if (!cbo_acb_is_finite(&a[m]) || cbo_acb_contains_zero(&a[m])) {
        acb_set_nan(out);
        return -1;
    }
    cbo_acb S, tc, w, U, l, v;
    cbo_acb_zero(&S);
    for (int k = 1; k <= m; k++) {
        if (cbo_acb_is_zero(&a[m - k])) continue;
        cbo_acb_div(&t, &a[m - k], &a[m], L);
        if (k == m) acb_mul_2exp(&t, -1);
        cbo_acb_neg(&tc, &t);
        tc.re = t.re; /* conj(t): re kept, im negated */
        cbo_acb_mul(&w, &t, &tc, L);
        cbo_acb_zero(&U);
        cbo_arb_upper_ball(&U.re, &w.re, L);
        if (U.re.flags & CBO_ZERO) continue; /* the coefficient is (numerically) exactly zero */
        cbo_acb_log(&l, &U, L);
        acb_div_si(&l, &l, 2 * (int64_t)k, L);
        cbo_acb_exp(&v, &l, L);
        cbo_acb_add(&S, &S, &v, L);
    }
    acb_mul_2exp(&S, 1);
    cbo_acb_zero(&U);
    cbo_arb_upper_ball(&U.re, &S.re, L);
    if (!cbo_acb_is_finite(&U) || (U.re.flags & CBO_ZERO)) {
        acb_set_nan(out);
        return -1;
    }
    cbo_acb R, X, one, E;
    cbo_acb_log(&l, &U, L);
    acb_mul_2exp(&l, -n_done);
    cbo_acb_exp(&R, &l, L);
    cbo_acb_one(&one, L);
    cbo_acb_div(&X, &one, &R, L);
    cbo_acb_set_si(&w, 2 * (int64_t)m * m + 1, L);
    cbo_acb_log(&l, &w, L);
    acb_mul_2exp(&l, -n_done);
    cbo_acb_exp(&E, &l, L);
    cbo_acb_add_si(&E, &E, -1, L);
    cbo_acb_mul(&w, &X, &E, L);
    *out = X;
    cbo_arb_add_error_arb(&out->re, &w.re, L);
    cbo_arb_zero(&out->im);
    if (!cbo_acb_is_finite(&X) || !cbo_acb_is_finite(&w)) {
        acb_set_nan(out);
        return -1;
    }
    return (int)n_done;
}
/* truncation_order, reference src/fuchs_solver.c:56-94: N = (log(1 - r) - bits log 2) / log r with r = eta / ((eta +
 * alpha) / 2), as a real ball; *nball receives it.  Returns ceil of its UPPER end (when the ball does not pin one
 * integer the reference's arb_get_unique_fmpz gives up; taking the larger count is the safe reading), 2 bits for
 * non-finite input, 0 when eta < alpha cannot be shown. */
int64_t cbo_truncation_order(cbo_acb *nball, const cbo_acb *eta, const cbo_acb *alpha, int64_t bits, int L) {
    cbo_acb r, N, tmp, one, two;
    acb_set_nan(nball);
    if (!cbo_acb_is_finite(eta) || !cbo_acb_is_finite(alpha)) return 2 * bits;
    cbo_acb_sub(&tmp, alpha, eta, L); /* eta < alpha  <=>  alpha - eta is positive */
    if ((tmp.re.flags & (CBO_ZERO | CBO_SIGN)) || cbo_arb_contains_zero(&tmp.re)) return 0;
    cbo_acb_add(&r, eta, alpha, L);
    acb_mul_2exp(&r, -1);
    cbo_acb_div(&r, eta, &r, L);
    cbo_acb_one(&one, L);
    cbo_acb_sub(&tmp, &one, &r, L);
    cbo_acb_log(&N, &tmp, L);
    cbo_acb_set_si(&two, 2, L);
    cbo_acb_log(&tmp, &two, L);
    cbo_acb_mul_si(&tmp, &tmp, bits, L);
    cbo_acb_sub(&N, &N, &tmp, L);
    cbo_acb_log(&tmp, &r, L);
    cbo_acb_div(&N, &N, &tmp, L);
    *nball = N;
    if (!cbo_acb_is_finite(&N)) return 2 * bits;
    /* ceil of the upper end: the upper ball is exact, its integer part is read from the top limbs */
    cbo_acb U;
    cbo_acb_zero(&U);
    cbo_arb_upper_ball(&U.re, &N.re, L);
    if (U.re.flags & CBO_ZERO) return 0;
    if (N.re.flags & CBO_SIGN) return 0; /* a negative count: nothing to add */
    if (U.re.exp > 62) return 2 * bits;
    if (U.re.exp <= 0) return 1;
    uint64_t top = ((uint64_t)U.re.m.w[L - 1] << 32) | U.re.m.w[L - 2];
    uint64_t ip = top >> (64 - U.re.exp);
    int frac = (top << U.re.exp) != 0 || U.re.exp > 64;
    for (int i = 0; i < L - 2 && !frac; i++) frac = U.re.m.w[i] != 0;
    return (int64_t)ip + (frac ? 1 : 0);
}
int cbo_radius_flat(int L, int order, int degree, int64_t batch, int shared, const uint32_t *ode_m, const int32_t *ode_t,
                    int64_t n, uint32_t *out_m, int32_t *out_t, int32_t *n_done) {
    int cnt = (order + 1) * (degree + 1);
    cbo_ode o;
    o.order = order;
    o.degree = degree;
    o.valuation = CBO_UNDEFINED;
    o.polys = (cbo_acb *)malloc(sizeof(cbo_acb) * cnt);
    for (int64_t s = 0; s < batch; s++) {
        for (int k = 0; k < cnt; k++) {
            int64_t i = (shared ? 0 : s) * cnt + k;
            cbo_load(&o.polys[k].re, ode_m + (2 * i) * L, ode_t + (2 * i) * 4, L);
            cbo_load(&o.polys[k].im, ode_m + (2 * i + 1) * L, ode_t + (2 * i + 1) * 4, L);
        }
        cbo_acb out;
        n_done[s] = cbo_radius_of_convergence(&out, &o, n, L);
        cbo_store(out_m + (2 * s) * L, out_t + (2 * s) * 4, &out.re, L);
        cbo_store(out_m + (2 * s + 1) * L, out_t + (2 * s + 1) * 4, &out.im, L);
    }
    free(o.polys);
    return 0;
}
int cbo_truncation_flat(int L, int64_t batch, const uint32_t *eta_m, const int32_t *eta_t, const uint32_t *al_m,
                        const int32_t *al_t, int64_t bits, uint32_t *n_m, int32_t *n_t, int64_t *n_out) {
    for (int64_t s = 0; s < batch; s++) {
        cbo_acb eta, al, nb;
        cbo_load(&eta.re, eta_m + (2 * s) * L, eta_t + (2 * s) * 4, L);
        cbo_load(&eta.im, eta_m + (2 * s + 1) * L, eta_t + (2 * s + 1) * 4, L);
        cbo_load(&al.re, al_m + (2 * s) * L, al_t + (2 * s) * 4, L);
        cbo_load(&al.im, al_m + (2 * s + 1) * L, al_t + (2 * s + 1) * 4, L);
        n_out[s] = cbo_truncation_order(&nb, &eta, &al, bits, L);
        cbo_store(n_m + (2 * s) * L, n_t + (2 * s) * 4, &nb.re, L);
        cbo_store(n_m + (2 * s + 1) * L, n_t + (2 * s + 1) * 4, &nb.im, L);
    }
    return 0;
}

/* elementwise transcendental ops for the unit tests: which 0 exp, 1 log */
int cbo_ew_trans(int which, int L, int64_t n, uint32_t *zm, int32_t *zt, const uint32_t *xm, const int32_t *xt) {
    for (int64_t i = 0; i < n; i++) {
        cbo_acb x, z;
        cbo_load(&x.re, xm + (2 * i) * L, xt + (2 * i) * 4, L);
        cbo_load(&x.im, xm + (2 * i + 1) * L, xt + (2 * i + 1) * 4, L);
        if (which == 0) cbo_acb_exp(&z, &x, L);
        else cbo_acb_log(&z, &x, L);
        cbo_store(zm + (2 * i) * L, zt + (2 * i) * 4, &z.re, L);
        cbo_store(zm + (2 * i + 1) * L, zt + (2 * i + 1) * 4, &z.im, L);
    }
    return 0;
}
