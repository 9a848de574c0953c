/* cbo_cascade.c -- ORACLE restatement of Cascade's hot-path loops on cbo ball arithmetic.
 * TEST INFRASTRUCTURE ONLY (see cbo.h): the product never links or calls this file.
 *
 * Each function follows the operation ORDER of the reference function it names, so that the
 * rounding sequence (and therefore every ball) is the one the reference would produce on an
 * Arb with the semantics stated in cbo_arith.c.
 */
#include "cbo.h"
#include <stdlib.h>
#include <string.h>

static int64_t clampi(int64_t x, int64_t lo, int64_t hi) { /* reference src/cascade.h:30-38 */
    return x < lo ? lo : (x > hi ? hi : x);
}

/* ---------------------------------------------------------------- operator (src/acb_ode.c) */
cbo_ode *cbo_ode_new(int degree, int order) { /* acb_ode_init_blank, src/acb_ode.c:28-41 */
    cbo_ode *o = (cbo_ode *)calloc(1, sizeof(cbo_ode));
    o->order = order;
    o->degree = degree;
    o->valuation = CBO_UNDEFINED;
    int cnt = (order + 1) * (degree + 1);
    o->polys = (cbo_acb *)malloc(sizeof(cbo_acb) * cnt);
    for (int i = 0; i < cnt; i++) cbo_acb_zero(o->polys + i);
    return o;
}
void cbo_ode_free(cbo_ode *o) {
    if (!o) return;
    free(o->polys);
    free(o);
}
#define COEFF(o, i, j) ((o)->polys + (i) * ((o)->degree + 1) + (j))

int cbo_ode_valuation(cbo_ode *o) { /* acb_ode_valuation, src/acb_ode.c:159-181 */
    if (o->valuation != CBO_UNDEFINED) return o->valuation;
    int val = o->degree;
    for (int i = 0; i <= o->order; i++) {
        int z = 0;
        while (z <= o->degree && cbo_acb_is_zero(COEFF(o, i, z))) z++;
        if (z - i < val) val = z - i;
    }
    o->valuation = val;
    return val;
}

/* examples, src/examples.c:3-44 */
void cbo_ode_legendre(cbo_ode *o, uint64_t n, int L) {
    cbo_acb_set_si(COEFF(o, 0, 0), (int64_t)(n * (n + 1)), L);
    cbo_acb_set_si(COEFF(o, 1, 1), -2, L);
    cbo_acb_set_si(COEFF(o, 2, 0), 1, L);
    cbo_acb_set_si(COEFF(o, 2, 2), -1, L);
}
void cbo_ode_bessel(cbo_ode *o, cbo_acb *nu, int L) {
    cbo_acb_set_si(COEFF(o, 2, 2), 1, L);
    cbo_acb_set_si(COEFF(o, 1, 1), 1, L);
    cbo_acb_set_si(COEFF(o, 0, 2), 1, L);
    cbo_acb_mul(nu, nu, nu, L); /* the reference squares its argument in place, :20 */
    cbo_acb_neg(COEFF(o, 0, 0), nu);
}
void cbo_ode_hypgeom(cbo_ode *o, const cbo_acb *a, const cbo_acb *b, const cbo_acb *c, int L) {
    cbo_acb t;
    cbo_acb_set_si(COEFF(o, 2, 1), 1, L);
    cbo_acb_set_si(COEFF(o, 2, 2), -1, L);
    *COEFF(o, 1, 0) = *c;
    cbo_acb_one(&t, L);
    cbo_acb_add(&t, &t, a, L);
    cbo_acb_add(&t, &t, b, L);
    cbo_acb_neg(COEFF(o, 1, 1), &t);
    cbo_acb_mul(&t, a, b, L);
    cbo_acb_neg(COEFF(o, 0, 0), &t);
}

/* acb_ode_shift, src/acb_ode.c:124-135.  Each row is Taylor-shifted with the Horner scheme
 * (Arb's _acb_poly_taylor_shift_horner, the algorithm Arb selects for these 3-term rows):
 *   for i = n-2 .. 0: for j = i .. n-2: p[j] += p[j+1] * a        (acb_addmul) */
void cbo_ode_shift(cbo_ode *out, cbo_ode *in, const cbo_acb *a, int L) {
    int cnt = (in->order + 1) * (in->degree + 1);
    if (out != in) {
        out->order = in->order;
        out->degree = in->degree;
        memcpy(out->polys, in->polys, sizeof(cbo_acb) * cnt);
        out->valuation = cbo_ode_valuation(in);
    }
    if (cbo_acb_is_zero(a) || in->degree == 0) return;
    int n = out->degree + 1;
    for (int row = 0; row <= out->order; row++) {
        cbo_acb *p = COEFF(out, row, 0);
        for (int i = n - 2; i >= 0; i--)
            for (int j = i; j <= n - 2; j++) cbo_acb_addmul(p + j, p + j + 1, a, L);
    }
    out->valuation = CBO_UNDEFINED;
}

/* ---------------------------------------------------------------- Fuchs (src/fuchs_solver.c:96-145) */
int cbo_solve_fuchs(cbo_acb *res, cbo_ode *o, int64_t n, int L) {
    int v = cbo_ode_valuation(o);
    if (v >= 0) return -1;
    cbo_acb t1, t2, acc;
    for (int64_t bmax = -v; bmax <= n; bmax++) {
        cbo_acb_zero(&acc);
        int64_t e = bmax + v;
        int64_t bmin = clampi(e - o->degree, 0, bmax);
        t2 = *COEFF(o, 0, e - bmin);
        int64_t b = bmin;
        do {
            t1 = res[b];
            cbo_acb_mul(&t2, &t1, &t2, L);   /* :118 */
            cbo_acb_sub(&acc, &acc, &t2, L); /* :119 */
            cbo_acb_zero(&t2);
            int64_t fac = 1;
            b++;
            int64_t imin = clampi(b - e, 0, -v);
            int64_t imax = clampi(b - bmin, 0, o->order);
            for (int64_t i = imin; i <= imax; i++) { /* :127-133 */
                t1 = *COEFF(o, i, i + e - b);
                cbo_acb_mul_si(&t1, &t1, fac, L);
                cbo_acb_add(&t2, &t2, &t1, L);
                fac *= (b - i);
            }
            /* rising factorial (b-imin+1)...(b), imin factors, :134 */
            fac = 1;
            for (int64_t k = 0; k < imin; k++) fac *= (b - imin + 1 + k);
            cbo_acb_mul_si(&t2, &t2, fac, L); /* :135 */
        } while (b != bmax);
        cbo_acb_div(&acc, &acc, &t2, L); /* :137 */
        res[bmax] = acc;
    }
    return 0;
}

/* ---------------------------------------------------------------- Frobenius (src/frobenius_solver.c) */
void cbo_indicial_polynomial_evaluate(cbo_acb *result, cbo_ode *o, int64_t nu, const cbo_acb *rho,
                                      int64_t shift, int L) { /* :41-70 */
    nu += cbo_ode_valuation(o);
    if (nu > o->degree) {
        cbo_acb_zero(result);
        return;
    }
    cbo_acb t, out;
    cbo_acb_zero(&out);
    for (int64_t lam = clampi(o->degree - nu, 0, o->order); lam >= 0; lam--) {
        cbo_acb_add_si(&t, rho, shift - lam, L);
        cbo_acb_mul(&out, &out, &t, L);
        if (lam + nu < 0) continue;
        cbo_acb_add(&out, &out, COEFF(o, lam, lam + nu), L);
    }
    *result = out;
}

/* _acb_ode_solve_frobenius, :72-121.  rhs == NULL or all-zero rhs polynomial: g_0 = 1. */
void cbo_solve_frobenius1(cbo_acb *res, cbo_ode *o, const cbo_acb *rho_in, const cbo_acb *rhs,
                          int64_t rhs_len, int64_t n, int L) {
    cbo_acb g_new, ind, g_i, rho = *rho_in, zero;
    cbo_acb_zero(&zero);
    int rhs_zero = 1;
    for (int64_t i = 0; i < rhs_len; i++)
        if (!cbo_acb_is_zero(rhs + i)) rhs_zero = 0;
    if (rhs_zero) {
        cbo_acb_one(res, L);
        for (int64_t i = 1; i <= n; i++) cbo_acb_zero(res + i);
    } else {
        int v = cbo_ode_valuation(o);
        cbo_acb_add_si(&rho, &rho, -(int64_t)v, L); /* acb_sub_si(rho, rho, v), :89 */
        g_new = rhs[0];
        cbo_indicial_polynomial_evaluate(&ind, o, 0, &rho, 0, L);
        cbo_acb_div(&g_new, &g_new, &ind, L);
        res[0] = g_new;
    }
    for (int64_t nu = 1; nu <= n; nu++) {
        g_new = (!rhs_zero && nu < rhs_len) ? rhs[nu] : zero;
        int64_t i = clampi(nu, 1, o->degree);
        cbo_indicial_polynomial_evaluate(&ind, o, i, &rho, nu - i, L);
        do {
            g_i = res[nu - i];
            cbo_acb_mul(&ind, &ind, &g_i, L);
            cbo_acb_sub(&g_new, &g_new, &ind, L);
            i--;
            cbo_indicial_polynomial_evaluate(&ind, o, i, &rho, nu - i, L);
        } while (i > 0);
        cbo_acb_div(&g_new, &g_new, &ind, L);
        res[nu] = g_new;
    }
}

/* ---------------------------------------------------------------- Horner
 * acb_poly_evaluate as used at src/acb_ode_solution.c:40,45,72,103 (Arb's
 * _acb_poly_evaluate_horner: t = f[len-1]; t = t*a + f[i]), extended with the first nder-1
 * normalised derivatives  out[k] = f^(k)(a)/k!  (the only part of acb_poly_taylor_shift at
 * src/fuchs_solver.c:159 that the next solve reads):
 *   for j = len-2 .. 0:  for k = nder-1 .. 1: out[k] = out[k]*a + out[k-1];  out[0] = out[0]*a + f[j] */
void cbo_poly_evaluate_horner(cbo_acb *out, const cbo_acb *f, int64_t len, const cbo_acb *a,
                              int nder, int L) {
    for (int k = 0; k < nder; k++) cbo_acb_zero(out + k);
    if (len <= 0) return;
    out[0] = f[len - 1];
    for (int64_t j = len - 2; j >= 0; j--) {
        for (int k = nder - 1; k >= 1; k--) {
            cbo_acb_mul(out + k, out + k, a, L);
            cbo_acb_add(out + k, out + k, out + k - 1, L);
        }
        cbo_acb_mul(out, out, a, L);
        cbo_acb_add(out, out, f + j, L);
    }
}

/* ---------------------------------------------------------------- residual check
 * acb_ode_apply / acb_ode_solves, src/acb_ode.c:185-240: out = sum_i p_i * in^(i) with classical
 * polynomial products; the reference's own acceptance criterion for both solvers. */
void cbo_ode_apply(cbo_acb *out, int64_t out_len, cbo_ode *o, const cbo_acb *in, int64_t len_in,
                   int L) {
    cbo_acb *der = (cbo_acb *)malloc(sizeof(cbo_acb) * (len_in + 1));
    cbo_acb t;
    memcpy(der, in, sizeof(cbo_acb) * len_in);
    int64_t dl = len_in;
    for (int64_t k = 0; k < out_len; k++) cbo_acb_zero(out + k);
    for (int i = 0; i <= o->order; i++) {
        for (int j = 0; j <= o->degree; j++) {
            const cbo_acb *p = COEFF(o, i, j);
            if (cbo_acb_is_zero(p)) continue;
            for (int64_t k = 0; k < dl && j + k < out_len; k++) {
                cbo_acb_mul(&t, p, der + k, L);
                cbo_acb_add(out + j + k, out + j + k, &t, L);
            }
        }
        /* derivative */
        for (int64_t k = 1; k < dl; k++) cbo_acb_mul_si(der + k - 1, der + k, k, L);
        if (dl > 0) dl--;
    }
    free(der);
}
int cbo_ode_solves(cbo_ode *o, const cbo_acb *res, int64_t len, int64_t deg, int L) {
    int64_t out_len = len + o->degree;
    cbo_acb *out = (cbo_acb *)malloc(sizeof(cbo_acb) * (out_len > 0 ? out_len : 1));
    cbo_ode_apply(out, out_len, o, res, len, L);
    int ok = 1;
    for (int64_t i = 0; i < deg && i < out_len; i++) {
        if (!cbo_acb_is_finite(out + i) || !cbo_acb_contains_zero(out + i)) {
            ok = 0;
            break;
        }
    }
    free(out);
    return ok;
}


/* ---------------------------------------------------------------- tiny pthread parallel-for */
#include <pthread.h>
typedef void (*pf_body)(int64_t idx, void *ctx);
typedef struct {
    pf_body body;
    void *ctx;
    int64_t n;
    int64_t next;
    pthread_mutex_t mu;
} pf_state;
static void *pf_worker(void *arg) {
    pf_state *st = (pf_state *)arg;
    for (;;) {
        pthread_mutex_lock(&st->mu);
        int64_t i = st->next++;
        pthread_mutex_unlock(&st->mu);
        if (i >= st->n) break;
        st->body(i, st->ctx);
    }
    return NULL;
}
static void parallel_for(int64_t n, int nthreads, pf_body body, void *ctx) {
    pf_state st;
    st.body = body;
    st.ctx = ctx;
    st.n = n;
    st.next = 0;
    pthread_mutex_init(&st.mu, NULL);
    if (nthreads <= 1 || n <= 1) {
        pf_worker(&st);
    } else {
        if (nthreads > 256) nthreads = 256;
        pthread_t th[256];
        for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, pf_worker, &st);
        for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    }
    pthread_mutex_destroy(&st.mu);
}

/* ---------------------------------------------------------------- flat (ctypes) entry points */
static void load_acb(cbo_acb *z, const uint32_t *m, const int32_t *t, int64_t i, int L) {
    cbo_load(&z->re, m + (2 * i) * L, t + (2 * i) * 4, L);
    cbo_load(&z->im, m + (2 * i + 1) * L, t + (2 * i + 1) * 4, L);
}
static void store_acb(uint32_t *m, int32_t *t, int64_t i, const cbo_acb *z, int L) {
    cbo_store(m + (2 * i) * L, t + (2 * i) * 4, &z->re, L);
    cbo_store(m + (2 * i + 1) * L, t + (2 * i + 1) * 4, &z->im, L);
}
static cbo_ode *load_ode(int L, int order, int degree, const uint32_t *m, const int32_t *t,
                         int64_t inst) {
    cbo_ode *o = cbo_ode_new(degree, order);
    int cnt = (order + 1) * (degree + 1);
    for (int k = 0; k < cnt; k++) load_acb(o->polys + k, m, t, inst * cnt + k, L);
    return o;
}

/* res arrays: [batch][n+1] complex balls; on entry the first -v of each row are the initial values.
 * ode arrays: [batch or 1][(order+1)(degree+1)] complex balls. */
typedef struct {
    int L, order, degree, shared_ode, rc;
    int64_t n;
    const uint32_t *ode_m, *rho_m;
    const int32_t *ode_t, *rho_t;
    uint32_t *res_m;
    int32_t *res_t;
} solve_ctx;

static void fuchs_body(int64_t s, void *vc) {
    solve_ctx *c = (solve_ctx *)vc;
    int L = c->L;
    int64_t n = c->n;
    cbo_ode *o = load_ode(L, c->order, c->degree, c->ode_m, c->ode_t, c->shared_ode ? 0 : s);
    cbo_acb *res = (cbo_acb *)malloc(sizeof(cbo_acb) * (n + 1));
    for (int64_t k = 0; k <= n; k++) load_acb(res + k, c->res_m, c->res_t, s * (n + 1) + k, L);
    if (cbo_solve_fuchs(res, o, n, L) != 0) c->rc = -1;
    for (int64_t k = 0; k <= n; k++) store_acb(c->res_m, c->res_t, s * (n + 1) + k, res + k, L);
    free(res);
    cbo_ode_free(o);
}
int cbo_fuchs_batch(int L, int order, int degree, int64_t n, int64_t batch, int shared_ode,
                    const uint32_t *ode_m, const int32_t *ode_t, uint32_t *res_m, int32_t *res_t,
                    int nthreads) {
    solve_ctx c = {L, order, degree, shared_ode, 0, n, ode_m, NULL, ode_t, NULL, res_m, res_t};
    parallel_for(batch, nthreads, fuchs_body, &c);
    return c.rc;
}

static void frob_body(int64_t s, void *vc) {
    solve_ctx *c = (solve_ctx *)vc;
    int L = c->L;
    int64_t n = c->n;
    cbo_ode *o = load_ode(L, c->order, c->degree, c->ode_m, c->ode_t, c->shared_ode ? 0 : s);
    cbo_acb *res = (cbo_acb *)malloc(sizeof(cbo_acb) * (n + 1));
    cbo_acb rho;
    load_acb(&rho, c->rho_m, c->rho_t, s, L);
    cbo_solve_frobenius1(res, o, &rho, NULL, 0, n, L);
    for (int64_t k = 0; k <= n; k++) store_acb(c->res_m, c->res_t, s * (n + 1) + k, res + k, L);
    free(res);
    cbo_ode_free(o);
}
int cbo_frobenius1_batch(int L, int order, int degree, int64_t n, int64_t batch, int shared_ode,
                         const uint32_t *ode_m, const int32_t *ode_t, const uint32_t *rho_m,
                         const int32_t *rho_t, uint32_t *res_m, int32_t *res_t, int nthreads) {
    solve_ctx c = {L, order, degree, shared_ode, 0, n, ode_m, rho_m, ode_t, rho_t, res_m, res_t};
    parallel_for(batch, nthreads, frob_body, &c);
    return 0;
}

/* f: one series of len complex balls shared by all points; out: [npts][nder] */
typedef struct {
    int L, nder;
    int64_t len;
    const cbo_acb *f;
    const uint32_t *a_m;
    const int32_t *a_t;
    uint32_t *out_m;
    int32_t *out_t;
} horner_ctx;
static void horner_body(int64_t p, void *vc) {
    horner_ctx *c = (horner_ctx *)vc;
    cbo_acb a, out[8];
    load_acb(&a, c->a_m, c->a_t, p, c->L);
    cbo_poly_evaluate_horner(out, c->f, c->len, &a, c->nder, c->L);
    for (int k = 0; k < c->nder; k++) store_acb(c->out_m, c->out_t, p * c->nder + k, out + k, c->L);
}
int cbo_horner_batch(int L, int64_t len, int64_t npts, int nder, const uint32_t *f_m,
                     const int32_t *f_t, const uint32_t *a_m, const int32_t *a_t, uint32_t *out_m,
                     int32_t *out_t, int nthreads) {
    if (nder < 1 || nder > 8) return -1;
    cbo_acb *f = (cbo_acb *)malloc(sizeof(cbo_acb) * (len > 0 ? len : 1));
    for (int64_t k = 0; k < len; k++) load_acb(f + k, f_m, f_t, k, L);
    horner_ctx c = {L, nder, len, f, a_m, a_t, out_m, out_t};
    parallel_for(npts, nthreads, horner_body, &c);
    free(f);
    return 0;
}

int cbo_ode_shift_flat(int L, int order, int degree, const uint32_t *in_m, const int32_t *in_t,
                       const uint32_t *a_m, const int32_t *a_t, uint32_t *out_m, int32_t *out_t) {
    cbo_ode *in = load_ode(L, order, degree, in_m, in_t, 0);
    cbo_ode *out = cbo_ode_new(degree, order);
    cbo_acb a;
    load_acb(&a, a_m, a_t, 0, L);
    cbo_ode_shift(out, in, &a, L);
    int cnt = (order + 1) * (degree + 1);
    for (int k = 0; k < cnt; k++) store_acb(out_m, out_t, k, out->polys + k, L);
    cbo_ode_free(in);
    cbo_ode_free(out);
    return 0;
}

int cbo_ode_solves_flat(int L, int order, int degree, const uint32_t *ode_m, const int32_t *ode_t,
                        const uint32_t *res_m, const int32_t *res_t, int64_t len, int64_t deg) {
    cbo_ode *o = load_ode(L, order, degree, ode_m, ode_t, 0);
    cbo_acb *res = (cbo_acb *)malloc(sizeof(cbo_acb) * (len > 0 ? len : 1));
    for (int64_t k = 0; k < len; k++) load_acb(res + k, res_m, res_t, k, L);
    int ok = cbo_ode_solves(o, res, len, deg, L);
    free(res);
    cbo_ode_free(o);
    return ok;
}

/* residual of one series: out[0 .. len + degree) */
int cbo_ode_apply_flat(int L, int order, int degree, const uint32_t *ode_m, const int32_t *ode_t,
                       const uint32_t *res_m, const int32_t *res_t, int64_t len, uint32_t *out_m, int32_t *out_t) {
    cbo_ode *o = load_ode(L, order, degree, ode_m, ode_t, 0);
    int64_t out_len = len + degree;
    cbo_acb *res = (cbo_acb *)malloc(sizeof(cbo_acb) * (len > 0 ? len : 1));
    cbo_acb *out = (cbo_acb *)malloc(sizeof(cbo_acb) * (out_len > 0 ? out_len : 1));
    for (int64_t k = 0; k < len; k++) load_acb(res + k, res_m, res_t, k, L);
    cbo_ode_apply(out, out_len, o, res, len, L);
    for (int64_t k = 0; k < out_len; k++) store_acb(out_m, out_t, k, out + k, L);
    free(res);
    free(out);
    cbo_ode_free(o);
    return 0;
}

/* acb_poly_taylor_shift (reference src/fuchs_solver.c:159) in the form Arb's _acb_poly_taylor_shift_horner has:
 *   for i = len-2 .. 0: for j = i .. len-2: p[j] += p[j+1] * a     (acb_addmul)
 * (Arb switches to divide-and-conquer / convolution for long series -- enclosures of the same numbers with other
 * rounding sequences; which one runs depends on the unpinned Arb version, so the Horner form is the stated one.)
 * The first nder coefficients are, operation for operation, what cbo_poly_evaluate_horner(nder) produces. */
void cbo_poly_taylor_shift_horner(cbo_acb *p, int64_t len, const cbo_acb *a, int L) {
    if (cbo_acb_is_zero(a)) return;
    for (int64_t i = len - 2; i >= 0; i--)
        for (int64_t j = i; j <= len - 2; j++) cbo_acb_addmul(p + j, p + j + 1, a, L);
}
/* f: [batch][len] in place; a: [batch] or one ball (shared_a) */
int cbo_taylor_shift_flat(int L, int64_t len, int64_t batch, uint32_t *f_m, int32_t *f_t, const uint32_t *a_m,
                          const int32_t *a_t, int shared_a) {
    cbo_acb *p = (cbo_acb *)malloc(sizeof(cbo_acb) * (len > 0 ? len : 1));
    for (int64_t s = 0; s < batch; s++) {
        cbo_acb a;
        load_acb(&a, a_m, a_t, shared_a ? 0 : s, L);
        for (int64_t k = 0; k < len; k++) load_acb(p + k, f_m, f_t, s * len + k, L);
        cbo_poly_taylor_shift_horner(p, len, &a, L);
        for (int64_t k = 0; k < len; k++) store_acb(f_m, f_t, s * len + k, p + k, L);
    }
    free(p);
    return 0;
}

/* analytic_continuation (reference src/fuchs_solver.c:147-163) for `batch` solutions of one operator
 * along one path.  y: [batch][order] Taylor coefficients at path[0] in, at the last corner out.
 * Per corner: acb_ode_shift, acb_ode_solve_fuchs (n terms; the valuation is that of the SHIFTED operator, as
 * acb_ode_solve_fuchs computes it at :105 -- so a corner where the operator is singular but of negative valuation
 * reads only -v initial values), then the first `order` coefficients of the series re-expanded at the next corner,
 * by Horner evaluation with derivatives (the rest of acb_poly_taylor_shift's output is overwritten by
 * the next solve before it is read).  full_m/full_t (may be NULL): [batch][n+1], the COMPLETE re-expanded series
 * the reference leaves in res after the last step (:159).  Returns -2 when a corner has valuation >= 0. */
int cbo_continuation_series_flat(int L, int order, int degree, int64_t n, int64_t batch, int64_t path_len,
                                 const uint32_t *ode_m, const int32_t *ode_t, const uint32_t *path_m,
                                 const int32_t *path_t, uint32_t *y_m, int32_t *y_t, uint32_t *full_m,
                                 int32_t *full_t) {
    cbo_ode *o = load_ode(L, order, degree, ode_m, ode_t, 0);
    cbo_ode *sh = cbo_ode_new(degree, order);
    cbo_acb *res = (cbo_acb *)malloc(sizeof(cbo_acb) * (n + 1));
    int rc = 0;
    for (int64_t t = 0; t + 1 < path_len && rc == 0; t++) {
        cbo_acb c0, c1, a, out[8];
        load_acb(&c0, path_m, path_t, t, L);
        load_acb(&c1, path_m, path_t, t + 1, L);
        cbo_ode_shift(sh, o, &c0, L);
        sh->valuation = CBO_UNDEFINED;
        if (cbo_ode_valuation(sh) >= 0) {
            rc = -2;
            break;
        }
        cbo_acb_sub(&a, &c1, &c0, L);
        const int last = (t + 2 == path_len);
        for (int64_t s = 0; s < batch; s++) {
            for (int k = 0; k < order; k++) load_acb(res + k, y_m, y_t, s * order + k, L);
            if (cbo_solve_fuchs(res, sh, n, L) != 0) rc = -1;
            if (last && full_m) {
                cbo_poly_taylor_shift_horner(res, n + 1, &a, L);
                for (int64_t k = 0; k <= n; k++) store_acb(full_m, full_t, s * (n + 1) + k, res + k, L);
                for (int k = 0; k < order; k++) store_acb(y_m, y_t, s * order + k, res + k, L);
            } else {
                cbo_poly_evaluate_horner(out, res, n + 1, &a, order, L);
                for (int k = 0; k < order; k++) store_acb(y_m, y_t, s * order + k, out + k, L);
            }
        }
    }
    free(res);
    cbo_ode_free(o);
    cbo_ode_free(sh);
    return rc;
}
int cbo_continuation_flat(int L, int order, int degree, int64_t n, int64_t batch, int64_t path_len,
                          const uint32_t *ode_m, const int32_t *ode_t, const uint32_t *path_m,
                          const int32_t *path_t, uint32_t *y_m, int32_t *y_t) {
    return cbo_continuation_series_flat(L, order, degree, n, batch, path_len, ode_m, ode_t, path_m, path_t, y_m, y_t,
                                        NULL, NULL);
}

/* which: 0 legendre(n), 1 bessel(par[0]) (par[0] is overwritten with nu^2), 2 hypgeom(par[0..2]) */
int cbo_example_flat(int which, int L, uint64_t n, const uint32_t *par_m, const int32_t *par_t,
                     uint32_t *out_m, int32_t *out_t) {
    cbo_ode *o = cbo_ode_new(2, 2);
    cbo_acb p[3];
    if (which == 0) {
        cbo_ode_legendre(o, n, L);
    } else if (which == 1) {
        load_acb(&p[0], par_m, par_t, 0, L);
        cbo_ode_bessel(o, &p[0], L);
    } else if (which == 2) {
        for (int k = 0; k < 3; k++) load_acb(&p[k], par_m, par_t, k, L);
        cbo_ode_hypgeom(o, &p[0], &p[1], &p[2], L);
    } else {
        cbo_ode_free(o);
        return -1;
    }
    for (int k = 0; k < 9; k++) store_acb(out_m, out_t, k, o->polys + k, L);
    cbo_ode_free(o);
    return 0;
}

/* ================================================================ general Frobenius (M > 1)
 * acb_ode_solve_frobenius, reference src/frobenius_solver.c:131-205, with indicial_polynomial
 * (:4-39) and the solution bookkeeping _acb_ode_solution_update / _extend / _normalize
 * (src/acb_ode_solution.c:60-123).  Polynomials in rho are acb_poly objects in the reference;
 * cbo_poly keeps Arb's conventions: a normalised length (trailing EXACT zeros stripped after
 * every operation that Arb normalises after), reads beyond the length give zero.
 *
 * Arb sub-algorithms the reference reaches through acb_poly_* (Arb version unpinned, so these are
 * restated choices -- each is an algorithm Arb ships):
 *   acb_poly_mul        -> longer operand first, _acb_poly_mullow_classical in its acb_mul /
 *                          acb_addmul form: res[i] = a[i] b[0]; res[la-1+j] = a[la-1] b[j];
 *                          then res[i+j] += a[i] b[j]  (i < la-1, 1 <= j < lb).  Arb picks the
 *                          classical algorithm whenever the shorter operand has <= 15 terms at
 *                          1024 bits; here it is the indicial polynomial, order+1 terms.
 *   acb_poly_evaluate   -> _acb_poly_evaluate_horner (cbo_poly_evaluate_horner above)
 *   acb_poly_scalar_mul/div, acb_poly_add/sub, acb_poly_derivative -> coefficientwise
 *   acb_div_si(x, n)    -> acb_div by the exact ball n */
typedef struct {
    cbo_acb *c;
    int64_t len, alloc;
} cbo_poly;

static void poly_init(cbo_poly *p) { p->c = NULL; p->len = p->alloc = 0; }
static void poly_clear(cbo_poly *p) { free(p->c); poly_init(p); }
static void poly_fit(cbo_poly *p, int64_t n) {
    if (n > p->alloc) {
        int64_t na = n > 2 * p->alloc ? n : 2 * p->alloc;
        p->c = (cbo_acb *)realloc(p->c, sizeof(cbo_acb) * na);
        for (int64_t i = p->alloc; i < na; i++) cbo_acb_zero(p->c + i);
        p->alloc = na;
    }
}
static void poly_normalise(cbo_poly *p) {
    while (p->len > 0 && cbo_acb_is_zero(p->c + p->len - 1)) p->len--;
}
static void poly_zero(cbo_poly *p) { p->len = 0; }
static void poly_one(cbo_poly *p, int L) {
    poly_fit(p, 1);
    cbo_acb_one(p->c, L);
    p->len = 1;
}
static void poly_get(cbo_acb *z, const cbo_poly *p, int64_t n) { /* acb_poly_get_coeff_acb */
    if (n < p->len) *z = p->c[n]; else cbo_acb_zero(z);
}
static void poly_set_coeff(cbo_poly *p, int64_t n, const cbo_acb *x) { /* acb_poly_set_coeff_acb */
    poly_fit(p, n + 1);
    if (n + 1 > p->len) {
        for (int64_t i = p->len; i < n; i++) cbo_acb_zero(p->c + i);
        p->len = n + 1;
    }
    p->c[n] = *x;
    poly_normalise(p);
}
static void poly_set(cbo_poly *z, const cbo_poly *x) {
    poly_fit(z, x->len);
    memcpy(z->c, x->c, sizeof(cbo_acb) * x->len);
    z->len = x->len;
}
static void poly_mul(cbo_poly *res, const cbo_poly *p1, const cbo_poly *p2, int L) { /* acb_poly_mul */
    if (p1->len == 0 || p2->len == 0) {
        poly_zero(res);
        return;
    }
    const cbo_poly *a = p1->len >= p2->len ? p1 : p2, *b = p1->len >= p2->len ? p2 : p1;
    int64_t la = a->len, lb = b->len, n = la + lb - 1;
    cbo_poly t;
    poly_init(&t);
    poly_fit(&t, n);
    for (int64_t i = 0; i < la; i++) cbo_acb_mul(t.c + i, a->c + i, b->c, L);
    for (int64_t j = 1; j < lb; j++) cbo_acb_mul(t.c + la - 1 + j, a->c + la - 1, b->c + j, L);
    for (int64_t i = 0; i < la - 1; i++)
        for (int64_t j = 1; j < lb; j++) cbo_acb_addmul(t.c + i + j, a->c + i, b->c + j, L);
    t.len = n;
    poly_normalise(&t);
    poly_clear(res); /* acb_poly_swap(res, t) */
    *res = t;
}
static void poly_addsub(cbo_poly *res, const cbo_poly *a, const cbo_poly *b, int sub, int L) {
    int64_t la = a->len, lb = b->len, mn = la < lb ? la : lb, mx = la < lb ? lb : la;
    poly_fit(res, mx);
    for (int64_t i = 0; i < mn; i++) {
        if (sub) cbo_acb_sub(res->c + i, a->c + i, b->c + i, L);
        else cbo_acb_add(res->c + i, a->c + i, b->c + i, L);
    }
    for (int64_t i = mn; i < la; i++) res->c[i] = a->c[i];
    for (int64_t i = mn; i < lb; i++) {
        if (sub) cbo_acb_neg(res->c + i, b->c + i);
        else res->c[i] = b->c[i];
    }
    res->len = mx;
    poly_normalise(res);
}
static void poly_scalar_mul(cbo_poly *res, const cbo_poly *p, const cbo_acb *c, int L) {
    poly_fit(res, p->len);
    for (int64_t i = 0; i < p->len; i++) cbo_acb_mul(res->c + i, p->c + i, c, L);
    res->len = p->len;
    poly_normalise(res);
}
static void poly_scalar_div(cbo_poly *res, const cbo_poly *p, const cbo_acb *c, int L) {
    poly_fit(res, p->len);
    for (int64_t i = 0; i < p->len; i++) cbo_acb_div(res->c + i, p->c + i, c, L);
    res->len = p->len;
    poly_normalise(res);
}
static void poly_derivative(cbo_poly *p, int L) { /* in place, as every call site uses it */
    for (int64_t i = 1; i < p->len; i++) cbo_acb_mul_si(p->c + i - 1, p->c + i, i, L);
    if (p->len > 0) p->len--;
    poly_normalise(p);
}
static void poly_evaluate(cbo_acb *y, const cbo_poly *p, const cbo_acb *x, int L) {
    if (p->len == 0) {
        cbo_acb_zero(y);
    } else if (p->len == 1 || cbo_acb_is_zero(x)) {
        *y = p->c[0];
    } else {
        cbo_acb out;
        cbo_poly_evaluate_horner(&out, p->c, p->len, x, 1, L);
        *y = out;
    }
}

/* indicial_polynomial, src/frobenius_solver.c:4-39 */
static void indicial_polynomial(cbo_poly *result, cbo_ode *o, int64_t nu, int64_t shift, int L) {
    poly_zero(result);
    nu += cbo_ode_valuation(o);
    if (nu > o->degree) return;
    cbo_acb t1, t2;
    poly_fit(result, o->order + 1);
    for (int64_t lam = clampi(o->degree - nu, 0, o->order); lam >= 0; lam--) {
        for (int64_t j = result->len; j > 0; j--) {
            poly_get(&t1, result, j);
            poly_get(&t2, result, j - 1);
            cbo_acb_mul_si(&t1, &t1, shift - lam, L);
            cbo_acb_add(&t1, &t2, &t1, L);
            poly_set_coeff(result, j, &t1);
        }
        poly_get(&t1, result, 0);
        cbo_acb_mul_si(&t1, &t1, shift - lam, L);
        if (lam + nu >= 0) cbo_acb_add(&t1, &t1, COEFF(o, lam, lam + nu), L);
        poly_set_coeff(result, 0, &t1);
    }
}

typedef struct {
    cbo_acb rho;
    int64_t mul, M;
    cbo_poly *gens;
} cbo_solution;

static void acb_div_si(cbo_acb *z, const cbo_acb *x, int64_t n, int L) {
    cbo_acb k;
    cbo_acb_set_si(&k, n, L);
    cbo_acb_div(z, x, &k, L);
}

/* _acb_ode_solution_update, src/acb_ode_solution.c:60-95 (consumes f) */
static void solution_update(cbo_solution *sol, cbo_poly *f, int L) {
    int64_t M = sol->M, binom = 1;
    cbo_acb *F = (cbo_acb *)malloc(sizeof(cbo_acb) * (M > 0 ? M : 1));
    for (int64_t k = 0; k < M; k++) {
        poly_evaluate(F + k, f, &sol->rho, L);
        poly_derivative(f, L);
        cbo_acb_mul_si(F + k, F + k, binom, L);
        binom = (binom * (M - 1 - k)) / (k + 1);
    }
    for (int64_t n = M - 1; n >= 0; n--) {
        poly_scalar_mul(sol->gens + n, sol->gens + n, F + 0, L);
        for (int64_t k = 1; k <= n; k++) {
            poly_scalar_mul(f, sol->gens + (n - k), F + k, L);
            poly_addsub(sol->gens + n, sol->gens + n, f, 0, L);
            cbo_acb_mul_si(F + k, F + k, n - k, L);
            acb_div_si(F + k, F + k, n, L);
        }
    }
    poly_zero(f);
    free(F);
}
/* _acb_ode_solution_extend, :97-109 (consumes g_nu) */
static void solution_extend(cbo_solution *sol, int64_t nu, cbo_poly *g_nu, int L) {
    cbo_acb t;
    for (int64_t i = 0; i < sol->M; i++) {
        poly_evaluate(&t, g_nu, &sol->rho, L);
        poly_set_coeff(sol->gens + i, nu, &t);
        poly_derivative(g_nu, L);
    }
    poly_zero(g_nu);
}
/* _acb_ode_solution_normalize, :111-123 */
static void solution_normalize(cbo_solution *sol, int L) {
    cbo_acb t;
    int64_t alpha = sol->M - sol->mul;
    poly_get(&t, sol->gens + alpha, 0);
    for (int64_t i = 0; i < sol->M; i++) poly_scalar_div(sol->gens + i, sol->gens + i, &t, L);
}

/* acb_ode_solve_frobenius for M > 1, src/frobenius_solver.c:131-205 */
static void solve_frobenius_general(cbo_solution *sol, cbo_ode *o, int64_t n, int L) {
    int d = o->degree;
    cbo_acb temp;
    cbo_poly indicial, g_new, *g_rho = (cbo_poly *)malloc(sizeof(cbo_poly) * (d > 0 ? d : 1));
    poly_init(&indicial);
    poly_init(&g_new);
    for (int i = 0; i < d; i++) poly_init(g_rho + i);
    poly_one(g_rho, L);
    for (int64_t i = 0; i < sol->M; i++) {
        poly_zero(sol->gens + i);
        poly_fit(sol->gens + i, n + 1);
    }
    poly_one(sol->gens, L);
    for (int64_t nu = 1; nu <= n; nu++) {
        int64_t i = clampi(nu, 1, d);
        indicial_polynomial(&indicial, o, i, nu - i, L); /* :156 */
        do {
            poly_mul(&indicial, &indicial, g_rho + (i - 1), L);
            poly_addsub(&g_new, &g_new, &indicial, 1, L);
            i--;
            indicial_polynomial(&indicial, o, i, nu - i, L);
        } while (i > 0);
        cbo_indicial_polynomial_evaluate(&temp, o, 0, &sol->rho, nu, L); /* :170 */
        if (!cbo_acb_contains_zero(&temp)) {
            poly_scalar_div(&indicial, &indicial, &temp, L);
            poly_scalar_div(&g_new, &g_new, &temp, L);
        }
        int all_zero = g_new.len == 0;
        for (i = d - 1; i > 0; i--) {
            poly_mul(g_rho + i, g_rho + (i - 1), &indicial, L);
            all_zero &= g_rho[i].len == 0;
        }
        poly_set(g_rho, &g_new);
        solution_update(sol, &indicial, L);
        solution_extend(sol, nu, &g_new, L);
        if (all_zero) break;
    }
    solution_normalize(sol, L);
    poly_clear(&indicial);
    poly_clear(&g_new);
    for (int i = 0; i < d; i++) poly_clear(g_rho + i);
    free(g_rho);
}

/* _acb_ode_solve_frobenius with right-hand sides: rhs [batch][rhs_len], rho [batch], res [batch][n+1] */
int cbo_frobenius1_rhs_batch(int L, int order, int degree, int64_t n, int64_t batch, int shared_ode,
                             const uint32_t *ode_m, const int32_t *ode_t, const uint32_t *rho_m,
                             const int32_t *rho_t, const uint32_t *rhs_m, const int32_t *rhs_t, int64_t rhs_len,
                             uint32_t *res_m, int32_t *res_t) {
    for (int64_t s = 0; s < batch; s++) {
        cbo_ode *o = load_ode(L, order, degree, ode_m, ode_t, shared_ode ? 0 : s);
        cbo_acb *res = (cbo_acb *)malloc(sizeof(cbo_acb) * (n + 1));
        cbo_acb *rhs = (cbo_acb *)malloc(sizeof(cbo_acb) * (rhs_len > 0 ? rhs_len : 1));
        cbo_acb rho;
        load_acb(&rho, rho_m, rho_t, s, L);
        for (int64_t k = 0; k < rhs_len; k++) load_acb(rhs + k, rhs_m, rhs_t, s * rhs_len + k, L);
        cbo_solve_frobenius1(res, o, &rho, rhs, rhs_len, n, L);
        for (int64_t k = 0; k <= n; k++) store_acb(res_m, res_t, s * (n + 1) + k, res + k, L);
        free(res);
        free(rhs);
        cbo_ode_free(o);
    }
    return 0;
}

/* ---- flat entry points.  Polynomials travel as fixed-capacity coefficient arrays whose length is
 * implied (trailing exact zeros do not count), as on the device interface. */
static void load_poly(cbo_poly *p, const uint32_t *m, const int32_t *t, int64_t first, int64_t cap, int L) {
    poly_fit(p, cap);
    for (int64_t j = 0; j < cap; j++) load_acb(p->c + j, m, t, first + j, L);
    p->len = cap;
    poly_normalise(p);
}
static void store_poly(uint32_t *m, int32_t *t, int64_t first, int64_t cap, const cbo_poly *p, int L) {
    cbo_acb z;
    cbo_acb_zero(&z);
    for (int64_t j = 0; j < cap; j++) store_acb(m, t, first + j, j < p->len ? p->c + j : &z, L);
}
static void sol_alloc(cbo_solution *s, int64_t mul, int64_t M) {
    s->mul = mul;
    s->M = M;
    s->gens = (cbo_poly *)malloc(sizeof(cbo_poly) * (M > 0 ? M : 1));
    for (int64_t i = 0; i < M; i++) poly_init(s->gens + i);
}
static void sol_free(cbo_solution *s) {
    for (int64_t i = 0; i < s->M; i++) poly_clear(s->gens + i);
    free(s->gens);
}

typedef struct {
    int L, order, degree, shared_ode, shared_rho;
    int64_t n, mul, M;
    const uint32_t *ode_m, *rho_m;
    const int32_t *ode_t, *rho_t;
    uint32_t *gens_m;
    int32_t *gens_t;
} frobg_ctx;
static void frobg_body(int64_t s, void *vc) {
    frobg_ctx *c = (frobg_ctx *)vc;
    int L = c->L;
    cbo_ode *o = load_ode(L, c->order, c->degree, c->ode_m, c->ode_t, c->shared_ode ? 0 : s);
    cbo_solution sol;
    sol_alloc(&sol, c->mul, c->M);
    load_acb(&sol.rho, c->rho_m, c->rho_t, c->shared_rho ? 0 : s, L);
    solve_frobenius_general(&sol, o, c->n, L);
    for (int64_t i = 0; i < c->M; i++)
        store_poly(c->gens_m, c->gens_t, (s * c->M + i) * (c->n + 1), c->n + 1, sol.gens + i, L);
    sol_free(&sol);
    cbo_ode_free(o);
}
/* gens: [batch][M][n+1] complex balls out */
int cbo_frobenius_general_batch(int L, int order, int degree, int64_t n, int64_t batch, int shared_ode,
                                const uint32_t *ode_m, const int32_t *ode_t, const uint32_t *rho_m,
                                const int32_t *rho_t, int shared_rho, int64_t mul, int64_t M,
                                uint32_t *gens_m, int32_t *gens_t, int nthreads) {
    frobg_ctx c = {L, order, degree, shared_ode, shared_rho, n, mul, M, ode_m, rho_m, ode_t, rho_t, gens_m, gens_t};
    parallel_for(batch, nthreads, frobg_body, &c);
    return 0;
}

/* indicial_polynomial for one operator: out has order+1 coefficients */
int cbo_indicial_polynomial_flat(int L, int order, int degree, const uint32_t *ode_m, const int32_t *ode_t,
                                 int64_t nu, int64_t shift, uint32_t *out_m, int32_t *out_t) {
    cbo_ode *o = load_ode(L, order, degree, ode_m, ode_t, 0);
    cbo_poly r;
    poly_init(&r);
    indicial_polynomial(&r, o, nu, shift, L);
    store_poly(out_m, out_t, 0, order + 1, &r, L);
    poly_clear(&r);
    cbo_ode_free(o);
    return 0;
}
int cbo_indicial_polynomial_evaluate_flat(int L, int order, int degree, const uint32_t *ode_m,
                                          const int32_t *ode_t, int64_t nu, const uint32_t *rho_m,
                                          const int32_t *rho_t, int64_t shift, uint32_t *out_m, int32_t *out_t) {
    cbo_ode *o = load_ode(L, order, degree, ode_m, ode_t, 0);
    cbo_acb rho, out;
    load_acb(&rho, rho_m, rho_t, 0, L);
    cbo_indicial_polynomial_evaluate(&out, o, nu, &rho, shift, L);
    store_acb(out_m, out_t, 0, &out, L);
    cbo_ode_free(o);
    return 0;
}

/* the three solution bookkeeping steps for a batch: gens [batch][M][cap], f/g [batch][fcap], rho [batch] */
int cbo_solution_op_flat(int which, int L, int64_t batch, int64_t mul, int64_t M, int64_t cap, int64_t nu,
                         const uint32_t *rho_m, const int32_t *rho_t, uint32_t *gens_m, int32_t *gens_t,
                         uint32_t *f_m, int32_t *f_t, int64_t fcap) {
    for (int64_t s = 0; s < batch; s++) {
        cbo_solution sol;
        sol_alloc(&sol, mul, M);
        load_acb(&sol.rho, rho_m, rho_t, s, L);
        for (int64_t i = 0; i < M; i++) load_poly(sol.gens + i, gens_m, gens_t, (s * M + i) * cap, cap, L);
        cbo_poly f;
        poly_init(&f);
        if (which != 2) load_poly(&f, f_m, f_t, s * fcap, fcap, L);
        if (which == 0) solution_extend(&sol, nu, &f, L);
        else if (which == 1) solution_update(&sol, &f, L);
        else solution_normalize(&sol, L);
        if (which != 2) store_poly(f_m, f_t, s * fcap, fcap, &f, L);
        for (int64_t i = 0; i < M; i++) store_poly(gens_m, gens_t, (s * M + i) * cap, cap, sol.gens + i, L);
        poly_clear(&f);
        sol_free(&sol);
    }
    return 0;
}
