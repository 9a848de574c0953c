/* cbo.h -- CPU ORACLE for cascade_b200 (TEST INFRASTRUCTURE, never part of the product path).
 *
 * A plain-C restatement of (i) the Arb ball-arithmetic semantics that Cascade's hot path
 * relies on and (ii) Cascade's own solver loops, used only by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs as the checker and the timed CPU arm.
 *
 * PARITY STATUS: "parity unpinned" at bit level.  Real Arb/FLINT are not present in this
 * image and /root/reference pins no Arb version (reference CMakeLists.txt:4-14), so the
 * arithmetic below is a restatement of Arb's documented semantics:
 *   - midpoints: exact result rounded toward zero to prec bits (ARB_RND = ARF_RND_DOWN),
 *   - radii: 30-bit mantissa magnitudes, every step rounded up,
 *   - one ulp added to the radius when (and only when) the midpoint op was inexact.
 * What IS pinned: the reference's literal known answers (README.md:66-77, tests/legendre.c,
 * tests/bessel.c, tests/hypgeom.c) and exact-rational evaluation of the same recurrences
 * (tests/test_oracle_*.py): every oracle ball must contain the exact Fraction result.
 *
 * Layout of one real ball on this interface ("slot"):
 *   mant : uint32_t[L]   little-endian limbs, normalised (bit 31 of limb L-1 set) or all zero
 *   meta : int32_t[4]    {exp, flags, rman, rexp}  (same meaning as include/cb_format.h)
 * A complex ball is two consecutive slots (re, im).  Arrays are slot-major ("AoS"):
 * mant[slot][limb], meta[slot][4].  prec = 32*L, L even, 2 <= L <= 128.
 */
#ifndef CBO_H
#define CBO_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CBO_MAXL 128 /* 32-bit limbs */

#define CBO_SIGN 1u
#define CBO_ZERO 2u
#define CBO_NAN 4u

typedef struct {
    int32_t exp;
    uint32_t flags;
    uint32_t rman; /* 0 or in [2^29, 2^30) */
    int32_t rexp;
    union {
        uint32_t w[CBO_MAXL];
        uint64_t d[CBO_MAXL / 2];
    } m;
} cbo_arb;

typedef struct {
    cbo_arb re, im;
} cbo_acb;

/* ---- scalar ball arithmetic (L = prec/32) ---- */
void cbo_arb_zero(cbo_arb *z);
void cbo_arb_set_si(cbo_arb *z, int64_t k, int L);
void cbo_arb_neg(cbo_arb *z, const cbo_arb *x);
void cbo_arb_add(cbo_arb *z, const cbo_arb *x, const cbo_arb *y, int L);
void cbo_arb_sub(cbo_arb *z, const cbo_arb *x, const cbo_arb *y, int L);
void cbo_arb_mul(cbo_arb *z, const cbo_arb *x, const cbo_arb *y, int L);
/* z = x1*y1 + (neg2 ? -1 : +1) * x2*y2, ONE rounding of the midpoint */
void cbo_arb_fdot2(cbo_arb *z, const cbo_arb *x1, const cbo_arb *y1, const cbo_arb *x2,
                   const cbo_arb *y2, int neg2, int L);
void cbo_arb_div(cbo_arb *z, const cbo_arb *x, const cbo_arb *y, int L);
void cbo_arb_mul_si(cbo_arb *z, const cbo_arb *x, int64_t k, int L);
void cbo_arb_add_si(cbo_arb *z, const cbo_arb *x, int64_t k, int L);
int cbo_arb_is_zero(const cbo_arb *x);       /* midpoint 0 and radius 0 */
int cbo_arb_contains_zero(const cbo_arb *x); /* |mid| <= rad (exact comparison) */
int cbo_arb_is_finite(const cbo_arb *x);

void cbo_acb_zero(cbo_acb *z);
void cbo_acb_one(cbo_acb *z, int L);
void cbo_acb_set_si(cbo_acb *z, int64_t k, int L);
void cbo_acb_neg(cbo_acb *z, const cbo_acb *x);
void cbo_acb_add(cbo_acb *z, const cbo_acb *x, const cbo_acb *y, int L);
void cbo_acb_sub(cbo_acb *z, const cbo_acb *x, const cbo_acb *y, int L);
void cbo_acb_mul(cbo_acb *z, const cbo_acb *x, const cbo_acb *y, int L);
void cbo_acb_inv(cbo_acb *z, const cbo_acb *x, int L);
void cbo_acb_div(cbo_acb *z, const cbo_acb *x, const cbo_acb *y, int L);
void cbo_acb_mul_si(cbo_acb *z, const cbo_acb *x, int64_t k, int L);
void cbo_acb_add_si(cbo_acb *z, const cbo_acb *x, int64_t k, int L);
void cbo_acb_addmul(cbo_acb *z, const cbo_acb *x, const cbo_acb *y, int L);
int cbo_acb_is_zero(const cbo_acb *x);
int cbo_acb_contains_zero(const cbo_acb *x);
int cbo_acb_is_finite(const cbo_acb *x);

/* ---- slot-array interface used through ctypes ---- */
void cbo_load(cbo_arb *z, const uint32_t *mant, const int32_t *meta, int L);
void cbo_store(uint32_t *mant, int32_t *meta, const cbo_arb *x, int L);

/* elementwise ops over n complex balls (2n slots each); op codes as in include/cascade_b200.h */
int cbo_ew_binary(int op, int L, int64_t n, uint32_t *zm, int32_t *zt, const uint32_t *xm,
                  const int32_t *xt, const uint32_t *ym, const int32_t *yt);
int cbo_ew_scalar_si(int op, int L, int64_t n, uint32_t *zm, int32_t *zt, const uint32_t *xm,
                     const int32_t *xt, const int64_t *k);

/* ---- Cascade restated (see cbo_cascade.c for the reference file:line of each) ---- */
typedef struct {
    int order, degree, valuation; /* valuation: CBO_UNDEFINED until computed */
    cbo_acb *polys;               /* (order+1)*(degree+1), row-major: polys[i*(degree+1)+j] */
} cbo_ode;
#define CBO_UNDEFINED (-0xFFFF)

cbo_ode *cbo_ode_new(int degree, int order);
void cbo_ode_free(cbo_ode *o);
int cbo_ode_valuation(cbo_ode *o);
void cbo_ode_legendre(cbo_ode *o, uint64_t n, int L);
void cbo_ode_bessel(cbo_ode *o, cbo_acb *nu, int L);
void cbo_ode_hypgeom(cbo_ode *o, const cbo_acb *a, const cbo_acb *b, const cbo_acb *c, int L);
void cbo_ode_shift(cbo_ode *out, cbo_ode *in, const cbo_acb *a, int L);

/* res has room for n+1 coefficients; res[0..-v-1] are the initial values. Returns 0, or
 * -1 when valuation >= 0 (the reference is undefined there, src/fuchs_solver.c:108-136). */
int cbo_solve_fuchs(cbo_acb *res, cbo_ode *o, int64_t n, int L);
void cbo_indicial_polynomial_evaluate(cbo_acb *result, cbo_ode *o, int64_t nu, const cbo_acb *rho,
                                      int64_t shift, int L);
/* M == 1 Frobenius; rhs may be NULL / rhs_len 0 for the homogeneous case */
void cbo_solve_frobenius1(cbo_acb *res, cbo_ode *o, const cbo_acb *rho_in, const cbo_acb *rhs,
                          int64_t rhs_len, int64_t n, int L);
/* Horner value of a series and its first nder-1 normalised derivatives at a */
void cbo_poly_evaluate_horner(cbo_acb *out, const cbo_acb *f, int64_t len, const cbo_acb *a,
                              int nder, int L);
/* out[0..len_in+degree) = L applied to the series in[0..len_in) */
void cbo_ode_apply(cbo_acb *out, int64_t out_len, cbo_ode *o, const cbo_acb *in, int64_t len_in,
                   int L);
int cbo_ode_solves(cbo_ode *o, const cbo_acb *res, int64_t len, int64_t deg, int L);

/* ---- flat ctypes entry points: all balls as slot arrays ---- */
int cbo_fuchs_batch(int L, int order, int degree, int64_t n, int64_t batch, int shared_ode,
                    const uint32_t *ode_m, const int32_t *ode_t, uint32_t *res_m, int32_t *res_t,
                    int nthreads);
int cbo_frobenius1_batch(int L, int order, int degree, int64_t n, int64_t batch, int shared_ode,
                         const uint32_t *ode_m, const int32_t *ode_t, const uint32_t *rho_m,
                         const int32_t *rho_t, uint32_t *res_m, int32_t *res_t, int nthreads);
int cbo_horner_batch(int L, int64_t len, int64_t npts, int nder, const uint32_t *f_m,
                     const int32_t *f_t, const uint32_t *a_m, const int32_t *a_t, uint32_t *out_m,
                     int32_t *out_t, int nthreads);
int cbo_ode_shift_flat(int L, int order, int degree, const uint32_t *in_m, const int32_t *in_t,
                       const uint32_t *a_m, const int32_t *a_t, uint32_t *out_m, int32_t *out_t);
int cbo_ode_solves_flat(int L, int order, int degree, const uint32_t *ode_m, const int32_t *ode_t,
                        const uint32_t *res_m, const int32_t *res_t, int64_t len, int64_t deg);
int cbo_ode_apply_flat(int L, int order, int degree, const uint32_t *ode_m, const int32_t *ode_t,
                       const uint32_t *res_m, const int32_t *res_t, int64_t len, uint32_t *out_m, int32_t *out_t);
int cbo_example_flat(int which, int L, uint64_t n, const uint32_t *par_m, const int32_t *par_t,
                     uint32_t *out_m, int32_t *out_t);
void cbo_poly_taylor_shift_horner(cbo_acb *p, int64_t len, const cbo_acb *a, int L);
int cbo_taylor_shift_flat(int L, int64_t len, int64_t batch, uint32_t *f_m, int32_t *f_t, const uint32_t *a_m,
                          const int32_t *a_t, int shared_a);
int cbo_continuation_series_flat(int L, int order, int degree, int64_t n, int64_t batch, int64_t path_len,
                                 const uint32_t *ode_m, const int32_t *ode_t, const uint32_t *path_m,
                                 const int32_t *path_t, uint32_t *y_m, int32_t *y_t, uint32_t *full_m,
                                 int32_t *full_t);
int cbo_continuation_flat(int L, int order, int degree, int64_t n, int64_t batch, int64_t path_len,
                          const uint32_t *ode_m, const int32_t *ode_t, const uint32_t *path_m,
                          const int32_t *path_t, uint32_t *y_m, int32_t *y_t);
int cbo_frobenius1_rhs_batch(int L, int order, int degree, int64_t n, int64_t batch, int shared_ode,
                             const uint32_t *ode_m, const int32_t *ode_t, const uint32_t *rho_m,
                             const int32_t *rho_t, const uint32_t *rhs_m, const int32_t *rhs_t, int64_t rhs_len,
                             uint32_t *res_m, int32_t *res_t);
/* general Frobenius (M = mul + alpha > 1), reference src/frobenius_solver.c:131-205.
 * gens out: [batch][M][n+1] complex balls */
int cbo_frobenius_general_batch(int L, int order, int degree, int64_t n, int64_t batch, int shared_ode,
                                const uint32_t *ode_m, const int32_t *ode_t, const uint32_t *rho_m,
                                const int32_t *rho_t, int shared_rho, int64_t mul, int64_t M,
                                uint32_t *gens_m, int32_t *gens_t, int nthreads);
/* indicial_polynomial (src/frobenius_solver.c:4-39): out = order+1 coefficients */
int cbo_indicial_polynomial_flat(int L, int order, int degree, const uint32_t *ode_m, const int32_t *ode_t,
                                 int64_t nu, int64_t shift, uint32_t *out_m, int32_t *out_t);
int cbo_indicial_polynomial_evaluate_flat(int L, int order, int degree, const uint32_t *ode_m,
                                          const int32_t *ode_t, int64_t nu, const uint32_t *rho_m,
                                          const int32_t *rho_t, int64_t shift, uint32_t *out_m, int32_t *out_t);
/* which: 0 _acb_ode_solution_extend, 1 _update, 2 _normalize (src/acb_ode_solution.c:60-123).
 * gens [batch][M][cap] in/out; f [batch][fcap] in/out (consumed: comes back zero); rho [batch] */
int cbo_solution_op_flat(int which, int L, int64_t batch, int64_t mul, int64_t M, int64_t cap, int64_t nu,
                         const uint32_t *rho_m, const int32_t *rho_t, uint32_t *gens_m, int32_t *gens_t,
                         uint32_t *f_m, int32_t *f_t, int64_t fcap);
/* transcendental part of acb_ode_solution_evaluate (reference src/acb_ode_solution.c:24-58); see cbo_trans.c */
void cbo_acb_exp(cbo_acb *res, const cbo_acb *z, int L);
void cbo_acb_log(cbo_acb *res, const cbo_acb *a, int L);
void cbo_acb_pow_mode(cbo_acb *res, const cbo_acb *a, const cbo_acb *rho, int mode, uint64_t e,
                      const cbo_acb *lg, int L);
int cbo_solution_evaluate_flat(int L, int64_t M, int64_t cap, int64_t npts, const uint32_t *gens_m,
                               const int32_t *gens_t, const uint32_t *rho_m, const int32_t *rho_t,
                               int pow_mode, uint64_t pow_e, const uint32_t *a_m, const int32_t *a_t,
                               uint32_t *out_m, int32_t *out_t);
int cbo_ew_trans(int which, int L, int64_t n, uint32_t *zm, int32_t *zt, const uint32_t *xm, const int32_t *xt);
const char *cbo_backend(void); /* "gmp-mpn" or "plain-c" */

#ifdef __cplusplus
}
#endif
#endif
