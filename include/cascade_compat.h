/* cascade_compat.h -- the reference's public API surface for the hot path, implemented on top of
 * the batched CUDA entry points of include/cascade_b200.h (each call is a batch of one).
 *
 * Same names, argument order and in-place/consuming behaviour as the reference headers
 * (reference src/acb_ode.h:9-67, src/cascade.h:8-26), so that callers such as the README program
 * (reference README.md:31-57) compile unchanged against this header and link against
 * libcascade_b200.so instead of libcascade.so + libarb + libflint.
 *
 * Arb/FLINT are not available in this build environment, so the handful of Arb TYPES that appear
 * in Cascade's signatures (acb_t, acb_poly_t, slong, ...) are provided here in a minimal,
 * self-contained form under #ifndef ACB_H; with real Arb present the marshalling functions in
 * INTEGRATION.md take their place.  Only data movement lives on the host: every ball OPERATION
 * (acb_mul, acb_add, ... and the solver loops) runs on the GPU.
 *
 * Precision: `bits` / `prec` arguments must be 128, 1024, 2048 or 4096 (a multiple of 32 that the
 * device kernels are instantiated for); other values abort with a message.
 */
#ifndef CASCADE_COMPAT_H
#define CASCADE_COMPAT_H

#include <stdint.h>
#include <stdio.h>
#include "cascade_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#ifndef ACB_H
typedef long slong;
typedef unsigned long ulong;

/* a real ball: midpoint mantissa kept top-aligned in 128 limbs (limb 127 is the most significant),
 * so a value stored at one precision can be handed to the device at any other */
typedef struct {
    cb_meta meta;
    uint32_t mant[CB_MAX_LIMBS];
} arb_struct;
typedef arb_struct arb_t[1];
typedef struct {
    arb_struct real, imag;
} acb_struct;
typedef acb_struct acb_t[1];
typedef acb_struct *acb_ptr;
typedef const acb_struct *acb_srcptr;
#define acb_realref(x) (&(x)->real)
#define acb_imagref(x) (&(x)->imag)

typedef struct {
    acb_struct *coeffs;
    slong length;
    slong alloc;
} acb_poly_struct;
typedef acb_poly_struct acb_poly_t[1];

/* --- the few Arb helpers Cascade's callers use around the hot path (no arithmetic) --- */
void acb_init(acb_t x);
void acb_clear(acb_t x);
void acb_zero(acb_t x);
void acb_one(acb_t x);
void acb_set(acb_t z, const acb_t x);
void acb_set_si(acb_t z, slong v);
void acb_set_d(acb_t z, double v);
void acb_set_d_d(acb_t z, double re, double im);
void acb_neg(acb_t z, const acb_t x);
int acb_is_zero(const acb_t x);
int acb_is_finite(const acb_t x);
int acb_contains_zero(const acb_t x);
double acb_get_mid_re_d(const acb_t x); /* midpoint of the real part, nearest double */
double acb_get_mid_im_d(const acb_t x);
double acb_get_rad_re_d(const acb_t x); /* radius of the real part as an (upper) double */
void acb_printd(const acb_t x, slong digits);
acb_ptr _acb_vec_init(slong n);
void _acb_vec_clear(acb_ptr v, slong n);

void acb_poly_init(acb_poly_t p);
void acb_poly_clear(acb_poly_t p);
void acb_poly_fit_length(acb_poly_t p, slong len);
void acb_poly_zero(acb_poly_t p);
void acb_poly_one(acb_poly_t p);
slong acb_poly_length(const acb_poly_t p);
void acb_poly_set_coeff_acb(acb_poly_t p, slong n, const acb_t c);
void acb_poly_set_coeff_si(acb_poly_t p, slong n, slong c);
void acb_poly_get_coeff_acb(acb_t c, const acb_poly_t p, slong n);
int acb_poly_is_zero(const acb_poly_t p);
void acb_poly_printd(const acb_poly_t p, slong digits);
void flint_cleanup(void);

/* ball arithmetic the example constructors and user code need: each is ONE GPU launch */
void acb_add(acb_t z, const acb_t x, const acb_t y, slong prec);
void acb_sub(acb_t z, const acb_t x, const acb_t y, slong prec);
void acb_mul(acb_t z, const acb_t x, const acb_t y, slong prec);
void acb_div(acb_t z, const acb_t x, const acb_t y, slong prec);
void acb_mul_si(acb_t z, const acb_t x, slong k, slong prec);
void acb_add_si(acb_t z, const acb_t x, slong k, slong prec);
void acb_poly_evaluate(acb_t y, const acb_poly_t f, const acb_t x, slong prec);
void acb_exp(acb_t z, const acb_t x, slong prec);
void acb_log(acb_t z, const acb_t x, slong prec);
void acb_pow(acb_t z, const acb_t x, const acb_t y, slong prec);
void arb_init(arb_t x);
void arb_clear(arb_t x);
void arb_set_d(arb_t x, double v);
double arb_get_mid_d(const arb_t x);
double arb_get_rad_d(const arb_t x);
int arb_is_finite(const arb_t x);
typedef struct {
    acb_ptr entries;
    slong r, c;
    acb_ptr *rows;
} acb_mat_struct;
typedef acb_mat_struct acb_mat_t[1];
#define acb_mat_entry(mat, i, j) ((mat)->rows[i] + (j))
void acb_mat_init(acb_mat_t m, slong r, slong c);
void acb_mat_clear(acb_mat_t m);
/* stand-in for FLINT's generator (used by acb_ode_random only) */
typedef struct {
    uint64_t s;
} flint_rand_struct;
typedef flint_rand_struct flint_rand_t[1];
void flint_randinit(flint_rand_t state);
void flint_randclear(flint_rand_t state);
ulong n_randint(flint_rand_t state, ulong limit);
#endif /* ACB_H */

/* ========================= Differential Operators (reference src/acb_ode.h:9-41) ============== */
typedef struct {
    slong order;
    slong degree;
    slong alloc;
    slong valuation;
    acb_ptr polys;
} acb_ode_struct;
typedef acb_ode_struct acb_ode_t[1];

#define degree(ODE) ((ODE)->degree)
#define order(ODE) ((ODE)->order)
#define acb_ode_poly(ODE, i) (((ODE)->polys) + (i) * (degree(ODE) + 1))
#define acb_ode_coeff(ODE, i, j) (acb_ode_poly(ODE, i) + (j))

void acb_ode_init_blank(acb_ode_t ODE, slong degree, slong order); /* src/acb_ode.c:28-41 */
void acb_ode_init(acb_ode_t ODE, acb_poly_t *polys, slong order);  /* src/acb_ode.c:43-57 */
void acb_ode_clear(acb_ode_t ODE);                                 /* src/acb_ode.c:59-68 */
void acb_ode_set(acb_ode_t ODE_out, acb_ode_t ODE_in);             /* src/acb_ode.c:70-81 */
void acb_ode_dump(acb_ode_t ODE, char *file);                      /* src/acb_ode.c:99-120 */
void acb_ode_shift(acb_ode_t ODE_out, acb_ode_t ODE_in, acb_srcptr a, slong bits); /* :124-135 */
slong acb_ode_valuation(acb_ode_t ODE);                            /* src/acb_ode.c:159-181 */
slong acb_ode_reduce(acb_ode_t ODE);                               /* src/acb_ode.c:137-157 */
void acb_ode_random(acb_ode_t ode, flint_rand_t state, slong prec); /* src/acb_ode.c:83-95 (own generator) */
/* the reference's acceptance test of both solvers, on the device (src/acb_ode.c:185-240) */
void acb_ode_apply(acb_poly_t out, acb_ode_t ODE, acb_poly_t in, slong prec);
int acb_ode_solves(acb_ode_t ODE, acb_poly_t res, slong deg, slong prec);

/* =============================== Solutions (reference src/acb_ode.h:45-61) ==================== */
typedef struct {
    acb_t rho;
    slong mul;
    slong M;
    acb_poly_struct *gens;
} acb_ode_solution_struct;
typedef acb_ode_solution_struct acb_ode_solution_t[1];

void acb_ode_solution_init(acb_ode_solution_t sol, acb_t rho, slong mul, slong alpha);
void acb_ode_solution_clear(acb_ode_solution_t sol);
/* a^rho * sum_i w_i gens[i](a) log(a)^(M-1-i) (reference src/acb_ode_solution.c:24-58): the Horner
 * evaluations, acb_log, acb_pow and the combination run in one device launch.  acb_log / acb_exp are
 * the enclosure algorithms of cascade_b200/csrc/cb_trans.cuh (Arb's are not correctly rounded
 * functions: the ball is a rigorous enclosure, a few tens of bits wider than the precision). */
void acb_ode_solution_evaluate(acb_t res, acb_ode_solution_t sol, acb_t x, slong prec);
/* the rho-derivative bookkeeping of the Frobenius solver (reference src/acb_ode_solution.c:60-123);
 * _update consumes f and _extend consumes g_nu (they come back as zero polynomials), as in the reference */
void _acb_ode_solution_update(acb_ode_solution_t sol, acb_poly_t f, slong prec);
void _acb_ode_solution_extend(acb_ode_solution_t sol, slong nu, acb_poly_t g_nu, slong prec);
void _acb_ode_solution_normalize(acb_ode_solution_t sol, slong prec);

/* ================================ Examples (reference src/examples.c) ========================= */
void acb_ode_legendre(acb_ode_t ODE, ulong n);
void acb_ode_bessel(acb_ode_t ODE, acb_t nu, slong bits); /* overwrites nu with nu^2, as the reference */
void acb_ode_hypgeom(acb_ode_t ODE, acb_t a, acb_t b, acb_t c, slong bits);

/* ============================== Solvers (reference src/cascade.h:13-26) ======================= */
void acb_ode_solve_fuchs(acb_poly_t res, acb_ode_t ODE, slong deg, slong bits);
/* Per path corner: shift, solve (with the valuation of the shifted operator, as src/fuchs_solver.c:105), and
 * re-expand at the next corner.  Between corners only the first `order` Taylor coefficients survive:
 * Horner evaluation with derivatives on the device.  After the LAST corner res holds, as in the reference
 * (src/fuchs_solver.c:159), the complete re-expanded series of deg + 1 coefficients (acb_poly_taylor_shift, Horner
 * form).  A corner where the shifted operator has valuation >= 0 aborts (undefined in the reference). */
void analytic_continuation(acb_poly_t res, acb_ode_t ODE, acb_srcptr path, slong len, slong deg, slong bits);
/* Monodromy around the origin along the 256-gon of half the convergence radius (reference
 * src/fuchs_solver.c:165-206): the `order` unit initial vectors continue together as one device-resident
 * batch.  radius_of_convergence / truncation_order (src/fuchs_solver.c:3-94) are control-plane scalars and
 * are formed in double precision on the host (the reference consumes only a midpoint and an integer). */
void radius_of_convergence(arb_t rad_of_conv, acb_ode_t ODE, slong n, slong bits);
slong truncation_order(arb_t eta, arb_t alpha, slong bits);
void find_monodromy_matrix(acb_mat_t mono, acb_ode_t ODE, slong bits);
void indicial_polynomial_evaluate(acb_t result, acb_ode_t ODE, slong nu, acb_t rho, slong shift, slong prec);
void _acb_ode_solve_frobenius(acb_poly_t res, acb_ode_t ODE, acb_ode_solution_t rhs, slong sol_degree, slong prec);
/* M == 1: the scalar recurrence; M > 1: the polynomial-in-rho (logarithmic) recurrence of
 * src/frobenius_solver.c:131-205, one device launch (M <= 16, 1 <= degree <= 8) */
void acb_ode_solve_frobenius(acb_ode_solution_t sol, acb_ode_t ODE, slong sol_degree, slong prec);
void indicial_polynomial(acb_poly_t result, acb_ode_t ODE, slong nu, slong shift, slong prec); /* :4-39 */

static inline slong clamp(slong in, slong min, slong max) {
    if (in < min)
        return min;
    else if (in > max)
        return max;
    else
        return in;
}

#ifdef __cplusplus
}
#endif
#endif /* CASCADE_COMPAT_H */
