/* The logarithmic Frobenius case through the reference's own names (reference tests/frobenius.c,
 * tests/solution_extend.c, tests/indicial_polynomial.c restated as a program whose printed values
 * tests/test_gpu_compat.py checks):
 *   - hypergeometric equation with c = 1: rho = 0 is a double indicial root; acb_ode_solution_init
 *     (sol, 0, 2, 0) + acb_ode_solve_frobenius give gens[0] = 2F1(a, b; 1; z) and gens[1];
 *   - indicial_polynomial(f, ODE, 0, 0) is rho^2 here, and agrees with indicial_polynomial_evaluate;
 *   - _acb_ode_solution_extend stores g(rho), g'(rho) and consumes g. */
#include <cascade_compat.h>

int main ()
{
    slong prec = 1024, n = 12;
    acb_t a, b, c, rho, y, z;
    acb_ode_t ODE;
    acb_ode_solution_t sol;
    acb_poly_t f, g;

    acb_set_d(a, 0.375); acb_set_d(b, 0.3125); acb_set_d(c, 1.0);
    acb_ode_hypgeom(ODE, a, b, c, prec);
    acb_zero(rho);
    acb_ode_solution_init(sol, rho, 2, 0);
    acb_ode_solve_frobenius(sol, ODE, n, prec);
    for (slong k = 0; k <= n; k++) {
        acb_poly_get_coeff_acb(y, sol->gens + 0, k);
        acb_poly_get_coeff_acb(z, sol->gens + 1, k);
        printf("gens %ld %.17g %.17g %.3g\n", k, acb_get_mid_re_d(y), acb_get_mid_re_d(z), acb_get_rad_re_d(z));
    }
    /* acb_ode_solution_evaluate with M = 2: gens[0](x) log(x) + 2 gens[1](x) at x = 1/8 (rho = 0), and
     * acb_log / acb_exp on their own */
    acb_set_d(z, 0.125);
    acb_ode_solution_evaluate(y, sol, z, prec);
    printf("eval2 %.17g %.17g %.3g\n", acb_get_mid_re_d(y), acb_get_mid_im_d(y), acb_get_rad_re_d(y));
    acb_set_d_d(z, -0.125, 0.0);
    acb_log(y, z, prec);
    printf("log %.17g %.17g %.3g\n", acb_get_mid_re_d(y), acb_get_mid_im_d(y), acb_get_rad_re_d(y));
    acb_set_d(c, 0.5);
    acb_pow(a, z, c, prec); /* (-1/8)^(1/2) = i / sqrt(8) */
    printf("pow %.17g %.17g %.3g\n", acb_get_mid_re_d(a), acb_get_mid_im_d(a), acb_get_rad_re_d(a));
    acb_exp(y, y, prec);
    printf("explog %.17g %.17g %.3g\n", acb_get_mid_re_d(y), acb_get_mid_im_d(y), acb_get_rad_re_d(y));
    acb_ode_solution_clear(sol);

    acb_poly_init(f);
    indicial_polynomial(f, ODE, 0, 3, prec); /* f_0(rho + 3) = (rho + 3)^2 = 9 + 6 rho + rho^2 */
    printf("indicial_len %ld\n", acb_poly_length(f));
    for (slong k = 0; k < acb_poly_length(f); k++) {
        acb_poly_get_coeff_acb(y, f, k);
        printf("indicial %ld %.17g\n", k, acb_get_mid_re_d(y));
    }
    acb_set_d(rho, 2.0);
    acb_poly_evaluate(y, f, rho, prec);
    indicial_polynomial_evaluate(z, ODE, 0, rho, 3, prec);
    printf("indicial_at_2 %.17g %.17g\n", acb_get_mid_re_d(y), acb_get_mid_re_d(z));

    /* _acb_ode_solution_extend: g = 1 + 2 r + 3 r^2 at rho = 1/2: g = 2.75, g' = 5, g'' = 6 */
    acb_poly_init(g);
    acb_poly_set_coeff_si(g, 0, 1); acb_poly_set_coeff_si(g, 1, 2); acb_poly_set_coeff_si(g, 2, 3);
    acb_set_d(rho, 0.5);
    acb_ode_solution_init(sol, rho, 3, 0);
    _acb_ode_solution_extend(sol, 4, g, prec);
    for (slong i = 0; i < 3; i++) {
        acb_poly_get_coeff_acb(y, sol->gens + i, 4);
        printf("extend %ld %.17g %ld\n", i, acb_get_mid_re_d(y), acb_poly_length(sol->gens + i));
    }
    printf("extend_consumed %ld\n", acb_poly_length(g));
    /* _update with f = rho (f(1/2) = 1/2, f' = 1): gens[m] <- sum_k C(m,k) f^(k) gens[m-k] */
    acb_poly_zero(f);
    acb_poly_set_coeff_si(f, 1, 1);
    _acb_ode_solution_update(sol, f, prec);
    for (slong i = 0; i < 3; i++) {
        acb_poly_get_coeff_acb(y, sol->gens + i, 4);
        printf("update %ld %.17g\n", i, acb_get_mid_re_d(y));
    }
    printf("update_consumed %ld\n", acb_poly_length(f));
    _acb_ode_solution_normalize(sol, prec); /* alpha = 0: divide by gens[0][0] = 0 -> indeterminate */
    acb_poly_get_coeff_acb(y, sol->gens + 0, 4);
    printf("normalize_finite %d\n", acb_is_finite(y));
    acb_ode_solution_clear(sol);
    acb_poly_clear(f);
    acb_poly_clear(g);
    acb_ode_clear(ODE);
    return 0;
}
