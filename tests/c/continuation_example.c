/* Drop-in use of the remaining hot-path entry points through the reference's own names:
 * acb_ode_hypgeom + acb_ode_solve_frobenius (M = 1), acb_ode_bessel + analytic_continuation,
 * acb_poly_evaluate.  Prints values that tests/test_gpu_compat.py checks. */
#include <cascade_compat.h>

int main ()
{
    slong prec = 1024;
    acb_t a, b, c, rho, x, y;
    acb_ode_t ODE;
    acb_ode_solution_t sol;

    /* 2F1(1/2, 1/4; 3/2; x) by Frobenius at rho = 0, summed at x = 1/8 */
    acb_set_d(a, 0.5); acb_set_d(b, 0.25); acb_set_d(c, 1.5);
    acb_ode_hypgeom(ODE, a, b, c, prec);
    acb_zero(rho);
    acb_ode_solution_init(sol, rho, 1, 0);
    acb_ode_solve_frobenius(sol, ODE, 120, prec);
    acb_set_d(x, 0.125);
    acb_poly_evaluate(y, sol->gens, x, prec);
    printf("hyp2f1 %.17g %.3g\n", acb_get_mid_re_d(y), acb_get_rad_re_d(y));
    acb_ode_solution_evaluate(y, sol, x, prec); /* rho = 0: a^0 * gens[0](a) */
    printf("hyp2f1_eval %.17g %.3g\n", acb_get_mid_re_d(y), acb_get_rad_re_d(y));
    acb_ode_solution_clear(sol);
    acb_ode_clear(ODE);

    /* Bessel nu = 0 continued from z = 1 to z = 1.5 in two steps, starting from (J0(1), J0'(1)) */
    acb_t nu;
    acb_ptr path = _acb_vec_init(3);
    acb_poly_t res;
    acb_zero(nu);
    acb_ode_bessel(ODE, nu, prec);
    acb_set_d(path + 0, 1.0); acb_set_d(path + 1, 1.25); acb_set_d(path + 2, 1.5);
    acb_poly_init(res);
    acb_set_d(x, 0.76519768655796655145); acb_poly_set_coeff_acb(res, 0, x);
    acb_set_d(x, -0.44005058574493351596); acb_poly_set_coeff_acb(res, 1, x);
    analytic_continuation(res, ODE, path, 3, 80, prec);
    acb_poly_get_coeff_acb(y, res, 0);
    printf("J0(1.5) %.15g\n", acb_get_mid_re_d(y));
    acb_poly_get_coeff_acb(y, res, 1);
    printf("J0'(1.5) %.15g\n", acb_get_mid_re_d(y));
    acb_poly_clear(res);
    _acb_vec_clear(path, 3);
    acb_ode_clear(ODE);
    return 0;
}
