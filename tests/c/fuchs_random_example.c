/* reference tests/fuchs.c restated with the reference's own names: random operators (acb_ode_random), random
 * initial values, acb_ode_solve_fuchs, and the acceptance test acb_ode_solves -- all on the device. */
#include <cascade_compat.h>

int main ()
{
    int failures = 0;
    flint_rand_t state;
    acb_ode_t ODE;
    acb_poly_t res;
    acb_t c;
    flint_randinit(state);
    acb_poly_init(res);
    for (slong iter = 0; iter < 16; iter++) {
        /* reference tests/fuchs.c:20 draws prec = 30 + n_randint(state, 128); every second iteration does the same
         * here (the compat layer runs such a call at the next compiled precision), the others run at 1024 bits */
        slong prec = (iter & 1) ? 30 + (slong)n_randint(state, 128) : 1024;
        acb_ode_random(ODE, state, prec);
        if (acb_ode_valuation(ODE) != -ODE->order) { acb_ode_clear(ODE); continue; } /* an ordinary point */
        slong n = ODE->order + (slong)n_randint(state, 24);
        acb_poly_zero(res);
        for (slong k = 0; k < ODE->order; k++) {
            acb_set_d_d(c, 0.25 + 0.125 * (double)n_randint(state, 8), -0.5 + 0.125 * (double)n_randint(state, 8));
            acb_poly_set_coeff_acb(res, k, c);
        }
        acb_ode_solve_fuchs(res, ODE, n, prec);
        int ok = acb_ode_solves(ODE, res, n - ODE->order, prec);
        printf("iter %ld prec %ld order %ld degree %ld n %ld solves %d\n", iter, prec, ODE->order, ODE->degree, n, ok);
        failures += !ok;
        acb_ode_clear(ODE);
    }
    printf("failures %d\n", failures);
    acb_poly_clear(res);
    flint_randclear(state);
    flint_cleanup();
    return failures != 0;
}
