/* find_monodromy_matrix through the reference's own names (reference src/fuchs_solver.c:165-206): the
 * hypergeometric equation with c = 3/4 has local exponents 0 and 1 - c at the origin, so the monodromy
 * around 0 has eigenvalues 1 and exp(2 pi i (1 - c)) = i: trace 1 + i, determinant i.  Also prints
 * radius_of_convergence (1: the singularity at z = 1) and truncation_order. */
#include <cascade_compat.h>

int main ()
{
    slong prec = 1024;
    acb_t a, b, c, tr, det, t;
    acb_ode_t ODE;
    acb_mat_t mono;
    arb_t rad, eta, alpha;

    acb_set_d(a, 0.5); acb_set_d(b, 0.25); acb_set_d(c, 0.75);
    acb_ode_hypgeom(ODE, a, b, c, prec);
    radius_of_convergence(rad, ODE, 40, prec);
    printf("radius %.15g %.3g\n", arb_get_mid_d(rad), arb_get_rad_d(rad));
    arb_set_d(eta, 0.5 * 0.02454); arb_set_d(alpha, 0.5);
    printf("truncation_order %ld\n", truncation_order(eta, alpha, prec));

    acb_mat_init(mono, 2, 2);
    find_monodromy_matrix(mono, ODE, prec);
    acb_add(tr, acb_mat_entry(mono, 0, 0), acb_mat_entry(mono, 1, 1), prec);
    acb_mul(det, acb_mat_entry(mono, 0, 0), acb_mat_entry(mono, 1, 1), prec);
    acb_mul(t, acb_mat_entry(mono, 0, 1), acb_mat_entry(mono, 1, 0), prec);
    acb_sub(det, det, t, prec);
    printf("trace %.15g %.15g %.3g\n", acb_get_mid_re_d(tr), acb_get_mid_im_d(tr), acb_get_rad_re_d(tr));
    printf("det %.15g %.15g %.3g\n", acb_get_mid_re_d(det), acb_get_mid_im_d(det), acb_get_rad_re_d(det));
    acb_mat_clear(mono);
    acb_ode_clear(ODE);
    return 0;
}
