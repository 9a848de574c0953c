/* reference tests/radius.c restated: the leading coefficient of the operator is a polynomial with known roots;
 * radius_of_convergence must not exceed any root modulus (and, being a Graeffe-quality bound in the reference,
 * sits at the smallest one).  Also acb_ode_reduce (reference tests/reduce.c: a common factor z^k is divided out). */
#include <cascade_compat.h>
#include <complex.h>

int main ()
{
    slong prec = 1024;
    acb_ode_t ode;
    arb_t rad;
    /* (z - 2)(z + 3i)(z - (1 + i)) */
    double complex r[3] = {2.0, -3.0 * I, 1.0 + 1.0 * I}, c[4] = {1, 0, 0, 0};
    for (int k = 0; k < 3; k++) { /* multiply by (z - r_k): coefficients low to high in c[0..k+1] */
        for (int j = k + 1; j >= 1; j--) c[j] = c[j - 1] - r[k] * c[j];
        c[0] = -r[k] * c[0];
    }
    acb_ode_init_blank(ode, 5, 1);
    for (int j = 0; j < 4; j++) acb_set_d_d(acb_ode_coeff(ode, 1, j), creal(c[j]), cimag(c[j]));
    acb_set_d(acb_ode_coeff(ode, 0, 0), 1.0);
    radius_of_convergence(rad, ode, 20, prec);
    printf("radius %.15g %.3g\n", arb_get_mid_d(rad), arb_get_rad_d(rad));
    acb_ode_clear(ode);

    /* z^2 * (same cubic) as leading coefficient, lower rows with a common factor z^2 as well */
    acb_ode_init_blank(ode, 5, 1);
    for (int j = 0; j < 4; j++) acb_set_d_d(acb_ode_coeff(ode, 1, j + 2), creal(c[j]), cimag(c[j]));
    acb_set_d(acb_ode_coeff(ode, 0, 2), 1.0);
    acb_set_d(acb_ode_coeff(ode, 0, 3), -0.5);
    slong k = acb_ode_reduce(ode);
    printf("reduced %ld degree %ld valuation %ld\n", k, degree(ode), acb_ode_valuation(ode));
    radius_of_convergence(rad, ode, 20, prec);
    printf("radius_after_reduce %.15g\n", arb_get_mid_d(rad));
    acb_ode_clear(ode);
    return 0;
}
