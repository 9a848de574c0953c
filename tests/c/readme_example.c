/* The reference's README program (reference README.md:31-57), compiled UNCHANGED in its calls
 * against include/cascade_compat.h and linked with libcascade_b200.so (the only edit: the two
 * #include lines name the compat header instead of <acb_poly.h>/<cascade.h>).  Expected midpoints
 * (reference README.md:72-77): 0.375, 0, -3.75, 0, 4.375. */
#include <cascade_compat.h>

int main ()
{
    acb_t z;
    acb_poly_t pol;
    acb_ode_t ODE;

    acb_init(z);
    acb_poly_init(pol);
    acb_ode_legendre(ODE, 4);

    acb_set_d(z, 0.375);
    acb_poly_set_coeff_acb(pol, 0, z);
    acb_ode_solve_fuchs(pol, ODE, 5, 1024);
    acb_ode_dump(ODE, NULL);
    acb_poly_printd(pol, 10);

    acb_clear(z);
    acb_poly_clear(pol);
    acb_ode_clear(ODE);
    flint_cleanup();
    return 0;
}
