/* A program that owns Arb's symbol names (as a Cascade built against the real libarb does) and links
 * libcascade_b200.so: the batch library exports cb200_* only, so nothing clashes.  (tests/test_abi_and_host.py) */
#include <stdio.h>
#include <cascade_b200.h>

/* stand-ins for "the real Arb is linked too": same names, deliberately different signatures */
int acb_mul(int a, int b) { return a * b; }
int acb_init(void) { return 1; }
int acb_poly_evaluate(void) { return 2; }
int acb_ode_solve_fuchs(void) { return 3; }
int flint_malloc(void) { return 4; }

int main(void) {
    printf("own acb_mul: %d\n", acb_mul(6, 7));
    printf("%s, launches so far %llu, init %d eval %d solve %d malloc %d\n", cb200_version(),
           (unsigned long long)cb200_kernel_launches(), acb_init(), acb_poly_evaluate(), acb_ode_solve_fuchs(), flint_malloc());
    return 0;
}
