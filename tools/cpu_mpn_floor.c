/* How far is the oracle port (bench.py's CPU baseline) from the bare GMP limb kernels Arb sits on?
 * Times, on one core, the mpn calls a tuned Arb would make for ONE coefficient of BASELINE configs[1] (1024 bits,
 * order = degree = 2): 5 complex multiplications = 20 mpn_mul_n of 16 x 16 limbs (the inverse of the shared divisor is
 * amortised over the batch as in the GPU path; without that add one mpn_sqr pair and two 32-by-16 limb divisions) plus
 * 10 mpn_add_n/sub_n for the fused dot products and subtractions.  No rounding, no radius arithmetic, no allocation:
 * a LOWER bound for Cascade + Arb per coefficient.  GMP 6.3 has no headers in this image: prototypes declared by hand.
 *   gcc -O2 -o tools/_bin/cpu_mpn_floor tools/cpu_mpn_floor.c -l:libgmp.so.10 && tools/_bin/cpu_mpn_floor */
#include <stdint.h>
#include <stdio.h>
#include <time.h>
typedef unsigned long mp_limb_t;
extern void __gmpn_mul_n(mp_limb_t *, const mp_limb_t *, const mp_limb_t *, long);
extern mp_limb_t __gmpn_add_n(mp_limb_t *, const mp_limb_t *, const mp_limb_t *, long);
extern void __gmpn_tdiv_qr(mp_limb_t *, mp_limb_t *, long, const mp_limb_t *, long, const mp_limb_t *, long);
int main(void) {
    enum { N = 16 };
    mp_limb_t a[N], b[N], p[2 * N], q[2 * N], s[2 * N], num[2 * N], qq[N + 1], rr[N];
    uint64_t x = 88172645463325252ull;
    for (int i = 0; i < N; i++) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17; a[i] = x;
        x ^= x << 13; x ^= x >> 7; x ^= x << 17; b[i] = x;
    }
    a[N - 1] |= 1ul << 63; b[N - 1] |= 1ul << 63;
    const long iters = 200000;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    mp_limb_t sink = 0;
    for (long it = 0; it < iters; it++) {
        for (int k = 0; k < 10; k++) {  /* 10 fused dot products: two products and one 2N-limb addition each */
            __gmpn_mul_n(p, a, b, N);
            __gmpn_mul_n(q, b, a, N);
            sink += __gmpn_add_n(s, p, q, 2 * N);
            a[0] += s[N];
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    double per_amortised = ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec)) / iters;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (long it = 0; it < iters; it++) {
        __gmpn_mul_n(p, a, a, N);
        __gmpn_mul_n(q, b, b, N);
        sink += __gmpn_add_n(s, p, q, 2 * N);
        for (int k = 0; k < 2; k++) {
            for (int i = 0; i < N; i++) { num[i] = 0; num[N + i] = a[i]; }
            s[2 * N - 1] |= 1ul << 63;
            __gmpn_tdiv_qr(qq, rr, 0, num, 2 * N, s + N, N);
            a[0] += qq[0];
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    double per_inv = ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec)) / iters;
    printf("{\"mpn_us_per_coeff_inverse_amortised\": %.3f, \"mpn_us_per_inverse\": %.3f, \"mpn_us_per_coeff_with_inverse\": %.3f, \"sink\": %lu}\n",
           per_amortised * 1e6, per_inv * 1e6, (per_amortised + per_inv) * 1e6, (unsigned long)(sink & 1));
    return 0;
}
