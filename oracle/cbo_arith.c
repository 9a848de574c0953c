/* cbo_arith.c -- ORACLE ball arithmetic (test infrastructure; parity unpinned, see cbo.h).
 *
 * Restates the Arb semantics reached from Cascade's hot path:
 *   acb_mul / acb_sub / acb_add / acb_div / acb_mul_fmpz   reference src/fuchs_solver.c:118-137
 *   acb_add_si / acb_mul / acb_add / acb_sub_si / acb_div  reference src/frobenius_solver.c:58-113
 *   acb_mul / acb_mul_si / acb_add (Horner)                reference src/acb_ode_solution.c:40-49
 *
 * Midpoints: every operation is computed EXACTLY on integer mantissas (struct xf below) and then
 * rounded toward zero to prec = 32 L bits counted from the leading bit; "inexact" is exact.
 * Because a correctly rounded result is unique, any other correct implementation (the CUDA one)
 * must agree bit for bit.  Radii: (30-bit mantissa, exponent) magnitudes, each step rounded up,
 * with the formulas stated at each function -- these ARE implementation defined, and the CUDA
 * code mirrors them step by step (cascade_b200/csrc/cb_mag.cuh).
 *
 * Limb products / divisions go through GMP 6.3.0 mpn_* (libgmp.so.10, hand-declared prototypes:
 * the image ships no gmp.h) unless built with -DCBO_NO_GMP, which selects plain C (__int128).
 */
#include "cbo.h"
#include <string.h>
#include <stdlib.h>

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------ mpn layer */
#ifndef CBO_NO_GMP
extern uint64_t __gmpn_add_n(uint64_t *, const uint64_t *, const uint64_t *, long);
extern uint64_t __gmpn_sub_n(uint64_t *, const uint64_t *, const uint64_t *, long);
extern uint64_t __gmpn_mul(uint64_t *, const uint64_t *, long, const uint64_t *, long);
extern void __gmpn_tdiv_qr(uint64_t *, uint64_t *, long, const uint64_t *, long, const uint64_t *,
                           long);
const char *cbo_backend(void) { return "gmp-mpn"; }
#else
const char *cbo_backend(void) { return "plain-c"; }
#endif

static uint64_t n_add(uint64_t *r, const uint64_t *a, const uint64_t *b, int n) {
#ifndef CBO_NO_GMP
    return n ? __gmpn_add_n(r, a, b, n) : 0;
#else
    uint64_t c = 0;
    for (int i = 0; i < n; i++) {
        u128 t = (u128)a[i] + b[i] + c;
        r[i] = (uint64_t)t;
        c = (uint64_t)(t >> 64);
    }
    return c;
#endif
}

static uint64_t n_sub(uint64_t *r, const uint64_t *a, const uint64_t *b, int n) {
#ifndef CBO_NO_GMP
    return n ? __gmpn_sub_n(r, a, b, n) : 0;
#else
    uint64_t c = 0;
    for (int i = 0; i < n; i++) {
        u128 t = (u128)a[i] - b[i] - c;
        r[i] = (uint64_t)t;
        c = (uint64_t)(t >> 64) & 1;
    }
    return c;
#endif
}

/* r[0..an+bn) = a * b ; r must not alias */
static void n_mul(uint64_t *r, const uint64_t *a, int an, const uint64_t *b, int bn) {
#ifndef CBO_NO_GMP
    if (an >= bn)
        __gmpn_mul(r, a, an, b, bn);
    else
        __gmpn_mul(r, b, bn, a, an);
#else
    memset(r, 0, sizeof(uint64_t) * (an + bn));
    for (int i = 0; i < bn; i++) {
        uint64_t c = 0;
        for (int j = 0; j < an; j++) {
            u128 t = (u128)a[j] * b[i] + r[i + j] + c;
            r[i + j] = (uint64_t)t;
            c = (uint64_t)(t >> 64);
        }
        r[i + an] = c;
    }
#endif
}

static int n_cmp(const uint64_t *a, const uint64_t *b, int n) {
    for (int i = n - 1; i >= 0; i--)
        if (a[i] != b[i]) return a[i] > b[i] ? 1 : -1;
    return 0;
}

/* q[0..nn-dn] , r[0..dn) = num / den ; den[dn-1] has its top bit set ; nn >= dn */
static void n_divrem(uint64_t *q, uint64_t *r, const uint64_t *num, int nn, const uint64_t *den,
                     int dn) {
#ifndef CBO_NO_GMP
    __gmpn_tdiv_qr(q, r, 0, num, nn, den, dn);
#else
    /* Knuth algorithm D on 64-bit limbs (divisor already normalised) */
    uint64_t u[2 * (CBO_MAXL / 2) + 4];
    memcpy(u, num, sizeof(uint64_t) * nn);
    u[nn] = 0;
    if (dn == 1) {
        uint64_t rem = 0;
        for (int i = nn - 1; i >= 0; i--) {
            u128 t = ((u128)rem << 64) | u[i];
            q[i] = (uint64_t)(t / den[0]);
            rem = (uint64_t)(t % den[0]);
        }
        r[0] = rem;
        return;
    }
    for (int j = nn - dn; j >= 0; j--) {
        u128 top = ((u128)u[j + dn] << 64) | u[j + dn - 1];
        u128 qh = top / den[dn - 1], rh = top % den[dn - 1];
        while (qh >> 64 || (u128)(uint64_t)qh * den[dn - 2] > ((rh << 64) | u[j + dn - 2])) {
            qh--;
            rh += den[dn - 1];
            if (rh >> 64) break;
        }
        /* multiply-subtract */
        uint64_t borrow = 0, carry = 0;
        for (int i = 0; i < dn; i++) {
            u128 p = (u128)(uint64_t)qh * den[i] + carry;
            carry = (uint64_t)(p >> 64);
            u128 t = (u128)u[i + j] - (uint64_t)p - borrow;
            u[i + j] = (uint64_t)t;
            borrow = (uint64_t)(t >> 64) & 1;
        }
        u128 t = (u128)u[j + dn] - carry - borrow;
        u[j + dn] = (uint64_t)t;
        if ((uint64_t)(t >> 64) & 1) {
            qh--;
            uint64_t c = 0;
            for (int i = 0; i < dn; i++) {
                u128 s = (u128)u[i + j] + den[i] + c;
                u[i + j] = (uint64_t)s;
                c = (uint64_t)(s >> 64);
            }
            u[j + dn] += c;
        }
        q[j] = (uint64_t)qh;
    }
    memcpy(r, u, sizeof(uint64_t) * dn);
#endif
}

/* ------------------------------------------------------------------ exact numbers */
#define XFN (4 * (CBO_MAXL / 2) + 16)
typedef struct {
    int sign;    /* 0, +1, -1 */
    int64_t exp; /* value = sign * D * 2^exp */
    int n;       /* limbs; d[n-1] != 0 when sign != 0 */
    uint64_t d[XFN];
} xf;

static int64_t xf_bitlen(const xf *a) { return 64 * (int64_t)a->n - __builtin_clzll(a->d[a->n - 1]); }

static void xf_strip(xf *a) {
    while (a->n > 0 && a->d[a->n - 1] == 0) a->n--;
    if (a->n == 0) a->sign = 0;
}

static void xf_from_arb(xf *r, const cbo_arb *x, int L64) {
    if (x->flags & CBO_ZERO) {
        r->sign = 0;
        r->n = 0;
        r->exp = 0;
        return;
    }
    /* drop trailing zero limbs so that short operands (small integers) stay short */
    int lo = 0;
    while (lo < L64 - 1 && x->m.d[lo] == 0) lo++;
    r->n = L64 - lo;
    memcpy(r->d, x->m.d + lo, sizeof(uint64_t) * r->n);
    r->exp = (int64_t)x->exp - 64 * (int64_t)L64 + 64 * (int64_t)lo;
    r->sign = (x->flags & CBO_SIGN) ? -1 : 1;
}

static void xf_from_si(xf *r, int64_t k) {
    if (k == 0) {
        r->sign = 0;
        r->n = 0;
        r->exp = 0;
        return;
    }
    r->sign = k < 0 ? -1 : 1;
    r->d[0] = k < 0 ? (uint64_t)0 - (uint64_t)k : (uint64_t)k;
    r->n = 1;
    r->exp = 0;
}

static void xf_mul(xf *r, const xf *a, const xf *b) {
    if (a->sign == 0 || b->sign == 0) {
        r->sign = 0;
        r->n = 0;
        r->exp = 0;
        return;
    }
    n_mul(r->d, a->d, a->n, b->d, b->n);
    r->n = a->n + b->n;
    r->sign = a->sign * b->sign;
    r->exp = a->exp + b->exp;
    xf_strip(r);
}

/* dst[0..N) = a->d shifted left by sh bits (sh >= 0), zero filled */
static void place(uint64_t *dst, int N, const xf *a, int64_t sh) {
    memset(dst, 0, sizeof(uint64_t) * N);
    int q = (int)(sh / 64), r = (int)(sh % 64);
    for (int i = 0; i < a->n; i++) {
        dst[i + q] |= a->d[i] << r;
        if (r) dst[i + q + 1] |= a->d[i] >> (64 - r);
    }
}

/* r = a + (negb ? -b : b), exact as far as rounding to prec bits can tell (see comment) */
static void xf_add(xf *r, const xf *a0, const xf *b0, int negb, int64_t prec) {
    xf bneg;
    const xf *a = a0, *b = b0;
    int sb = negb ? -b0->sign : b0->sign;
    if (b->sign == 0) {
        *r = *a;
        return;
    }
    if (a->sign == 0) {
        *r = *b;
        r->sign = sb;
        return;
    }
    int sa = a->sign;
    int64_t ta = a->exp + xf_bitlen(a), tb = b->exp + xf_bitlen(b);
    if (tb > ta) {
        const xf *t = a;
        a = b;
        b = t;
        int s = sa;
        sa = sb;
        sb = s;
        int64_t tt = ta;
        ta = tb;
        tb = tt;
    }
    /* If b lies entirely below both a's last bit and the rounding position (with 3 bits of
     * slack for cancellation), any nonzero value of b's sign and of magnitude < 2^cut rounds
     * the same way and is inexact: substitute 2^(cut-1).  This bounds the aligned length. */
    int64_t cut = a->exp < ta - prec - 3 ? a->exp : ta - prec - 3;
    if (tb <= cut - 1) {
        bneg.sign = sb;
        bneg.n = 1;
        bneg.d[0] = 1;
        bneg.exp = cut - 1;
        b = &bneg;
    }
    int64_t lsb = a->exp < b->exp ? a->exp : b->exp;
    int N = (int)((ta - lsb) / 64) + 2;
    static __thread uint64_t A[XFN], B[XFN];
    place(A, N, a, a->exp - lsb);
    place(B, N, b, b->exp - lsb);
    if (sa == sb) {
        n_add(r->d, A, B, N);
        r->sign = sa;
    } else {
        int c = n_cmp(A, B, N);
        if (c == 0) {
            r->sign = 0;
            r->n = 0;
            r->exp = 0;
            return;
        }
        if (c > 0) {
            n_sub(r->d, A, B, N);
            r->sign = sa;
        } else {
            n_sub(r->d, B, A, N);
            r->sign = sb;
        }
    }
    r->n = N;
    r->exp = lsb;
    xf_strip(r);
}

#define EXP_LIMIT (1 << 28)

static void arb_set_nan(cbo_arb *z) {
    memset(z, 0, sizeof(int32_t) * 4);
    memset(z->m.w, 0, sizeof(z->m.w));
    z->flags = CBO_NAN;
    z->rman = 0xFFFFFFFFu;
}

/* midpoint of z := a rounded toward zero to 64*L64 bits; returns inexact; radius untouched.
 * Returns -1 if the exponent leaves the supported range (caller makes the ball indeterminate). */
static int xf_round(cbo_arb *z, const xf *a, int L64) {
    memset(z->m.w, 0, sizeof(z->m.w));
    if (a->sign == 0) {
        z->exp = 0;
        z->flags = CBO_ZERO;
        return 0;
    }
    int64_t prec = 64 * (int64_t)L64, bl = xf_bitlen(a);
    int inexact = 0;
    int64_t sh = bl - prec; /* right shift amount */
    if (sh > 0) {
        int q = (int)(sh / 64), r = (int)(sh % 64);
        for (int i = 0; i < q; i++) inexact |= a->d[i] != 0;
        if (r) inexact |= (a->d[q] << (64 - r)) != 0;
        for (int i = 0; i < L64; i++) {
            uint64_t lo = a->d[i + q] >> r;
            uint64_t hi = (r && i + q + 1 < a->n) ? a->d[i + q + 1] << (64 - r) : 0;
            z->m.d[i] = lo | hi;
        }
    } else {
        int64_t ls = -sh;
        int q = (int)(ls / 64), r = (int)(ls % 64);
        for (int i = 0; i < a->n; i++) {
            z->m.d[i + q] |= a->d[i] << r;
            if (r && i + q + 1 < L64) z->m.d[i + q + 1] |= a->d[i] >> (64 - r);
        }
    }
    int64_t e = a->exp + bl;
    if (e > EXP_LIMIT || e < -EXP_LIMIT) return -1;
    z->exp = (int32_t)e;
    z->flags = a->sign < 0 ? CBO_SIGN : 0;
    return inexact;
}

/* ------------------------------------------------------------------ magnitudes */
typedef struct {
    uint32_t man; /* 0 or [2^29, 2^30) */
    int64_t exp;
} mg;
#define MAG_ONE_HALF (1u << 29)

static mg mg_zero(void) {
    mg z = {0, 0};
    return z;
}
static mg mg_of(const cbo_arb *x) {
    mg z = {x->rman, x->rexp};
    return z;
}
/* upper bound of x*y: ((xm*ym) >> 30) + 1, renormalised */
static mg mg_mul(mg x, mg y) {
    mg z;
    if (x.man == 0 || y.man == 0) return mg_zero();
    uint32_t m = (uint32_t)(((uint64_t)x.man * y.man) >> 30) + 1;
    z.exp = x.exp + y.exp;
    if (!(m >> 29)) {
        m <<= 1;
        z.exp -= 1;
    }
    z.man = m;
    return z;
}
/* lower bound of x*y */
static mg mg_mul_lower(mg x, mg y) {
    mg z;
    if (x.man == 0 || y.man == 0) return mg_zero();
    uint32_t m = (uint32_t)(((uint64_t)x.man * y.man) >> 30);
    z.exp = x.exp + y.exp;
    if (!(m >> 29)) {
        m <<= 1;
        z.exp -= 1;
    }
    z.man = m;
    return z;
}
/* upper bound of x+y */
static mg mg_add(mg x, mg y) {
    if (x.man == 0) return y;
    if (y.man == 0) return x;
    if (x.exp < y.exp) {
        mg t = x;
        x = y;
        y = t;
    }
    int64_t sh = x.exp - y.exp;
    uint32_t m;
    if (sh == 0)
        m = x.man + y.man;
    else if (sh >= 30)
        m = x.man + 1;
    else
        m = x.man + (y.man >> sh) + 1;
    mg z;
    z.exp = x.exp;
    if (m >> 30) {
        m = (m >> 1) + (m & 1);
        z.exp += 1;
        /* m <= 2^30 - 1 here: inputs are < 2^30 each */
    }
    z.man = m;
    return z;
}
/* lower bound of x-y, 0 when it cannot be shown positive */
static mg mg_sub_lower(mg x, mg y) {
    if (y.man == 0) return x;
    if (x.man == 0) return mg_zero();
    int64_t sh = x.exp - y.exp;
    if (sh < 0) return mg_zero();
    int64_t m;
    if (sh >= 30)
        m = (int64_t)x.man - 1;
    else
        m = (int64_t)x.man - (int64_t)(y.man >> sh) - 1;
    if (m <= 0) return mg_zero();
    mg z;
    z.exp = x.exp;
    int s = __builtin_clz((uint32_t)m) - 2; /* bring the top bit to bit 29 */
    z.man = (uint32_t)m << s;
    z.exp -= s;
    return z;
}
/* upper bound of x / y (y > 0) */
static mg mg_div(mg x, mg y) {
    if (x.man == 0) return mg_zero();
    uint32_t q = (uint32_t)((((uint64_t)x.man) << 30) / y.man) + 1; /* in (2^29, 2^31] */
    mg z;
    z.exp = x.exp - y.exp;
    while (q >> 30) {
        q = (q >> 1) + (q & 1);
        z.exp += 1;
    }
    z.man = q;
    return z;
}
/* upper / lower bound of |mid| */
static mg mg_mid_upper(const cbo_arb *x, int L) {
    if (x->flags & CBO_ZERO) return mg_zero();
    mg z;
    z.man = (x->m.w[L - 1] >> 2) + 1;
    z.exp = x->exp;
    if (z.man >> 30) {
        z.man = MAG_ONE_HALF;
        z.exp += 1;
    }
    return z;
}
static mg mg_mid_lower(const cbo_arb *x, int L) {
    if (x->flags & CBO_ZERO) return mg_zero();
    mg z;
    z.man = x->m.w[L - 1] >> 2;
    z.exp = x->exp;
    return z;
}
/* 2^(exp - prec) */
static mg mg_ulp(const cbo_arb *z, int L) {
    mg u;
    u.man = MAG_ONE_HALF;
    u.exp = (int64_t)z->exp - 32 * (int64_t)L + 1;
    return u;
}
static void set_rad(cbo_arb *z, mg r) {
    if (z->flags & CBO_NAN) return;
    if (r.man != 0 && (r.exp > 4 * (int64_t)EXP_LIMIT || r.exp < -4 * (int64_t)EXP_LIMIT)) {
        arb_set_nan(z); /* radius exponent outside the 32-bit device range */
        return;
    }
    z->rman = r.man;
    z->rexp = r.man ? (int32_t)r.exp : 0;
}
static int is_nan(const cbo_arb *x) { return (x->flags & CBO_NAN) != 0; }

/* finish: midpoint from exact value, radius = rad (+ ulp if inexact) */
static void finish(cbo_arb *z, const xf *v, mg rad, int L) {
    cbo_arb t;
    int inex = xf_round(&t, v, L / 2);
    if (inex < 0) {
        arb_set_nan(z);
        return;
    }
    if (inex) rad = mg_add(rad, mg_ulp(&t, L));
    t.rman = 0;
    t.rexp = 0;
    set_rad(&t, rad);
    *z = t;
}

/* ------------------------------------------------------------------ arb */
void cbo_arb_zero(cbo_arb *z) {
    memset(z, 0, sizeof(int32_t) * 4);
    memset(z->m.w, 0, sizeof(z->m.w));
    z->flags = CBO_ZERO;
}
void cbo_arb_set_si(cbo_arb *z, int64_t k, int L) {
    xf v;
    xf_from_si(&v, k);
    cbo_arb_zero(z);
    finish(z, &v, mg_zero(), L);
}
void cbo_arb_neg(cbo_arb *z, const cbo_arb *x) {
    if (z != x) *z = *x;
    if (!(z->flags & (CBO_ZERO | CBO_NAN))) z->flags ^= CBO_SIGN;
}
int cbo_arb_is_zero(const cbo_arb *x) { return (x->flags & CBO_ZERO) && x->rman == 0; }
int cbo_arb_is_finite(const cbo_arb *x) { return !is_nan(x); }

/* exact |mid| <= rad */
static int arb_contains_zero_L(const cbo_arb *x, int L) {
    if (is_nan(x)) return 1;
    if (x->flags & CBO_ZERO) return 1;
    if (x->rman == 0) return 0;
    if (x->exp < x->rexp) return 1;
    if (x->exp > x->rexp) return 0;
    uint32_t top30 = x->m.w[L - 1] >> 2;
    if (top30 != x->rman) return top30 < x->rman;
    if (x->m.w[L - 1] & 3u) return 0;
    for (int i = 0; i < L - 1; i++)
        if (x->m.w[i]) return 0;
    return 1;
}
int cbo_arb_contains_zero(const cbo_arb *x) {
    /* the mantissa length is not stored: find the top limb (normalised => top bit set) */
    int L = CBO_MAXL;
    while (L > 1 && x->m.w[L - 1] == 0) L--;
    return arb_contains_zero_L(x, L);
}

/* arb_add / arb_sub: mid = trunc(x +- y); rad = xr + yr (+ ulp) */
static void arb_addsub(cbo_arb *z, const cbo_arb *x, const cbo_arb *y, int neg, int L) {
    if (is_nan(x) || is_nan(y)) {
        arb_set_nan(z);
        return;
    }
    xf a, b, s;
    xf_from_arb(&a, x, L / 2);
    xf_from_arb(&b, y, L / 2);
    xf_add(&s, &a, &b, neg, 32 * (int64_t)L);
    finish(z, &s, mg_add(mg_of(x), mg_of(y)), L);
}
void cbo_arb_add(cbo_arb *z, const cbo_arb *x, const cbo_arb *y, int L) { arb_addsub(z, x, y, 0, L); }
void cbo_arb_sub(cbo_arb *z, const cbo_arb *x, const cbo_arb *y, int L) { arb_addsub(z, x, y, 1, L); }

/* radius of a product x*y: |x|^ yr + |y|^ xr + xr yr, in this order, zero terms drop out */
static mg mul_rad(mg acc, const cbo_arb *x, const cbo_arb *y, int L) {
    acc = mg_add(acc, mg_mul(mg_mid_upper(x, L), mg_of(y)));
    acc = mg_add(acc, mg_mul(mg_mid_upper(y, L), mg_of(x)));
    acc = mg_add(acc, mg_mul(mg_of(x), mg_of(y)));
    return acc;
}

void cbo_arb_fdot2(cbo_arb *z, const cbo_arb *x1, const cbo_arb *y1, const cbo_arb *x2,
                   const cbo_arb *y2, int neg2, int L) {
    if (is_nan(x1) || is_nan(y1) || is_nan(x2) || is_nan(y2)) {
        arb_set_nan(z);
        return;
    }
    xf a, b, p1, p2, s;
    xf_from_arb(&a, x1, L / 2);
    xf_from_arb(&b, y1, L / 2);
    xf_mul(&p1, &a, &b);
    xf_from_arb(&a, x2, L / 2);
    xf_from_arb(&b, y2, L / 2);
    xf_mul(&p2, &a, &b);
    xf_add(&s, &p1, &p2, neg2, 32 * (int64_t)L);
    mg rad = mul_rad(mg_zero(), x1, y1, L);
    rad = mul_rad(rad, x2, y2, L);
    finish(z, &s, rad, L);
}

void cbo_arb_mul(cbo_arb *z, const cbo_arb *x, const cbo_arb *y, int L) {
    if (is_nan(x) || is_nan(y)) {
        arb_set_nan(z);
        return;
    }
    xf a, b, p;
    xf_from_arb(&a, x, L / 2);
    xf_from_arb(&b, y, L / 2);
    xf_mul(&p, &a, &b);
    finish(z, &p, mul_rad(mg_zero(), x, y, L), L);
}

void cbo_arb_mul_si(cbo_arb *z, const cbo_arb *x, int64_t k, int L) {
    cbo_arb kk;
    cbo_arb_set_si(&kk, k, L);
    cbo_arb_mul(z, x, &kk, L);
}
void cbo_arb_add_si(cbo_arb *z, const cbo_arb *x, int64_t k, int L) {
    cbo_arb kk;
    cbo_arb_set_si(&kk, k, L);
    cbo_arb_add(z, x, &kk, L);
}

/* arb_div: indeterminate when y's ball reaches zero; mid = trunc(x/y);
 * rad = (|x|^ yr + |y|^ xr) / (|y|_ (|y|_ - yr)) (+ ulp); exact operands: rad = ulp or 0 */
void cbo_arb_div(cbo_arb *z, const cbo_arb *x, const cbo_arb *y, int L) {
    if (is_nan(x) || is_nan(y) || (y->flags & CBO_ZERO)) {
        arb_set_nan(z);
        return;
    }
    int L64 = L / 2;
    mg yl = mg_mid_lower(y, L);
    mg gap = mg_sub_lower(yl, mg_of(y));
    if (gap.man == 0) {
        arb_set_nan(z);
        return;
    }
    mg rad = mg_zero();
    if (x->rman != 0 || y->rman != 0) {
        mg num = mg_add(mg_mul(mg_mid_upper(x, L), mg_of(y)), mg_mul(mg_mid_upper(y, L), mg_of(x)));
        rad = mg_div(num, mg_mul_lower(yl, gap));
    }
    if (x->flags & CBO_ZERO) {
        xf zero;
        zero.sign = 0;
        zero.n = 0;
        zero.exp = 0;
        finish(z, &zero, rad, L);
        return;
    }
    /* Q = floor(mx * 2^prec / my), prec or prec+1 bits */
    uint64_t num[CBO_MAXL + 2], q[CBO_MAXL / 2 + 2], r[CBO_MAXL / 2 + 2];
    memset(num, 0, sizeof(uint64_t) * L64);
    memcpy(num + L64, x->m.d, sizeof(uint64_t) * L64);
    n_divrem(q, r, num, 2 * L64, y->m.d, L64);
    int remnz = 0;
    for (int i = 0; i < L64; i++) remnz |= r[i] != 0;
    xf v;
    v.sign = ((x->flags ^ y->flags) & CBO_SIGN) ? -1 : 1;
    v.n = L64 + 1;
    memcpy(v.d, q, sizeof(uint64_t) * (L64 + 1));
    v.exp = (int64_t)x->exp - y->exp - 64 * (int64_t)L64;
    xf_strip(&v);
    if (remnz) {
        /* append a sticky bit below: the exact quotient is not an integer */
        for (int i = v.n; i > 0; i--) v.d[i] = v.d[i - 1];
        v.d[0] = 1;
        v.n += 1;
        v.exp -= 64;
    }
    finish(z, &v, rad, L);
}

/* ------------------------------------------------------------------ acb */
void cbo_acb_zero(cbo_acb *z) {
    cbo_arb_zero(&z->re);
    cbo_arb_zero(&z->im);
}
void cbo_acb_set_si(cbo_acb *z, int64_t k, int L) {
    cbo_arb_set_si(&z->re, k, L);
    cbo_arb_zero(&z->im);
}
void cbo_acb_one(cbo_acb *z, int L) { cbo_acb_set_si(z, 1, L); }
void cbo_acb_neg(cbo_acb *z, const cbo_acb *x) {
    cbo_arb_neg(&z->re, &x->re);
    cbo_arb_neg(&z->im, &x->im);
}
void cbo_acb_add(cbo_acb *z, const cbo_acb *x, const cbo_acb *y, int L) {
    cbo_arb_add(&z->re, &x->re, &y->re, L);
    cbo_arb_add(&z->im, &x->im, &y->im, L);
}
void cbo_acb_sub(cbo_acb *z, const cbo_acb *x, const cbo_acb *y, int L) {
    cbo_arb_sub(&z->re, &x->re, &y->re, L);
    cbo_arb_sub(&z->im, &x->im, &y->im, L);
}
/* acb_mul: re = a c - b d, im = a d + b c, each rounded ONCE (arf_complex_mul semantics) */
void cbo_acb_mul(cbo_acb *z, const cbo_acb *x, const cbo_acb *y, int L) {
    cbo_acb t;
    cbo_arb_fdot2(&t.re, &x->re, &y->re, &x->im, &y->im, 1, L);
    cbo_arb_fdot2(&t.im, &x->re, &y->im, &x->im, &y->re, 0, L);
    *z = t;
}
/* acb_inv: t = c^2 + d^2 (one rounding, arf_sosq), re = c / t, im = -(d / t) */
void cbo_acb_inv(cbo_acb *z, const cbo_acb *x, int L) {
    cbo_arb t;
    cbo_acb r;
    cbo_arb_fdot2(&t, &x->re, &x->re, &x->im, &x->im, 0, L);
    cbo_arb_div(&r.re, &x->re, &t, L);
    cbo_arb_div(&r.im, &x->im, &t, L);
    cbo_arb_neg(&r.im, &r.im);
    *z = r;
}
/* acb_div: real divisor (imaginary part exactly zero) -> componentwise; else x * inv(y) */
void cbo_acb_div(cbo_acb *z, const cbo_acb *x, const cbo_acb *y, int L) {
    cbo_acb t;
    if (cbo_arb_is_zero(&y->im)) {
        cbo_arb_div(&t.re, &x->re, &y->re, L);
        cbo_arb_div(&t.im, &x->im, &y->re, L);
        *z = t;
        return;
    }
    cbo_acb_inv(&t, y, L);
    cbo_acb_mul(z, x, &t, L);
}
void cbo_acb_mul_si(cbo_acb *z, const cbo_acb *x, int64_t k, int L) {
    cbo_arb_mul_si(&z->re, &x->re, k, L);
    cbo_arb_mul_si(&z->im, &x->im, k, L);
}
void cbo_acb_add_si(cbo_acb *z, const cbo_acb *x, int64_t k, int L) {
    cbo_arb_add_si(&z->re, &x->re, k, L);
    if (z != x) z->im = x->im;
}
void cbo_acb_addmul(cbo_acb *z, const cbo_acb *x, const cbo_acb *y, int L) {
    cbo_acb t;
    cbo_acb_mul(&t, x, y, L);
    cbo_acb_add(z, z, &t, L);
}
int cbo_acb_is_zero(const cbo_acb *x) { return cbo_arb_is_zero(&x->re) && cbo_arb_is_zero(&x->im); }
int cbo_acb_contains_zero(const cbo_acb *x) {
    return cbo_arb_contains_zero(&x->re) && cbo_arb_contains_zero(&x->im);
}
int cbo_acb_is_finite(const cbo_acb *x) { return cbo_arb_is_finite(&x->re) && cbo_arb_is_finite(&x->im); }

/* ------------------------------------------------------------------ slot arrays */
void cbo_load(cbo_arb *z, const uint32_t *mant, const int32_t *meta, int L) {
    memset(z->m.w, 0, sizeof(z->m.w));
    memcpy(z->m.w, mant, sizeof(uint32_t) * L);
    z->exp = meta[0];
    z->flags = (uint32_t)meta[1];
    z->rman = (uint32_t)meta[2];
    z->rexp = meta[3];
}
void cbo_store(uint32_t *mant, int32_t *meta, const cbo_arb *x, int L) {
    memcpy(mant, x->m.w, sizeof(uint32_t) * L);
    meta[0] = x->exp;
    meta[1] = (int32_t)x->flags;
    meta[2] = (int32_t)x->rman;
    meta[3] = x->rexp;
}
static void load_acb(cbo_acb *z, const uint32_t *m, const int32_t *t, int64_t i, int L) {
    cbo_load(&z->re, m + (2 * i) * L, t + (2 * i) * 4, L);
    cbo_load(&z->im, m + (2 * i + 1) * L, t + (2 * i + 1) * 4, L);
}
static void store_acb(uint32_t *m, int32_t *t, int64_t i, const cbo_acb *z, int L) {
    cbo_store(m + (2 * i) * L, t + (2 * i) * 4, &z->re, L);
    cbo_store(m + (2 * i + 1) * L, t + (2 * i + 1) * 4, &z->im, L);
}

/* op codes shared with include/cascade_b200.h (CB200_OP_*) */
int cbo_ew_binary(int op, int L, int64_t n, uint32_t *zm, int32_t *zt, const uint32_t *xm,
                  const int32_t *xt, const uint32_t *ym, const int32_t *yt) {
    for (int64_t i = 0; i < n; i++) {
        cbo_acb x, y, z;
        load_acb(&x, xm, xt, i, L);
        load_acb(&y, ym, yt, i, L);
        switch (op) {
            case 0: cbo_acb_add(&z, &x, &y, L); break;
            case 1: cbo_acb_sub(&z, &x, &y, L); break;
            case 2: cbo_acb_mul(&z, &x, &y, L); break;
            case 3: cbo_acb_div(&z, &x, &y, L); break;
            case 4: cbo_acb_inv(&z, &x, L); break;
            default: return -1;
        }
        store_acb(zm, zt, i, &z, L);
    }
    return 0;
}
int cbo_ew_scalar_si(int op, int L, int64_t n, uint32_t *zm, int32_t *zt, const uint32_t *xm,
                     const int32_t *xt, const int64_t *k) {
    for (int64_t i = 0; i < n; i++) {
        cbo_acb x, z;
        load_acb(&x, xm, xt, i, L);
        switch (op) {
            case 16: cbo_acb_mul_si(&z, &x, k[i], L); break;
            case 17: cbo_acb_add_si(&z, &x, k[i], L); break;
            default: return -1;
        }
        store_acb(zm, zt, i, &z, L);
    }
    return 0;
}

/* ------------------------------------------------------------------ helpers of cbo_trans.c */
void cbo_arb_set_nan(cbo_arb *z) { arb_set_nan(z); }
/* radius += 2^e (rounded up) */
void cbo_arb_add_error_2exp(cbo_arb *z, int64_t e) {
    if (is_nan(z)) return;
    mg u;
    u.man = MAG_ONE_HALF;
    u.exp = e + 1;
    set_rad(z, mg_add(mg_of(z), u));
}
/* z = the exact ball of a 30-bit upper bound of |x| (|mid| rounded up to 30 bits, plus the radius) */
void cbo_arb_upper_ball(cbo_arb *z, const cbo_arb *x, int L) {
    if (is_nan(x)) {
        arb_set_nan(z);
        return;
    }
    mg u = mg_add(mg_mid_upper(x, L), mg_of(x));
    cbo_arb_zero(z);
    if (u.man == 0) return;
    z->m.w[L - 1] = u.man << 2;
    z->exp = u.exp;
    z->flags = 0;
}
/* z.rad += a 30-bit upper bound of |err| (arb_add_error) */
void cbo_arb_add_error_arb(cbo_arb *z, const cbo_arb *err, int L) {
    if (is_nan(z)) return;
    if (is_nan(err)) {
        arb_set_nan(z);
        return;
    }
    set_rad(z, mg_add(mg_of(z), mg_add(mg_mid_upper(err, L), mg_of(err))));
}
/* z *= 2^k, exact */
void cbo_arb_mul_2exp(cbo_arb *z, int64_t k) {
    if (is_nan(z)) return;
    if (!(z->flags & CBO_ZERO)) z->exp = (int32_t)(z->exp + k);
    if (z->rman) z->rexp = (int32_t)(z->rexp + k);
}
