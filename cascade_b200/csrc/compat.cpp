// compat.cpp -- host side of include/cascade_compat.h: Cascade's own API names over the batched
// CUDA entry points (batch of one).  Marshalling and control flow only; every ball operation is a
// call into the device library (there is no CPU arithmetic path).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/cascade_compat.h"

namespace {

const slong UNDEFINED_VAL = -0xFFFF;  // reference src/acb_ode.c:3

[[noreturn]] void die(const char* what, int rc) {
    fprintf(stderr, "cascade_b200: %s failed (rc = %d); there is no CPU fallback\n", what, rc);
    abort();
}
// Any precision 2 <= prec <= 4096 (the reference's own tests run at 30..157 bits, reference tests/fuchs.c:20): the
// device arithmetic is compiled for 128, 1024, 2048 and 4096 bits, and a call at `prec` bits runs at the smallest of
// those that is >= prec.  Every result is then a rigorous ball computed with AT LEAST the requested precision -- an
// enclosure of the same exact quantity, at least as tight as a prec-bit computation -- which is what every property
// the reference tests (containment, overlap, residual contains zero) needs.  Bit patterns differ from a prec-bit run,
// as they do between Arb versions.
int limbs_of(slong prec) {
    if (prec >= 2 && prec <= 128) return 4;
    if (prec <= 1024 && prec > 128) return 32;
    if (prec <= 2048 && prec > 1024) return 64;
    if (prec <= 4096 && prec > 2048) return 128;
    fprintf(stderr, "cascade_b200: precision %ld is not supported on the device (2 <= prec <= 4096)\n", prec);
    abort();
}

// compat ball (128 limbs, top aligned) -> device slot with L limbs (entry layout [L][S], slot s)
void to_slot(const arb_struct* x, int L, uint32_t* mant, cb_meta* meta, size_t S, size_t s) {
    bool lost = false;
    for (int k = 0; k < CB_MAX_LIMBS - L; k++) lost |= x->mant[k] != 0;
    for (int k = 0; k < L; k++) mant[(size_t)k * S + s] = x->mant[CB_MAX_LIMBS - L + k];
    meta[s] = x->meta;
    if (lost && !(x->meta.flags & CB_FLAG_NAN)) {
        // the value carries more bits than the target precision: truncate and widen the radius
        // by one ulp (what arb_set_round does); a 30-bit magnitude add, rounded up
        uint32_t uman = 1u << 29;
        int32_t uexp = x->meta.exp - 32 * L + 1;
        cb_meta& t = meta[s];
        if (t.rman == 0) {
            t.rman = uman;
            t.rexp = uexp;
        } else {
            int32_t sh = t.rexp - uexp;
            uint32_t m;
            int32_t e = t.rexp;
            if (sh >= 0) {
                m = t.rman + (sh >= 30 ? 0 : (uman >> sh)) + 1;
            } else {
                m = uman + ((-sh) >= 30 ? 0 : (t.rman >> (-sh))) + 1;
                e = uexp;
            }
            if (m >> 30) {
                m = (m >> 1) + (m & 1);
                e += 1;
            }
            t.rman = m;
            t.rexp = e;
        }
    }
}
void from_slot(arb_struct* x, int L, const uint32_t* mant, const cb_meta* meta, size_t S, size_t s) {
    memset(x->mant, 0, sizeof(x->mant));
    for (int k = 0; k < L; k++) x->mant[CB_MAX_LIMBS - L + k] = mant[(size_t)k * S + s];
    x->meta = meta[s];
}

// one entry array of n complex balls (S = 2n)
struct HostArr {
    int L;
    size_t entries, S;
    std::vector<uint32_t> m;
    std::vector<cb_meta> t;
    HostArr(int L_, size_t entries_, size_t S_) : L(L_), entries(entries_), S(S_), m(entries_ * L_ * S_, 0u), t(entries_ * S_) {
        for (auto& x : t) x = cb_meta{0, CB_FLAG_ZERO, 0, 0};
    }
    void put(size_t entry, size_t inst, const acb_struct* x) {
        to_slot(&x->real, L, m.data() + entry * L * S, t.data() + entry * S, S, 2 * inst);
        to_slot(&x->imag, L, m.data() + entry * L * S, t.data() + entry * S, S, 2 * inst + 1);
    }
    void get(size_t entry, size_t inst, acb_struct* x) const {
        from_slot(&x->real, L, m.data() + entry * L * S, t.data() + entry * S, S, 2 * inst);
        from_slot(&x->imag, L, m.data() + entry * L * S, t.data() + entry * S, S, 2 * inst + 1);
    }
};

void ew(int op, acb_struct* z, const acb_struct* x, const acb_struct* y, const int64_t* k, slong prec) {
    int L = limbs_of(prec);
    HostArr hx(L, 1, 2), hy(L, 1, 2), hz(L, 1, 2);
    hx.put(0, 0, x);
    if (y) hy.put(0, 0, y);
    int rc = cb200_ew_host(op, L, 1, hz.m.data(), hz.t.data(), hx.m.data(), hx.t.data(), y ? hy.m.data() : nullptr,
                           y ? hy.t.data() : nullptr, k, 0);
    if (rc) die("cb200_ew_host", rc);
    hz.get(0, 0, z);
}

void arb_set_exact_u64(arb_struct* x, uint64_t v, int neg, int32_t exp2) {
    // value = (-1)^neg * v * 2^exp2
    memset(x, 0, sizeof(*x));
    if (v == 0) {
        x->meta.flags = CB_FLAG_ZERO;
        return;
    }
    int sh = __builtin_clzll(v);
    uint64_t n = v << sh;
    x->mant[CB_MAX_LIMBS - 1] = (uint32_t)(n >> 32);
    x->mant[CB_MAX_LIMBS - 2] = (uint32_t)n;
    x->meta.exp = 64 - sh + exp2;
    x->meta.flags = neg ? CB_FLAG_SIGN : 0;
}
void arb_set_double(arb_struct* x, double d) {
    if (d == 0.0 || !isfinite(d)) {
        memset(x, 0, sizeof(*x));
        x->meta.flags = isfinite(d) ? CB_FLAG_ZERO : CB_FLAG_NAN;
        if (!isfinite(d)) x->meta.rman = CB_MAG_INF_MAN;
        return;
    }
    int e;
    double f = frexp(fabs(d), &e);              // f in [0.5, 1)
    uint64_t m = (uint64_t)ldexp(f, 53);        // 53-bit integer, exact
    arb_set_exact_u64(x, m, d < 0, e - 53);
}
double arb_mid_d(const arb_struct* x) {
    if (x->meta.flags & CB_FLAG_NAN) return NAN;
    if (x->meta.flags & CB_FLAG_ZERO) return 0.0;
    uint64_t top = ((uint64_t)x->mant[CB_MAX_LIMBS - 1] << 32) | x->mant[CB_MAX_LIMBS - 2];
    double v = ldexp((double)top, x->meta.exp - 64);
    return (x->meta.flags & CB_FLAG_SIGN) ? -v : v;
}
double arb_rad_d(const arb_struct* x) {
    if (x->meta.flags & CB_FLAG_NAN) return INFINITY;
    if (x->meta.rman == 0) return 0.0;
    return ldexp((double)x->meta.rman, x->meta.rexp - 30);
}
bool arb_exact_zero(const arb_struct* x) { return (x->meta.flags & CB_FLAG_ZERO) && x->meta.rman == 0; }
bool arb_has_zero(const arb_struct* x) {
    if (x->meta.flags & (CB_FLAG_NAN | CB_FLAG_ZERO)) return true;
    if (x->meta.rman == 0) return false;
    if (x->meta.exp != x->meta.rexp) return x->meta.exp < x->meta.rexp;
    uint32_t top30 = x->mant[CB_MAX_LIMBS - 1] >> 2;
    if (top30 != x->meta.rman) return top30 < x->meta.rman;
    if (x->mant[CB_MAX_LIMBS - 1] & 3u) return false;
    for (int k = 0; k < CB_MAX_LIMBS - 1; k++)
        if (x->mant[k]) return false;
    return true;
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------ minimal acb / acb_poly
void acb_init(acb_t x) { acb_zero(x); }
void acb_clear(acb_t) {}
void acb_zero(acb_t x) {
    memset(x, 0, sizeof(acb_struct));
    x->real.meta.flags = CB_FLAG_ZERO;
    x->imag.meta.flags = CB_FLAG_ZERO;
}
void acb_one(acb_t x) { acb_set_si(x, 1); }
void acb_set(acb_t z, const acb_t x) {
    if (z != x) memcpy(z, x, sizeof(acb_struct));
}
void acb_set_si(acb_t z, slong v) {
    acb_zero(z);
    arb_set_exact_u64(&z->real, v < 0 ? (uint64_t)0 - (uint64_t)v : (uint64_t)v, v < 0, 0);
}
void acb_set_d(acb_t z, double v) { acb_set_d_d(z, v, 0.0); }
void acb_set_d_d(acb_t z, double re, double im) {
    arb_set_double(&z->real, re);
    arb_set_double(&z->imag, im);
}
void acb_neg(acb_t z, const acb_t x) {
    acb_set(z, x);
    if (!(z->real.meta.flags & (CB_FLAG_ZERO | CB_FLAG_NAN))) z->real.meta.flags ^= CB_FLAG_SIGN;
    if (!(z->imag.meta.flags & (CB_FLAG_ZERO | CB_FLAG_NAN))) z->imag.meta.flags ^= CB_FLAG_SIGN;
}
int acb_is_zero(const acb_t x) { return arb_exact_zero(&x->real) && arb_exact_zero(&x->imag); }
int acb_is_finite(const acb_t x) { return !((x->real.meta.flags | x->imag.meta.flags) & CB_FLAG_NAN); }
int acb_contains_zero(const acb_t x) { return arb_has_zero(&x->real) && arb_has_zero(&x->imag); }
double acb_get_mid_re_d(const acb_t x) { return arb_mid_d(&x->real); }
double acb_get_mid_im_d(const acb_t x) { return arb_mid_d(&x->imag); }
double acb_get_rad_re_d(const acb_t x) { return arb_rad_d(&x->real); }
void acb_printd(const acb_t x, slong digits) {
    printf("(%.*g + %.*gj)  +/-  (%.3g, %.3gj)", (int)digits, arb_mid_d(&x->real), (int)digits, arb_mid_d(&x->imag),
           arb_rad_d(&x->real), arb_rad_d(&x->imag));
}
acb_ptr _acb_vec_init(slong n) {
    acb_ptr v = (acb_ptr)malloc(sizeof(acb_struct) * (n > 0 ? n : 1));
    for (slong i = 0; i < n; i++) acb_zero(v + i);
    return v;
}
void _acb_vec_clear(acb_ptr v, slong) { free(v); }

void acb_poly_init(acb_poly_t p) {
    p->coeffs = nullptr;
    p->length = p->alloc = 0;
}
void acb_poly_clear(acb_poly_t p) {
    free(p->coeffs);
    p->coeffs = nullptr;
    p->length = p->alloc = 0;
}
void acb_poly_fit_length(acb_poly_t p, slong len) {
    if (len <= p->alloc) return;
    slong na = len > 2 * p->alloc ? len : 2 * p->alloc;
    p->coeffs = (acb_struct*)realloc(p->coeffs, sizeof(acb_struct) * na);
    for (slong i = p->alloc; i < na; i++) acb_zero(p->coeffs + i);
    p->alloc = na;
}
static void poly_normalise(acb_poly_t p) {
    while (p->length > 0 && acb_is_zero(p->coeffs + p->length - 1)) p->length--;
}
void acb_poly_zero(acb_poly_t p) {
    for (slong i = 0; i < p->length; i++) acb_zero(p->coeffs + i);
    p->length = 0;
}
void acb_poly_one(acb_poly_t p) {
    acb_poly_zero(p);
    acb_poly_set_coeff_si(p, 0, 1);
}
slong acb_poly_length(const acb_poly_t p) { return p->length; }
void acb_poly_set_coeff_acb(acb_poly_t p, slong n, const acb_t c) {
    acb_poly_fit_length(p, n + 1);
    if (n + 1 > p->length) {
        for (slong i = p->length; i < n; i++) acb_zero(p->coeffs + i);
        p->length = n + 1;
    }
    acb_set(p->coeffs + n, c);
    poly_normalise(p);
}
void acb_poly_set_coeff_si(acb_poly_t p, slong n, slong c) {
    acb_t t;
    acb_set_si(t, c);
    acb_poly_set_coeff_acb(p, n, t);
}
void acb_poly_get_coeff_acb(acb_t c, const acb_poly_t p, slong n) {
    if (n < 0 || n >= p->length)
        acb_zero(c);
    else
        acb_set(c, p->coeffs + n);
}
int acb_poly_is_zero(const acb_poly_t p) { return p->length == 0; }
void acb_poly_printd(const acb_poly_t p, slong digits) {
    printf("[");
    for (slong i = 0; i < p->length; i++) {
        acb_printd(p->coeffs + i, digits);
        if (i + 1 < p->length) printf("\n");
    }
    printf("]\n");
}
void flint_cleanup(void) {}

void acb_add(acb_t z, const acb_t x, const acb_t y, slong prec) { ew(CB200_OP_ADD, z, x, y, nullptr, prec); }
void acb_sub(acb_t z, const acb_t x, const acb_t y, slong prec) { ew(CB200_OP_SUB, z, x, y, nullptr, prec); }
void acb_mul(acb_t z, const acb_t x, const acb_t y, slong prec) { ew(CB200_OP_MUL, z, x, y, nullptr, prec); }
void acb_div(acb_t z, const acb_t x, const acb_t y, slong prec) { ew(CB200_OP_DIV, z, x, y, nullptr, prec); }
void acb_mul_si(acb_t z, const acb_t x, slong k, slong prec) {
    int64_t kk = k;
    ew(CB200_OP_MUL_SI, z, x, nullptr, &kk, prec);
}
void acb_add_si(acb_t z, const acb_t x, slong k, slong prec) {
    int64_t kk = k;
    ew(CB200_OP_ADD_SI, z, x, nullptr, &kk, prec);
}
void acb_poly_evaluate(acb_t y, const acb_poly_t f, const acb_t x, slong prec) {
    int L = limbs_of(prec);
    slong len = f->length;
    HostArr hf(L, len > 0 ? len : 1, 2), hp(L, 1, 2), ho(L, 1, 2);
    for (slong i = 0; i < len; i++) hf.put(i, 0, f->coeffs + i);
    hp.put(0, 0, x);
    int rc = cb200_horner_batch_host(L, len, 1, 1, hf.m.data(), hf.t.data(), 1, hp.m.data(), hp.t.data(), ho.m.data(),
                                     ho.t.data(), 0);
    if (rc) die("cb200_horner_batch_host", rc);
    ho.get(0, 0, y);
}

// ------------------------------------------------------------------ operators
void acb_ode_init_blank(acb_ode_t ODE, slong degree, slong order) {
    ODE->order = order;
    ODE->degree = degree;
    ODE->polys = nullptr;
    ODE->alloc = 0;
    ODE->valuation = UNDEFINED_VAL;
    if (degree < 0 || order <= 0) return;
    ODE->alloc = (order + 1) * (degree + 1);
    ODE->polys = _acb_vec_init(ODE->alloc);
}
void acb_ode_init(acb_ode_t ODE, acb_poly_t* polys, slong order) {
    while (order > 0 && acb_poly_is_zero(polys[order])) order--;
    slong deg = 0;
    for (slong i = 0; i <= order; i++)
        if (acb_poly_length(polys[i]) - 1 > deg) deg = acb_poly_length(polys[i]) - 1;
    acb_ode_init_blank(ODE, order <= 0 ? -1 : deg, order);
    for (slong i = 0; i <= order && ODE->polys; i++) {
        for (slong j = 0; j < acb_poly_length(polys[i]); j++) acb_poly_get_coeff_acb(acb_ode_coeff(ODE, i, j), polys[i], j);
    }
}
void acb_ode_clear(acb_ode_t ODE) {
    if (ODE->alloc <= 0) return;
    free(ODE->polys);
    ODE->polys = nullptr;
    ODE->alloc = 0;
}
void acb_ode_set(acb_ode_t out, acb_ode_t in) {
    if (out->alloc < in->alloc) {
        acb_ode_clear(out);
        acb_ode_init_blank(out, in->degree, in->order);  // the intended argument order
    }
    memcpy(out->polys, in->polys, sizeof(acb_struct) * in->alloc);
    out->degree = in->degree;
    out->order = in->order;
    out->valuation = acb_ode_valuation(in);
}
void acb_ode_dump(acb_ode_t ODE, char* file) {
    FILE* out = file ? fopen(file, "w") : stdout;
    if (!out) return;
    fprintf(out, "Order: %ld\nDegree: %ld\n", ODE->order, ODE->degree);
    for (slong i = 0; i <= ODE->order; i++) {
        fprintf(out, "acb_ode_poly(ODE, %ld) = ", i);
        for (slong j = 0; j <= ODE->degree; j++) {
            const acb_struct* c = acb_ode_coeff(ODE, i, j);
            fprintf(out, "(%.20g + %.20gj)  +/-  (%.3g, %.3gj)\t", arb_mid_d(&c->real), arb_mid_d(&c->imag),
                    arb_rad_d(&c->real), arb_rad_d(&c->imag));
        }
        fprintf(out, "\n");
    }
    if (out != stdout) fclose(out);
}
// reference src/acb_ode.c:137-157: divide every polynomial by the largest common power of z.  The reference
// scans only the rows below the leading one (i < order) and repacks the rows at the new stride in place.
slong acb_ode_reduce(acb_ode_t ODE) {
    slong reduced = 0;
    int all_zero = 1;
    do {
        for (slong i = 0; i < ODE->order; i++) {
            if (reduced > ODE->degree) { all_zero = 0; break; }
            all_zero &= acb_is_zero(acb_ode_coeff(ODE, i, reduced));
        }
        reduced++;
    } while (all_zero && reduced <= ODE->degree + 1);
    reduced -= 1;
    if (reduced <= 0) return 0;
    slong new_deg = ODE->degree - reduced;
    if (new_deg < 0) return 0;
    for (slong i = 0; i <= ODE->order; i++)
        for (slong j = 0; j <= new_deg; j++)
            acb_set(ODE->polys + i * (new_deg + 1) + j, ODE->polys + i * (ODE->degree + 1) + j + reduced);
    ODE->degree = new_deg;
    ODE->valuation = UNDEFINED_VAL;
    return reduced;
}

slong acb_ode_valuation(acb_ode_t ODE) {
    if (ODE->valuation != UNDEFINED_VAL) return ODE->valuation;
    slong val = ODE->degree;
    for (slong i = 0; i <= ODE->order; i++) {
        slong v = 0;
        while (v <= ODE->degree && acb_is_zero(acb_ode_coeff(ODE, i, v))) v++;
        if (v - i < val) val = v - i;
    }
    ODE->valuation = val;
    return val;
}
void acb_ode_shift(acb_ode_t out, acb_ode_t in, acb_srcptr a, slong bits) {
    // reference src/acb_ode.c:124-135: copy, then the Horner Taylor shift of every row -- one launch (ode_shift_kernel)
    acb_ode_set(out, in);
    if (acb_is_zero(a) || in->degree == 0) return;
    int L = limbs_of(bits);
    slong ne = (in->order + 1) * (in->degree + 1);
    HostArr hi(L, ne, 2), ha(L, 1, 2), ho(L, ne, 2);
    for (slong e = 0; e < ne; e++) hi.put(e, 0, in->polys + e);
    ha.put(0, 0, a);
    int rc = cb200_ode_shift_host(L, (int)in->order, (int)in->degree, 1, hi.m.data(), hi.t.data(), 1, ha.m.data(), ha.t.data(), 1,
                                  ho.m.data(), ho.t.data(), 0);
    if (rc) die("acb_ode_shift (cb200_ode_shift_host)", rc);
    for (slong e = 0; e < ne; e++) ho.get(e, 0, out->polys + e);
    out->valuation = UNDEFINED_VAL;
}

// ------------------------------------------------------------------ solutions
static void gens_put(HostArr& h, const acb_ode_solution_struct* sol, slong cap);
static slong gens_cap(const acb_ode_solution_struct* sol, slong at_least);
void acb_ode_solution_init(acb_ode_solution_t sol, acb_t rho, slong mul, slong alpha) {
    sol->mul = mul;
    sol->M = mul + alpha;
    acb_set(sol->rho, rho);
    sol->gens = (acb_poly_struct*)malloc(sizeof(acb_poly_struct) * (sol->M > 0 ? sol->M : 1));
    for (slong i = 0; i < sol->M; i++) acb_poly_init(sol->gens + i);
}
void acb_ode_solution_clear(acb_ode_solution_t sol) {
    for (slong i = 0; i < sol->M; i++) acb_poly_clear(sol->gens + i);
    free(sol->gens);
    sol->gens = nullptr;
}

// how acb_pow treats the exponent: 0 exact zero, 1 exact integer 0 <= e < 2^31, 2 general
static int pow_mode_of(const acb_struct* rho, uint64_t* e_out) {
    const arb_struct* r = &rho->real;
    *e_out = 0;
    bool exact = arb_exact_zero(&rho->imag) && r->meta.rman == 0 && !(r->meta.flags & CB_FLAG_NAN);
    if (!exact) return 2;
    if (r->meta.flags & CB_FLAG_ZERO) return 0;
    if ((r->meta.flags & CB_FLAG_SIGN) || r->meta.exp < 1 || r->meta.exp > 31) return 2;
    for (int k = 0; k < CB_MAX_LIMBS - 1; k++)
        if (r->mant[k]) return 2;
    uint32_t top = r->mant[CB_MAX_LIMBS - 1];
    uint64_t e = top >> (32 - r->meta.exp);
    if ((e << (32 - r->meta.exp)) != top) return 2;
    *e_out = e;
    return 1;
}
// reference src/acb_ode_solution.c:24-58: Horner evaluation of the M series, acb_log, acb_pow and the
// combination all run in ONE device launch (cb200_solution_evaluate_host)
void acb_ode_solution_evaluate(acb_t out, acb_ode_solution_t sol, acb_t a, slong prec) {
    int L = limbs_of(prec);
    if (sol->M < 1 || sol->M > 16) die("acb_ode_solution_evaluate: M outside 1..16", -1);
    slong cap = gens_cap(sol, 1);
    HostArr hg(L, (size_t)(sol->M * cap), 2), hrho(L, 1, 2), hp(L, 1, 2), ho(L, 1, 2);
    gens_put(hg, sol, cap);
    hrho.put(0, 0, sol->rho);
    hp.put(0, 0, a);
    uint64_t e = 0;
    int mode = pow_mode_of(sol->rho, &e);
    int rc = cb200_solution_evaluate_host(L, (int)sol->M, cap, 1, hg.m.data(), hg.t.data(), hrho.m.data(), hrho.t.data(), mode,
                                          e, hp.m.data(), hp.t.data(), ho.m.data(), ho.t.data(), 0);
    if (rc) die("acb_ode_solution_evaluate (cb200_solution_evaluate_host)", rc);
    ho.get(0, 0, out);
}
static void trans1(int which, acb_struct* z, const acb_struct* x, slong prec, const char* what) {
    int L = limbs_of(prec);
    HostArr hx(L, 1, 2), hz(L, 1, 2);
    hx.put(0, 0, x);
    int rc = cb200_trans_host(which, L, 1, hz.m.data(), hz.t.data(), hx.m.data(), hx.t.data(), 0);
    if (rc) die(what, rc);
    hz.get(0, 0, z);
}
// acb_pow(z, x, y): 1 for y = 0, binary exponentiation for an exact integer 0 <= y < 2^31, exp(y log x) otherwise --
// the evaluate kernel with the constant series 1 (out = 1 * x^y; the product with the exact 1 is exact)
void acb_pow(acb_t z, const acb_t x, const acb_t y, slong prec) {
    int L = limbs_of(prec);
    HostArr hg(L, 1, 2), hrho(L, 1, 2), hp(L, 1, 2), ho(L, 1, 2);
    acb_t one;
    acb_one(one);
    hg.put(0, 0, one);
    hrho.put(0, 0, y);
    hp.put(0, 0, x);
    uint64_t e = 0;
    int mode = pow_mode_of(y, &e);
    int rc = cb200_solution_evaluate_host(L, 1, 1, 1, hg.m.data(), hg.t.data(), hrho.m.data(), hrho.t.data(), mode, e,
                                          hp.m.data(), hp.t.data(), ho.m.data(), ho.t.data(), 0);
    if (rc) die("acb_pow (cb200_solution_evaluate_host)", rc);
    ho.get(0, 0, z);
}
void acb_exp(acb_t z, const acb_t x, slong prec) { trans1(CB200_TRANS_EXP, z, x, prec, "acb_exp (cb200_trans_host)"); }
void acb_log(acb_t z, const acb_t x, slong prec) { trans1(CB200_TRANS_LOG, z, x, prec, "acb_log (cb200_trans_host)"); }

// ------------------------------------------------------------------ examples
void acb_ode_legendre(acb_ode_t ODE, ulong n) {
    acb_ode_init_blank(ODE, 2, 2);
    acb_set_si(acb_ode_coeff(ODE, 0, 0), (slong)(n * (n + 1)));
    acb_set_si(acb_ode_coeff(ODE, 1, 1), -2);
    acb_set_si(acb_ode_coeff(ODE, 2, 0), 1);
    acb_set_si(acb_ode_coeff(ODE, 2, 2), -1);
}
void acb_ode_bessel(acb_ode_t ODE, acb_t nu, slong bits) {
    acb_ode_init_blank(ODE, 2, 2);
    acb_set_si(acb_ode_coeff(ODE, 2, 2), 1);
    acb_set_si(acb_ode_coeff(ODE, 1, 1), 1);
    acb_set_si(acb_ode_coeff(ODE, 0, 2), 1);
    acb_mul(nu, nu, nu, bits);
    acb_neg(acb_ode_coeff(ODE, 0, 0), nu);
}
void acb_ode_hypgeom(acb_ode_t ODE, acb_t a, acb_t b, acb_t c, slong bits) {
    acb_ode_init_blank(ODE, 2, 2);
    acb_t temp;
    acb_set_si(acb_ode_coeff(ODE, 2, 1), 1);
    acb_set_si(acb_ode_coeff(ODE, 2, 2), -1);
    acb_set(acb_ode_coeff(ODE, 1, 0), c);
    acb_one(temp);
    acb_add(temp, temp, a, bits);
    acb_add(temp, temp, b, bits);
    acb_neg(acb_ode_coeff(ODE, 1, 1), temp);
    acb_mul(temp, a, b, bits);
    acb_neg(acb_ode_coeff(ODE, 0, 0), temp);
}

// ------------------------------------------------------------------ solvers
void acb_ode_solve_fuchs(acb_poly_t res, acb_ode_t ODE, slong deg, slong bits) {
    int L = limbs_of(bits);
    slong v = acb_ode_valuation(ODE);
    slong ne = (ODE->order + 1) * (ODE->degree + 1);
    HostArr ho(L, ne, 2), hr(L, deg + 1, 2);
    for (slong e = 0; e < ne; e++) ho.put(e, 0, ODE->polys + e);
    for (slong k = 0; k < -v && k <= deg; k++) {
        acb_t c;
        acb_poly_get_coeff_acb(c, res, k);
        hr.put(k, 0, c);
    }
    int rc = cb200_fuchs_batch_host(L, (int)ODE->order, (int)ODE->degree, (int)v, deg, 1, ho.m.data(), ho.t.data(), 1,
                                    hr.m.data(), hr.t.data(), 0);
    if (rc) die("acb_ode_solve_fuchs (cb200_fuchs_batch_host)", rc);
    acb_poly_fit_length(res, deg + 1);
    for (slong k = 0; k <= deg; k++) hr.get(k, 0, res->coeffs + k);
    if (res->length < deg + 1) res->length = deg + 1;
    poly_normalise(res);
}

void analytic_continuation(acb_poly_t res, acb_ode_t ODE, acb_srcptr path, slong len, slong deg, slong bits) {
    int L = limbs_of(bits);
    int ord = (int)ODE->order;
    if (len < 2) return;
    if (ord > 4) die("analytic_continuation: order > 4", -1);
    slong ne = (ODE->order + 1) * (ODE->degree + 1);
    HostArr ho(L, ne, 2), hp(L, len, 2), hy(L, ord, 2);
    for (slong e = 0; e < ne; e++) ho.put(e, 0, ODE->polys + e);
    for (slong t = 0; t < len; t++) hp.put(t, 0, path + t);
    for (int k = 0; k < ord; k++) {
        acb_t c;
        acb_poly_get_coeff_acb(c, res, k);
        hy.put(k, 0, c);
    }
    // the reference leaves the COMPLETE re-expanded series (deg + 1 coefficients) in res after the last corner
    // (src/fuchs_solver.c:159): the _series entry point returns it
    HostArr hr(L, deg + 1, 2);
    int rc = cb200_continuation_series_host(L, ord, (int)ODE->degree, deg, 1, len, ho.m.data(), ho.t.data(), hp.m.data(),
                                            hp.t.data(), hy.m.data(), hy.t.data(), hr.m.data(), hr.t.data(), 0);
    if (rc) die("analytic_continuation (cb200_continuation_series_host)", rc);
    acb_poly_zero(res);
    acb_poly_fit_length(res, deg + 1);
    for (slong k = 0; k <= deg; k++) hr.get(k, 0, res->coeffs + k);
    res->length = deg + 1;
    poly_normalise(res);
}

void indicial_polynomial_evaluate(acb_t result, acb_ode_t ODE, slong nu, acb_t rho, slong shift, slong prec) {
    nu += acb_ode_valuation(ODE);
    if (nu > ODE->degree) {
        acb_zero(result);
        return;
    }
    acb_t temp1, out, prod;
    acb_zero(out);
    for (slong lambda = clamp(ODE->degree - nu, 0, ODE->order); lambda >= 0; lambda--) {
        acb_add_si(temp1, rho, shift - lambda, prec);
        acb_mul(prod, out, temp1, prec);
        acb_set(out, prod);
        if (lambda + nu < 0) continue;
        acb_add(out, out, acb_ode_coeff(ODE, lambda, lambda + nu), prec);
    }
    acb_set(result, out);
}

void _acb_ode_solve_frobenius(acb_poly_t res, acb_ode_t ODE, acb_ode_solution_t rhs, slong sol_degree, slong prec) {
    int L = limbs_of(prec);
    slong v = acb_ode_valuation(ODE);
    slong ne = (ODE->order + 1) * (ODE->degree + 1);
    // the right-hand side is marshalled before res is written: res may alias rhs->gens
    // (reference src/frobenius_solver.c:127 passes sol->gens as both)
    slong hl = rhs->gens->length;
    HostArr ho(L, ne, 2), hrho(L, 1, 2), hr(L, sol_degree + 1, 2), hh(L, hl > 0 ? hl : 1, 2);
    for (slong e = 0; e < ne; e++) ho.put(e, 0, ODE->polys + e);
    for (slong k = 0; k < hl; k++) hh.put(k, 0, rhs->gens->coeffs + k);
    hrho.put(0, 0, rhs->rho);
    int rc = cb200_frobenius1_rhs_batch_host(L, (int)ODE->order, (int)ODE->degree, (int)v, sol_degree, 1, ho.m.data(),
                                             ho.t.data(), 1, hrho.m.data(), hrho.t.data(), 1, hh.m.data(), hh.t.data(), hl, 1,
                                             hr.m.data(), hr.t.data(), 0);
    if (rc) die("_acb_ode_solve_frobenius (cb200_frobenius1_rhs_batch_host)", rc);
    acb_poly_zero(res);
    acb_poly_fit_length(res, sol_degree + 1);
    for (slong k = 0; k <= sol_degree; k++) hr.get(k, 0, res->coeffs + k);
    res->length = sol_degree + 1;
    poly_normalise(res);
}

// sol->gens <-> device array [M * cap entries][S = 2]
static void gens_put(HostArr& h, const acb_ode_solution_struct* sol, slong cap) {
    for (slong i = 0; i < sol->M; i++)
        for (slong j = 0; j < cap && j < sol->gens[i].length; j++) h.put((size_t)(i * cap + j), 0, sol->gens[i].coeffs + j);
}
static void gens_get(const HostArr& h, acb_ode_solution_struct* sol, slong cap) {
    for (slong i = 0; i < sol->M; i++) {
        acb_poly_struct* g = sol->gens + i;
        acb_poly_zero(g);
        acb_poly_fit_length(g, cap);
        for (slong j = 0; j < cap; j++) h.get((size_t)(i * cap + j), 0, g->coeffs + j);
        g->length = cap;
        poly_normalise(g);
    }
}
static slong gens_cap(const acb_ode_solution_struct* sol, slong at_least) {
    slong cap = at_least;
    for (slong i = 0; i < sol->M; i++)
        if (sol->gens[i].length > cap) cap = sol->gens[i].length;
    return cap > 0 ? cap : 1;
}
static void solution_op(int which, acb_ode_solution_struct* sol, slong nu, acb_poly_struct* f, slong prec, const char* what) {
    int L = limbs_of(prec);
    if (sol->M > 16) die("solution bookkeeping with M > 16", -1);
    slong cap = gens_cap(sol, which == CB200_SOL_EXTEND ? nu + 1 : 1);
    slong fcap = (f && f->length > 0) ? f->length : 1;
    HostArr hg(L, (size_t)(sol->M > 0 ? sol->M * cap : 1), 2), hf(L, (size_t)fcap, 2), hrho(L, 1, 2);
    gens_put(hg, sol, cap);
    hrho.put(0, 0, sol->rho);
    if (f)
        for (slong j = 0; j < f->length; j++) hf.put((size_t)j, 0, f->coeffs + j);
    int rc = cb200_solution_op_host(which, L, 1, (int)sol->mul, (int)sol->M, cap, nu, hrho.m.data(), hrho.t.data(),
                                    hg.m.data(), hg.t.data(), hf.m.data(), hf.t.data(), fcap, 0);
    if (rc) die(what, rc);
    gens_get(hg, sol, cap);
    if (f) acb_poly_zero(f);  // consumed, as in the reference (src/acb_ode_solution.c:93,107)
}
void _acb_ode_solution_update(acb_ode_solution_t sol, acb_poly_t f, slong prec) {
    solution_op(CB200_SOL_UPDATE, sol, 0, f, prec, "_acb_ode_solution_update (cb200_solution_op_host)");
}
void _acb_ode_solution_extend(acb_ode_solution_t sol, slong nu, acb_poly_t g_nu, slong prec) {
    solution_op(CB200_SOL_EXTEND, sol, nu, g_nu, prec, "_acb_ode_solution_extend (cb200_solution_op_host)");
}
void _acb_ode_solution_normalize(acb_ode_solution_t sol, slong prec) {
    solution_op(CB200_SOL_NORMALIZE, sol, 0, nullptr, prec, "_acb_ode_solution_normalize (cb200_solution_op_host)");
}

void indicial_polynomial(acb_poly_t result, acb_ode_t ODE, slong nu, slong shift, slong prec) {
    int L = limbs_of(prec);
    slong v = acb_ode_valuation(ODE);
    slong ne = (ODE->order + 1) * (ODE->degree + 1), ord = ODE->order;
    HostArr ho(L, ne, 2), hr(L, ord + 2, 2);
    for (slong e = 0; e < ne; e++) ho.put(e, 0, ODE->polys + e);
    int rc = cb200_indicial_polynomial_host(L, (int)ord, (int)ODE->degree, (int)v, 1, ho.m.data(), ho.t.data(), 1, nu, shift,
                                            hr.m.data(), hr.t.data(), 0);
    if (rc) die("indicial_polynomial (cb200_indicial_polynomial_host)", rc);
    acb_poly_zero(result);
    acb_poly_fit_length(result, ord + 1);
    for (slong k = 0; k <= ord; k++) hr.get(k, 0, result->coeffs + k);
    result->length = ord + 1;
    poly_normalise(result);
}

void acb_ode_solve_frobenius(acb_ode_solution_t sol, acb_ode_t ODE, slong sol_degree, slong prec) {
    if (sol->M == 1) {
        _acb_ode_solve_frobenius(sol->gens, ODE, sol, sol_degree, prec);
        return;
    }
    // M > 1: the polynomial-in-rho recurrence (reference src/frobenius_solver.c:131-205), one launch
    int L = limbs_of(prec);
    if (sol->M > 16 || ODE->degree > 8 || ODE->degree < 1) die("acb_ode_solve_frobenius: M > 16 or degree outside 1..8", -1);
    slong v = acb_ode_valuation(ODE);
    slong ne = (ODE->order + 1) * (ODE->degree + 1), cap = sol_degree + 1;
    HostArr ho(L, ne, 2), hrho(L, 1, 2), hg(L, (size_t)(sol->M * cap), 2);
    for (slong e = 0; e < ne; e++) ho.put(e, 0, ODE->polys + e);
    hrho.put(0, 0, sol->rho);
    int rc = cb200_frobenius_general_batch_host(L, (int)ODE->order, (int)ODE->degree, (int)v, sol_degree, 1, ho.m.data(),
                                                ho.t.data(), 1, hrho.m.data(), hrho.t.data(), 1, (int)sol->mul, (int)sol->M,
                                                hg.m.data(), hg.t.data(), 0);
    if (rc) die("acb_ode_solve_frobenius (cb200_frobenius_general_batch_host)", rc);
    gens_get(hg, sol, cap);
}


// ------------------------------------------------------------------ monodromy (reference src/fuchs_solver.c:3-94,165-206)
// Control-plane scalars (the convergence radius that sizes the path and the truncation order) are
// formed in double precision on the host: the reference uses them only through arb_get_mid_arb
// (:178) and as an integer count (:190), and its own continuation adds no truncation bound, so no
// rigour hangs on them.  Every BALL of the computation -- the 256 path corners exp(2 pi i k / 256),
// the shifted operators, the series and their re-expansions -- is computed on the device.
void arb_init(arb_t x) { memset(x, 0, sizeof(arb_struct)); x->meta.flags = CB_FLAG_ZERO; }
void arb_clear(arb_t) {}
void arb_set_d(arb_t x, double v) { arb_set_double(x, v); }
double arb_get_mid_d(const arb_t x) { return arb_mid_d(x); }
double arb_get_rad_d(const arb_t x) { return arb_rad_d(x); }
int arb_is_finite(const arb_t x) { return !(x->meta.flags & CB_FLAG_NAN); }

void acb_mat_init(acb_mat_t m, slong r, slong c) {
    m->r = r;
    m->c = c;
    m->entries = _acb_vec_init(r * c);
    m->rows = (acb_ptr*)malloc(sizeof(acb_ptr) * (r > 0 ? r : 1));
    for (slong i = 0; i < r; i++) m->rows[i] = m->entries + i * c;
}
void acb_mat_clear(acb_mat_t m) {
    _acb_vec_clear(m->entries, m->r * m->c);
    free(m->rows);
}

void radius_of_convergence(arb_t rad_of_conv, acb_ode_t ODE, slong n, slong bits) {
    // reference src/fuchs_solver.c:3-54, rigorous on the device ball type (cb200_radius_of_convergence_host: Graeffe
    // iterations, Fujiwara's bound in its sum form, the sharpness of the bound as the error bar)
    int L = limbs_of(bits);
    slong ne = (ODE->order + 1) * (ODE->degree + 1);
    HostArr ho(L, ne, 2), hr(L, 1, 2);
    for (slong e = 0; e < ne; e++) ho.put(e, 0, ODE->polys + e);
    int32_t n_done = 0;
    int rc = cb200_radius_of_convergence_host(L, (int)ODE->order, (int)ODE->degree, 1, ho.m.data(), ho.t.data(), 1, n,
                                              hr.m.data(), hr.t.data(), &n_done, 0);
    if (rc) die("radius_of_convergence (cb200_radius_of_convergence_host)", rc);
    acb_t z;
    hr.get(0, 0, z);
    *rad_of_conv = z->real;
}

slong truncation_order(arb_t eta, arb_t alpha, slong bits) {  // reference :56-94, on the device ball type
    int L = limbs_of(bits);
    HostArr he(L, 1, 2), ha(L, 1, 2), hn(L, 1, 2);
    acb_t e, a;
    memset(e, 0, sizeof(acb_t));
    memset(a, 0, sizeof(acb_t));
    e->real = *eta;
    e->imag.meta.flags = CB_FLAG_ZERO;
    a->real = *alpha;
    a->imag.meta.flags = CB_FLAG_ZERO;
    he.put(0, 0, e);
    ha.put(0, 0, a);
    int64_t n = 0;
    int rc = cb200_truncation_order_host(L, 1, he.m.data(), he.t.data(), ha.m.data(), ha.t.data(), bits, hn.m.data(),
                                         hn.t.data(), &n, 0);
    if (rc) die("truncation_order (cb200_truncation_order_host)", rc);
    return (slong)n;
}

void find_monodromy_matrix(acb_mat_t mono, acb_ode_t ODE, slong bits) {
    int L = limbs_of(bits);
    arb_t rad;
    radius_of_convergence(rad, ODE, 40, bits);
    double r;
    if (arb_exact_zero(rad)) return;
    if (!arb_is_finite(rad)) r = 1.0;                // reference :173-174
    else r = arb_mid_d(rad) / 2;                      // :176-178 (midpoint of half the radius)
    const slong steps = 256;
    slong ord = ODE->order;
    if (ord > 4) die("find_monodromy_matrix: order > 4", -1);
    // corners r * exp(2 pi i k / steps): i pi = log(-1), theta_k = (i pi) * k / 128, one batched exp launch
    acb_t minus1, ipi;
    acb_set_si(minus1, -1);
    acb_log(ipi, minus1, bits);
    HostArr hx(L, 1, 2 * steps), hk(L, 1, 2 * steps), hz(L, 1, 2 * steps), hr(L, 1, 2 * steps);
    std::vector<int64_t> kk(steps);
    acb_t rb;
    acb_set_d(rb, r);
    for (slong k = 0; k < steps; k++) {
        hx.put(0, k, ipi);
        hr.put(0, k, rb);
        kk[k] = k;
    }
    int rc = cb200_ew_host(CB200_OP_MUL_SI, L, steps, hk.m.data(), hk.t.data(), hx.m.data(), hx.t.data(), nullptr, nullptr,
                           kk.data(), 0);
    if (rc) die("find_monodromy_matrix (cb200_ew_host)", rc);
    for (auto& t : hk.t) {  // divide by 128: an exact exponent shift
        if (!(t.flags & (CB_FLAG_ZERO | CB_FLAG_NAN))) t.exp -= 7;
        if (t.rman) t.rexp -= 7;
    }
    rc = cb200_trans_host(CB200_TRANS_EXP, L, steps, hz.m.data(), hz.t.data(), hk.m.data(), hk.t.data(), 0);
    if (rc) die("find_monodromy_matrix (cb200_trans_host)", rc);
    rc = cb200_ew_host(CB200_OP_MUL, L, steps, hx.m.data(), hx.t.data(), hz.m.data(), hz.t.data(), hr.m.data(), hr.t.data(),
                       nullptr, 0);
    if (rc) die("find_monodromy_matrix (cb200_ew_host)", rc);
    acb_ptr path = _acb_vec_init(steps + 1);
    for (slong k = 0; k < steps; k++) hx.get(0, k, path + k);
    acb_set(path + steps, path);  // closed loop (:191)
    // number of coefficients: eta = |path[0] - path[1]|, alpha = r (:187-190)
    arb_t eta, alpha;
    arb_set_double(eta, 2 * r * sin(3.14159265358979323846 / steps));
    arb_set_double(alpha, r);
    slong n = truncation_order(eta, alpha, bits);
    if (n < ord) n = ord;
    // the `order` unit initial vectors continue together: one batch along the path (:195-202)
    slong ne = (ODE->order + 1) * (ODE->degree + 1);
    HostArr ho(L, ne, 2), hp(L, steps + 1, 2), hy(L, ord, 2 * ord);
    for (slong e = 0; e < ne; e++) ho.put(e, 0, ODE->polys + e);
    for (slong t = 0; t <= steps; t++) hp.put(t, 0, path + t);
    acb_t one;
    acb_one(one);
    for (slong i = 0; i < ord; i++) hy.put(i, i, one);  // instance i starts from res = z^i
    rc = cb200_continuation_host(L, (int)ord, (int)ODE->degree, n, ord, steps + 1, ho.m.data(), ho.t.data(), hp.m.data(),
                                 hp.t.data(), hy.m.data(), hy.t.data(), 0);
    if (rc) die("find_monodromy_matrix (cb200_continuation_host)", rc);
    // row i of the reference's matrix before the transpose = the order Taylor coefficients of solution i
    for (slong i = 0; i < ord; i++)
        for (slong k = 0; k < ord; k++) hy.get(k, i, mono->rows[k] + i);
    _acb_vec_clear(path, steps + 1);
}


// ------------------------------------------------------------------ residual check (reference src/acb_ode.c:185-240)
// out = sum_i p_i * in^(i) with all len(in) + degree coefficients; the products run on the device
void acb_ode_apply(acb_poly_t out, acb_ode_t ODE, acb_poly_t in, slong prec) {
    int L = limbs_of(prec);
    slong ne = (ODE->order + 1) * (ODE->degree + 1), len = in->length, out_len = len + ODE->degree;
    HostArr ho(L, ne, 2), hi(L, len > 0 ? len : 1, 2), hr(L, out_len > 0 ? out_len : 1, 2);
    for (slong e = 0; e < ne; e++) ho.put(e, 0, ODE->polys + e);
    for (slong k = 0; k < len; k++) hi.put(k, 0, in->coeffs + k);  // marshalled first: out may alias in
    int rc = cb200_ode_apply_host(L, (int)ODE->order, (int)ODE->degree, 1, ho.m.data(), ho.t.data(), 1, hi.m.data(),
                                  hi.t.data(), len, hr.m.data(), hr.t.data(), out_len, 0, nullptr, 0);
    if (rc) die("acb_ode_apply (cb200_ode_apply_host)", rc);
    acb_poly_zero(out);
    acb_poly_fit_length(out, out_len);
    for (slong k = 0; k < out_len; k++) hr.get(k, 0, out->coeffs + k);
    out->length = out_len;
    poly_normalise(out);
}
int acb_ode_solves(acb_ode_t ODE, acb_poly_t res, slong deg, slong prec) {
    int L = limbs_of(prec);
    slong ne = (ODE->order + 1) * (ODE->degree + 1), len = res->length, out_len = len + ODE->degree;
    HostArr ho(L, ne, 2), hi(L, len > 0 ? len : 1, 2), hr(L, out_len > 0 ? out_len : 1, 2);
    for (slong e = 0; e < ne; e++) ho.put(e, 0, ODE->polys + e);
    for (slong k = 0; k < len; k++) hi.put(k, 0, res->coeffs + k);
    int32_t solved = 0;
    int rc = cb200_ode_apply_host(L, (int)ODE->order, (int)ODE->degree, 1, ho.m.data(), ho.t.data(), 1, hi.m.data(),
                                  hi.t.data(), len, hr.m.data(), hr.t.data(), out_len, deg, &solved, 0);
    if (rc) die("acb_ode_solves (cb200_ode_apply_host)", rc);
    return solved;
}

// ------------------------------------------------------------------ acb_ode_random (reference src/acb_ode.c:83-95)
// FLINT's generator is not available: a splitmix64 stream stands in (same shape of the distribution:
// degree 3..9, order 2..degree-1, coefficients with up to prec significant bits and exponents within +-16)
void flint_randinit(flint_rand_t state) { state->s = 0x9E3779B97F4A7C15ull; }
void flint_randclear(flint_rand_t) {}
static uint64_t next64(flint_rand_struct* st) {
    uint64_t z = (st->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
ulong n_randint(flint_rand_t state, ulong limit) { return limit ? (ulong)(next64(state) % limit) : 0; }
static void arb_random(arb_struct* x, flint_rand_struct* st, slong prec) {
    memset(x, 0, sizeof(*x));
    if (next64(st) % 8 == 0) {  // exact zeros are part of acb_randtest's distribution
        x->meta.flags = CB_FLAG_ZERO;
        return;
    }
    slong bits = 1 + (slong)(next64(st) % (uint64_t)(prec > 4096 ? 4096 : prec));
    slong limbs = (bits + 31) / 32;
    for (slong k = 0; k < limbs; k++) x->mant[CB_MAX_LIMBS - 1 - k] = (uint32_t)next64(st);
    x->mant[CB_MAX_LIMBS - 1] |= 0x80000000u;
    slong drop = 32 * limbs - bits;  // clear the bits below the chosen length
    if (drop) x->mant[CB_MAX_LIMBS - limbs] &= ~((1u << drop) - 1u);
    if (limbs == 1 && drop) x->mant[CB_MAX_LIMBS - 1] |= 0x80000000u;
    x->meta.exp = (int32_t)(next64(st) % 33) - 16;
    x->meta.flags = (next64(st) & 1) ? CB_FLAG_SIGN : 0;
    if (next64(st) % 2 == 0) {  // a radius of about one ulp at the working precision
        x->meta.rman = 1u << 29;
        x->meta.rexp = x->meta.exp - (int32_t)prec + 1;
    }
}
void acb_ode_random(acb_ode_t ode, flint_rand_t state, slong prec) {
    slong degree = 3 + (slong)n_randint(state, 7);
    slong order = 2 + (slong)n_randint(state, (ulong)(degree - 2));
    acb_ode_init_blank(ode, degree, order);
    for (slong i = 0; i <= order; i++)
        for (slong j = 0; j <= degree; j++) {
            arb_random(&acb_ode_coeff(ode, i, j)->real, state, prec);
            arb_random(&acb_ode_coeff(ode, i, j)->imag, state, prec);
        }
}

}  // extern "C"
