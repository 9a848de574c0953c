// cb_arith.cuh -- device ball arithmetic for sm_100a: one CUDA thread owns one complex ODE
// instance and runs plain multi-limb code on 32-bit limbs (IMAD.WIDE carry chains, see
// cb_mul.cuh); per-thread scratch lives in shared memory in a [word][thread] layout so that
// data-dependent limb offsets (alignment, renormalisation) stay bank-conflict free.
//
// This is the device replacement for the Arb calls Cascade's hot path makes:
//   acb_mul / acb_sub / acb_add / acb_div / acb_mul_fmpz   reference src/fuchs_solver.c:118-137
//   acb_add_si / acb_mul / acb_add / acb_div               reference src/frobenius_solver.c:58-113
//   acb_mul / acb_add (Horner inside acb_poly_evaluate)    reference src/acb_ode_solution.c:40-49
//
// Semantics (DESIGN.md "Ball arithmetic"): midpoints are the EXACT result rounded toward zero
// to 32 L bits from the leading bit, with an exact "inexact" flag; radii are 30-bit magnitudes
// rounded up, formulas as written at each function.  The same header compiles for the host
// (tests/host_sim) so the logic can be checked against the oracle without a GPU; the product
// library only ever contains the device build.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <type_traits>
#include "cb_mul.cuh"

#if defined(__CUDACC__)
#define CB_HD __host__ __device__ __forceinline__
#define CB_NOINLINE __host__ __device__ __noinline__
#else
#define CB_HD inline
#define CB_NOINLINE __attribute__((noinline))
#endif

#if defined(CB_COUNT_PATHS) && !defined(__CUDA_ARCH__)
extern "C" long long cb_path_counts[8];  // host-sim statistics: [0] fdot fast, [1] fdot general, [2] add fast, [3] add general,
                                         // [4] blocked fdot accepted, [5] refused, [6] two-row (integer-like operand) products
#define CB_COUNT(i) (cb_path_counts[i]++)
#else
#define CB_COUNT(i) ((void)0)
#endif

namespace cb {

constexpr uint32_t F_SIGN = 1u, F_ZERO = 2u, F_NAN = 4u;
// Internal hint, never part of a result: bits 8..15 of flags may hold a lower bound on the number
// of trailing zero LIMBS of the mantissa (set by the table kernel on the shared multipliers).
constexpr int F_TZ_SHIFT = 8;
CB_HD int tz_hint(uint32_t flags) { return (int)((flags >> F_TZ_SHIFT) & 0xFFu); }
// fill in the hint of an operand that has none: dense operands cost one load.  Short ones (integers: up to L - 1
// zero limbs) are scanned sixteen (from 2048 bits: sixty-four) limbs at a time with independent loads -- one limb after the other every load
// waited for the one before it, a global-memory round trip per zero limb (the 4096-bit hypergeometric sweep with
// rho = 0 spent most of its stall time there, profiles/ncu_frobenius1_128_r02_before_summary.txt).
template <int L, class O>
CB_NOINLINE void tz_scan(O& o) {  // (out of line: it is reached from every fused dot product and rarely runs)
    if (tz_hint(o.t.flags) != 0 || (o.t.flags & (F_ZERO | F_NAN))) return;
    if (o.r.ld(0) != 0u) return;
    int tz = L - 1;
    constexpr int B = L < 16 ? L : (L >= 64 ? 64 : 16);  // (4096 bits: an integer-like operand costs two round trips)
#pragma unroll 1
    for (int base = 0; base < L; base += B) {
        uint32_t w[B];
#pragma unroll
        for (int k = 0; k < B; k++) w[k] = (base + k < L) ? o.r.ld(base + k) : 1u;
        int first = B;
#pragma unroll
        for (int k = B - 1; k >= 0; k--) first = w[k] ? k : first;
        if (first < B) {
            tz = base + first < L - 1 ? base + first : L - 1;
            break;
        }
    }
    o.t.flags |= (uint32_t)tz << F_TZ_SHIFT;
}
// the four operands of a fused dot product at once: the first limbs of those without a hint are loaded together (one
// memory round trip instead of four); only operands whose first limb is zero go on to the scan
template <int L, class O1, class O2>
CB_HD void tz_scan4(O1& x1, O2& y1, O1& x2, O2& y2) {
    auto open = [](uint32_t f) { return tz_hint(f) == 0 && !(f & (F_ZERO | F_NAN)); };
    const bool ox1 = open(x1.t.flags), oy1 = open(y1.t.flags), ox2 = open(x2.t.flags), oy2 = open(y2.t.flags);
    const uint32_t a = ox1 ? x1.r.ld(0) : 1u, b = oy1 ? y1.r.ld(0) : 1u, c = ox2 ? x2.r.ld(0) : 1u, d = oy2 ? y2.r.ld(0) : 1u;
    if (a == 0u) tz_scan<L>(x1);
    if (b == 0u) tz_scan<L>(y1);
    if (c == 0u) tz_scan<L>(x2);
    if (d == 0u) tz_scan<L>(y2);
}
template <int L, class O1, class O2>
CB_HD void tz_scan2(O1& x, O2& y) {
    auto open = [](uint32_t f) { return tz_hint(f) == 0 && !(f & (F_ZERO | F_NAN)); };
    const bool ox = open(x.t.flags), oy = open(y.t.flags);
    const uint32_t a = ox ? x.r.ld(0) : 1u, b = oy ? y.r.ld(0) : 1u;
    if (a == 0u) tz_scan<L>(x);
    if (b == 0u) tz_scan<L>(y);
}
constexpr int32_t EXP_LIMIT = 1 << 28;

struct Meta {
    int32_t exp;
    uint32_t flags;
    uint32_t rman;
    int32_t rexp;
};

// ---------------------------------------------------------------- small helpers
CB_HD int clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}
CB_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}
// (hi:lo) >> r, low word; r in [0,31]
CB_HD uint32_t shr_pair(uint32_t lo, uint32_t hi, int r) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, (uint32_t)r);
#else
    return r ? (lo >> r) | (hi << (32 - r)) : lo;
#endif
}
CB_HD bool warp_all(bool p) {
#if defined(__CUDA_ARCH__)
    return __all_sync(__activemask(), p);
#else
    return p;
#endif
}
CB_HD bool warp_any(bool p) {
#if defined(__CUDA_ARCH__)
    return __any_sync(__activemask(), p);
#else
    return p;
#endif
}

// Per-thread scratch column.  Device: word k of this thread is at p[k * STRIDE] (shared memory,
// STRIDE = threads per block, so lane t always hits bank t).  Host: STRIDE = 1.
// On the device the column is addressed through ld.shared/st.shared on a 32-bit shared-window
// address, so the address space survives the (non-inlined) call boundaries below.
template <int STRIDE>
struct Scratch {
#if defined(__CUDA_ARCH__)
    uint32_t a;  // shared-space byte address of word 0 of this thread's column
    __device__ __forceinline__ uint32_t ld(int k) const {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a + (uint32_t)(k * (STRIDE * 4))));
        return v;
    }
    __device__ __forceinline__ void st(int k, uint32_t v) const {
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(a + (uint32_t)(k * (STRIDE * 4))), "r"(v));
    }
    __device__ __forceinline__ uint32_t a_or_p() const { return a; }
    // generic-address view of the column, for code that is shared with global-memory operands
    __device__ __forceinline__ uint32_t* generic() const {
        return (uint32_t*)__cvta_shared_to_generic((size_t)a);
    }
    static constexpr size_t GSTRIDE = STRIDE;
#else
    uint32_t* p;
    uint32_t ld(int k) const { return p[(ptrdiff_t)k * STRIDE]; }
    void st(int k, uint32_t v) const { p[(ptrdiff_t)k * STRIDE] = v; }
    const uint32_t* a_or_p() const { return p; }
    uint32_t* generic() const { return p; }
    static constexpr size_t GSTRIDE = STRIDE;
#endif
};

// A scratch column whose first KS words live in shared memory and whose tail lives in a global-memory column.
// The paths every operation takes (truncated blocked products, aligned sums of L-limb operands: at most L + 5
// words) see only the shared part -- they are handed lo(), a plain Scratch --; the rare exact paths (full
// 2L-limb products when the guard-bit proof fails, long division) spill their upper words to global memory.
// This is what lets the 2048/4096-bit kernels keep 8 / 6 warps per SM instead of 6 / 3.5.
template <int STRIDE, int KS>
struct HybridScratch {
#if defined(__CUDA_ARCH__)
    uint32_t a;    // shared-space byte address of word 0
    uint32_t* g;   // word k >= KS at g[(k - KS) * gs]
    uint32_t gs;
    __device__ __forceinline__ uint32_t ld(int k) const {
        if (k < KS) {
            uint32_t v;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a + (uint32_t)(k * (STRIDE * 4))));
            return v;
        }
        return g[(size_t)(k - KS) * gs];
    }
    __device__ __forceinline__ void st(int k, uint32_t v) const {
        if (k < KS)
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(a + (uint32_t)(k * (STRIDE * 4))), "r"(v));
        else
            g[(size_t)(k - KS) * gs] = v;
    }
    __device__ __forceinline__ Scratch<STRIDE> lo() const { return Scratch<STRIDE>{a}; }
#else
    uint32_t* p;
    uint32_t* g;
    uint32_t gs;
    uint32_t ld(int k) const { return k < KS ? p[(ptrdiff_t)k * STRIDE] : g[(size_t)(k - KS) * gs]; }
    void st(int k, uint32_t v) const {
        if (k < KS) p[(ptrdiff_t)k * STRIDE] = v;
        else g[(size_t)(k - KS) * gs] = v;
    }
    Scratch<STRIDE> lo() const { return Scratch<STRIDE>{p}; }
#endif
};
template <class S> struct ScrLo {
    using type = S;
    static constexpr bool HYBRID = false;
    static constexpr int LO_WORDS = 1 << 30;
    static CB_HD S get(S s) { return s; }
};
template <int ST, int KS> struct ScrLo<HybridScratch<ST, KS>> {
    using type = Scratch<ST>;
    static constexpr bool HYBRID = true;
    static constexpr int LO_WORDS = KS;
    static CB_HD Scratch<ST> get(HybridScratch<ST, KS> s) { return s.lo(); }
};
// the shared-memory part of a column (the column itself unless it is a HybridScratch)
template <class S>
CB_HD typename ScrLo<S>::type lo_view(S s) { return ScrLo<S>::get(s); }

// ---------------------------------------------------------------- magnitudes (radii)
// value = man * 2^(exp-30), man in [2^29, 2^30) or 0.  Every function mirrors oracle/cbo_arith.c.
struct Mag {
    uint32_t man;
    int32_t exp;
};
CB_HD Mag mag_zero() { return Mag{0u, 0}; }
CB_HD Mag mag_of(const Meta& t) { return Mag{t.rman, t.rexp}; }
CB_HD Mag mag_mul(Mag x, Mag y) {  // upper bound
    if (x.man == 0 || y.man == 0) return mag_zero();
    uint32_t m = (uint32_t)(((uint64_t)x.man * y.man) >> 30) + 1;
    int32_t e = x.exp + y.exp;
    if (!(m >> 29)) {
        m <<= 1;
        e -= 1;
    }
    return Mag{m, e};
}
CB_HD Mag mag_mul_lower(Mag x, Mag y) {
    if (x.man == 0 || y.man == 0) return mag_zero();
    uint32_t m = (uint32_t)(((uint64_t)x.man * y.man) >> 30);
    int32_t e = x.exp + y.exp;
    if (!(m >> 29)) {
        m <<= 1;
        e -= 1;
    }
    return Mag{m, e};
}
CB_HD Mag mag_add(Mag x, Mag y) {  // upper bound
    if (x.man == 0) return y;
    if (y.man == 0) return x;
    if (x.exp < y.exp) {
        Mag t = x;
        x = y;
        y = t;
    }
    int32_t sh = x.exp - y.exp;
    uint32_t m;
    if (sh == 0)
        m = x.man + y.man;
    else if (sh >= 30)
        m = x.man + 1;
    else
        m = x.man + (y.man >> sh) + 1;
    int32_t e = x.exp;
    if (m >> 30) {
        m = (m >> 1) + (m & 1);
        e += 1;
    }
    return Mag{m, e};
}
CB_HD Mag mag_sub_lower(Mag x, Mag y) {  // lower bound of x - y, 0 if not provably positive
    if (y.man == 0) return x;
    if (x.man == 0) return mag_zero();
    int32_t sh = x.exp - y.exp;
    if (sh < 0) return mag_zero();
    int32_t m;
    if (sh >= 30)
        m = (int32_t)x.man - 1;
    else
        m = (int32_t)x.man - (int32_t)(y.man >> sh) - 1;
    if (m <= 0) return mag_zero();
    int s = clz32((uint32_t)m) - 2;
    return Mag{(uint32_t)m << s, x.exp - s};
}
CB_HD Mag mag_div(Mag x, Mag y) {  // upper bound of x / y, y > 0
    if (x.man == 0) return mag_zero();
    uint32_t q = (uint32_t)((((uint64_t)x.man) << 30) / y.man) + 1;
    int32_t e = x.exp - y.exp;
    while (q >> 30) {
        q = (q >> 1) + (q & 1);
        e += 1;
    }
    return Mag{q, e};
}
CB_HD Mag mag_mid_upper(const Meta& t, uint32_t top_limb) {
    if (t.flags & F_ZERO) return mag_zero();
    uint32_t m = (top_limb >> 2) + 1;
    int32_t e = t.exp;
    if (m >> 30) {
        m = 1u << 29;
        e += 1;
    }
    return Mag{m, e};
}
CB_HD Mag mag_mid_lower(const Meta& t, uint32_t top_limb) {
    if (t.flags & F_ZERO) return mag_zero();
    return Mag{top_limb >> 2, t.exp};
}
template <int L>
CB_HD Mag mag_ulp(int32_t exp) {
    return Mag{1u << 29, exp - 32 * L + 1};
}
// |x|^ yr + |y|^ xr + xr yr added onto acc, in this order
CB_HD Mag mul_rad(Mag acc, const Meta& x, uint32_t xtop, const Meta& y, uint32_t ytop) {
    acc = mag_add(acc, mag_mul(mag_mid_upper(x, xtop), mag_of(y)));
    acc = mag_add(acc, mag_mul(mag_mid_upper(y, ytop), mag_of(x)));
    acc = mag_add(acc, mag_mul(mag_of(x), mag_of(y)));
    return acc;
}

CB_HD Meta meta_nan() { return Meta{0, F_NAN, 0xFFFFFFFFu, 0}; }
CB_HD Meta meta_zero() { return Meta{0, F_ZERO, 0u, 0}; }
CB_HD bool is_nan(const Meta& t) { return (t.flags & F_NAN) != 0; }
CB_HD bool is_zero_mid(const Meta& t) { return (t.flags & F_ZERO) != 0; }
CB_HD bool is_exact_zero(const Meta& t) { return (t.flags & F_ZERO) && t.rman == 0; }

// a real ball in memory: limb k at m[k*stride]; meta by value
struct RRef {
    const uint32_t* m;
    size_t stride;
    CB_HD uint32_t ld(int j) const { return m[(size_t)j * stride]; }
};
// limb k at m[k * STRIDE] with a compile-time stride (data shared by the batch: STRIDE = 2), so
// every access is base + immediate
template <int STRIDE>
struct RFix {
    const uint32_t* m;
    CB_HD uint32_t ld(int j) const { return m[(size_t)j * STRIDE]; }
    CB_HD RRef dyn() const { return RRef{m, (size_t)STRIDE}; }
};
// load L consecutive limbs of a source into registers; for a runtime stride the address is
// advanced by pointer increments (two ALU instructions per limb) instead of a 64-bit multiply per
// limb on the FMA pipe, which the limb products need
template <int L>
CB_HD void load_limbs(uint32_t (&a)[L], RRef x) {
    const uint32_t* p = x.m;
#pragma unroll
    for (int j = 0; j < L; j++) {
        a[j] = *p;
        p += x.stride;
#if defined(__CUDA_ARCH__)
        asm volatile("" : "+l"(p));  // keep the increment chain (no re-materialised j * stride)
#endif
    }
}
template <int L, int STRIDE>
CB_HD void load_limbs(uint32_t (&a)[L], RFix<STRIDE> x) {
#pragma unroll
    for (int j = 0; j < L; j++) a[j] = x.ld(j);
}

// B limbs starting at limb `base` (runtime): pointer walk for a runtime stride, immediates otherwise
template <int B, class XS>
CB_HD void load_block(uint32_t (&a)[B], const XS& x, int base) {
#pragma unroll
    for (int j = 0; j < B; j++) a[j] = x.ld(base + j);
}
template <int B>
CB_HD void load_block(uint32_t (&a)[B], const RRef& x, int base) {
    load_limbs<B>(a, RRef{x.m + (size_t)base * x.stride, x.stride});
}

struct WRef {
    uint32_t* m;
    size_t stride;
    CB_HD void st(int j, uint32_t v) const { m[(size_t)j * stride] = v; }
};

// put(j, word(j)) for j in [0, N): the words of a block are all LOADED before the first one is stored.  Written as
// `S.st(j, x.ld(j))` in one loop the store (a volatile shared-memory asm, or a store that may alias the source) pins
// every load behind the store before it, and each word pays a full global-memory round trip (measured: a third of
// the stall samples of the 4096-bit Frobenius kernel, profiles/ncu_frobenius1_128_r02_lines.txt).
template <int N, class GET, class PUT>
CB_HD void staged_copy(GET word, PUT put) {
    constexpr int B = N < 32 ? N : 32;
#pragma unroll 1
    for (int base = 0; base < N; base += B) {
        uint32_t w[B];
#pragma unroll
        for (int k = 0; k < B; k++) w[k] = (base + k < N) ? word(base + k) : 0u;
#pragma unroll
        for (int k = 0; k < B; k++)
            if (base + k < N) put(base + k, w[k]);
    }
}

// Rounded midpoint result description produced by extract()
struct Rounded {
    int32_t exp;
    bool zero;
    bool inexact;
    bool overflow;
};

// ---------------------------------------------------------------- extract
// R[0..hi] (scratch words, R[hi] != 0 unless the value is zero) is an unsigned integer V whose
// bit 0 has weight 2^frame.  Writes the top 32 L bits of V to dst (truncation), returns exponent
// and the exact inexact flag (found by looking at the discarded low words, normally just R[0]).
template <int L, class DST, class SCR>
CB_HD Rounded extract(DST dst, SCR R, int hi, bool sticky, int64_t frame) {
    Rounded out;
    uint32_t top = (hi >= 0) ? R.ld(hi) : 0u;
    if (hi < 0 || top == 0) {
#pragma unroll 4
        for (int j = 0; j < L; j++) dst.st(j, 0u);
        out.exp = 0;
        out.zero = true;
        out.inexact = sticky;  // cannot happen with sticky (see align_add), kept for safety
        out.overflow = false;
        return out;
    }
    int bl = 32 * hi + 32 - clz32(top);
    int off = bl - 32 * L;
    int q = off >> 5;  // arithmetic shift: floor
    int r = off & 31;
    uint32_t cur = ((unsigned)q <= (unsigned)hi) ? R.ld(q) : 0u;
    // inexact iff any discarded bit is set: low r bits of R[q], or any of R[0..q-1]
    bool inex = sticky || (r != 0 && (cur & ((1u << r) - 1u)) != 0u);
    for (int k = 0; !inex && k < q; k++) inex = R.ld(k) != 0u;
#pragma unroll 4
    for (int j = 0; j < L; j++) {
        int idx = q + j + 1;
        uint32_t nxt = ((unsigned)idx <= (unsigned)hi) ? R.ld(idx) : 0u;
        dst.st(j, shr_pair(cur, nxt, r));
        cur = nxt;
    }
    int64_t e = frame + bl;
    out.overflow = (e > EXP_LIMIT || e < -EXP_LIMIT);
    out.exp = (int32_t)e;
    out.zero = false;
    out.inexact = inex;
    return out;
}

// ---------------------------------------------------------------- align_add
// A and B hold N-limb magnitudes (words 0..N-1), top aligned: |a| = A * 2^(ea - 32 N); the top
// word of a nonzero operand is nonzero.  Computes |R| = | sa*a + sb*b | exactly as far as
// truncation to 32 L bits can tell, into the scratch column of the larger operand
// (words 0..N+1), and rounds it into dst.  Returns the rounded description and the sign via *neg.
// TO_SMALL: the rounded result is written to words 0..L-1 of the SMALLER operand's column (free
// once the aligned sum has been formed) and *where reports that column; dst is then unused.
template <int L, int N, bool TO_SMALL = false, class DST, class SCR>
CB_HD Rounded align_add(DST dst, SCR A, SCR B, int64_t ea, int64_t eb, bool nega, bool negb,
                        bool za, bool zb, bool* neg, SCR* where = nullptr) {
    // pick "big": the nonzero one, else the larger exponent (ties: A)
    bool swap = za ? true : (zb ? false : (eb > ea));
    SCR big = swap ? B : A, small = swap ? A : B;
    int64_t ebig = swap ? eb : ea, esm = swap ? ea : eb;
    bool nbig = swap ? negb : nega, nsm = swap ? nega : negb;
    bool zbig = swap ? zb : za, zsm = swap ? za : zb;
    if constexpr (TO_SMALL) *where = small;
    if (zbig) {  // both zero
        *neg = false;
        if constexpr (TO_SMALL)
            return extract<L>(small, big, -1, false, 0);
        else
            return extract<L>(dst, big, -1, false, 0);
    }
    int64_t d64 = zsm ? 0 : (ebig - esm);
    bool far = d64 > 32 * (N + 2);  // small lies entirely below the window
    int d = far ? 32 * (N + 2) : (int)d64;
    int q = d >> 5, r = d & 31;
    bool sub = (nbig != nsm) && !zsm;
    // T = small * 2^32 (N+1 words, T[0] = 0, T[i] = small[i-1]); S = T >> d.
    // T[i] for i in 1..N is small[i-1]; everything else is 0 (also when small is zero).
    const unsigned nlim = zsm ? 0u : (unsigned)N;
    auto Tw = [&](int idx) -> uint32_t { return ((unsigned)(idx - 1) < nlim) ? small.ld(idx - 1) : 0u; };
    // sticky: nonzero bits of T below bit d
    bool sticky = false;
    if (!zsm) {
        if (r != 0) sticky = (Tw(q) & ((1u << r) - 1u)) != 0u;
        int lim = q - 1 < N ? q - 1 : N;  // words T[1..lim] lie entirely below
        for (int j = 0; !sticky && j < lim; j++) sticky = small.ld(j) != 0u;
    }
    // aligned top words decide the operand order of a subtraction (ties resolved by a fix-up)
    bool swapped = false;
    if (sub) {
        uint32_t t2 = Tw(N + q + 1), t1 = Tw(N + q), t0 = Tw(N + q - 1);
        uint64_t s64 = ((uint64_t)shr_pair(t1, t2, r) << 32) | shr_pair(t0, t1, r);
        uint64_t w64 = ((uint64_t)big.ld(N - 1) << 32) | ((N >= 2) ? big.ld(N - 2) : 0u);
        swapped = s64 > w64;
    }
    const uint32_t mw = (sub && swapped) ? 0xFFFFFFFFu : 0u;
    const uint32_t ms = (sub && !swapped) ? 0xFFFFFFFFu : 0u;
    uint32_t carry = (sub && !sticky) ? 1u : 0u;
    uint32_t wprev = 0u;    // W[k] = big[k-1]
    uint32_t tcur = Tw(q);  // T[k+q]
    // in the loop T[k+q+1] = small[k+q] is in range iff k+q < nlim
    const int tl = (int)nlim - q;
#pragma unroll 4
    for (int k = 0; k < N; k++) {
        uint32_t wnext = big.ld(k);  // read before the slot is overwritten
        uint32_t tnxt = (k < tl) ? small.ld(k + q) : 0u;
        uint32_t sv = shr_pair(tcur, tnxt, r);
        uint64_t t = (uint64_t)(wprev ^ mw) + (uint64_t)(sv ^ ms) + carry;
        carry = (uint32_t)(t >> 32);
        big.st(k, (uint32_t)t);
        wprev = wnext;
        tcur = tnxt;
    }
    {   // k = N (top word of the window)
        uint32_t tnxt = (N < tl) ? small.ld(N + q) : 0u;
        uint32_t sv = shr_pair(tcur, tnxt, r);
        uint64_t t = (uint64_t)(wprev ^ mw) + (uint64_t)(sv ^ ms) + carry;
        carry = (uint32_t)(t >> 32);
        big.st(N, (uint32_t)t);
    }
    bool rneg = nbig;
    int hi;
    if (!sub) {
        // W[N] != 0, so without a carry R[N] != 0; with one the top word is the carry itself
        big.st(N + 1, carry);
        hi = carry ? N + 1 : N;
    } else {
        big.st(N + 1, 0u);
        if (swapped) rneg = nsm;
        if (carry == 0u) {
            // the order guess was wrong (top 64 aligned bits tied): negate in place
            rneg = !rneg;
            uint32_t c = 1u;
#pragma unroll 1
            for (int k = 0; k <= N; k++) {
                uint64_t t = (uint64_t)(~big.ld(k)) + c;
                c = (uint32_t)(t >> 32);
                big.st(k, (uint32_t)t);
            }
        }
        hi = N;
        while (hi >= 0 && big.ld(hi) == 0u) hi--;
    }
    *neg = rneg;
    if constexpr (TO_SMALL)
        return extract<L>(small, big, hi, sticky, ebig - 32 * (int64_t)N - 32);
    else
        return extract<L>(dst, big, hi, sticky, ebig - 32 * (int64_t)N - 32);
}

// ---------------------------------------------------------------- products into scratch
// P[0..2L) = x * y (exact), limbs read from memory
template <int L, class SCR, class XS, class YS>
CB_HD void product(SCR P, XS x, YS y) {
    if constexpr (L <= 32) {
        uint32_t a[L], b[L], r[2 * L];
#pragma unroll
        for (int j = 0; j < L; j++) a[j] = x.ld(j);
#pragma unroll
        for (int j = 0; j < L; j++) b[j] = y.ld(j);
        mul_full<L>(r, a, b);
#pragma unroll
        for (int k = 0; k < 2 * L; k++) P.st(k, r[k]);
    } else {
        constexpr int NB = L / 32;
#pragma unroll 1
        for (int k = 0; k < 2 * L; k++) P.st(k, 0u);
#pragma unroll 1
        for (int bi = 0; bi < NB; bi++) {
            uint32_t a[32];
#pragma unroll
            for (int j = 0; j < 32; j++) a[j] = x.ld(32 * bi + j);
#pragma unroll 1
            for (int bj = 0; bj < NB; bj++) {
                uint32_t b[32], r[64];
#pragma unroll
                for (int j = 0; j < 32; j++) b[j] = y.ld(32 * bj + j);
                mul_full<32>(r, a, b);
                int off = 32 * (bi + bj);
                uint32_t o[64];
#pragma unroll
                for (int k = 0; k < 64; k++) o[k] = P.ld(off + k);
                uint32_t c = add_chain<64>(o, r);
#pragma unroll
                for (int k = 0; k < 64; k++) P.st(off + k, o[k]);
#pragma unroll 1
                for (int k = off + 64; c != 0u && k < 2 * L; k++) {
                    uint64_t t = (uint64_t)P.ld(k) + c;
                    P.st(k, (uint32_t)t);
                    c = (uint32_t)(t >> 32);
                }
            }
        }
    }
}

// one real operand: where its limbs are and its meta record
template <class SRC>
struct OpdT {
    SRC r;  // anything with ld(j): RRef (global memory) or a Scratch column
    Meta t;
};
using Opd = OpdT<RRef>;
template <int L, class SRC>
CB_HD uint32_t top_limb(const OpdT<SRC>& o) {
    return o.r.ld(L - 1);
}

template <int L>
CB_HD Meta finish(const Rounded& rr, bool neg, Mag rad) {
    if (rr.overflow) return meta_nan();
    Meta t;
    if (rr.zero) {
        t.exp = 0;
        t.flags = F_ZERO;
    } else {
        t.exp = rr.exp;
        t.flags = neg ? F_SIGN : 0u;
        if (rr.inexact) rad = mag_add(rad, mag_ulp<L>(rr.exp));
    }
    t.rman = rad.man;
    t.rexp = rad.man ? rad.exp : 0;
    return t;
}
template <int L, class DST>
CB_HD void store_nan(DST dst) {
#pragma unroll 8
    for (int j = 0; j < L; j++) dst.st(j, 0u);
}
// finish + canonical form of an indeterminate result (exponent overflow): all-zero mantissa
template <int L, class DST>
CB_HD Meta finish_to(DST dst, const Rounded& rr, bool neg, Mag rad) {
    Meta t = finish<L>(rr, neg, rad);
    if (is_nan(t)) store_nan<L>(dst);
    return t;
}

// destinations that live in registers (cb_intop.cuh RegDst) take their stores with a STATIC loop index
template <class T, class = void> struct DstTraits { static constexpr bool REG_DST = false; };
template <class T> struct DstTraits<T, std::enable_if_t<T::REG_DST>> { static constexpr bool REG_DST = true; };
template <class DST, int QMIN, int COUNT, int I>
struct StoreStatic {
    template <class RX>
    static CB_HD void run(const DST& dst, int c, int r, RX Rx) {
        if constexpr (I < COUNT) {
            dst.template st_static<I>(c, shr_pair(Rx(QMIN + I), Rx(QMIN + I + 1), r));
            StoreStatic<DST, QMIN, COUNT, I + 1>::run(dst, c, r, Rx);
        }
    }
};

// ---------------------------------------------------------------- in-register fast path
// The common case of an aligned add: both operands within 31 bits of each other (or one of them
// zero) and at most about a limb of cancellation.  Everything is statically indexed, so the sum
// lives in registers and only the final stores use a data-dependent offset.
//   V_i = X_i * 2^(E - s_i - 32 N),  X_i given word by word by x_i(k), k in [0, N) (0 if zero).
// Returns false without touching dst when the case is not the easy one (operand order guess of a
// subtraction wrong, or more than a limb cancelled): the caller then takes the general path.
// KLOW > 0 (truncated products): words k < KLOW of both operands are not available -- the operands
// are the high products of mul_high with K = KLOW dropped columns, i.e. each true operand exceeds the
// given one by less than (K+1) 2^(32(K+1)) -- and the result is only accepted when the guard bits
// just below the rounding position prove that the missing low part cannot change the truncated
// mantissa (then the result is certainly inexact); otherwise false is returned.
// trusted: the caller knows that nothing nonzero was dropped (both products have at least KLOW
// trailing zero limbs in total), so the available words are exact and no guard check is needed.
template <int L, int N, int KLOW = 0, class DST, class X1, class X2>
CB_HD bool fast_add_round(DST dst, X1 x1, X2 x2, int s1, int s2, bool sub, bool n1, bool n2, int64_t E,
                          Rounded* rr, bool* neg, bool trusted = false) {
    constexpr bool TRUNC = KLOW > 0;
    uint32_t Rs[N + 2 - KLOW];
#define R_(i) Rs[(i) - KLOW]
    bool swapped = false;
    if (sub) {  // compare the top 64 aligned bits: W_i[N], W_i[N-1] with T_i[k] = X_i[k-1]
        uint32_t a2 = x1(N - 1), a1 = x1(N - 2), a0 = (N >= 3) ? x1(N - 3) : 0u;
        uint32_t b2 = x2(N - 1), b1 = x2(N - 2), b0 = (N >= 3) ? x2(N - 3) : 0u;
        uint64_t w1 = ((uint64_t)shr_pair(a2, 0u, s1) << 32) | shr_pair(a1, a2, s1);
        uint64_t w2 = ((uint64_t)shr_pair(b2, 0u, s2) << 32) | shr_pair(b1, b2, s2);
        (void)a0;
        (void)b0;
        swapped = w2 > w1;
    }
    const uint32_t m1 = (sub && swapped) ? 0xFFFFFFFFu : 0u;
    const uint32_t m2 = (sub && !swapped) ? 0xFFFFFFFFu : 0u;
    uint32_t carry = sub ? 1u : 0u;
    uint32_t t1c = 0u, t2c = 0u;  // T_i[k]; T_i[KLOW] = 0 (guard limb / first dropped word)
#pragma unroll
    for (int k = KLOW; k <= N; k++) {
        uint32_t t1n = (k < N) ? x1(k) : 0u, t2n = (k < N) ? x2(k) : 0u;
        uint32_t w1 = shr_pair(t1c, t1n, s1), w2 = shr_pair(t2c, t2n, s2);
        uint64_t t = (uint64_t)(w1 ^ m1) + (uint64_t)(w2 ^ m2) + carry;
        R_(k) = (uint32_t)t;
        carry = (uint32_t)(t >> 32);
        t1c = t1n;
        t2c = t2n;
    }
    if (sub && carry == 0u) return false;  // |W2| > |W1| after all (tie in the top 64 bits)
    R_(N + 1) = sub ? 0u : carry;
    int hi;
    uint32_t top;
    if (R_(N + 1)) {
        hi = N + 1;
        top = R_(N + 1);
    } else if (R_(N)) {
        hi = N;
        top = R_(N);
    } else if (R_(N - 1)) {
        hi = N - 1;
        top = R_(N - 1);
    } else {
        return false;  // a limb or more cancelled (or the result is zero)
    }
    const int bl = 32 * hi + 32 - clz32(top);
    const int off = bl - 32 * L;
    const int q = off >> 5, r = off & 31;
    constexpr int QMIN = N - 1 - L;  // q is in [QMIN, QMIN + 3]
    auto Rx = [&](int idx) -> uint32_t { return (idx >= KLOW && idx <= N + 1) ? R_(idx) : 0u; };
    uint32_t low = 0u;
    if constexpr (TRUNC) {
        // Error of the sum: each truncated product is short by < (K+1) 2^(32(K+2)) in this frame, and
        // an operand moved down by whole words additionally loses < 2^(32 K): in total
        // < 4 (K+1) 2^(32(K+2)) <= 2^B.  The 24 bits just below the rounding position must be
        // neither all zero nor all one; then the missing part cannot carry into the kept bits and
        // the result is certainly inexact.
        constexpr int LOGK = (KLOW + 1 <= 2) ? 1 : (KLOW + 1 <= 4) ? 2 : (KLOW + 1 <= 8) ? 3 : (KLOW + 1 <= 16) ? 4
                             : (KLOW + 1 <= 32) ? 5 : (KLOW + 1 <= 64) ? 6 : 7;
        constexpr int B = 32 * (KLOW + 2) + 2 + LOGK;
        static_assert(32 * (N - 1) + 1 - 32 * L - 24 >= B, "not enough guard bits for this truncation");
        static_assert(QMIN - 1 >= KLOW, "guard word must be available");
        uint32_t lo_w = 0u, hi_w = 0u;  // words q-1 and q
#pragma unroll
        for (int i = QMIN - 1; i < QMIN + 4; i++) {
            uint32_t w = Rx(i);
            lo_w = (i == q - 1) ? w : lo_w;
            hi_w = (i == q) ? w : hi_w;
        }
        uint32_t g = shr_pair(lo_w, hi_w, r) >> 8;  // bits [off-24, off)
        if (!trusted) {
            if (g == 0u || g == 0xFFFFFFu) return false;
            low = 1u;  // proven inexact
        } else {
            // exact words: discarded = words KLOW..q-1 and the low r bits of word q
#pragma unroll
            for (int i = KLOW; i < QMIN; i++) low |= R_(i);
#pragma unroll
            for (int i = QMIN; i < QMIN + 4; i++) {
                uint32_t w = Rx(i);
                low |= (i < q) ? w : 0u;
                low |= (i == q) ? (w & ((1u << r) - 1u)) : 0u;
            }
        }
    } else {
        // discarded bits: words below q, and the low r bits of word q
#pragma unroll
        for (int i = 0; i < QMIN; i++) low |= R_(i);
#pragma unroll
        for (int i = QMIN; i < QMIN + 4; i++) {
            uint32_t w = Rx(i);
            low |= (i < q) ? w : 0u;
            low |= (i == q) ? (w & ((1u << r) - 1u)) : 0u;
        }
    }
    // dst[i - q] = bits [32 i + r, 32 i + r + 32) of R, for the L words with 0 <= i - q < L
    if constexpr (DstTraits<DST>::REG_DST) {
        StoreStatic<DST, QMIN, L + 3, 0>::run(dst, q - QMIN, r, Rx);  // register destination: static indices only
    } else {
        if constexpr (std::is_same<DST, WRef>::value) {
            // runtime stride: walk a pointer (word i goes to limb i - q; the first q - QMIN steps are before limb 0
            // and store nothing) instead of forming j * stride for every store
            uint32_t* pw = dst.m - (ptrdiff_t)(q - QMIN) * (ptrdiff_t)dst.stride;
#pragma unroll
            for (int i = QMIN; i < QMIN + L + 3; i++) {
                uint32_t v = shr_pair(Rx(i), Rx(i + 1), r);
                if ((unsigned)(i - q) < (unsigned)L) *pw = v;
                pw += dst.stride;
#if defined(__CUDA_ARCH__)
                asm volatile("" : "+l"(pw));
#endif
            }
        } else {
#pragma unroll
            for (int i = QMIN; i < QMIN + L + 3; i++) {
                uint32_t v = shr_pair(Rx(i), Rx(i + 1), r);
                int j = i - q;
                if ((unsigned)j < (unsigned)L) dst.st(j, v);
            }
        }
    }
    int64_t e = E - 32 * (int64_t)N - 32 + bl;
    rr->overflow = (e > EXP_LIMIT || e < -EXP_LIMIT);
    rr->exp = (int32_t)e;
    rr->zero = false;
    rr->inexact = low != 0u;
    *neg = (sub && swapped) ? n2 : n1;
    return true;
#undef R_
}

// ---------------------------------------------------------------- arb level
// dst = x1*y1 + (neg2 ? -1 : 1) * x2*y2, one rounding.  S0/S1: two scratch columns of 2L+2 words.
template <int L, class SCR, class XS = RRef, class YS = RRef, class DST = WRef>
CB_NOINLINE Meta arb_fdot2(DST dst, OpdT<XS> x1, OpdT<YS> y1, OpdT<XS> x2, OpdT<YS> y2, bool neg2, SCR S0,
                           SCR S1) {
    bool nan = is_nan(x1.t) || is_nan(y1.t) || is_nan(x2.t) || is_nan(y2.t);
    bool z1 = is_zero_mid(x1.t) || is_zero_mid(y1.t) || nan;
    bool z2 = is_zero_mid(x2.t) || is_zero_mid(y2.t) || nan;
    Mag rad = mul_rad(mag_zero(), x1.t, top_limb<L>(x1), y1.t, top_limb<L>(y1));
    rad = mul_rad(rad, x2.t, top_limb<L>(x2), y2.t, top_limb<L>(y2));
    bool n1 = ((x1.t.flags ^ y1.t.flags) & F_SIGN) != 0;
    bool n2 = (((x2.t.flags ^ y2.t.flags) & F_SIGN) != 0) != neg2;
    const int64_t e1 = (int64_t)x1.t.exp + y1.t.exp, e2 = (int64_t)x2.t.exp + y2.t.exp;
    bool skip1 = warp_all(z1), skip2 = warp_all(z2);
    bool neg;
    Rounded rr;
    if constexpr (L <= 32) {
        // products in registers; the first one is parked in S0 while the second is formed
        uint32_t r[2 * L];
#pragma unroll 1
        for (int term = 0; term < 2; term++) {  // a loop, so that the product code exists once
            if (term == 1) {
#pragma unroll
                for (int k = 0; k < 2 * L; k++) S0.st(k, r[k]);
            }
            if (term ? skip2 : skip1) {
#pragma unroll
                for (int k = 0; k < 2 * L; k++) r[k] = 0u;
                continue;
            }
            uint32_t a[L], b[L];
            XS xs = term ? x2.r : x1.r;
            YS ys = term ? y2.r : y1.r;
#pragma unroll
            for (int j = 0; j < L; j++) a[j] = xs.ld(j);
#pragma unroll
            for (int j = 0; j < L; j++) b[j] = ys.ld(j);
            mul_full<L>(r, a, b);
        }
        // easy: one term zero, or exponents within 31 bits
        int64_t d = e1 - e2;
        bool easy = !(z1 && z2) && (z1 || z2 || (d <= 31 && d >= -31));
        bool done = false;  // per lane: lanes that are not easy take the general path below
        if (easy) {
            int64_t E = z1 ? e2 : (z2 ? e1 : (e1 > e2 ? e1 : e2));
            int s1 = z1 ? 0 : (int)(E - e1), s2 = z2 ? 0 : (int)(E - e2);
            bool sub = (n1 != n2) && !z1 && !z2;
            // a zero term contributes zero words (its product registers / column may hold the
            // product of a zero-flagged operand's stale limbs)
            auto w1 = [&](int k) -> uint32_t { return z1 ? 0u : S0.ld(k); };
            auto w2 = [&](int k) -> uint32_t { return z2 ? 0u : r[k]; };
            done = fast_add_round<L, 2 * L>(dst, w1, w2, s1, s2, sub, z1 ? n2 : n1, z1 ? n1 : n2, E, &rr, &neg);
        }
        CB_COUNT(done ? 0 : 1);
        if (!done) {
#pragma unroll
            for (int k = 0; k < 2 * L; k++) S1.st(k, r[k]);
            rr = align_add<L, 2 * L>(dst, S0, S1, e1, e2, n1, n2, z1, z2, &neg);
        }
    } else {
#pragma unroll 1
        for (int term = 0; term < 2; term++) {
            if (term ? skip2 : skip1) continue;
            product<L>(term ? S1 : S0, term ? x2.r : x1.r, term ? y2.r : y1.r);
        }
        bool done = false;
        if constexpr (L <= 64) {
            // same in-register tail as above, reading both products from their columns
            int64_t d = e1 - e2;
            bool easy = !(z1 && z2) && (z1 || z2 || (d <= 31 && d >= -31));
            if (easy) {
                int64_t E = z1 ? e2 : (z2 ? e1 : (e1 > e2 ? e1 : e2));
                int s1 = z1 ? 0 : (int)(E - e1), s2 = z2 ? 0 : (int)(E - e2);
                bool sub = (n1 != n2) && !z1 && !z2;
                auto w1 = [&](int k) -> uint32_t { return z1 ? 0u : S0.ld(k); };
                auto w2 = [&](int k) -> uint32_t { return z2 ? 0u : S1.ld(k); };
                done = fast_add_round<L, 2 * L>(dst, w1, w2, s1, s2, sub, z1 ? n2 : n1, z1 ? n1 : n2, E, &rr, &neg);
            }
        }
        if (!done) rr = align_add<L, 2 * L>(dst, S0, S1, e1, e2, n1, n2, z1, z2, &neg);
    }
    if (nan) return meta_nan();
    return finish_to<L>(dst, rr, neg, rad);
}

// ---- the recurrence kernel's variant: truncated products first, exact general path as fallback
// A column in global memory ([word][thread] layout like the shared ones): only the rarely taken
// general path of the recurrence kernel uses it.
struct GScratch {
    uint32_t* p;
    size_t stride;
    CB_HD uint32_t ld(int k) const { return p[(size_t)k * stride]; }
    CB_HD void st(int k, uint32_t v) const { p[(size_t)k * stride] = v; }
};

// P[0..2L) = x * y with BLK-limb register blocks accumulated in the column (few registers, slow)
template <int L, int BLK, class SCR, class XS, class YS>
CB_HD void product_blocked(SCR P, XS x, YS y) {
    constexpr int NB = L / BLK;
#pragma unroll 1
    for (int k = 0; k < 2 * L; k++) P.st(k, 0u);
#pragma unroll 1
    for (int bi = 0; bi < NB; bi++) {
        uint32_t a[BLK];
#pragma unroll
        for (int j = 0; j < BLK; j++) a[j] = x.ld(BLK * bi + j);
#pragma unroll 1
        for (int bj = 0; bj < NB; bj++) {
            uint32_t b[BLK], r[2 * BLK];
#pragma unroll
            for (int j = 0; j < BLK; j++) b[j] = y.ld(BLK * bj + j);
            mul_full<BLK>(r, a, b);
            int off = BLK * (bi + bj);
            uint32_t c = 0u;
#pragma unroll
            for (int k = 0; k < 2 * BLK; k++) {
                uint64_t t = (uint64_t)P.ld(off + k) + r[k] + c;
                P.st(off + k, (uint32_t)t);
                c = (uint32_t)(t >> 32);
            }
#pragma unroll 1
            for (int k = off + 2 * BLK; c != 0u && k < 2 * L; k++) {
                uint64_t t = (uint64_t)P.ld(k) + c;
                P.st(k, (uint32_t)t);
                c = (uint32_t)(t >> 32);
            }
        }
    }
}

// exact general fdot2 on two global-memory columns (any exponent gap, any cancellation)
template <int L, class XS = RRef, class YS = RRef, class DST = WRef>
CB_NOINLINE Meta arb_fdot2_general(DST dst, OpdT<XS> x1, OpdT<YS> y1, OpdT<XS> x2, OpdT<YS> y2, bool neg2,
                                   GScratch G0, GScratch G1) {
    bool nan = is_nan(x1.t) || is_nan(y1.t) || is_nan(x2.t) || is_nan(y2.t);
    bool z1 = is_zero_mid(x1.t) || is_zero_mid(y1.t) || nan;
    bool z2 = is_zero_mid(x2.t) || is_zero_mid(y2.t) || nan;
    Mag rad = mul_rad(mag_zero(), x1.t, top_limb<L>(x1), y1.t, top_limb<L>(y1));
    rad = mul_rad(rad, x2.t, top_limb<L>(x2), y2.t, top_limb<L>(y2));
    bool n1 = ((x1.t.flags ^ y1.t.flags) & F_SIGN) != 0;
    bool n2 = (((x2.t.flags ^ y2.t.flags) & F_SIGN) != 0) != neg2;
    const int64_t e1 = (int64_t)x1.t.exp + y1.t.exp, e2 = (int64_t)x2.t.exp + y2.t.exp;
    constexpr int BLK = (L % 8 == 0) ? 8 : 4;
#pragma unroll 1
    for (int term = 0; term < 2; term++) {
        if (term ? z2 : z1) continue;
        product_blocked<L, BLK>(term ? G1 : G0, term ? x2.r : x1.r, term ? y2.r : y1.r);
    }
    bool neg;
    Rounded rr = align_add<L, 2 * L>(dst, G0, G1, e1, e2, n1, n2, z1, z2, &neg);
    if (nan) return meta_nan();
    return finish_to<L>(dst, rr, neg, rad);
}

// dst = x1*y1 +- x2*y2 from TRUNCATED products (columns below K = L-4 dropped: 40% fewer limb
// products); *ok = false (dst untouched... or to be overwritten) when the result cannot be proven
// equal to the exact one -- the caller then runs arb_fdot2_general.  S0: a shared column of
// 2L-K+1 words.
// SCAN_TZ: look for short operands in the loaded limbs (general kernels); the recurrence kernel
// gets the hints of its shared multipliers from the table kernel and leaves this off.
template <int L, class SCR, class XS = RRef, class YS = RRef, class DST = WRef, bool SCAN_TZ = false>
CB_NOINLINE Meta arb_fdot2_trunc(DST dst, OpdT<XS> x1, OpdT<YS> y1, OpdT<XS> x2, OpdT<YS> y2, bool neg2, SCR S0,
                                 bool* ok) {
    static_assert(L >= 16 && L <= 32, "truncated products are used for 512..1024-bit mantissas");
    constexpr int K = L - 4, PW = 2 * L - K;
    bool nan = is_nan(x1.t) || is_nan(y1.t) || is_nan(x2.t) || is_nan(y2.t);
    bool z1 = is_zero_mid(x1.t) || is_zero_mid(y1.t) || nan;
    bool z2 = is_zero_mid(x2.t) || is_zero_mid(y2.t) || nan;
    Mag rad = mul_rad(mag_zero(), x1.t, top_limb<L>(x1), y1.t, top_limb<L>(y1));
    rad = mul_rad(rad, x2.t, top_limb<L>(x2), y2.t, top_limb<L>(y2));
    bool n1 = ((x1.t.flags ^ y1.t.flags) & F_SIGN) != 0;
    bool n2 = (((x2.t.flags ^ y2.t.flags) & F_SIGN) != 0) != neg2;
    const int64_t e1 = (int64_t)x1.t.exp + y1.t.exp, e2 = (int64_t)x2.t.exp + y2.t.exp;
    // Per lane, term A is the one that may need a limb shift (the zero one, else the smaller
    // exponent); its product is parked in the shared column, where a data-dependent word offset
    // is cheap.  Term B (the larger) stays in registers and needs no shift at all.
    const bool a_is_1 = z1 ? true : (z2 ? false : (e1 <= e2));
    const bool zA = a_is_1 ? z1 : z2, zB = a_is_1 ? z2 : z1;
    const bool nA = a_is_1 ? n1 : n2, nB = a_is_1 ? n2 : n1;
    const int64_t eA = a_is_1 ? e1 : e2, eB = a_is_1 ? e2 : e1;
    const XS xA = a_is_1 ? x1.r : x2.r, xB = a_is_1 ? x2.r : x1.r;
    const YS yA = a_is_1 ? y1.r : y2.r, yB = a_is_1 ? y2.r : y1.r;
    int tzA = a_is_1 ? tz_hint(x1.t.flags) + tz_hint(y1.t.flags) : tz_hint(x2.t.flags) + tz_hint(y2.t.flags);
    int tzB = a_is_1 ? tz_hint(x2.t.flags) + tz_hint(y2.t.flags) : tz_hint(x1.t.flags) + tz_hint(y1.t.flags);
    bool skipA = warp_all(zA), skipB = warp_all(zB);
    uint32_t r[PW];
#pragma unroll 1
    for (int term = 0; term < 2; term++) {  // a loop, so that the product code exists once
        if (term ? skipB : skipA) {
#pragma unroll
            for (int k = 0; k < PW; k++) r[k] = 0u;
        } else {
            uint32_t a[L], b[L];
            load_limbs<L>(a, term ? xB : xA);
            load_limbs<L>(b, term ? yB : yA);
            // short (integer-like) operands: count their trailing zero limbs while they are in
            // registers (exact products carry no information in their guard bits)
            if (SCAN_TZ && (a[0] == 0u || b[0] == 0u)) {
                int ta = L - 1, tb = L - 1;
#pragma unroll
                for (int j = L - 2; j >= 0; j--) {
                    ta = a[j] ? j : ta;
                    tb = b[j] ? j : tb;
                }
                int t = ta + tb;
                if (term) tzB = tzB > t ? tzB : t; else tzA = tzA > t ? tzA : t;
            }
            mul_high<L, K>(r, a, b);
        }
        if (term == 0) {  // park product A; nothing of r is live across the back edge
#pragma unroll
            for (int k = 0; k < PW; k++) S0.st(k, r[k]);
        }
    }
    bool done = false, neg = false;
    Rounded rr;
    rr.exp = 0;
    rr.zero = true;
    rr.inexact = false;
    rr.overflow = false;
    if (zA && zB) {  // exact zero midpoint (also the indeterminate case): nothing to round
#pragma unroll 4
        for (int j = 0; j < L; j++) dst.st(j, 0u);
        done = true;
    } else {
        // zB implies zA by construction of the order, so B is nonzero here
        int64_t d = zA ? 0 : (eB - eA);              // >= 0
        int m = d > 32 * 70 ? 70 : (int)(d >> 5);   // limb shift of A (70: entirely below the window)
        int sA = (int)(d & 31);
        bool sub = (nA != nB) && !zA;
        // word k (k in [K, 2L)) of the product B, and of the product A moved down by m words
        // zero-flagged operands are masked explicitly, so a non-canonical zero (flag set, stale
        // limbs) behaves like the oracle's
        auto wB = [&](int k) -> uint32_t { return r[k - K]; };
        auto wA = [&](int k) -> uint32_t {
            int i = k + m - K;
            return (zA || (unsigned)i >= (unsigned)PW) ? 0u : S0.ld(i);
        };
        // nothing is lost to truncation when every dropped column is zero: both products have
        // >= K trailing zero limbs in total and A is not moved below the window
        bool trusted = (zA || (tzA >= K && m == 0)) && tzB >= K;
        done = fast_add_round<L, 2 * L, K>(dst, wB, wA, 0, sA, sub, nB, nA, eB, &rr, &neg, trusted);
    }
    *ok = done;
    if (nan) return meta_nan();
    return finish_to<L>(dst, rr, neg, rad);
}

// ---- above 1024 bits: truncated product from 32-limb blocks, both products in shared columns
// Ph[k] = word K + k of the truncated product, K = L - 4, k in [0, L + 4): every limb product whose
// column is below K is dropped, exactly the set mul_high drops for an unblocked product, so the
// same error bound (K+1) 2^(32(K+1)) holds.  With NB = L/32 blocks per operand and s = bi + bj:
//   s >= NB      full 32x32 block                     -> product words 32 s .. 32 s + 63
//   s == NB - 1  high block (mul_high<32, 28>)        -> product words K .. K + 35
//   s == NB - 2  only the 6 limb products with i + j >= 60 inside the block -> words K .. K + 3
//   s <  NB - 2  dropped entirely
// (2048 bits: 2 266 limb multiplies instead of 4 096; 4096 bits: 8 634 instead of 16 384.)
template <class SCR>
CB_HD void ripple(SCR Ph, int k, int top, uint32_t c) {
#pragma unroll 1
    for (; c != 0u && k < top; k++) {
        uint64_t s = (uint64_t)Ph.ld(k) + c;
        Ph.st(k, (uint32_t)s);
        c = (uint32_t)(s >> 32);
    }
}
template <int L, class SCR, class XS, class YS>
CB_HD void product_trunc_blocked(SCR Ph, XS x, YS y) {
    constexpr int NB = L / 32, K = L - 4, PW = L + 4;
    static_assert(L % 32 == 0 && NB >= 2, "blocked truncated product");
#pragma unroll 8
    for (int k = 0; k < PW; k++) Ph.st(k, 0u);
    // full blocks
#pragma unroll 1
    for (int bi = 1; bi < NB; bi++) {
        uint32_t a[32];
        load_block<32>(a, x, 32 * bi);
#pragma unroll 1
        for (int bj = NB - bi; bj < NB; bj++) {
            uint32_t b[32], r[64], o[64];
            load_block<32>(b, y, 32 * bj);
            mul_full<32>(r, a, b);
            int off = 32 * (bi + bj) - K;
#pragma unroll
            for (int k = 0; k < 64; k++) o[k] = Ph.ld(off + k);
            uint32_t c = add_chain<64>(o, r);
#pragma unroll
            for (int k = 0; k < 64; k++) Ph.st(off + k, o[k]);
            ripple(Ph, off + 64, PW, c);
        }
    }
    // high blocks: bi + bj == NB - 1
#pragma unroll 1
    for (int bi = 0; bi < NB; bi++) {
        int bj = NB - 1 - bi;
        uint32_t a[32], b[32], r[36], o[36];
        load_block<32>(a, x, 32 * bi);
        load_block<32>(b, y, 32 * bj);
        mul_high<32, 28>(r, a, b);
#pragma unroll
        for (int k = 0; k < 36; k++) o[k] = Ph.ld(k);
        uint32_t c = add_chain<36>(o, r);
#pragma unroll
        for (int k = 0; k < 36; k++) Ph.st(k, o[k]);
        ripple(Ph, 36, PW, c);
    }
    // corners: bi + bj == NB - 2, limb products with i + j in {60, 61, 62}
#pragma unroll 1
    for (int bi = 0; bi <= NB - 2; bi++) {
        int bj = NB - 2 - bi;
        uint32_t x29 = x.ld(32 * bi + 29), x30 = x.ld(32 * bi + 30), x31 = x.ld(32 * bi + 31);
        uint32_t y29 = y.ld(32 * bj + 29), y30 = y.ld(32 * bj + 30), y31 = y.ld(32 * bj + 31);
        uint64_t p, c0 = 0, c1 = 0, c2 = 0, c3 = 0;  // column sums (each < 2^35)
        p = (uint64_t)x29 * y31; c0 += (uint32_t)p; c1 += p >> 32;
        p = (uint64_t)x30 * y30; c0 += (uint32_t)p; c1 += p >> 32;
        p = (uint64_t)x31 * y29; c0 += (uint32_t)p; c1 += p >> 32;
        p = (uint64_t)x30 * y31; c1 += (uint32_t)p; c2 += p >> 32;
        p = (uint64_t)x31 * y30; c1 += (uint32_t)p; c2 += p >> 32;
        p = (uint64_t)x31 * y31; c2 += (uint32_t)p; c3 += p >> 32;
        uint64_t cols[4] = {c0, c1, c2, c3};
        uint64_t carry = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint64_t sum = (uint64_t)Ph.ld(k) + (uint32_t)cols[k] + (uint32_t)carry;
            carry = (carry >> 32) + (cols[k] >> 32) + (sum >> 32);
            Ph.st(k, (uint32_t)sum);
        }
#pragma unroll 1
        for (int k = 4; carry != 0u && k < PW; k++) {
            uint64_t sum = (uint64_t)Ph.ld(k) + (uint32_t)carry;
            carry = (carry >> 32) + (sum >> 32);
            Ph.st(k, (uint32_t)sum);
        }
    }
}

// The same product column for an INTEGER-LIKE y (every limb below L - 2 zero: rho + k with an integer rho, a small
// rational coefficient, g_0 = 1 ...): two rows instead of L.  All of its limb products lie in columns >= L - 2 > K, so
// nothing is dropped and the words are those of the exact product -- the same words product_trunc_blocked would give
// (its dropped columns hold zeros only), at 2 L limb multiplies instead of about L^2 / 2.
template <int L, class SCR, class XS, class YS>
CB_HD void product_short_y(SCR Ph, XS x, YS y) {
    CB_COUNT(6);
    const uint32_t y0 = y.ld(L - 2), y1 = y.ld(L - 1);
    Ph.st(0, 0u);
    Ph.st(1, 0u);
    // product word (L - 2) + j = Ph[j + 2] = lo(x_j y0) + hi(x_{j-1} y0) + lo(x_{j-1} y1) + hi(x_{j-2} y1) + carry
    uint32_t carry = 0u, hi0p = 0u, lo1p = 0u, hi1p = 0u, hi1pp = 0u;
    constexpr int B = 32;
#pragma unroll 1
    for (int base = 0; base < L + 2; base += B) {
        uint32_t xv[B];
        if (base < L) load_block<B>(xv, x, base);  // (L is a multiple of 32 here)
#pragma unroll
        for (int k = 0; k < B; k++) {
            const int j = base + k;
            if (j < L + 2) {
                const uint32_t xj = (base < L) ? xv[k] : 0u;
                const uint64_t p0 = (uint64_t)xj * y0, p1 = (uint64_t)xj * y1;
                const uint64_t t = (uint64_t)carry + (uint32_t)p0 + hi0p + lo1p + hi1pp;
                Ph.st(j + 2, (uint32_t)t);
                carry = (uint32_t)(t >> 32);
                hi0p = (uint32_t)(p0 >> 32);
                hi1pp = hi1p;
                hi1p = (uint32_t)(p1 >> 32);
                lo1p = (uint32_t)p1;
            }
        }
    }
}
// one truncated product of a fused dot product: the short form when one operand is known (trailing-zero hint) to be
// integer-like -- for every lane of the warp when UNIFORM (thread-per-instance kernels: a divergent choice would run
// both forms), for this lane alone otherwise (cooperative kernels, whose lanes diverge anyway).  An operand flagged
// zero counts as short: its term is masked by the caller.
template <int L, bool UNIFORM, class SCR, class XS, class YS>
CB_HD void product_trunc_auto(SCR Ph, const OpdT<XS>& x, const OpdT<YS>& y) {
    auto is_short = [](uint32_t f) { return tz_hint(f) >= L - 2 || (f & (F_ZERO | F_NAN)) != 0u; };
    bool sy = is_short(y.t.flags), sx = is_short(x.t.flags);
    if constexpr (UNIFORM) {
        sy = warp_all(sy);
        sx = warp_all(sx);
    }
    if (sy) product_short_y<L>(Ph, x.r, y.r);
    else if (sx) product_short_y<L>(Ph, y.r, x.r);
    else product_trunc_blocked<L>(Ph, x.r, y.r);
}

// dst = x1*y1 +- x2*y2 above 1024 bits from truncated blocked products;
// *ok = false when the guard bits cannot prove the result (the caller then runs the exact path).
// S0, S1: shared columns of at least L + 4 words.
// The work is split into what depends on the operand records only (FdotPrep), the two products (term 1 -> P1, term 2
// -> P2: independent of each other -- the cooperative kernels of cb_coop.cuh form them on two lanes) and the tail
// that aligns, adds and rounds them.
struct FdotPrep {
    bool nan, z1, z2, n1, n2;
    int64_t e1, e2;
    Mag rad;
    int tz1, tz2;  // trailing zero limbs of the two products (hints; 0 when unknown)
};
template <int L, class XS, class YS>
CB_HD FdotPrep fdot2_prep(const OpdT<XS>& x1, const OpdT<YS>& y1, const OpdT<XS>& x2, const OpdT<YS>& y2, bool neg2) {
    FdotPrep f;
    f.nan = is_nan(x1.t) || is_nan(y1.t) || is_nan(x2.t) || is_nan(y2.t);
    f.z1 = is_zero_mid(x1.t) || is_zero_mid(y1.t) || f.nan;
    f.z2 = is_zero_mid(x2.t) || is_zero_mid(y2.t) || f.nan;
    f.rad = mul_rad(mag_zero(), x1.t, top_limb<L>(x1), y1.t, top_limb<L>(y1));
    f.rad = mul_rad(f.rad, x2.t, top_limb<L>(x2), y2.t, top_limb<L>(y2));
    f.n1 = ((x1.t.flags ^ y1.t.flags) & F_SIGN) != 0;
    f.n2 = (((x2.t.flags ^ y2.t.flags) & F_SIGN) != 0) != neg2;
    f.e1 = (int64_t)x1.t.exp + y1.t.exp;
    f.e2 = (int64_t)x2.t.exp + y2.t.exp;
    f.tz1 = tz_hint(x1.t.flags) + tz_hint(y1.t.flags);
    f.tz2 = tz_hint(x2.t.flags) + tz_hint(y2.t.flags);
    return f;
}
// P1 / P2: the truncated products of term 1 / term 2 (L + 4 words each, untouched for a zero term)
template <int L, class SCR, class DST>
CB_NOINLINE Meta fdot2_tb_tail(DST dst, const FdotPrep& f, SCR P1, SCR P2, bool* ok) {
    constexpr int K = L - 4, PW = L + 4;
    // term A (zero, else smaller exponent) may be read with a word offset
    const bool a_is_1 = f.z1 ? true : (f.z2 ? false : (f.e1 <= f.e2));
    const bool zA = a_is_1 ? f.z1 : f.z2, zB = a_is_1 ? f.z2 : f.z1;
    const bool nA = a_is_1 ? f.n1 : f.n2, nB = a_is_1 ? f.n2 : f.n1;
    const int64_t eA = a_is_1 ? f.e1 : f.e2, eB = a_is_1 ? f.e2 : f.e1;
    const SCR SA = a_is_1 ? P1 : P2, SB = a_is_1 ? P2 : P1;
    bool done = false, neg = false;
    Rounded rr;
    rr.exp = 0;
    rr.zero = true;
    rr.inexact = false;
    rr.overflow = false;
    if (zA && zB) {
#pragma unroll 8
        for (int j = 0; j < L; j++) dst.st(j, 0u);
        done = true;
    } else {
        int64_t d = zA ? 0 : (eB - eA);
        int m = d > 32 * (2 * L + 8) ? (2 * L + 8) : (int)(d >> 5);
        int sA = (int)(d & 31);
        bool sub = (nA != nB) && !zA;
        auto wB = [&](int k) -> uint32_t { return zB ? 0u : SB.ld(k - K); };
        auto wA = [&](int k) -> uint32_t {
            int i = k + m - K;
            return (zA || (unsigned)i >= (unsigned)PW) ? 0u : SA.ld(i);
        };
        const int tzA = a_is_1 ? f.tz1 : f.tz2, tzB = a_is_1 ? f.tz2 : f.tz1;
        bool trusted = (zA || (tzA >= K && m == 0)) && tzB >= K;
        done = fast_add_round<L, 2 * L, K>(dst, wB, wA, 0, sA, sub, nB, nA, eB, &rr, &neg, trusted);
    }
    *ok = done;
    if (f.nan) {
        store_nan<L>(dst);
        return meta_nan();
    }
    return finish_to<L>(dst, rr, neg, f.rad);
}
template <int L, class SCR, class XS = RRef, class YS = RRef, class DST = WRef>
CB_NOINLINE Meta arb_fdot2_trunc_blocked(DST dst, OpdT<XS> x1, OpdT<YS> y1, OpdT<XS> x2, OpdT<YS> y2, bool neg2,
                                         SCR S0, SCR S1, bool* ok) {
    const FdotPrep f = fdot2_prep<L>(x1, y1, x2, y2, neg2);
    bool skip1 = warp_all(f.z1), skip2 = warp_all(f.z2);
#pragma unroll 1
    for (int term = 0; term < 2; term++) {
        if (term ? skip2 : skip1) continue;
        product_trunc_auto<L, true>(term ? S1 : S0, term ? x2 : x1, term ? y2 : y1);
    }
    return fdot2_tb_tail<L>(dst, f, S0, S1, ok);
}

// dst = x + (negy ? -y : y)
template <int L, class SCR>
CB_NOINLINE Meta arb_addsub_impl(WRef dst, Opd x, Opd y, bool negy, SCR S0, SCR S1);
template <int L, class SCR>
CB_HD Meta arb_addsub(WRef dst, Opd x, Opd y, bool negy, SCR S0, SCR S1) {  // touches words 0..L+1 only
    return arb_addsub_impl<L>(dst, x, y, negy, lo_view(S0), lo_view(S1));
}
template <int L, class SCR>
CB_NOINLINE Meta arb_addsub_impl(WRef dst, Opd x, Opd y, bool negy, SCR S0, SCR S1) {
    bool nan = is_nan(x.t) || is_nan(y.t);
    bool zx = is_zero_mid(x.t) || nan, zy = is_zero_mid(y.t) || nan;
    bool nx = (x.t.flags & F_SIGN) != 0, ny = ((y.t.flags & F_SIGN) != 0) != negy;
    bool neg;
    Rounded rr;
    bool done = false;
    {
        // operands within 31 bits of each other: statically indexed sum straight from memory
        // (all limbs are read before dst is written, so dst may alias x or y)
        int64_t d = (int64_t)x.t.exp - y.t.exp;
        bool easy = !(zx && zy) && (zx || zy || (d <= 31 && d >= -31));
        if (easy) {
            int64_t E = zx ? y.t.exp : (zy ? x.t.exp : (x.t.exp > y.t.exp ? x.t.exp : y.t.exp));
            int s1 = zx ? 0 : (int)(E - x.t.exp), s2 = zy ? 0 : (int)(E - y.t.exp);
            bool sub = (nx != ny) && !zx && !zy;
            auto w1 = [&](int k) -> uint32_t { return zx ? 0u : x.r.ld(k); };
            auto w2 = [&](int k) -> uint32_t { return zy ? 0u : y.r.ld(k); };
            done = fast_add_round<L, L>(dst, w1, w2, s1, s2, sub, zx ? ny : nx, zx ? nx : ny, E, &rr, &neg);
        }
    }
    if (!done) {
        // (32 independent loads in flight per round: these operands come from global memory)
        staged_copy<L>([&](int j) { return x.r.ld(j); }, [&](int j, uint32_t v) { S0.st(j, v); });
        staged_copy<L>([&](int j) { return y.r.ld(j); }, [&](int j, uint32_t v) { S1.st(j, v); });
        rr = align_add<L, L>(dst, S0, S1, x.t.exp, y.t.exp, nx, ny, zx, zy, &neg);
    }
    if (nan) return meta_nan();
    return finish_to<L>(dst, rr, neg, mag_add(mag_of(x.t), mag_of(y.t)));
}

// Column form used by the register/shared-memory resident recurrence: A and B are scratch columns
// holding L-limb mantissas (words 0..L-1, room for L+2); the rounded x + (negb ? -b : b) lands in
// words 0..L-1 of *where (the column of the smaller operand); the other column is left clobbered.
template <int L, class SCR>
CB_NOINLINE Meta arb_addsub_cols(SCR A, Meta ta, SCR B, Meta tb, bool negb, SCR* where) {
    bool nan = is_nan(ta) || is_nan(tb);
    bool zx = is_zero_mid(ta) || nan, zy = is_zero_mid(tb) || nan;
    bool na = (ta.flags & F_SIGN) != 0, nb = ((tb.flags & F_SIGN) != 0) != negb;
    bool neg;
    Rounded rr;
    bool done = false;
    if constexpr (L <= 32) {
        int64_t d = (int64_t)ta.exp - tb.exp;
        bool easy = !(zx && zy) && (zx || zy || (d <= 31 && d >= -31));
        if (easy) {  // per lane; a lane that fails keeps its columns intact for the general path
            int64_t E = zx ? tb.exp : (zy ? ta.exp : (ta.exp > tb.exp ? ta.exp : tb.exp));
            int s1 = zx ? 0 : (int)(E - ta.exp), s2 = zy ? 0 : (int)(E - tb.exp);
            bool sub = (na != nb) && !zx && !zy;
            auto w1 = [&](int k) -> uint32_t { return zx ? 0u : A.ld(k); };
            auto w2 = [&](int k) -> uint32_t { return zy ? 0u : B.ld(k); };
            // both columns are fully read before anything is stored, so the result can go to B
            done = fast_add_round<L, L>(B, w1, w2, s1, s2, sub, zx ? nb : na, zx ? na : nb, E, &rr, &neg);
            if (done) *where = B;
        }
    }
    CB_COUNT(done ? 2 : 3);
    if (!done)
        rr = align_add<L, L, true>(A, A, B, ta.exp, tb.exp, na, nb, zx, zy, &neg, where);
    if (nan) return meta_nan();
    return finish_to<L>(*where, rr, neg, mag_add(mag_of(ta), mag_of(tb)));
}

// dst = x * k, k a signed 64-bit integer (acb_mul_fmpz / acb_mul_si with a small integer)
template <int L, class SCR>
CB_NOINLINE Meta arb_mul_si_impl(WRef dst, Opd x, int64_t k, SCR S0);
template <int L, class SCR>
CB_HD Meta arb_mul_si(WRef dst, Opd x, int64_t k, SCR S0) {  // touches words 0..L+1 only
    return arb_mul_si_impl<L>(dst, x, k, lo_view(S0));
}
template <int L, class SCR>
CB_NOINLINE Meta arb_mul_si_impl(WRef dst, Opd x, int64_t k, SCR S0) {
    bool nan = is_nan(x.t);
    uint64_t ka = k < 0 ? (uint64_t)0 - (uint64_t)k : (uint64_t)k;
    uint32_t k0 = (uint32_t)ka, k1 = (uint32_t)(ka >> 32);
    // radius: |k|^ * xr  (|x|^ * 0 and xr * 0 drop out)
    Mag km = mag_zero();
    if (ka) {
        int sh = clz64(ka);
        uint64_t kn = ka << sh;  // top bit at 63
        Meta kt{64 - sh, 0u, 0u, 0};
        km = mag_mid_upper(kt, (uint32_t)(kn >> 32));
    }
    Mag rad = mag_mul(km, mag_of(x.t));
    bool z = is_zero_mid(x.t) || ka == 0 || nan;
    // P[0..L+2) = x * ka, column by column:
    //   col j = lo(x_j k0) + hi(x_{j-1} k0) + lo(x_{j-1} k1) + hi(x_{j-2} k1) + carry
    uint32_t hi0_prev = 0u, lo1_prev = 0u, hi1_prev = 0u, hi1_prev2 = 0u, carry = 0u;
    int hi = -1;
#pragma unroll
    for (int j = 0; j < L + 2; j++) {
        uint32_t xj = (j < L) ? x.r.ld(j) : 0u;
        uint64_t p0 = (uint64_t)xj * k0, p1 = (uint64_t)xj * k1;
        uint64_t t = (uint64_t)carry + (uint32_t)p0 + hi0_prev + lo1_prev + hi1_prev2;
        uint32_t v = z ? 0u : (uint32_t)t;
        carry = (uint32_t)(t >> 32);
        S0.st(j, v);
        if (v != 0u) hi = j;
        hi0_prev = (uint32_t)(p0 >> 32);
        hi1_prev2 = hi1_prev;
        hi1_prev = (uint32_t)(p1 >> 32);
        lo1_prev = (uint32_t)p1;
    }
    Rounded rr = extract<L>(dst, S0, hi, false, (int64_t)x.t.exp - 32 * L);
    if (nan) return meta_nan();
    bool neg = ((x.t.flags & F_SIGN) != 0) != (k < 0);
    return finish_to<L>(dst, rr, neg, rad);
}

// One multiply-subtract row of the long division below: u[0 .. B) -= qd * d[0 .. B) with the running multiply carry
// `mc` and the running borrow, on register blocks (the B products are independent IMAD.WIDE; two short carry chains
// fold them into the window).  On entry / exit: mc = carry of the products so far, borrow in {0, 1}.
template <int B>
CB_HD void div_row_block(uint32_t (&u)[B], const uint32_t (&d)[B], uint32_t qd, uint32_t& mc, uint32_t& borrow) {
#if defined(__CUDA_ARCH__)
    uint32_t lo[B], hi[B];
#pragma unroll
    for (int k = 0; k < B; k++) {
        uint64_t p = (uint64_t)qd * d[k];
        lo[k] = (uint32_t)p;
        hi[k] = (uint32_t)(p >> 32);
    }
    // s[k] = lo[k] + hi[k-1] (hi[-1] = mc), carries chained; the last carry joins hi[B-1] (no overflow: the exact
    // product qd * d + mc fits B + 1 words)
    asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(lo[0]) : "r"(mc));
#pragma unroll
    for (int k = 1; k < B; k++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(lo[k]) : "r"(hi[k - 1]));
    asm volatile("addc.u32 %0, %1, 0;" : "=r"(mc) : "r"(hi[B - 1]));
    // u -= s with the incoming borrow
    uint32_t t;
    asm volatile("sub.cc.u32 %0, 0, %1;" : "=r"(t) : "r"(borrow));  // CF <- borrow
#pragma unroll
    for (int k = 0; k < B; k++) asm volatile("subc.cc.u32 %0, %0, %1;" : "+r"(u[k]) : "r"(lo[k]));
    asm volatile("subc.u32 %0, 0, 0;" : "=r"(t));
    borrow = t & 1u;
#else
    for (int k = 0; k < B; k++) {
        uint64_t p = (uint64_t)qd * d[k] + mc;
        mc = (uint32_t)(p >> 32);
        uint64_t t = (uint64_t)u[k] - (uint32_t)p - borrow;
        u[k] = (uint32_t)t;
        borrow = (uint32_t)(t >> 32) & 1u;
    }
#endif
}

// dst = x / y  (Knuth D on 32-bit limbs; numerator x * 2^(32L) in S0, divisor in S1)
template <int L, class SCR, class XS = RRef, class YS = RRef, class DST = WRef>
CB_NOINLINE Meta arb_div(DST dst, OpdT<XS> x, OpdT<YS> y, SCR S0, SCR S1) {
    bool nan = is_nan(x.t) || is_nan(y.t) || is_zero_mid(y.t);
    uint32_t ytop = top_limb<L>(y), xtop = top_limb<L>(x);
    Mag yl = mag_mid_lower(y.t, ytop);
    Mag gap = mag_sub_lower(yl, mag_of(y.t));
    if (gap.man == 0) nan = true;
    Mag rad = mag_zero();
    if (!nan && (x.t.rman != 0 || y.t.rman != 0)) {
        Mag num = mag_add(mag_mul(mag_mid_upper(x.t, xtop), mag_of(y.t)),
                          mag_mul(mag_mid_upper(y.t, ytop), mag_of(x.t)));
        rad = mag_div(num, mag_mul_lower(yl, gap));
    }
    bool zx = is_zero_mid(x.t);
    // U[0..2L] = x << 32L (U[2L] = 0), D[0..L) = y (made a harmless normalised dummy when nan).
    // Step j works on the window U[j .. j+L]; the words above it have cancelled to zero and the words below it
    // are still the zeros of the shift, so U lives in a ring of R = L + 2 words of S0 (U[k] at word k mod R, the
    // word entering the window cleared as it enters): the division touches words 0..L+1 of S0 only -- with a
    // HybridScratch that is the shared-memory part.  After step 0 the window is U[0..L] at words 0..L.
    // On an all-shared column the ring is as long as the numerator (no wrap, one unsplit inner loop: the split
    // loops cost the 1024-bit logarithmic Frobenius kernel 9 %).
    constexpr bool RING = ScrLo<SCR>::HYBRID;
    constexpr int R = RING ? L + 2 : 2 * L + 2;
    auto slot = [](int k) { return (RING && k >= R) ? k - R : k; };  // k < 2R
    // the ring and the divisor sit in the shared-memory part of a hybrid column: addressed there directly (a
    // HybridScratch access decides between shared and global memory word by word)
    static_assert(!RING || ScrLo<SCR>::LO_WORDS >= R, "the ring must fit the shared part of a hybrid column");
    const auto U = lo_view(S0), D = lo_view(S1);
    staged_copy<L>([&](int j) { return (zx || nan) ? 0u : x.r.ld(j); }, [&](int j, uint32_t v) { U.st(slot(L + j), v); });
    U.st(slot(2 * L), 0u);
    staged_copy<L>([&](int j) { return nan ? (j == L - 1 ? 0x80000000u : 0u) : y.r.ld(j); },
                   [&](int j, uint32_t v) { D.st(j, v); });
    uint32_t d1 = D.ld(L - 1), d0 = D.ld(L - 2);
    // reciprocal of the (normalised) top divisor word, v = floor((2^64 - 1) / d1) - 2^32: ONE 64-bit division per
    // arb_div; every quotient-digit estimate below is then Moeller-Granlund's 2-by-1 division (two multiplies and
    // two corrections, exact) instead of a 64-bit division subroutine
    const uint32_t dinv = (uint32_t)(0xFFFFFFFFFFFFFFFFull / d1 - 0x100000000ull);
    // quotient words q_j are kept in the upper half of S1 (words L..2L): written once, read once
    int hi = -1;
#pragma unroll 1
    for (int j = L; j >= 0; j--) {
        if (j < L) U.st(j, 0u);  // U[j] enters the window (j < R: its own word)
        uint32_t u2 = U.ld(slot(j + L)), u1 = U.ld(slot(j + L - 1)), u0 = U.ld(slot(j + L - 2));
        uint64_t num = ((uint64_t)u2 << 32) | u1;
        uint64_t qh, rh;
        if (u2 >= d1) {
            qh = 0xFFFFFFFFull;
            rh = num - qh * d1;
        } else {  // u2 < d1: (qh, rh) = divmod(u2:u1, d1)
            uint64_t q = (uint64_t)dinv * u2 + num;
            uint32_t q1 = (uint32_t)(q >> 32) + 1u, q0 = (uint32_t)q;
            uint32_t r = u1 - q1 * d1;
            if (r > q0) {
                q1--;
                r += d1;
            }
            if (r >= d1) {
                q1++;
                r -= d1;
            }
            qh = q1;
            rh = r;
        }
        while (rh <= 0xFFFFFFFFull && qh * d0 > ((rh << 32) | u0)) {
            qh--;
            rh += d1;
        }
        // U[j .. j+L] -= qh * D   (words j .. R-1 of the ring, then words 0 ..), sixteen limbs at a time in registers
        uint32_t qd = (uint32_t)qh, borrow = 0u, mc = 0u;
        {
            constexpr int B = (L % 16 == 0) ? 16 : 4;
#pragma unroll 1
            for (int blk = 0; blk < L; blk += B) {
                uint32_t dd[B], uu[B];
#pragma unroll
                for (int k = 0; k < B; k++) dd[k] = D.ld(blk + k);
                // a block that does not straddle the end of the ring is one run of words (the wrap is then paid once
                // per block, not once per word)
                const int base = j + blk;
                const bool flat = !RING || base + B <= R || base >= R;
                const int o = (RING && base >= R) ? base - R : base;
                if (flat) {
#pragma unroll
                    for (int k = 0; k < B; k++) uu[k] = U.ld(o + k);
                } else {
#pragma unroll
                    for (int k = 0; k < B; k++) uu[k] = U.ld(slot(base + k));
                }
                div_row_block<B>(uu, dd, qd, mc, borrow);
                if (flat) {
#pragma unroll
                    for (int k = 0; k < B; k++) U.st(o + k, uu[k]);
                } else {
#pragma unroll
                    for (int k = 0; k < B; k++) U.st(slot(base + k), uu[k]);
                }
            }
        }
        const int top = slot(j + L);
        uint64_t t = (uint64_t)U.ld(top) - mc - borrow;
        U.st(top, (uint32_t)t);
        if ((uint32_t)(t >> 32) & 1u) {  // one too large: add the divisor back
            qd--;
            uint32_t c = 0u;
#pragma unroll 8
            for (int i = 0; i < L; i++) {
                const int k = slot(i + j);
                uint64_t s = (uint64_t)U.ld(k) + D.ld(i) + c;
                U.st(k, (uint32_t)s);
                c = (uint32_t)(s >> 32);
            }
            U.st(top, U.ld(top) + c);
        }
        S1.st(L + j, qd);
        if (qd != 0u && hi < 0) hi = j;
    }
    // remainder in U[0..L)
    uint32_t rem = 0u;
#pragma unroll 8
    for (int i = 0; i < L; i++) rem |= U.ld(i);
    // Q = S1[L .. 2L] (L+1 words), presented to extract() as a column starting at word L
    struct View {
        SCR s;
        CB_HD uint32_t ld(int k) const { return s.ld(L + k); }
    } V{S1};
    Rounded rr = extract<L>(dst, V, hi, rem != 0u, (int64_t)x.t.exp - y.t.exp - 32 * L);
    if (nan) {
        store_nan<L>(dst);
        return meta_nan();
    }
    bool neg = ((x.t.flags ^ y.t.flags) & F_SIGN) != 0;
    return finish_to<L>(dst, rr, neg, rad);
}

// ---------------------------------------------------------------- division by a 64-bit integer
// reciprocal of the normalised two-limb divisor (d1:d0), d1 >= 2^31 (Moller-Granlund 3-by-2)
CB_HD uint32_t recip_3by2(uint32_t d1, uint32_t d0) {
    uint32_t v = (uint32_t)((((uint64_t)(~d1) << 32) | 0xFFFFFFFFull) / d1);  // floor((B^2-1)/d1) - B
    uint32_t p = d1 * v + d0;
    if (p < d0) {
        v--;
        uint32_t mask = (p >= d1) ? 0xFFFFFFFFu : 0u;
        p -= d1;
        v += mask;
        p -= mask & d1;
    }
    uint64_t t = (uint64_t)d0 * v;
    uint32_t t1 = (uint32_t)(t >> 32), t0 = (uint32_t)t;
    p += t1;
    if (p < t1) {
        v--;
        if (p >= d1 && (p > d1 || t0 >= d0)) v--;
    }
    return v;
}
// (rem : u) / D with rem < D: returns the quotient digit, updates rem (64-bit remainder)
CB_HD uint32_t div_3by2(uint64_t& rem, uint32_t u, uint64_t D, uint32_t v) {
    const uint32_t d1 = (uint32_t)(D >> 32), d0 = (uint32_t)D;
    const uint32_t n2 = (uint32_t)(rem >> 32), n1 = (uint32_t)rem;
    uint64_t qq = (uint64_t)n2 * v + rem;
    uint32_t q = (uint32_t)(qq >> 32), q0 = (uint32_t)qq;
    uint32_t r1 = n1 - d1 * q;
    uint64_t R = (((uint64_t)r1 << 32) | u) - D;
    R -= (uint64_t)d0 * q;
    q++;
    if ((uint32_t)(R >> 32) >= q0) {
        q--;
        R += D;
    }
    if (R >= D) {
        q++;
        R -= D;
    }
    rem = R;
    return q;
}

// dst = x / d for an exact integer divisor d (|d| < 2^63 given as magnitude + sign): the same ball
// arb_div(x, ball(d)) produces, in O(L) instead of O(L^2).  S0: L+3 words.
template <int L, class SCR, class XS = RRef, class DST = WRef>
CB_NOINLINE Meta arb_div_u64(DST dst, OpdT<XS> x, uint64_t d, bool dneg, SCR S0) {
    bool nan = is_nan(x.t) || d == 0;
    int sh = d ? clz64(d) : 0;
    uint64_t D = d ? (d << sh) : (1ull << 63);
    uint32_t d1 = (uint32_t)(D >> 32);
    Meta yt{64 - sh, dneg ? F_SIGN : 0u, 0u, 0};
    Mag yl = mag_mid_lower(yt, d1);
    Mag rad = mag_zero();
    if (!nan && x.t.rman != 0) {
        Mag num = mag_mul(mag_mid_upper(yt, d1), mag_of(x.t));
        rad = mag_div(num, mag_mul_lower(yl, yl));
    }
    bool zx = is_zero_mid(x.t) || nan;
    uint32_t v = recip_3by2(d1, (uint32_t)D);
    // quotient digits L .. 0 of floor(mx * 2^64 / D) (digit L+1 is always 0)
    uint64_t rem = 0;
    int hi = -1;
    (void)div_3by2(rem, zx ? 0u : x.r.ld(L - 1), D, v);
#pragma unroll 4
    for (int j = L; j >= 0; j--) {
        uint32_t u = (j >= 2 && !zx) ? x.r.ld(j - 2) : 0u;
        uint32_t q = div_3by2(rem, u, D, v);
        S0.st(j, q);
        if (q != 0u && hi < 0) hi = j;
    }
    Rounded rr = extract<L>(dst, S0, hi, rem != 0, (int64_t)x.t.exp - yt.exp - 32 * L);
    if (nan) {
        store_nan<L>(dst);
        return meta_nan();
    }
    bool neg = ((x.t.flags & F_SIGN) != 0) != dneg;
    return finish_to<L>(dst, rr, neg, rad);
}

}  // namespace cb

#include "cb_rr.cuh"
