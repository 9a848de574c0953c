// cb_solvers.cuh -- per-thread bodies of the batched solvers: complex ball ops on top of
// cb_arith.cuh and the restated Cascade loops, one CUDA thread per ODE instance.
//
//   fuchs_table_thread + fuchs_shared_thread  <- acb_ode_solve_fuchs   reference src/fuchs_solver.c:96-145
//   frobenius1_thread  <- _acb_ode_solve_frobenius       reference src/frobenius_solver.c:72-121
//                         indicial_polynomial_evaluate   reference src/frobenius_solver.c:41-70
//   horner_thread      <- acb_poly_evaluate (Horner) and the first `order` Taylor coefficients
//                         of acb_poly_taylor_shift       reference src/acb_ode_solution.c:40,
//                                                                  src/fuchs_solver.c:159
//   ew_thread          <- the scalar Arb calls themselves (unit-test surface)
//
// Memory layout: every ball array is [entry][limb][S] mantissa words + [entry][S] meta records,
// S = 2 * batch (slot = 2*instance + part) for per-instance data and S = 2 for data shared by
// the whole batch.  A thread reads/writes only its own slots (and shared data), so no
// synchronisation between threads is needed anywhere.
#pragma once
#include <type_traits>
#include "cb_arith.cuh"

// Block-wide phase alignment of the recurrence kernel: the hot loop is ~64 KB of straight-line code, twice the
// 32 KB L1.5 instruction cache, and warps that drift apart each stream it through the GPC-level instruction
// cache on their own.  A barrier before every long call keeps the warps of a block on the same cache lines.
// (Exited threads of a ragged last block do not take part in the barrier.)  Measured:
// +10 % at 12 warps per SM, nothing at 7 -- so only the 384-thread variant is built with it.
template <bool SYNC>
CB_HD void phase_sync() {
#if defined(__CUDA_ARCH__)
    if constexpr (SYNC) __syncthreads();
#endif
}
#define CB_PHASE_SYNC() phase_sync<SYNC>()

namespace cb {

// a complex ball in memory
struct CRef {
    uint32_t* m;  // re limb k at m[k*stride], im limb k at m[k*stride + 1]
    size_t stride;
    Meta* t;  // t[0] re, t[1] im
};
// an array of complex balls
struct CArr {
    uint32_t* m;
    Meta* t;
    size_t S;  // slots per entry
};
template <int L>
CB_HD CRef at(const CArr& a, int64_t entry, size_t slot) {
    return CRef{a.m + (size_t)entry * L * a.S + slot, a.S, a.t + (size_t)entry * a.S + slot};
}

CB_HD Meta ld_meta(const Meta* p) {
#if defined(__CUDA_ARCH__)
    int4 v = *reinterpret_cast<const int4*>(p);
    return Meta{v.x, (uint32_t)v.y, (uint32_t)v.z, v.w};
#else
    return *p;
#endif
}
CB_HD void st_meta(Meta* p, const Meta& t) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<int4*>(p) = make_int4(t.exp, (int)t.flags, (int)t.rman, t.rexp);
#else
    *p = t;
#endif
}
// Balls that may come from the caller: only the three flag bits of include/cb_format.h are taken from memory; bits
// 8..15 of flags are an INTERNAL trailing-zero hint (F_TZ_SHIFT) that only this library's own tables may carry --
// a caller-supplied value there must never be trusted (it would switch off the guard-bit proof of the truncated
// products).
CB_HD Meta ld_meta_user(const Meta* p) {
    Meta t = ld_meta(p);
    t.flags &= (F_SIGN | F_ZERO | F_NAN);
    return t;
}
CB_HD Opd re_of(const CRef& x) { return Opd{RRef{x.m, x.stride}, ld_meta_user(x.t)}; }
CB_HD Opd im_of(const CRef& x) { return Opd{RRef{x.m + 1, x.stride}, ld_meta_user(x.t + 1)}; }
// internal table operands (written by fuchs_table_thread / acb_mark_tz): hints kept
CB_HD Opd re_of_tbl(const CRef& x) { return Opd{RRef{x.m, x.stride}, ld_meta(x.t)}; }
CB_HD Opd im_of_tbl(const CRef& x) { return Opd{RRef{x.m + 1, x.stride}, ld_meta(x.t + 1)}; }
// shared-format entries (S = 2): fixed-stride views
using OpdS = OpdT<RFix<2>>;
CB_HD OpdS re_of_shared(const CRef& x) { return OpdS{RFix<2>{x.m}, ld_meta(x.t)}; }
CB_HD OpdS im_of_shared(const CRef& x) { return OpdS{RFix<2>{x.m + 1}, ld_meta(x.t + 1)}; }
CB_HD Opd dyn(const OpdS& o) { return Opd{o.r.dyn(), o.t}; }
CB_HD WRef wre(const CRef& x) { return WRef{x.m, x.stride}; }
CB_HD WRef wim(const CRef& x) { return WRef{x.m + 1, x.stride}; }

// a table operand: fixed-stride view for a table shared by the batch, strided view for a per-instance table
template <class YO> CB_HD YO y_re(const CRef& x);
template <class YO> CB_HD YO y_im(const CRef& x);
template <> CB_HD OpdS y_re<OpdS>(const CRef& x) { return re_of_shared(x); }
template <> CB_HD OpdS y_im<OpdS>(const CRef& x) { return im_of_shared(x); }
template <> CB_HD Opd y_re<Opd>(const CRef& x) { return re_of_tbl(x); }
template <> CB_HD Opd y_im<Opd>(const CRef& x) { return im_of_tbl(x); }

// does this precision use truncated products (with the exact path as per-lane fallback)?
template <int L>
constexpr bool use_trunc() { return L >= 16 && L <= 32; }
// does the recurrence kernel try the reduced-radix fused dot product (cb_rr.cuh) first?
template <int L>
constexpr bool use_rr() { return L == 32; }

// fused dot product of the general kernels: truncated products first (S0 parks the first one),
// the exact shared-memory version when the guard-bit proof fails
template <int L, class SCR>
CB_HD Meta fdot2_auto(WRef dst, Opd x1, Opd y1, Opd x2, Opd y2, bool neg2, SCR S0, SCR S1) {
    if constexpr (L == 64 || L == 128) {
        // short (integer-like) operands make exact products, whose guard bits prove nothing:
        // knowing their trailing zero limbs lets the truncated path accept them as exact
        // (at 1024 bits arb_fdot2_trunc finds them itself from the limbs it has in registers)
        tz_scan4<L>(x1, y1, x2, y2);
    }
    if constexpr (use_trunc<L>()) {
        bool ok;
        Meta t = arb_fdot2_trunc<L, SCR, RRef, RRef, WRef, true>(dst, x1, y1, x2, y2, neg2, S0, &ok);
        if (!ok) t = arb_fdot2<L>(dst, x1, y1, x2, y2, neg2, S0, S1);
        return t;
    } else if constexpr (L == 64 || L == 128) {
        bool ok;
        Meta t = arb_fdot2_trunc_blocked<L>(dst, x1, y1, x2, y2, neg2, lo_view(S0), lo_view(S1), &ok);  // L + 4 words
        CB_COUNT(ok ? 4 : 5);
        if (!ok) t = arb_fdot2<L>(dst, x1, y1, x2, y2, neg2, S0, S1);
        return t;
    } else {
        return arb_fdot2<L>(dst, x1, y1, x2, y2, neg2, S0, S1);
    }
}

template <int L>
CB_HD void acb_set(const CRef& z, const CRef& x) {
    staged_copy<2 * L>([&](int j) { return x.m[(size_t)(j >> 1) * x.stride + (j & 1)]; },
                       [&](int j, uint32_t v) { z.m[(size_t)(j >> 1) * z.stride + (j & 1)] = v; });
    st_meta(z.t, ld_meta(x.t));
    st_meta(z.t + 1, ld_meta(x.t + 1));
}
template <int L>
CB_HD void acb_zero(const CRef& z) {
#pragma unroll 8
    for (int k = 0; k < L; k++) {
        z.m[k * z.stride] = 0u;
        z.m[k * z.stride + 1] = 0u;
    }
    st_meta(z.t, meta_zero());
    st_meta(z.t + 1, meta_zero());
}
CB_HD bool acb_is_exact_zero(const CRef& x) { return is_exact_zero(ld_meta(x.t)) && is_exact_zero(ld_meta(x.t + 1)); }
template <int L>
CB_HD void acb_one(const CRef& z) {
    acb_zero<L>(z);
    z.m[(L - 1) * z.stride] = 0x80000000u;
    st_meta(z.t, Meta{1, 0u, 0u, 0});
}

// z = x * y.  z must not alias x or y.  (acb_mul: one rounding per component)
template <int L, class SCR>
CB_HD void acb_mul(const CRef& z, const CRef& x, const CRef& y, SCR S0, SCR S1) {
    Opd a = re_of(x), b = im_of(x), c = re_of(y), d = im_of(y);
    Meta tr = fdot2_auto<L>(wre(z), a, c, b, d, true, S0, S1);
    Meta ti = fdot2_auto<L>(wim(z), a, d, b, c, false, S0, S1);
    st_meta(z.t, tr);
    st_meta(z.t + 1, ti);
}
// z = x +- y (componentwise; z may alias x or y)
template <int L, class SCR>
CB_HD void acb_addsub(const CRef& z, const CRef& x, const CRef& y, bool sub, SCR S0, SCR S1) {
    Meta tr = arb_addsub<L>(wre(z), re_of(x), re_of(y), sub, S0, S1);
    Meta ti = arb_addsub<L>(wim(z), im_of(x), im_of(y), sub, S0, S1);
    st_meta(z.t, tr);
    st_meta(z.t + 1, ti);
}
template <int L, class SCR>
CB_HD void acb_mul_si(const CRef& z, const CRef& x, int64_t k, SCR S0) {
    Meta tr = arb_mul_si<L>(wre(z), re_of(x), k, S0);
    Meta ti = arb_mul_si<L>(wim(z), im_of(x), k, S0);
    st_meta(z.t, tr);
    st_meta(z.t + 1, ti);
}
// ask the L2 for a ball that will be read a few thousand instructions from now (operands that come around once per
// coefficient -- the per-instance operator, the earlier coefficients -- have left the L2 by then when the batch's
// working set exceeds it): one request per limb covers the real and the imaginary word
template <int L>
CB_HD void acb_prefetch(const CRef& x) {
#if defined(__CUDA_ARCH__)
    const uint32_t* q = x.m;
#pragma unroll 16
    for (int j = 0; j < L; j++) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
        q += x.stride;
    }
    asm volatile("prefetch.global.L2 [%0];" ::"l"(x.t));
#else
    (void)x;
#endif
}
template <int L>
CB_HD void acb_copy_im(const CRef& z, const CRef& x) {
    staged_copy<L>([&](int j) { return x.m[(size_t)j * x.stride + 1]; }, [&](int j, uint32_t v) { z.m[(size_t)j * z.stride + 1] = v; });
    st_meta(z.t + 1, ld_meta(x.t + 1));
}
// z = x + k: real part gets the integer, imaginary part is copied (acb_add_si).  COPY_IM = false: the caller has put
// the imaginary part of x into z already (a loop adding integers to the same x over and over)
template <int L, bool COPY_IM = true, class SCR>
CB_HD void acb_add_si(const CRef& z, const CRef& x, int64_t k, const CRef& kball, SCR S0, SCR S1) {
    // materialise k as an exact ball in kball.re (scratch ball owned by the caller)
    uint64_t ka = k < 0 ? (uint64_t)0 - (uint64_t)k : (uint64_t)k;
    int sh = clz64(ka);
    uint64_t kn = ka ? ka << sh : 0;
#pragma unroll 8
    for (int j = 0; j < L; j++) kball.m[j * kball.stride] = 0u;
    kball.m[(L - 1) * kball.stride] = (uint32_t)(kn >> 32);
    kball.m[(L - 2) * kball.stride] = (uint32_t)kn;
    Meta kt = ka ? Meta{64 - sh, k < 0 ? F_SIGN : 0u, 0u, 0} : meta_zero();
    Opd ko{RRef{kball.m, kball.stride}, kt};
    Meta tr = arb_addsub<L>(wre(z), re_of(x), ko, false, S0, S1);
    st_meta(z.t, tr);
    if (COPY_IM && z.m != x.m) acb_copy_im<L>(z, x);
}
// z = 1 / y; z must not alias y; uses tmp.re as a real temporary
template <int L, class SCR>
CB_HD void acb_inv(const CRef& z, const CRef& y, const CRef& tmp, SCR S0, SCR S1) {
    Opd c = re_of(y), d = im_of(y);
    Meta tt = fdot2_auto<L>(wre(tmp), c, c, d, d, false, S0, S1);
    Opd t{RRef{tmp.m, tmp.stride}, tt};
    Meta tr = arb_div<L>(wre(z), c, t, S0, S1);
    Meta ti = arb_div<L>(wim(z), d, t, S0, S1);
    if (!is_nan(ti) && !is_zero_mid(ti)) ti.flags ^= F_SIGN;
    st_meta(z.t, tr);
    st_meta(z.t + 1, ti);
}
// z = x / y; z must not alias x or y; inv: a complex temporary
template <int L, class SCR>
CB_HD void acb_div(const CRef& z, const CRef& x, const CRef& y, const CRef& inv, SCR S0, SCR S1) {
    Meta yim = ld_meta(y.t + 1);
    if (is_exact_zero(yim)) {
        Opd c = re_of(y);
        Meta tr = arb_div<L>(wre(z), re_of(x), c, S0, S1);
        Meta ti = arb_div<L>(wim(z), im_of(x), c, S0, S1);
        st_meta(z.t, tr);
        st_meta(z.t + 1, ti);
    } else {
        acb_inv<L>(inv, y, z, S0, S1);
        acb_mul<L>(z, x, inv, S0, S1);
    }
}

// record the number of trailing zero limbs of both parts in the hint bits of their flags
// (internal tables only; see F_TZ_SHIFT)
template <int L>
CB_HD void acb_mark_tz(const CRef& z) {
    for (int part = 0; part < 2; part++) {
        Meta t = ld_meta(z.t + part);
        int tz = 0;
        if (!(t.flags & (F_ZERO | F_NAN))) {
            while (tz < L - 1 && z.m[(size_t)tz * z.stride + part] == 0u) tz++;
        }
        t.flags = (t.flags & 0xFFu) | ((uint32_t)tz << F_TZ_SHIFT);
        st_meta(z.t + part, t);
    }
}

CB_HD int64_t clampi(int64_t x, int64_t lo, int64_t hi) { return x < lo ? lo : (x > hi ? hi : x); }

// ================================================================ Fuchs (acb_ode_solve_fuchs, reference src/fuchs_solver.c:96-145)
// When every instance has the same operator, the integer-weighted multipliers
//   cur(bmax, b) = sum_i P[i][i+e-b] * fac ...   (reference src/fuchs_solver.c:114,121-135)
// and the divisor are identical for all instances.  fuchs_table_thread computes them ONCE per
// bmax with exactly the reference's operations (so every ball is bit-identical to what each
// instance would have computed), including acb_inv of a complex divisor -- acb_div(x, y) is
// acb_mul(x, acb_inv(y)) for complex y.  fuchs_shared_thread then runs only the per-instance
// part: W complex multiply-subtracts and one multiply (or a division by a real divisor) per
// coefficient, with the accumulator held in shared-memory columns.
//
// When the operators differ between instances the same two threads run with a table row PER INSTANCE (S = 2 * batch
// table slots, built for a range of rows at a time): one table thread per (row, instance).
//
// Table layout: entry (bmax + v - tbl_row0) * (W + 2) + slot, W = degree - v (S = 2: one table for the batch):
//   slot k < cnt    multiplier for b = bmin + k          (cnt = bmax - bmin)
//   slot cnt        the divisor
//   slot W + 1      its inverse when kind[bmax + v] == 1 (complex divisor), unused for kind 0 (real)
struct FuchsTableParams {
    int order, degree, valuation;
    int64_t n;
    CArr ode;   // S = 2 (one operator for the batch) or 2 * ninst (an operator per instance)
    CArr tbl;   // (rows held) * (W + 2) entries, S = 2 or 2 * ninst
    CArr tmp;   // 2 entries, S = 2 * (rows of this launch * ninst)   (per table thread)
    int32_t* kind;  // [rows held][ninst]
    int64_t row0;   // first table row computed by this launch (rows row0 .. row0 + nrows - 1)
    int64_t nrows;
    int64_t ninst;     // 1: shared operator; else the batch size (per-instance operators)
    int64_t tbl_row0;  // table row stored at entry 0 (0 for the whole-solve table; the range start for a range table)
};

// what follows the multipliers of a table row: the divisor's kind, the inverse of a complex divisor, the
// trailing-zero hints of every slot (slots 0 .. cnt hold the multipliers and the divisor)
template <int L, class SCR>
CB_HD void fuchs_table_tail(const FuchsTableParams& p, int64_t idx, int64_t row, int64_t inst, int64_t base, size_t qslot,
                            int cnt, bool hints_done, SCR S0, SCR S1) {
    const int W = p.degree - p.valuation;
    CRef scratch = at<L>(p.tmp, 1, 2 * (size_t)idx);
    CRef t2 = at<L>(p.tbl, base + cnt, qslot);
    Meta dim = ld_meta(t2.t + 1);
    int32_t* kind = p.kind + (row - p.tbl_row0) * p.ninst + inst;
    if (is_exact_zero(dim)) {
        *kind = 0;
    } else {
        *kind = 1;
        acb_inv<L>(at<L>(p.tbl, base + W + 1, qslot), t2, scratch, S0, S1);
        acb_mark_tz<L>(at<L>(p.tbl, base + W + 1, qslot));
    }
    if (!hints_done)
        for (int s = 0; s <= cnt; s++) acb_mark_tz<L>(at<L>(p.tbl, base + s, qslot));
}

// one thread per (table row, instance): idx = (row - row0) * ninst + instance
template <int L, class SCR>
CB_HD void fuchs_table_thread(const FuchsTableParams& p, int64_t idx, SCR S0, SCR S1) {
    const int v = p.valuation, D1 = p.degree + 1, W = p.degree - v;
    const int64_t inst = idx % p.ninst, row = p.row0 + idx / p.ninst;
    const size_t oslot = (p.ode.S == 2) ? 0 : 2 * (size_t)inst, qslot = (p.tbl.S == 2) ? 0 : 2 * (size_t)inst;
    const int64_t bmax = row - v;
    const size_t tslot = 2 * (size_t)idx;
    CRef t1 = at<L>(p.tmp, 0, tslot);
    int64_t e = bmax + v;
    int64_t bmin = clampi(e - p.degree, 0, bmax);
    int64_t base = (row - p.tbl_row0) * (W + 2);
    acb_set<L>(at<L>(p.tbl, base, qslot), at<L>(p.ode, 0 * D1 + (e - bmin), oslot));
    int64_t b = bmin;
    int k = 0;
    CRef t2 = at<L>(p.tbl, base, qslot);
#pragma unroll 1
    do {
        k++;
        b++;
        t2 = at<L>(p.tbl, base + k, qslot);
        acb_zero<L>(t2);
        int64_t fac = 1;
        int64_t imin = clampi(b - e, 0, -v);
        int64_t imax = clampi(b - bmin, 0, p.order);
#pragma unroll 1
        for (int64_t i = imin; i <= imax; i++) {
            acb_mul_si<L>(t1, at<L>(p.ode, i * D1 + (i + e - b), oslot), fac, S0);
            acb_addsub<L>(t2, t2, t1, false, S0, S1);
            fac *= (b - i);
        }
        fac = 1;
        for (int64_t q = 0; q < imin; q++) fac *= (b - imin + 1 + q);
        acb_mul_si<L>(t2, t2, fac, S0);
    } while (b != bmax);
    fuchs_table_tail<L>(p, idx, row, inst, base, qslot, k, false, S0, S1);
}

struct FuchsSharedParams {
    int order, degree, valuation;
    int64_t n, batch;
    CArr tbl;   // see above
    CArr res;   // n+1 entries, S = 2*batch
    const int32_t* kind;
    uint32_t* gcols;  // 2 * (2L+3) words per instance, [word][instance]: columns of the general path
    int no_rr;        // development switch (CB200_NO_RR=1): skip the reduced-radix products
    unsigned long long* stats;  // development counters (CB200_RR_STATS=1) or nullptr: [0] rr accepted, [1] rr rejected
    int64_t n_begin;  // first coefficient this launch computes (>= -valuation); earlier ones are already in res
    // persistent launch (cb_kernels.cuh): work units of range_len coefficients x one block of instances
    int64_t range_len;
    int* sched;  // [0] unit counter, [1 + b] ranges completed for instance block b (zeroed before the launch)
    // table addressing: entry 0 holds row tbl_row0; kind is [rows held][kind_stride] (kind_stride = 1 or batch)
    int64_t tbl_row0, kind_stride;
    int per_inst;  // the table holds a row per instance (launches the PI instantiation)
};
struct RecCtl {
    bool rr;
    unsigned long long* stats;
};


// one fused dot product of the recurrence: truncated products with the exact general path as the
// per-lane fallback (L = 16..32), or the exact in-shared-memory version (other precisions)
CB_HD Opd dyn(const Opd& o) { return o; }
template <class YO> struct y_src;
template <> struct y_src<OpdS> { using type = RFix<2>; };
template <> struct y_src<Opd> { using type = RRef; };

template <int L, bool RR, class SCR, class DST, class YO>
CB_HD Meta rec_fdot2(DST dst, Opd x1, YO y1, Opd x2, YO y2, bool neg2, SCR S0, SCR S1, GScratch G0, GScratch G1,
                     RecCtl ctl = RecCtl{true, nullptr}) {
    using YS = typename y_src<YO>::type;
    if constexpr (use_trunc<L>()) {
        bool ok = false;
        Meta t;
        if constexpr (RR && use_rr<L>() && std::is_same<YO, OpdS>::value) {
            // reduced-radix products (plain IMAD.WIDE, no carry chains) first -- unless a shared
            // multiplier is zero or short (exact products: the carry-chain path knows how to trust
            // them); y1, y2 are the same for every lane, so this choice does not diverge
            const bool dense = !(is_zero_mid(y1.t) || is_zero_mid(y2.t) || is_nan(y1.t) || is_nan(y2.t) ||
                                 tz_hint(y1.t.flags) + tz_hint(y2.t.flags) > 0);
            if (dense && ctl.rr) {
                t = arb_fdot2_rr<L, RRef, RFix<2>>(dst, x1, y1, x2, y2, neg2, &ok);
                CB_COUNT(ok ? 6 : 7);
#if defined(__CUDA_ARCH__)
                if (ctl.stats) atomicAdd(ctl.stats + (ok ? 0 : 1), 1ull);
#endif
            }
        }
        if (!ok) {
            t = arb_fdot2_trunc<L, SCR, RRef, YS>(dst, x1, y1, x2, y2, neg2, S0, &ok);
            CB_COUNT(ok ? 4 : 5);
            if (!ok) t = arb_fdot2_general<L>(dst, x1, dyn(y1), x2, dyn(y2), neg2, G0, G1);
        }
        return t;
    } else {
        return fdot2_auto<L>(dst, x1, dyn(y1), x2, dyn(y2), neg2, S0, S1);
    }
}

// AC: three small scratch columns (L+2 words each) rotating between acc.re, acc.im and "free"
// RR: instantiate the reduced-radix attempt (a separate kernel: merely having the call in the code costs the
// default kernel 50 registers and spills)
// PI: the table holds rows per instance (operators that differ between instances: parameter sweeps)
template <int L, bool RR = false, bool SYNC = false, bool PI = false, class SCR>
CB_HD void fuchs_shared_thread(const FuchsSharedParams& p, int64_t inst, SCR S0, SCR S1, SCR A0, SCR A1,
                               SCR A2, GScratch G0, GScratch G1, int64_t n_from = -1, int64_t n_to = -1) {
    const size_t slot = 2 * (size_t)inst, qslot = PI ? slot : 0;
    using YO = typename std::conditional<PI, Opd, OpdS>::type;
    const int v = p.valuation, W = p.degree - v;
    SCR accRe = A0, accIm = A1, spare = A2;
    const RecCtl ctl{p.no_rr == 0, p.stats};
    // every coefficient is formed from the W previous ones read back from res, so the recurrence can be
    // resumed at any index: a caller may solve in coefficient ranges and stream finished rows out meanwhile
    // (n_from, n_to): the coefficient range of a persistent work unit; otherwise the launch's own range
    const int64_t b_first = n_from >= 0 ? n_from : (p.n_begin > -v ? p.n_begin : -v);
    const int64_t b_last = n_from >= 0 ? n_to : p.n;
#pragma unroll 1
    for (int64_t bmax = b_first; bmax <= b_last; bmax++) {
        int64_t e = bmax + v;
        int64_t bmin = clampi(e - p.degree, 0, bmax);
        int cnt = (int)(bmax - bmin);
        int64_t base = (bmax + v - p.tbl_row0) * (W + 2);
        Meta tRe = meta_zero(), tIm = meta_zero();
#pragma unroll 1
        for (int k = 0; k < cnt; k++) {
            CRef x = at<L>(p.res, bmin + k, slot), y = at<L>(p.tbl, base + k, qslot);
            Opd a = re_of(x), b = im_of(x);
            YO c = y_re<YO>(y), d = y_im<YO>(y);
            // real part: acc.re -= a c - b d
            CB_PHASE_SYNC();
            Meta pr = rec_fdot2<L, RR>(WRef{spare.generic(), SCR::GSTRIDE}, a, c, b, d, true, use_trunc<L>() ? spare : S0,
                                   S1, G0, G1, ctl);
            if (k == 0) {  // 0 - x = -x exactly: adopt the column, flip the sign
                SCR t = accRe;
                accRe = spare;
                spare = t;
                tRe = pr;
                if (!is_nan(tRe) && !is_zero_mid(tRe)) tRe.flags ^= F_SIGN;
            } else {
                SCR where;
                CB_PHASE_SYNC();
                tRe = arb_addsub_cols<L>(accRe, tRe, spare, pr, true, &where);
                if (where.a_or_p() == spare.a_or_p()) {
                    SCR t = accRe;
                    accRe = spare;
                    spare = t;
                }
            }
            // imaginary part: acc.im -= a d + b c
            CB_PHASE_SYNC();
            Meta pi = rec_fdot2<L, RR>(WRef{spare.generic(), SCR::GSTRIDE}, a, d, b, c, false,
                                   use_trunc<L>() ? spare : S0, S1, G0, G1, ctl);
            if (k == 0) {
                SCR t = accIm;
                accIm = spare;
                spare = t;
                tIm = pi;
                if (!is_nan(tIm) && !is_zero_mid(tIm)) tIm.flags ^= F_SIGN;
            } else {
                SCR where;
                CB_PHASE_SYNC();
                tIm = arb_addsub_cols<L>(accIm, tIm, spare, pi, true, &where);
                if (where.a_or_p() == spare.a_or_p()) {
                    SCR t = accIm;
                    accIm = spare;
                    spare = t;
                }
            }
        }
        CRef out = at<L>(p.res, bmax, slot);
        Opd xa{RRef{accRe.generic(), SCR::GSTRIDE}, tRe}, xb{RRef{accIm.generic(), SCR::GSTRIDE}, tIm};
        if (p.kind[(bmax + v - p.tbl_row0) * p.kind_stride + (PI ? inst : 0)] == 1) {
            CRef y = at<L>(p.tbl, base + W + 1, qslot);
            YO c = y_re<YO>(y), d = y_im<YO>(y);
            CB_PHASE_SYNC();
            Meta tr = rec_fdot2<L, RR>(wre(out), xa, c, xb, d, true, use_trunc<L>() ? spare : S0, S1, G0, G1, ctl);
            CB_PHASE_SYNC();
            Meta ti = rec_fdot2<L, RR>(wim(out), xa, d, xb, c, false, use_trunc<L>() ? spare : S0, S1, G0, G1, ctl);
            st_meta(out.t, tr);
            st_meta(out.t + 1, ti);
        } else {
            CRef y = at<L>(p.tbl, base + cnt, qslot);
            Opd c = re_of(y);
            // real divisor: long division needs two full-size columns; with truncated products the
            // shared columns are short, so the global ones serve (this path is integer-operator
            // territory, where the division dominates anyway)
            Meta tr, ti;
            if constexpr (use_trunc<L>()) {
                tr = arb_div<L>(wre(out), xa, c, G0, G1);
                ti = arb_div<L>(wim(out), xb, c, G0, G1);
            } else {
                tr = arb_div<L>(wre(out), xa, c, S0, S1);
                ti = arb_div<L>(wim(out), xb, c, S0, S1);
            }
            st_meta(out.t, tr);
            st_meta(out.t + 1, ti);
        }
    }
}

// ================================================================ Fuchs, integer operators
// Operators whose coefficients are exact real integers (Legendre: reference src/examples.c:3-11).
// The multipliers of src/fuchs_solver.c:121-135 are then exact integers too -- the reference's
// acb_mul_fmpz / acb_add produce them without rounding -- so they are formed in 64-bit integer
// arithmetic per instance, acb_mul(res[b], temp2) becomes a ball x integer product and the final
// acb_div a division by an integer: O(L) limb operations per coefficient instead of O(L^2), which
// turns the sweep into a streaming problem.  The balls are the same as the general kernel's.
// (The host checks that every multiplier stays below 2^62.)
struct FuchsIntParams {
    int order, degree, valuation;
    int64_t n, batch;
    const int64_t* ode;  // [(order+1)(degree+1)][batch] (ostride = batch) or [..][1] shared (ostride = 0)
    size_t ostride;
    CArr res;
    // persistent launch (cb_kernels.cuh): see FuchsSharedParams
    int64_t range_len;
    int* sched;
    // the product column of the general body lives in GLOBAL memory ([word][lane], L + 3 words per lane): only the
    // coefficients the register fast path hands back use it, and without it in shared memory a CTA holds 16 warps
    uint32_t* gs0;
    size_t gs0_stride;
};
constexpr int INTOP_MAXENT = 36;  // (order+1)(degree+1) <= 36

template <int L, class S0T, class SCR>
CB_HD void fuchs_intop_thread(const FuchsIntParams& p, int64_t inst, S0T S0, SCR A0, SCR A1, SCR A2,
                              int64_t n_from = -1, int64_t n_to = -1) {
    const size_t slot = 2 * (size_t)inst;
    const int v = p.valuation, D1 = p.degree + 1;
    const int ne = (p.order + 1) * D1;
    int64_t P[INTOP_MAXENT];
    uint64_t pmax = 0;
    for (int e = 0; e < ne; e++) {
        P[e] = p.ode[(size_t)e * (p.ostride ? p.ostride : 1) + (p.ostride ? (size_t)inst : 0)];
        uint64_t a = P[e] < 0 ? (uint64_t)0 - (uint64_t)P[e] : (uint64_t)P[e];
        pmax = a > pmax ? a : pmax;
    }
    // every multiplier formed below is bounded by (order+1) * max|P| * (n+1)^order; if that does not stay below
    // 2^62 the int64 arithmetic could wrap, so this instance's coefficients are made indeterminate instead of wrong
    // (the host-buffer entry point rejects such a batch up front with CB200_ERANGE)
    {
        double bound = (double)pmax * (double)(p.order + 1);
        for (int i = 0; i < p.order; i++) bound *= (double)(p.n + 1);
        if (!(bound < 4.0e18)) {
            const int64_t b0 = n_from >= 0 ? n_from : -p.valuation, b1 = n_from >= 0 ? n_to : p.n;
            for (int64_t b = b0; b <= b1; b++) {
                CRef out = at<L>(p.res, b, slot);
                store_nan<L>(wre(out));
                store_nan<L>(wim(out));
                st_meta(out.t, meta_nan());
                st_meta(out.t + 1, meta_nan());
            }
            return;
        }
    }
    SCR accRe = A0, accIm = A1, spare = A2;
    // (n_from, n_to): the coefficient range of a persistent work unit (the recurrence resumes from res)
    const int64_t b_first = n_from >= 0 ? n_from : -v, b_last = n_from >= 0 ? n_to : p.n;
#pragma unroll 1
    for (int64_t bmax = b_first; bmax <= b_last; bmax++) {
        Meta tRe = meta_zero(), tIm = meta_zero();
        int64_t e = bmax + v;
        int64_t bmin = clampi(e - p.degree, 0, bmax);
        int64_t cur = P[0 * D1 + (e - bmin)];
        int64_t b = bmin;
#pragma unroll 1
        do {
            if (!warp_all(cur == 0)) {  // x * 0 is an exact zero and acc - 0 leaves acc untouched
                CRef x = at<L>(p.res, b, slot);
                for (int part = 0; part < 2; part++) {
                    Opd xp = part ? im_of(x) : re_of(x);
                    Meta& tacc = part ? tIm : tRe;
                    SCR& acc = part ? accIm : accRe;
                    if (warp_all(is_exact_zero(xp.t))) continue;  // 0 * cur = 0
                    Meta pr = arb_mul_si<L>(WRef{spare.generic(), SCR::GSTRIDE}, xp, cur, S0);
                    if (warp_all(is_exact_zero(tacc))) {  // 0 - x = -x exactly
                        SCR t = acc;
                        acc = spare;
                        spare = t;
                        tacc = pr;
                        if (!is_nan(tacc) && !is_zero_mid(tacc)) tacc.flags ^= F_SIGN;
                    } else {
                        SCR where;
                        tacc = arb_addsub_cols<L>(acc, tacc, spare, pr, true, &where);
                        if (where.a_or_p() == spare.a_or_p()) {
                            SCR t = acc;
                            acc = spare;
                            spare = t;
                        }
                    }
                }
            }
            // next multiplier, exactly as the reference forms temp2 (all integers)
            int64_t t2 = 0, fac = 1;
            b++;
            int64_t imin = clampi(b - e, 0, -v);
            int64_t imax = clampi(b - bmin, 0, p.order);
            for (int64_t i = imin; i <= imax; i++) {
                t2 += P[i * D1 + (i + e - b)] * fac;
                fac *= (b - i);
            }
            fac = 1;
            for (int64_t q = 0; q < imin; q++) fac *= (b - imin + 1 + q);
            cur = t2 * fac;
        } while (b != bmax);
        CRef out = at<L>(p.res, bmax, slot);
        uint64_t dmag = cur < 0 ? (uint64_t)0 - (uint64_t)cur : (uint64_t)cur;
        for (int part = 0; part < 2; part++) {
            Meta tacc = part ? tIm : tRe;
            SCR acc = part ? accIm : accRe;
            WRef dst = part ? wim(out) : wre(out);
            Meta to;
            if (warp_all(is_exact_zero(tacc) && cur != 0)) {
                store_nan<L>(dst);  // zeros
                to = meta_zero();
            } else {
                Opd xa{RRef{acc.generic(), SCR::GSTRIDE}, tacc};
                to = arb_div_u64<L>(dst, xa, dmag, cur < 0, S0);
            }
            st_meta(out.t + part, to);
        }
    }
}

}  // namespace cb
#include "cb_intop.cuh"
namespace cb {

// a table row with the register fast path for its multipliers (<= 1024 bits), the general thread otherwise
template <int L, class SCR>
CB_HD void fuchs_table_fast_thread(const FuchsTableParams& p, int64_t idx, SCR S0, SCR S1) {
    if constexpr (L <= 32) {
        const int v = p.valuation, W = p.degree - v;
        const int64_t inst = idx % p.ninst, row = p.row0 + idx / p.ninst;
        const size_t oslot = (p.ode.S == 2) ? 0 : 2 * (size_t)inst, qslot = (p.tbl.S == 2) ? 0 : 2 * (size_t)inst;
        const int64_t base = (row - p.tbl_row0) * (W + 2);
        int cnt = 0;
        if (fuchs_table_fast_multipliers<L>(p, row, oslot, qslot, base, &cnt)) {
            fuchs_table_tail<L>(p, idx, row, inst, base, qslot, cnt, true, S0, S1);
            return;
        }
    }
    fuchs_table_thread<L>(p, idx, S0, S1);
}

// ================================================================ Frobenius (M = 1)
struct FrobParams {
    int order, degree, valuation;
    int64_t n, batch;
    CArr ode;  // S = 2*batch or 2
    CArr rho;  // 1 entry, S = 2*batch or 2
    CArr res;  // n+1 entries
    CArr tmp;  // FROB_NTMP entries
    CArr rhs;  // right-hand side series: rhs_len entries, S = 2*batch or 2 (ignored when rhs_len == 0)
    int64_t rhs_len;
    uint32_t* gcols;  // 2048 bits and up: global tails of the scratch columns (2 per instance; 4 per instance for the
                      // cooperative kernel, whose lanes own one column each), (L - 2) words per column
    int* sched;       // cooperative (work-unit) kernel: [0] unit counter, [1 + b] ranges completed for instance block b
    size_t gcols_words;  // size of gcols in words (the cooperative kernel needs more than the one-thread kernels)
};
constexpr int FROB_NTMP = 7;

// result may be one of (o1, o2): returns which holds f_nu(rho + shift).  IM_READY: t holds the imaginary part of rho
// already (acb_copy_im by the caller), so rho + k only writes the real part of t.  active = false: a thread without an
// instance (ragged last block of the persistent kernel) walks the same loops and meets the same phase barriers but
// touches no memory.
template <int L, bool SYNC = false, bool IM_READY = false, class SCR, class PRM>
CB_HD CRef indicial_eval(const PRM& p, size_t oslot, int64_t nu, const CRef& rho, int64_t shift,
                         CRef o1, CRef o2, const CRef& t, const CRef& kb, SCR S0, SCR S1, bool active = true) {
    const int D1 = p.degree + 1;
    nu += p.valuation;
    if (active) acb_zero<L>(o1);
    if (nu > p.degree) return o1;
#pragma unroll 1
    for (int64_t lam = clampi(p.degree - nu, 0, p.order); lam >= 0; lam--) {
        if (active) {
            if (L >= 64 && lam + nu >= 0) acb_prefetch<L>(at<L>(p.ode, lam * D1 + (lam + nu), oslot));
            acb_add_si<L, !IM_READY>(t, rho, shift - lam, kb, S0, S1);
        }
        CB_PHASE_SYNC();
        if (active) acb_mul<L>(o2, o1, t, S0, S1);
        CRef sw = o1;
        o1 = o2;
        o2 = sw;
        if (lam + nu < 0) continue;
        if (active) acb_addsub<L>(o1, o1, at<L>(p.ode, lam * D1 + (lam + nu), oslot), false, S0, S1);
    }
    return o1;
}

// SYNC: phase barriers before the long calls of the main loop (its trip counts are launch parameters; the
// right-hand-side branch before the loop is per instance and has none), see phase_sync
// nu_from .. nu_to (nu_to < 0: p.n): the coefficients this call computes -- the recurrence reads the earlier ones back
// from res, so a homogeneous solve can be cut into ranges (the persistent kernel's work units); nu_from = 0 starts the
// series (g_0, the right-hand side branch).  active: see indicial_eval (homogeneous solves only).
template <int L, bool SYNC = false, class SCR>
CB_HD void frobenius1_thread(const FrobParams& p, int64_t inst, SCR S0, SCR S1, int64_t nu_from = 0, int64_t nu_to = -1,
                             bool active = true) {
    const size_t slot = 2 * (size_t)inst;
    const size_t oslot = (p.ode.S == 2) ? 0 : slot;
    const size_t rslot = (p.rho.S == 2) ? 0 : slot;
    CRef gnew = at<L>(p.tmp, 0, slot), o1 = at<L>(p.tmp, 1, slot), o2 = at<L>(p.tmp, 2, slot),
         t = at<L>(p.tmp, 3, slot), kb = at<L>(p.tmp, 4, slot), prod = at<L>(p.tmp, 5, slot);
    CRef rho = at<L>(p.rho, 0, rslot);
    // length of the right-hand side as acb_poly sees it (trailing exact zeros do not count)
    const size_t hslot = (p.rhs_len > 0 && p.rhs.S == 2) ? 0 : slot;
    int64_t rl = p.rhs_len;
    while (rl > 0 && acb_is_exact_zero(at<L>(p.rhs, rl - 1, hslot))) rl--;
    if (nu_to < 0) nu_to = p.n;
    if (nu_from > 0) {
        // a later range of a homogeneous solve: nothing to start
    } else if (rl == 0) {
        if (active) acb_one<L>(at<L>(p.res, 0, slot));  // homogeneous case: g_0 = 1 (src/frobenius_solver.c:84-85)
    } else {
        // rho -= v; g_0 = rhs_0 / f_0(rho)   (src/frobenius_solver.c:88-95)
        CRef radj = at<L>(p.tmp, 6, slot);
        acb_add_si<L>(radj, rho, -(int64_t)p.valuation, kb, S0, S1);
        rho = radj;
        acb_copy_im<L>(t, rho);
        CRef ind = indicial_eval<L, false, true>(p, oslot, 0, rho, 0, o1, o2, t, kb, S0, S1);
        CRef spare = (ind.m == o1.m) ? o2 : o1;
        acb_div<L>(at<L>(p.res, 0, slot), at<L>(p.rhs, 0, hslot), ind, spare, S0, S1);
    }
    // every f(rho + k) below adds an integer to the same rho: its imaginary part goes into t once
    if (rl == 0 && active) acb_copy_im<L>(t, rho);
#pragma unroll 1
    for (int64_t nu = (nu_from > 1 ? nu_from : 1); nu <= nu_to; nu++) {
        if (active) {
            if (nu < rl) acb_set<L>(gnew, at<L>(p.rhs, nu, hslot));
            else acb_zero<L>(gnew);
        }
        int64_t i = clampi(nu, 1, p.degree);
        if (L >= 64 && active) acb_prefetch<L>(at<L>(p.res, nu - i, slot));
        CRef ind = indicial_eval<L, SYNC, true>(p, oslot, i, rho, nu - i, o1, o2, t, kb, S0, S1, active);
#pragma unroll 1
        do {
            CB_PHASE_SYNC();
            if (active) acb_mul<L>(prod, ind, at<L>(p.res, nu - i, slot), S0, S1);
            CB_PHASE_SYNC();
            if (active) acb_addsub<L>(gnew, gnew, prod, true, S0, S1);
            i--;
            if (L >= 64 && i > 0 && active) acb_prefetch<L>(at<L>(p.res, nu - i, slot));
            ind = indicial_eval<L, SYNC, true>(p, oslot, i, rho, nu - i, o1, o2, t, kb, S0, S1, active);
        } while (i > 0);
        // g_nu = g_new / f_0(rho + nu); `prod` and the unused one of (o1,o2) are free temporaries
        CRef spare = (ind.m == o1.m) ? o2 : o1;
        CB_PHASE_SYNC();
        if (active) acb_div<L>(at<L>(p.res, nu, slot), gnew, ind, spare, S0, S1);
    }
}

// ================================================================ Frobenius, general (M > 1)
// acb_ode_solve_frobenius for M = mul + alpha > 1 (reference src/frobenius_solver.c:131-205):
// the recurrence carried as polynomials in rho, with indicial_polynomial (:4-39) and the solution
// bookkeeping _acb_ode_solution_update / _extend / _normalize (src/acb_ode_solution.c:60-123).
// One thread per instance; every polynomial in rho is a run of `plen` entries of the workspace
// array plus a length held by the thread, with Arb's conventions: the length is normalised
// (trailing EXACT zeros stripped) wherever Arb normalises, and reads beyond it give zero.
// The Arb sub-algorithms behind acb_poly_* are the ones restated in oracle/cbo_cascade.c
// (classical product, longer operand first; Horner evaluation; coefficientwise scalar ops).
constexpr int FROBG_MAXM = 16, FROBG_MAXD = 8;
constexpr int FROBG_NTMP = 6;  // t1, t2, kb, inv, tmp, f0
struct FrobGenParams {
    int order, degree, valuation, M, mul;
    int64_t n, batch;
    int64_t plen;  // capacity of a polynomial in rho: order * n + order + 2
    CArr ode;      // S = 2*batch or 2
    CArr rho;      // 1 entry, S = 2*batch or 2
    CArr gens;     // M * (n + 1) entries (gens[i][j] at i * (n + 1) + j), S = 2*batch: all written
    CArr wp;       // (degree + 3) * plen + FROBG_NTMP + M entries, S = 2*batch
};

// exact |mid| <= rad (arb_contains_zero); an indeterminate ball contains everything
template <int L>
CB_HD bool arb_contains_zero(const uint32_t* m, size_t stride, const Meta& t) {
    if (t.flags & (F_NAN | F_ZERO)) return true;
    if (t.rman == 0) return false;
    if (t.exp != t.rexp) return t.exp < t.rexp;
    uint32_t top = m[(size_t)(L - 1) * stride];
    if ((top >> 2) != t.rman) return (top >> 2) < t.rman;
    if (top & 3u) return false;
    for (int k = 0; k < L - 1; k++)
        if (m[(size_t)k * stride]) return false;
    return true;
}
template <int L>
CB_HD bool acb_contains_zero(const CRef& x) {
    return arb_contains_zero<L>(x.m, x.stride, ld_meta(x.t)) && arb_contains_zero<L>(x.m + 1, x.stride, ld_meta(x.t + 1));
}
CB_HD void acb_neg_inplace(const CRef& z) {
    for (int part = 0; part < 2; part++) {
        Meta t = ld_meta(z.t + part);
        if (!(t.flags & (F_ZERO | F_NAN))) {
            t.flags ^= F_SIGN;
            st_meta(z.t + part, t);
        }
    }
}
// z = x / k for an integer k (acb_div_si as acb_div by the exact ball k); z may alias x
template <int L, class SCR>
CB_HD void acb_div_si(const CRef& z, const CRef& x, int64_t k, SCR S0) {
    uint64_t ka = k < 0 ? (uint64_t)0 - (uint64_t)k : (uint64_t)k;
    Meta tr = arb_div_u64<L>(wre(z), re_of(x), ka, k < 0, S0);
    Meta ti = arb_div_u64<L>(wim(z), im_of(x), ka, k < 0, S0);
    st_meta(z.t, tr);
    st_meta(z.t + 1, ti);
}

// the polynomial workspace of one instance
template <int L, class SCR>
struct PolyCtx {
    CArr A;
    size_t slot;
    SCR S0, S1;
    CRef t1, t2, kb, inv, tmp;
    CB_HD CRef c(int64_t e0, int64_t j) const { return at<L>(A, e0 + j, slot); }
    CB_HD int64_t norm(int64_t e0, int64_t len) const {
        while (len > 0 && acb_is_exact_zero(c(e0, len - 1))) len--;
        return len;
    }
};

// acb_poly_mul(dst, a, b), dst distinct from both; returns the normalised length
template <int L, class SCR>
CB_HD int64_t poly_mul(const PolyCtx<L, SCR>& pc, int64_t eD, int64_t eA, int64_t la, int64_t eB, int64_t lb) {
    if (la == 0 || lb == 0) return 0;
    if (la < lb) {
        int64_t t = eA; eA = eB; eB = t;
        t = la; la = lb; lb = t;
    }
#pragma unroll 1
    for (int64_t i = 0; i < la; i++) acb_mul<L>(pc.c(eD, i), pc.c(eA, i), pc.c(eB, 0), pc.S0, pc.S1);
#pragma unroll 1
    for (int64_t j = 1; j < lb; j++) acb_mul<L>(pc.c(eD, la - 1 + j), pc.c(eA, la - 1), pc.c(eB, j), pc.S0, pc.S1);
#pragma unroll 1
    for (int64_t i = 0; i < la - 1; i++)
#pragma unroll 1
        for (int64_t j = 1; j < lb; j++) {
            acb_mul<L>(pc.tmp, pc.c(eA, i), pc.c(eB, j), pc.S0, pc.S1);
            acb_addsub<L>(pc.c(eD, i + j), pc.c(eD, i + j), pc.tmp, false, pc.S0, pc.S1);
        }
    return pc.norm(eD, la + lb - 1);
}
// a -= b in place (acb_poly_sub(a, a, b)); returns the new length of a
template <int L, class SCR>
CB_HD int64_t poly_sub_inplace(const PolyCtx<L, SCR>& pc, int64_t eA, int64_t la, int64_t eB, int64_t lb) {
    int64_t mn = la < lb ? la : lb, mx = la < lb ? lb : la;
#pragma unroll 1
    for (int64_t i = 0; i < mn; i++) acb_addsub<L>(pc.c(eA, i), pc.c(eA, i), pc.c(eB, i), true, pc.S0, pc.S1);
#pragma unroll 1
    for (int64_t i = mn; i < lb; i++) {
        acb_set<L>(pc.c(eA, i), pc.c(eB, i));
        acb_neg_inplace(pc.c(eA, i));
    }
    return pc.norm(eA, mx);
}
// p /= y coefficientwise (acb_poly_scalar_div): acb_div(x, y) is x * acb_inv(y) for a complex y and
// the inverse is the same ball every time, so it is formed once
template <int L, class SCR>
CB_HD int64_t poly_scalar_div(const PolyCtx<L, SCR>& pc, int64_t e0, int64_t len, const CRef& y) {
    if (len == 0) return 0;
    bool real = is_exact_zero(ld_meta(y.t + 1));
    if (!real) acb_inv<L>(pc.inv, y, pc.t1, pc.S0, pc.S1);
#pragma unroll 1
    for (int64_t i = 0; i < len; i++) {
        CRef x = pc.c(e0, i);
        if (real) {
            Opd c = re_of(y);
            Meta tr = arb_div<L>(wre(pc.t1), re_of(x), c, pc.S0, pc.S1);
            Meta ti = arb_div<L>(wim(pc.t1), im_of(x), c, pc.S0, pc.S1);
            st_meta(pc.t1.t, tr);
            st_meta(pc.t1.t + 1, ti);
        } else {
            acb_mul<L>(pc.t1, x, pc.inv, pc.S0, pc.S1);
        }
        acb_set<L>(x, pc.t1);
    }
    return pc.norm(e0, len);
}
// p *= y coefficientwise (acb_poly_scalar_mul), in place
template <int L, class SCR>
CB_HD int64_t poly_scalar_mul(const PolyCtx<L, SCR>& pc, const CArr& A, int64_t e0, int64_t len, const CRef& y) {
#pragma unroll 1
    for (int64_t i = 0; i < len; i++) {
        CRef x = at<L>(A, e0 + i, pc.slot);
        acb_mul<L>(pc.t1, x, y, pc.S0, pc.S1);
        acb_set<L>(x, pc.t1);
    }
    while (len > 0 && acb_is_exact_zero(at<L>(A, e0 + len - 1, pc.slot))) len--;
    return len;
}
// out = p(x) (acb_poly_evaluate -> Horner; out must not be t1).  SYNC: phase barriers in the Horner loop -- only for
// callers whose polynomial length is the same for every thread of the block (see phase_sync)
template <int L, bool SYNC = false, class SCR>
CB_HD void poly_eval(const PolyCtx<L, SCR>& pc, const CRef& out, int64_t e0, int64_t len, const CRef& x) {
    if (len == 0) {
        acb_zero<L>(out);
        return;
    }
    acb_set<L>(out, pc.c(e0, len - 1));
    if (len == 1 || acb_is_exact_zero(x)) {
        if (len > 1) acb_set<L>(out, pc.c(e0, 0));
        return;
    }
#pragma unroll 1
    for (int64_t j = len - 2; j >= 0; j--) {
        CB_PHASE_SYNC();
        acb_mul<L>(pc.t1, out, x, pc.S0, pc.S1);
        CB_PHASE_SYNC();
        acb_addsub<L>(out, pc.t1, pc.c(e0, j), false, pc.S0, pc.S1);
    }
}
// p = p' in place (acb_poly_derivative); returns the new length
template <int L, class SCR>
CB_HD int64_t poly_derivative(const PolyCtx<L, SCR>& pc, int64_t e0, int64_t len) {
#pragma unroll 1
    for (int64_t i = 1; i < len; i++) acb_mul_si<L>(pc.c(e0, i - 1), pc.c(e0, i), i, pc.S0);
    return pc.norm(e0, len > 0 ? len - 1 : 0);
}
// indicial_polynomial (src/frobenius_solver.c:4-39) into the run at eR; returns its length
template <int L, class SCR, class PRM>
CB_HD int64_t indicial_poly(const PolyCtx<L, SCR>& pc, const PRM& p, size_t oslot, int64_t eR, int64_t nu, int64_t shift) {
    const int D1 = p.degree + 1;
    int64_t len = 0;
    nu += p.valuation;
    if (nu > p.degree) return 0;
    // entries at and beyond the length are kept as exact zeros, so "reads as zero" is a plain read
    for (int j = 0; j <= p.order + 1; j++) acb_zero<L>(pc.c(eR, j));
#pragma unroll 1
    for (int64_t lam = clampi(p.degree - nu, 0, p.order); lam >= 0; lam--) {
#pragma unroll 1
        for (int64_t j = len; j > 0; j--) {
            // result[j] = result[j-1] + result[j] * (shift - lam)
            acb_mul_si<L>(pc.t1, pc.c(eR, j), shift - lam, pc.S0);
            acb_addsub<L>(pc.c(eR, j), pc.c(eR, j - 1), pc.t1, false, pc.S0, pc.S1);
            if (j + 1 > len) len = j + 1;
            len = pc.norm(eR, len);
        }
        CRef r0 = pc.c(eR, 0);
        acb_mul_si<L>(r0, r0, shift - lam, pc.S0);
        if (lam + nu >= 0) acb_addsub<L>(r0, r0, at<L>(p.ode, lam * D1 + (lam + nu), oslot), false, pc.S0, pc.S1);
        if (len == 0) len = 1;
        len = pc.norm(eR, len);
    }
    return len;
}

// the solution bookkeeping of one instance: gens[i] is the run i*cap .. i*cap+cap-1 of G
template <int L, class SCR>
struct SolCtx {
    CArr G;
    int64_t cap;
    int M, mul;
    CRef rho;
    int64_t glen[FROBG_MAXM];
    CB_HD CRef g(const PolyCtx<L, SCR>& pc, int i, int64_t j) const { return at<L>(G, (int64_t)i * cap + j, pc.slot); }
};
// _acb_ode_solution_extend (src/acb_ode_solution.c:97-109): consumes the polynomial at eG
template <int L, class SCR>
CB_HD void solution_extend(const PolyCtx<L, SCR>& pc, SolCtx<L, SCR>& sc, int64_t nu, int64_t eG, int64_t lenG) {
#pragma unroll 1
    for (int i = 0; i < sc.M; i++) {
        poly_eval<L>(pc, pc.t2, eG, lenG, sc.rho);
        // acb_poly_set_coeff_acb(gens + i, nu, t2)
        bool z = acb_is_exact_zero(pc.t2);
        if (nu < sc.glen[i] || !z) {
            acb_set<L>(sc.g(pc, i, nu), pc.t2);
            if (nu + 1 > sc.glen[i]) sc.glen[i] = nu + 1;
            while (sc.glen[i] > 0 && acb_is_exact_zero(sc.g(pc, i, sc.glen[i] - 1))) sc.glen[i]--;
        }
        lenG = poly_derivative<L>(pc, eG, lenG);
    }
}
// _acb_ode_solution_update (:60-95), second half: the Leibniz combination with F[k] = C(M-1,k) f^(k)(rho)
// held in the M balls at eFs of pc.A
template <int L, class SCR>
CB_HD void solution_update_apply(const PolyCtx<L, SCR>& pc, SolCtx<L, SCR>& sc, int64_t eFs) {
#pragma unroll 1
    for (int n = sc.M - 1; n >= 0; n--) {
        sc.glen[n] = poly_scalar_mul<L>(pc, sc.G, (int64_t)n * sc.cap, sc.glen[n], pc.c(eFs, 0));
#pragma unroll 1
        for (int k = 1; k <= n; k++) {
            CRef Fk = pc.c(eFs, k);
            // f = gens[n-k] * F[k]; gens[n] += f  (coefficientwise, so fused per coefficient)
            int64_t lf = sc.glen[n - k], lg = sc.glen[n];
            // (f's trailing coefficients that come out exactly zero are not part of f; adding an
            // exact zero leaves gens[n] unchanged, so only the length bookkeeping sees the difference)
            int64_t top = -1;
#pragma unroll 1
            for (int64_t j = 0; j < lf; j++) {
                acb_mul<L>(pc.t1, sc.g(pc, n - k, j), Fk, pc.S0, pc.S1);
                bool fz = acb_is_exact_zero(pc.t1);
                if (!fz) top = j;
                if (j < lg) {
                    acb_addsub<L>(sc.g(pc, n, j), sc.g(pc, n, j), pc.t1, false, pc.S0, pc.S1);
                } else {
                    acb_set<L>(sc.g(pc, n, j), pc.t1);
                }
            }
            int64_t lfn = top + 1;
            int64_t mx = lg > lfn ? lg : lfn;
            while (mx > 0 && acb_is_exact_zero(sc.g(pc, n, mx - 1))) mx--;
            sc.glen[n] = mx;
            acb_mul_si<L>(Fk, Fk, n - k, pc.S0);
            acb_div_si<L>(Fk, Fk, n, pc.S0);
        }
    }
}
// _acb_ode_solution_update (:60-95): consumes the polynomial at eF; F: M scratch balls at eFs
template <int L, class SCR>
CB_HD void solution_update(const PolyCtx<L, SCR>& pc, SolCtx<L, SCR>& sc, int64_t eF, int64_t lenF, int64_t eFs) {
    int64_t binom = 1;
#pragma unroll 1
    for (int k = 0; k < sc.M; k++) {
        CRef Fk = pc.c(eFs, k);
        poly_eval<L>(pc, Fk, eF, lenF, sc.rho);
        lenF = poly_derivative<L>(pc, eF, lenF);
        acb_mul_si<L>(Fk, Fk, binom, pc.S0);
        binom = (binom * (sc.M - 1 - k)) / (k + 1);
    }
    solution_update_apply<L>(pc, sc, eFs);
}
// _acb_ode_solution_normalize (:111-123)
template <int L, class SCR>
CB_HD void solution_normalize(const PolyCtx<L, SCR>& pc, SolCtx<L, SCR>& sc) {
    int alpha = sc.M - sc.mul;
    if (sc.glen[alpha] > 0) acb_set<L>(pc.t2, sc.g(pc, alpha, 0));
    else acb_zero<L>(pc.t2);
    PolyCtx<L, SCR> gc = pc;
    gc.A = sc.G;
#pragma unroll 1
    for (int i = 0; i < sc.M; i++) sc.glen[i] = poly_scalar_div<L>(gc, (int64_t)i * sc.cap, sc.glen[i], pc.t2);
}

template <int L, class SCR>
CB_HD void frobenius_general_thread(const FrobGenParams& p, int64_t inst, SCR S0, SCR S1) {
    const size_t slot = 2 * (size_t)inst;
    const size_t oslot = (p.ode.S == 2) ? 0 : slot;
    const int d = p.degree;
    const int64_t tb = (int64_t)(d + 3) * p.plen;  // first temporary
    PolyCtx<L, SCR> pc{p.wp, slot, S0, S1, at<L>(p.wp, tb + 0, slot), at<L>(p.wp, tb + 1, slot),
                       at<L>(p.wp, tb + 2, slot), at<L>(p.wp, tb + 3, slot), at<L>(p.wp, tb + 4, slot)};
    CRef f0 = at<L>(p.wp, tb + 5, slot);
    const int64_t eFs = tb + FROBG_NTMP;
    SolCtx<L, SCR> sc;
    sc.G = p.gens;
    sc.cap = p.n + 1;
    sc.M = p.M;
    sc.mul = p.mul;
    sc.rho = at<L>(p.rho, 0, (p.rho.S == 2) ? 0 : slot);
    int64_t eInd = 0, eNew = p.plen, eProd = 2 * p.plen, lInd = 0, lNew = 0;
    int64_t eRho[FROBG_MAXD], lRho[FROBG_MAXD];
    for (int i = 0; i < d; i++) {
        eRho[i] = (int64_t)(3 + i) * p.plen;
        lRho[i] = 0;
    }
    acb_one<L>(pc.c(eRho[0], 0));
    lRho[0] = 1;
    for (int i = 0; i < p.M; i++) {
        sc.glen[i] = 0;
#pragma unroll 1
        for (int64_t j = 0; j <= p.n; j++) acb_zero<L>(sc.g(pc, i, j));
    }
    acb_one<L>(sc.g(pc, 0, 0));
    sc.glen[0] = 1;
#pragma unroll 1
    for (int64_t nu = 1; nu <= p.n; nu++) {
        int64_t i = clampi(nu, 1, d);
        lInd = indicial_poly<L>(pc, p, oslot, eInd, i, nu - i);
#pragma unroll 1
        do {
            int64_t lp = poly_mul<L>(pc, eProd, eInd, lInd, eRho[i - 1], lRho[i - 1]);
            {  // acb_poly_mul(indicial, indicial, ...) lands in indicial
                int64_t t = eInd; eInd = eProd; eProd = t;
                lInd = lp;
            }
            lNew = poly_sub_inplace<L>(pc, eNew, lNew, eInd, lInd);
            i--;
            lInd = indicial_poly<L>(pc, p, oslot, eInd, i, nu - i);
        } while (i > 0);
        // rescale by f_0(rho + nu) (:170-176)
        {
            CRef o1 = pc.t2, o2 = pc.tmp;
            CRef r = indicial_eval<L>(p, oslot, 0, sc.rho, nu, o1, o2, pc.t1, pc.kb, S0, S1);
            acb_set<L>(f0, r);
        }
        if (!acb_contains_zero<L>(f0)) {
            lInd = poly_scalar_div<L>(pc, eInd, lInd, f0);
            lNew = poly_scalar_div<L>(pc, eNew, lNew, f0);
        }
        bool all_zero = lNew == 0;
#pragma unroll 1
        for (int k = d - 1; k > 0; k--) {
            lRho[k] = poly_mul<L>(pc, eRho[k], eRho[k - 1], lRho[k - 1], eInd, lInd);
            all_zero = all_zero && lRho[k] == 0;
        }
#pragma unroll 1
        for (int64_t j = 0; j < lNew; j++) acb_set<L>(pc.c(eRho[0], j), pc.c(eNew, j));
        lRho[0] = lNew;
        solution_update<L>(pc, sc, eInd, lInd, eFs);
        lInd = 0;
        solution_extend<L>(pc, sc, nu, eNew, lNew);
        lNew = 0;
        if (all_zero) break;
    }
    solution_normalize<L>(pc, sc);
}

// ---- the bookkeeping steps and indicial_polynomial as entry points of their own (the reference
// exports them: src/acb_ode.h:58-60, src/cascade.h:22).  Polynomials cross the interface as
// fixed-capacity coefficient arrays; their length is implied (trailing exact zeros do not count).
enum { SOLOP_EXTEND = 0, SOLOP_UPDATE = 1, SOLOP_NORMALIZE = 2 };
struct SolOpParams {
    int which, M, mul;
    int64_t batch, cap, fcap, nu;
    CArr rho;   // 1 entry, S = 2*batch or 2
    CArr gens;  // M * cap entries, S = 2*batch, in/out
    CArr f;     // fcap entries, S = 2*batch, in/out (consumed: comes back zero)
    CArr wp;    // FROBG_NTMP + M entries, S = 2*batch
};
template <int L, class SCR>
CB_HD void solution_op_thread(const SolOpParams& p, int64_t inst, SCR S0, SCR S1) {
    const size_t slot = 2 * (size_t)inst;
    PolyCtx<L, SCR> pc{p.f, slot, S0, S1, at<L>(p.wp, 0, slot), at<L>(p.wp, 1, slot), at<L>(p.wp, 2, slot),
                       at<L>(p.wp, 3, slot), at<L>(p.wp, 4, slot)};
    SolCtx<L, SCR> sc;
    sc.G = p.gens;
    sc.cap = p.cap;
    sc.M = p.M;
    sc.mul = p.mul;
    sc.rho = at<L>(p.rho, 0, (p.rho.S == 2) ? 0 : slot);
    for (int i = 0; i < p.M; i++) {
        int64_t len = p.cap;
        while (len > 0 && acb_is_exact_zero(sc.g(pc, i, len - 1))) len--;
        sc.glen[i] = len;
    }
    int64_t lf = 0;
    if (p.which != SOLOP_NORMALIZE) lf = pc.norm(0, p.fcap);
    if (p.which == SOLOP_EXTEND) {
        solution_extend<L>(pc, sc, p.nu, 0, lf);
    } else if (p.which == SOLOP_UPDATE) {
        // F[k] = C(M-1,k) f^(k)(rho) go behind the temporaries of the workspace array
        int64_t binom = 1;
#pragma unroll 1
        for (int k = 0; k < sc.M; k++) {
            CRef Fk = at<L>(p.wp, FROBG_NTMP + k, slot);
            poly_eval<L>(pc, Fk, 0, lf, sc.rho);
            lf = poly_derivative<L>(pc, 0, lf);
            acb_mul_si<L>(Fk, Fk, binom, S0);
            binom = (binom * (sc.M - 1 - k)) / (k + 1);
        }
        PolyCtx<L, SCR> wc = pc;
        wc.A = p.wp;
        solution_update_apply<L>(wc, sc, FROBG_NTMP);
    } else {
        solution_normalize<L>(pc, sc);
    }
    if (p.which != SOLOP_NORMALIZE) {  // acb_poly_zero(f)
#pragma unroll 1
        for (int64_t j = 0; j < p.fcap; j++) acb_zero<L>(pc.c(0, j));
    }
}

struct IndicialParams {
    int order, degree, valuation;
    int64_t batch, nu, shift;
    CArr ode;  // S = 2*batch or 2
    CArr out;  // order + 2 entries, S = 2*batch (the last one is scratch and comes back zero)
    CArr wp;   // FROBG_NTMP entries
};
template <int L, class SCR>
CB_HD void indicial_poly_thread(const IndicialParams& p, int64_t inst, SCR S0, SCR S1) {
    const size_t slot = 2 * (size_t)inst;
    PolyCtx<L, SCR> pc{p.out, slot, S0, S1, at<L>(p.wp, 0, slot), at<L>(p.wp, 1, slot), at<L>(p.wp, 2, slot),
                       at<L>(p.wp, 3, slot), at<L>(p.wp, 4, slot)};
    for (int j = 0; j <= p.order + 1; j++) acb_zero<L>(pc.c(0, j));
    indicial_poly<L>(pc, p, (p.ode.S == 2) ? 0 : slot, 0, p.nu, p.shift);
}

// ================================================================ acb_ode_shift (reference src/acb_ode.c:124-135)
// out = in with every row p_i(z) replaced by p_i(z + a): copy, then the Horner Taylor shift of each row
//   for i = D-2 .. 0: for j = i .. D-2: p[j] += p[j+1] * a        (Arb's _acb_poly_taylor_shift_horner: acb_addmul)
// One thread per operator; an exact-zero a (or degree 0) leaves the copy (oracle: cbo_ode_shift).
struct ShiftParams {
    int order, degree;
    int64_t batch;
    CArr in;   // (order+1)(degree+1) entries, S = 2*batch or 2 (one operator shifted to `batch` points)
    CArr a;    // 1 entry, S = 2*batch or 2
    CArr out;  // S = 2*batch
    CArr tmp;  // 1 entry, S = 2*batch
};
template <int L, class SCR>
CB_HD void ode_shift_thread(const ShiftParams& p, int64_t inst, SCR S0, SCR S1) {
    const size_t slot = 2 * (size_t)inst;
    const size_t islot = (p.in.S == 2) ? 0 : slot, aslot = (p.a.S == 2) ? 0 : slot;
    const int D1 = p.degree + 1, ne = (p.order + 1) * D1;
    for (int e = 0; e < ne; e++) acb_set<L>(at<L>(p.out, e, slot), at<L>(p.in, e, islot));
    CRef a = at<L>(p.a, 0, aslot), t = at<L>(p.tmp, 0, slot);
    if (p.degree == 0 || acb_is_exact_zero(a)) return;
#pragma unroll 1
    for (int row = 0; row <= p.order; row++)
#pragma unroll 1
        for (int i = D1 - 2; i >= 0; i--)
#pragma unroll 1
            for (int j = i; j <= D1 - 2; j++) {
                CRef pj = at<L>(p.out, row * D1 + j, slot);
                acb_mul<L>(t, at<L>(p.out, row * D1 + j + 1, slot), a, S0, S1);
                acb_addsub<L>(pj, pj, t, false, S0, S1);
            }
}

// ================================================================ acb_poly_taylor_shift (reference src/fuchs_solver.c:159)
// The complete re-expansion f(x + a) of a series, in the Horner form (oracle: cbo_poly_taylor_shift_horner)
//   for i = len-2 .. 0: for j = i .. len-2: f[j] += f[j+1] * a
// Inside one sweep i every update reads the value f[j+1] had BEFORE the sweep, so a sweep is two parallel phases
// over j: tmp[j] = f[j+1] * a, then f[j] += tmp[j].  One CTA per series runs the sweeps with a barrier between the
// phases; the entries of a phase are dealt out to its threads.
struct TShiftParams {
    int64_t len, batch;
    CArr f;    // len entries, S = 2*batch, in place
    CArr a;    // 1 entry, S = 2*batch or 2
    CArr tmp;  // len entries, S = 2*batch
};
template <int L, class SCR>
CB_HD void taylor_shift_mul(const TShiftParams& p, int64_t inst, int64_t j, SCR S0, SCR S1) {
    const size_t slot = 2 * (size_t)inst;
    acb_mul<L>(at<L>(p.tmp, j, slot), at<L>(p.f, j + 1, slot), at<L>(p.a, 0, (p.a.S == 2) ? 0 : slot), S0, S1);
}
template <int L, class SCR>
CB_HD void taylor_shift_add(const TShiftParams& p, int64_t inst, int64_t j, SCR S0, SCR S1) {
    const size_t slot = 2 * (size_t)inst;
    CRef fj = at<L>(p.f, j, slot);
    acb_addsub<L>(fj, fj, at<L>(p.tmp, j, slot), false, S0, S1);
}

// ================================================================ Horner
struct HornerParams {
    int nder;       // number of Taylor coefficients wanted (1 = plain evaluation)
    int64_t len;    // series length
    int64_t npts;   // evaluation points == threads
    CArr f;         // len entries; S = 2 (one series for all points) or 2*npts
    CArr pts;       // 1 entry, S = 2*npts, or S = 2: one point for all series (continuation step)
    CArr out;       // nder entries, S = 2*npts
    CArr tmp;       // 1 entry, S = 2*npts
    uint32_t* gcols;  // 2048 bits and up: global tails of the scratch columns, (L - 2) words per column: 2 columns per
                      // point for the one-thread kernels, 4 (one per lane) for the cooperative kernel
    size_t gcols_words;
};
constexpr int HORNER_MAXDER = 4;

// SYNC: phase barriers before the long calls (every thread of the block runs the same loop: len and nder are
// launch parameters), see phase_sync
template <int L, bool SYNC = false, class SCR>
CB_HD void horner_thread(const HornerParams& p, int64_t pt, SCR S0, SCR S1) {
    const size_t slot = 2 * (size_t)pt;
    const size_t fslot = (p.f.S == 2) ? 0 : slot;
    CRef a = at<L>(p.pts, 0, (p.pts.S == 2) ? 0 : slot), tmp = at<L>(p.tmp, 0, slot);
    for (int k = 0; k < p.nder; k++) acb_zero<L>(at<L>(p.out, k, slot));
    if (p.len <= 0) return;
    acb_set<L>(at<L>(p.out, 0, slot), at<L>(p.f, p.len - 1, fslot));
#pragma unroll 1
    for (int64_t j = p.len - 2; j >= 0; j--) {
#pragma unroll 1
        for (int k = p.nder - 1; k >= 0; k--) {
            CRef o = at<L>(p.out, k, slot);
            CB_PHASE_SYNC();
            acb_mul<L>(tmp, o, a, S0, S1);
            CB_PHASE_SYNC();
            if (k > 0)
                acb_addsub<L>(o, tmp, at<L>(p.out, k - 1, slot), false, S0, S1);
            else
                acb_addsub<L>(o, tmp, at<L>(p.f, j, fslot), false, S0, S1);
        }
    }
}

}  // namespace cb
#include "cb_coop.cuh"
namespace cb {

// ================================================================ elementwise (unit-test surface)
enum { OP_ADD = 0, OP_SUB = 1, OP_MUL = 2, OP_DIV = 3, OP_INV = 4, OP_MUL_SI = 16, OP_ADD_SI = 17 };
struct EwParams {
    int op;
    int64_t n;
    CArr z, x, y;  // 1 entry each, S = 2*n
    CArr tmp;      // 2 entries
    const int64_t* k;
};
template <int L, class SCR>
CB_HD void ew_thread(const EwParams& p, int64_t i, SCR S0, SCR S1) {
    const size_t slot = 2 * (size_t)i;
    CRef z = at<L>(p.z, 0, slot), x = at<L>(p.x, 0, slot), y = at<L>(p.y, 0, slot);
    CRef t0 = at<L>(p.tmp, 0, slot), t1 = at<L>(p.tmp, 1, slot);
    switch (p.op) {
        case OP_ADD: acb_addsub<L>(z, x, y, false, S0, S1); break;
        case OP_SUB: acb_addsub<L>(z, x, y, true, S0, S1); break;
        case OP_MUL: acb_mul<L>(z, x, y, S0, S1); break;
        case OP_DIV: acb_div<L>(z, x, y, t0, S0, S1); break;
        case OP_INV: acb_inv<L>(z, x, t0, S0, S1); break;
        case OP_MUL_SI: acb_mul_si<L>(z, x, p.k[i], S0); break;
        case OP_ADD_SI: acb_add_si<L>(z, x, p.k[i], t1, S0, S1); break;
        default: break;
    }
}

}  // namespace cb

#include "cb_trans.cuh"
