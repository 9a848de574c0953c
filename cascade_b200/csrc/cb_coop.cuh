// cb_coop.cuh -- cooperating lanes: FOUR lanes per ODE instance for 2048 / 4096 bits and small batches.
//
// One thread per instance (cb_solvers.cuh) leaves a 4096-bit sweep of 16 384 instances (BASELINE configs[2]) with
// 111 threads per SM -- less than one warp per scheduler, every global-memory and carry-chain latency exposed (ncu:
// issue slots 10 % busy, profiles/ncu_frobenius1_128_r02_before.txt).  Here the four real products of a complex
// multiplication (reference acb_mul at src/frobenius_solver.c:59,107: re = a c - b d, im = a d + b c) run on four
// lanes, each into its own shared-memory column; lanes 0 and 2 then align / add / round the real and the imaginary
// part (the tail of arb_fdot2_trunc_blocked) side by side, and every componentwise operation (acb_add, acb_sub,
// acb_mul_si, the two long divisions of acb_inv / a real-divisor acb_div) runs its real part on lane 0 and its
// imaginary part on lane 2.  Same operations on the same operands in the same order as the one-thread code:
// results are BIT-IDENTICAL.  Lanes of a quad communicate through memory and __syncwarp on the quad's mask only, so
// quads of a warp may diverge (per-instance branches such as "is the divisor real?").
//
// The same source compiles for the host simulation: CB_COOP_EACH runs the phase of every lane of the quad one after
// the other there (phases are separated by quad barriers on the device, by sequence on the host).
#pragma once

namespace cb {

#if defined(__CUDA_ARCH__)
template <int TPB, int KS>
struct Quad {
    using Col = HybridScratch<TPB, KS>;
    int q;            // this lane's index in its quad
    unsigned qmask;   // the four lanes of the quad
    uint32_t a;       // own column: shared-space byte address of word 0
    uint32_t* g;      // own column: global tail
    uint32_t gs;
    __device__ __forceinline__ Col col(int qq) const { return Col{a + 4u * (uint32_t)(qq - q), g + (qq - q), gs}; }
    __device__ __forceinline__ void sync() const { __syncwarp(qmask); }
};
#define CB_COOP_EACH(Q, qv) for (int qv = (Q).q, qv##_go = 1; qv##_go; qv##_go = 0)
#else
template <int TPB, int KS>
struct Quad {
    using Col = HybridScratch<TPB, KS>;
    uint32_t* p[4];
    uint32_t* g[4];
    Col col(int qq) const { return Col{p[qq], g[qq], 1u}; }
    void sync() const {}
};
#define CB_COOP_EACH(Q, qv) for (int qv = 0; qv < 4; qv++)
#endif

// ---- copies and constants, limbs dealt out to the four lanes
template <int L, class QD>
CB_HD void coop_set(const QD& Q, const CRef& z, const CRef& x) {
    CB_COOP_EACH(Q, q) {
        const int part = q & 1;
#pragma unroll 8
        for (int k = (q >> 1); k < L; k += 2) z.m[k * z.stride + part] = x.m[k * x.stride + part];
        if (q < 2) st_meta(z.t + q, ld_meta(x.t + q));
    }
    Q.sync();
}
template <int L, class QD>
CB_HD void coop_zero(const QD& Q, const CRef& z) {
    CB_COOP_EACH(Q, q) {
        const int part = q & 1;
#pragma unroll 8
        for (int k = (q >> 1); k < L; k += 2) z.m[k * z.stride + part] = 0u;
        if (q < 2) st_meta(z.t + q, meta_zero());
    }
    Q.sync();
}
template <int L, class QD>
CB_HD void coop_one(const QD& Q, const CRef& z) {
    CB_COOP_EACH(Q, q) {
        const int part = q & 1;
#pragma unroll 8
        for (int k = (q >> 1); k < L; k += 2) z.m[k * z.stride + part] = (part == 0 && k == L - 1) ? 0x80000000u : 0u;
        if (q == 0) st_meta(z.t, Meta{1, 0u, 0u, 0});
        if (q == 1) st_meta(z.t + 1, meta_zero());
    }
    Q.sync();
}

// one fused dot product pair: lane 2h forms P1 = x1*y1, lane 2h+1 forms P2 = x2*y2 (h = 0: the pair given first,
// h = 1: the second), then lane 2h rounds x1*y1 +- x2*y2 into its destination.  Pairs with `have[h] == false` idle.
template <int L>
struct FdotJob {
    Opd x1, y1, x2, y2;
    bool neg2;
    WRef dst;
    Meta* out;  // where the meta record of the result goes
};
template <int L, class QD>
CB_HD void coop_fdot2_pair(const QD& Q, const FdotJob<L>& j0, const FdotJob<L>& j1, bool have1) {
    CB_COOP_EACH(Q, q) {  // the products
        const FdotJob<L>& j = (q >> 1) ? j1 : j0;
        if ((q >> 1) && !have1) continue;
        const Opd& xo = (q & 1) ? j.x2 : j.x1;
        const Opd& yo = (q & 1) ? j.y2 : j.y1;
        if (!((xo.t.flags | yo.t.flags) & (F_ZERO | F_NAN))) product_trunc_blocked<L>(lo_view(Q.col(q)), xo.r, yo.r);
    }
    Q.sync();
    CB_COOP_EACH(Q, q) {  // align, add, round
        if (q & 1) continue;
        if ((q >> 1) && !have1) continue;
        const FdotJob<L>& j = (q >> 1) ? j1 : j0;
        Opd x1 = j.x1, y1 = j.y1, x2 = j.x2, y2 = j.y2;
        // short (integer-like) operands make exact products, whose guard bits prove nothing: knowing their trailing
        // zero limbs lets the truncated path accept them as exact (as fdot2_auto does)
        tz_scan<L>(x1);
        tz_scan<L>(y1);
        tz_scan<L>(x2);
        tz_scan<L>(y2);
        const FdotPrep f = fdot2_prep<L>(x1, y1, x2, y2, j.neg2);
        bool ok;
        Meta t = fdot2_tb_tail<L>(j.dst, f, lo_view(Q.col(q)), lo_view(Q.col(q + 1)), &ok);
        CB_COUNT(ok ? 4 : 5);
        if (!ok) t = arb_fdot2<L>(j.dst, x1, y1, x2, y2, j.neg2, Q.col(q), Q.col(q + 1));  // exact, on this lane alone
        st_meta(j.out, t);
    }
    Q.sync();
}

// z = x * y (acb_mul); z must not alias x or y
template <int L, class QD>
CB_HD void coop_mul(const QD& Q, const CRef& z, const CRef& x, const CRef& y) {
    Opd a = re_of(x), b = im_of(x), c = re_of(y), d = im_of(y);
    FdotJob<L> jr{a, c, b, d, true, wre(z), z.t}, ji{a, d, b, c, false, wim(z), z.t + 1};
    coop_fdot2_pair<L>(Q, jr, ji, true);
}
// z = x +- y (componentwise; z may alias x or y): real part on lane 0, imaginary part on lane 2
template <int L, class QD>
CB_HD void coop_addsub(const QD& Q, const CRef& z, const CRef& x, const CRef& y, bool sub) {
    CB_COOP_EACH(Q, q) {
        if (q & 1) continue;
        const int part = q >> 1;
        Meta t = part ? arb_addsub<L>(wim(z), im_of(x), im_of(y), sub, Q.col(q), Q.col(q + 1))
                      : arb_addsub<L>(wre(z), re_of(x), re_of(y), sub, Q.col(q), Q.col(q + 1));
        st_meta(z.t + part, t);
    }
    Q.sync();
}
// z = x + k (acb_add_si): the real part gets the integer (lane 0), the imaginary part is copied (lane 2)
template <int L, class QD>
CB_HD void coop_add_si(const QD& Q, const CRef& z, const CRef& x, int64_t k, const CRef& kball) {
    CB_COOP_EACH(Q, q) {
        if (q == 0) {
            uint64_t ka = k < 0 ? (uint64_t)0 - (uint64_t)k : (uint64_t)k;
            int sh = clz64(ka);
            uint64_t kn = ka ? ka << sh : 0;
#pragma unroll 8
            for (int j = 0; j < L; j++) kball.m[j * kball.stride] = 0u;
            kball.m[(L - 1) * kball.stride] = (uint32_t)(kn >> 32);
            kball.m[(L - 2) * kball.stride] = (uint32_t)kn;
            Meta kt = ka ? Meta{64 - sh, k < 0 ? F_SIGN : 0u, 0u, 0} : meta_zero();
            Opd ko{RRef{kball.m, kball.stride}, kt};
            Meta tr = arb_addsub<L>(wre(z), re_of(x), ko, false, Q.col(0), Q.col(1));
            st_meta(z.t, tr);
        } else if (q == 2 && z.m != x.m) {
#pragma unroll 8
            for (int j = 0; j < L; j++) z.m[j * z.stride + 1] = x.m[j * x.stride + 1];
            st_meta(z.t + 1, ld_meta(x.t + 1));
        }
    }
    Q.sync();
}
// z = 1 / y; z must not alias y; uses tmp.re as a real temporary (acb_inv: t = c^2 + d^2 rounded once, then c/t, -d/t)
template <int L, class QD>
CB_HD void coop_inv(const QD& Q, const CRef& z, const CRef& y, const CRef& tmp) {
    Opd c = re_of(y), d = im_of(y);
    FdotJob<L> jt{c, c, d, d, false, wre(tmp), tmp.t};
    coop_fdot2_pair<L>(Q, jt, jt, false);
    CB_COOP_EACH(Q, q) {
        if (q & 1) continue;
        Opd t{RRef{tmp.m, tmp.stride}, ld_meta(tmp.t)};
        if (q == 0) {
            st_meta(z.t, arb_div<L>(wre(z), c, t, Q.col(0), Q.col(1)));
        } else {
            Meta ti = arb_div<L>(wim(z), d, t, Q.col(2), Q.col(3));
            if (!is_nan(ti) && !is_zero_mid(ti)) ti.flags ^= F_SIGN;
            st_meta(z.t + 1, ti);
        }
    }
    Q.sync();
}
// z = x / y; z must not alias x or y; inv: a complex temporary (acb_div: componentwise for a real divisor, else x * inv(y))
template <int L, class QD>
CB_HD void coop_div(const QD& Q, const CRef& z, const CRef& x, const CRef& y, const CRef& inv) {
    if (is_exact_zero(ld_meta(y.t + 1))) {
        CB_COOP_EACH(Q, q) {
            if (q & 1) continue;
            Opd c = re_of(y);
            if (q == 0) st_meta(z.t, arb_div<L>(wre(z), re_of(x), c, Q.col(0), Q.col(1)));
            else st_meta(z.t + 1, arb_div<L>(wim(z), im_of(x), c, Q.col(2), Q.col(3)));
        }
        Q.sync();
    } else {
        coop_inv<L>(Q, inv, y, z);
        coop_mul<L>(Q, z, x, inv);
    }
}

// indicial_polynomial_evaluate (reference src/frobenius_solver.c:41-70), as indicial_eval of cb_solvers.cuh
template <int L, class QD, class PRM>
CB_HD CRef coop_indicial_eval(const QD& Q, const PRM& p, size_t oslot, int64_t nu, const CRef& rho, int64_t shift, CRef o1,
                              CRef o2, const CRef& t, const CRef& kb) {
    const int D1 = p.degree + 1;
    nu += p.valuation;
    coop_zero<L>(Q, o1);
    if (nu > p.degree) return o1;
#pragma unroll 1
    for (int64_t lam = clampi(p.degree - nu, 0, p.order); lam >= 0; lam--) {
        coop_add_si<L>(Q, t, rho, shift - lam, kb);
        coop_mul<L>(Q, o2, o1, t);
        CRef sw = o1;
        o1 = o2;
        o2 = sw;
        if (lam + nu < 0) continue;
        coop_addsub<L>(Q, o1, o1, at<L>(p.ode, lam * D1 + (lam + nu), oslot), false);
    }
    return o1;
}

// _acb_ode_solve_frobenius, homogeneous case (reference src/frobenius_solver.c:72-121), coefficients nu_from ..
// nu_to of one instance on the four lanes of a quad (the recurrence reads earlier coefficients back from res, so
// it resumes anywhere; nu = 0 writes g_0 = 1)
template <int L, class QD>
CB_HD void frobenius1_coop_quad(const FrobParams& p, int64_t inst, const QD& Q, int64_t nu_from, int64_t nu_to) {
    const size_t slot = 2 * (size_t)inst;
    const size_t oslot = (p.ode.S == 2) ? 0 : slot;
    const size_t rslot = (p.rho.S == 2) ? 0 : slot;
    CRef gnew = at<L>(p.tmp, 0, slot), o1 = at<L>(p.tmp, 1, slot), o2 = at<L>(p.tmp, 2, slot),
         t = at<L>(p.tmp, 3, slot), kb = at<L>(p.tmp, 4, slot), prod = at<L>(p.tmp, 5, slot);
    CRef rho = at<L>(p.rho, 0, rslot);
    if (nu_from == 0) {
        coop_one<L>(Q, at<L>(p.res, 0, slot));  // g_0 = 1 (src/frobenius_solver.c:84-85)
        nu_from = 1;
    }
#pragma unroll 1
    for (int64_t nu = nu_from; nu <= nu_to; nu++) {
        coop_zero<L>(Q, gnew);
        int64_t i = clampi(nu, 1, p.degree);
        CRef ind = coop_indicial_eval<L>(Q, p, oslot, i, rho, nu - i, o1, o2, t, kb);
#pragma unroll 1
        do {
            coop_mul<L>(Q, prod, ind, at<L>(p.res, nu - i, slot));
            coop_addsub<L>(Q, gnew, gnew, prod, true);
            i--;
            ind = coop_indicial_eval<L>(Q, p, oslot, i, rho, nu - i, o1, o2, t, kb);
        } while (i > 0);
        CRef spare = (ind.m == o1.m) ? o2 : o1;
        coop_div<L>(Q, at<L>(p.res, nu, slot), gnew, ind, spare);
    }
}

// Horner evaluation of one series (acb_poly_evaluate; with nder > 1 the normalised derivatives kept by
// analytic_continuation), as horner_thread of cb_solvers.cuh, for one point on the four lanes of a quad
template <int L, class QD>
CB_HD void horner_coop_quad(const HornerParams& p, int64_t pt, const QD& Q) {
    const size_t slot = 2 * (size_t)pt;
    const size_t fslot = (p.f.S == 2) ? 0 : slot;
    CRef a = at<L>(p.pts, 0, (p.pts.S == 2) ? 0 : slot), tmp = at<L>(p.tmp, 0, slot);
    for (int k = 0; k < p.nder; k++) coop_zero<L>(Q, at<L>(p.out, k, slot));
    if (p.len <= 0) return;
    coop_set<L>(Q, at<L>(p.out, 0, slot), at<L>(p.f, p.len - 1, fslot));
#pragma unroll 1
    for (int64_t j = p.len - 2; j >= 0; j--) {
#pragma unroll 1
        for (int k = p.nder - 1; k >= 0; k--) {
            CRef o = at<L>(p.out, k, slot);
            coop_mul<L>(Q, tmp, o, a);
            if (k > 0)
                coop_addsub<L>(Q, o, tmp, at<L>(p.out, k - 1, slot), false);
            else
                coop_addsub<L>(Q, o, tmp, at<L>(p.f, j, fslot), false);
        }
    }
}

}  // namespace cb
