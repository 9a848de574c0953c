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

// The four ROLES of an instance (role q forms product q and owns column q; roles 0 and 2 are the real / imaginary
// workers) are played by four lanes (one role each) or by two lanes (lane 0: roles 0 and 1 -- both products and
// everything else of the real part --, lane 1: roles 2 and 3).  Two lanes waste no issue slots in the O(L) phases and
// are the form for batches that fill the SMs; four lanes halve the latency of a product pair once more and are the
// form for a handful of instances.  CB_COOP_EACH runs the body for every role this lane plays.
#if defined(__CUDA_ARCH__)
template <int STRIDE, int KS>
struct Quad {
    using Col = HybridScratch<STRIDE, KS>;
    int q0, per;      // first role and number of roles of this lane
    unsigned qmask;   // the lanes of this instance
    uint32_t a0;      // column of role 0: shared-space byte address of word 0
    uint32_t cstep;   // columns between the columns of consecutive roles (1: [word][4 i + role]; G: [word][role][i])
    uint32_t* g0;     // global tail of the column of role 0 (same column order), gs words between its words
    uint32_t gs;
    uint32_t a20 = 0;        // column of role 0 in a second region (L + 2 words: accumulators of the recurrence kernel)
    uint32_t* st = nullptr;  // 24 state words of the instance (shared memory, 16-byte aligned)
    bool active = true;      // false: lanes without an instance (ragged last unit): they only meet the phase barriers
    bool cta_sync = false;   // phase barriers of the whole CTA before the long phases (see phase_sync of cb_solvers.cuh)
    __device__ __forceinline__ void phase() const {
        if (cta_sync) __syncthreads();
    }
    __device__ __forceinline__ Col col(int c) const { return Col{a0 + 4u * cstep * (uint32_t)c, g0 + (size_t)cstep * c, gs}; }
    __device__ __forceinline__ Scratch<STRIDE> acol(int c) const { return Scratch<STRIDE>{a20 + 4u * cstep * (uint32_t)c}; }
    __device__ __forceinline__ void sync() const { __syncwarp(qmask); }
};
#else
template <int STRIDE, int KS>
struct Quad {
    using Col = HybridScratch<STRIDE, KS>;
    int q0 = 0, per = 4;  // the simulation plays all four roles, phase by phase
    uint32_t* p[4];
    uint32_t* g[4];
    uint32_t* p2[4] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t* st = nullptr;
    bool active = true, cta_sync = false;
    void phase() const {}
    Col col(int qq) const { return Col{p[qq], g[qq], 1u}; }
    Scratch<STRIDE> acol(int qq) const { return Scratch<STRIDE>{p2[qq]}; }
    void sync() const {}
};
#endif
#define CB_COOP_EACH(Q, qv) for (int qv = (Q).q0; qv < (Q).q0 + (Q).per; qv++)

// ---- copies and constants, limbs dealt out to the four lanes
template <int L, class QD>
CB_HD void coop_set(const QD& Q, const CRef& z, const CRef& x) {
    if (!Q.active) return;
    CB_COOP_EACH(Q, q) {
        const int part = q & 1, k0 = q >> 1;
        staged_copy<L / 2>([&](int j) { return x.m[(size_t)(2 * j + k0) * x.stride + part]; },
                           [&](int j, uint32_t v) { z.m[(size_t)(2 * j + k0) * z.stride + part] = v; });
        if (q < 2) st_meta(z.t + q, ld_meta(x.t + q));
    }
    Q.sync();
}
template <int L, class QD>
CB_HD void coop_zero(const QD& Q, const CRef& z) {
    if (!Q.active) return;
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
    if (!Q.active) return;
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
CB_HD void coop_fdot2_pair(const QD& Q, const FdotJob<L>& j0, const FdotJob<L>& j1, bool have1) {  // (inlined: out of line its job records live in local memory, 2048-bit two-lane kernel 16.8 -> 30.7 ms)
    if (!Q.active) return;
    // the products: every lane works through the non-zero ones among its roles in step with the others (a lane playing
    // two roles whose first product is zero starts on its second at once)
    unsigned todo = 0u;
    CB_COOP_EACH(Q, q) {
        const FdotJob<L>& j = (q >> 1) ? j1 : j0;
        if ((q >> 1) && !have1) continue;
        const uint32_t fl = (q & 1) ? (j.x2.t.flags | j.y2.t.flags) : (j.x1.t.flags | j.y1.t.flags);
        if (!(fl & (F_ZERO | F_NAN))) todo |= 1u << q;
    }
#pragma unroll 1
    while (todo) {
        int q = 0;
        while (!((todo >> q) & 1u)) q++;
        todo &= todo - 1u;
        const FdotJob<L>& j = (q >> 1) ? j1 : j0;
        Opd xo = (q & 1) ? j.x2 : j.x1, yo = (q & 1) ? j.y2 : j.y1;
        tz_scan2<L>(xo, yo);  // an integer-like operand takes the two-row product
        product_trunc_auto<L, false>(lo_view(Q.col(q)), xo, yo);
    }
    Q.sync();
    CB_COOP_EACH(Q, q) {  // align, add, round
        if (q & 1) continue;
        if ((q >> 1) && !have1) continue;
        const FdotJob<L>& j = (q >> 1) ? j1 : j0;
        Opd x1 = j.x1, y1 = j.y1, x2 = j.x2, y2 = j.y2;
        // short (integer-like) operands make exact products, whose guard bits prove nothing: knowing their trailing
        // zero limbs lets the truncated path accept them as exact (as fdot2_auto does)
        tz_scan4<L>(x1, y1, x2, y2);
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
    if (!Q.active) return;
    Opd a = re_of(x), b = im_of(x), c = re_of(y), d = im_of(y);
    FdotJob<L> jr{a, c, b, d, true, wre(z), z.t}, ji{a, d, b, c, false, wim(z), z.t + 1};
    coop_fdot2_pair<L>(Q, jr, ji, true);
}
// part 0 / 1 of a complex ball as an operand / a destination.  The componentwise operations below run ONE call with the
// operands chosen by the part, so the real and the imaginary workers of a warp stay converged (two calls, one per
// part, would run one after the other).
CB_HD Opd part_of(const CRef& x, int part) { return Opd{RRef{x.m + part, x.stride}, ld_meta_user(x.t + part)}; }
CB_HD WRef wpart(const CRef& z, int part) { return WRef{z.m + part, z.stride}; }
// z = x +- y (componentwise; z may alias x or y): real part on role 0, imaginary part on role 2
template <int L, class QD>
CB_HD void coop_addsub(const QD& Q, const CRef& z, const CRef& x, const CRef& y, bool sub) {
    if (!Q.active) return;
    CB_COOP_EACH(Q, q) {
        if (q & 1) continue;
        const int part = q >> 1;
        Meta t = arb_addsub<L>(wpart(z, part), part_of(x, part), part_of(y, part), sub, Q.col(q), Q.col(q + 1));
        st_meta(z.t + part, t);
    }
    Q.sync();
}
// the imaginary part of x into z (role 2)
template <int L, class QD>
CB_HD void coop_copy_im(const QD& Q, const CRef& z, const CRef& x) {
    if (!Q.active) return;
    CB_COOP_EACH(Q, q) {
        if (q == 2) acb_copy_im<L>(z, x);
    }
    Q.sync();
}
// z = x + k (acb_add_si): the real part gets the integer (role 0), the imaginary part is copied (role 2) unless the
// caller has put it there already (COPY_IM = false)
template <int L, bool COPY_IM = true, class QD>
CB_HD void coop_add_si(const QD& Q, const CRef& z, const CRef& x, int64_t k, const CRef& kball) {
    if (!Q.active) return;
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
        } else if (COPY_IM && q == 2 && z.m != x.m) {
            acb_copy_im<L>(z, x);
        }
    }
    Q.sync();
}
// z = 1 / y; z must not alias y; uses tmp.re as a real temporary (acb_inv: t = c^2 + d^2 rounded once, then c/t, -d/t)
template <int L, class QD>
CB_HD void coop_inv(const QD& Q, const CRef& z, const CRef& y, const CRef& tmp) {
    if (!Q.active) return;
    Opd c = re_of(y), d = im_of(y);
    FdotJob<L> jt{c, c, d, d, false, wre(tmp), tmp.t};
    coop_fdot2_pair<L>(Q, jt, jt, false);
    CB_COOP_EACH(Q, q) {
        if (q & 1) continue;
        const int part = q >> 1;
        Opd t{RRef{tmp.m, tmp.stride}, ld_meta(tmp.t)};
        Meta r = arb_div<L>(wpart(z, part), part ? d : c, t, Q.col(q), Q.col(q + 1));
        if (part && !is_nan(r) && !is_zero_mid(r)) r.flags ^= F_SIGN;
        st_meta(z.t + part, r);
    }
    Q.sync();
}
// z = x / y; z must not alias x or y; inv: a complex temporary (acb_div: componentwise for a real divisor, else x * inv(y))
template <int L, class QD>
CB_HD void coop_div(const QD& Q, const CRef& z, const CRef& x, const CRef& y, const CRef& inv) {
    if (!Q.active) return;
    if (is_exact_zero(ld_meta(y.t + 1))) {
        CB_COOP_EACH(Q, q) {
            if (q & 1) continue;
            const int part = q >> 1;
            st_meta(z.t + part, arb_div<L>(wpart(z, part), part_of(x, part), re_of(y), Q.col(q), Q.col(q + 1)));
        }
        Q.sync();
    } else {
        coop_inv<L>(Q, inv, y, z);
        coop_mul<L>(Q, z, x, inv);
    }
}

// indicial_polynomial_evaluate (reference src/frobenius_solver.c:41-70), as indicial_eval of cb_solvers.cuh
template <int L, bool IM_READY = false, class QD, class PRM>
CB_HD CRef coop_indicial_eval(const QD& Q, const PRM& p, size_t oslot, int64_t nu, const CRef& rho, int64_t shift, CRef o1,
                              CRef o2, const CRef& t, const CRef& kb) {
    const int D1 = p.degree + 1;
    nu += p.valuation;
    coop_zero<L>(Q, o1);
    if (nu > p.degree) return o1;
#pragma unroll 1
    for (int64_t lam = clampi(p.degree - nu, 0, p.order); lam >= 0; lam--) {
        coop_add_si<L, !IM_READY>(Q, t, rho, shift - lam, kb);
        Q.phase();
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
    coop_copy_im<L>(Q, t, rho);  // every f(rho + k) below adds an integer to the same rho
#pragma unroll 1
    for (int64_t nu = nu_from; nu <= nu_to; nu++) {
        coop_zero<L>(Q, gnew);
        int64_t i = clampi(nu, 1, p.degree);
        CRef ind = coop_indicial_eval<L, true>(Q, p, oslot, i, rho, nu - i, o1, o2, t, kb);
#pragma unroll 1
        do {
            Q.phase();
            coop_mul<L>(Q, prod, ind, at<L>(p.res, nu - i, slot));
            coop_addsub<L>(Q, gnew, gnew, prod, true);
            i--;
            ind = coop_indicial_eval<L, true>(Q, p, oslot, i, rho, nu - i, o1, o2, t, kb);
        } while (i > 0);
        CRef spare = (ind.m == o1.m) ? o2 : o1;
        Q.phase();
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
            Q.phase();
            coop_mul<L>(Q, tmp, o, a);
            if (k > 0)
                coop_addsub<L>(Q, o, tmp, at<L>(p.out, k - 1, slot), false);
            else
                coop_addsub<L>(Q, o, tmp, at<L>(p.f, j, fslot), false);
        }
    }
}

// acb_ode_solve_fuchs with one operator for the batch (fuchs_shared_thread of cb_solvers.cuh, the recurrence of
// reference src/fuchs_solver.c:108-139 on the precomputed multiplier table), coefficients b_first .. b_last of one
// instance on the four lanes of a quad: the analytic continuation of a FEW solutions (reference src/fuchs_solver.c:
// 147-206: one path, `order` initial vectors) is this loop on one or two instances.  Lanes 0 / 2 own the accumulator
// and spare columns of the real / imaginary part (acol(0), acol(1) / acol(2), acol(3)); which of the two holds the
// accumulator after an aligned add depends on the data, so the choice and the meta records travel through the quad's
// state words.
template <int L, class QD>
CB_HD void fuchs_shared_coop_quad(const FuchsSharedParams& p, int64_t inst, const QD& Q, int64_t b_first, int64_t b_last) {
    const size_t slot = 2 * (size_t)inst;
    const int v = p.valuation, W = p.degree - v;
    // state words: [0] re accumulator column (0 or 1), [1] im accumulator column (2 or 3), [4..7] meta re, [8..11] meta im
    uint32_t* st = Q.st;
    Meta* stm = reinterpret_cast<Meta*>(st + 4);
    auto colref = [&](int c) { auto s = Q.acol(c); return RRef{s.generic(), decltype(s)::GSTRIDE}; };
#pragma unroll 1
    for (int64_t bmax = b_first; bmax <= b_last; bmax++) {
        const int64_t e = bmax + v;
        const int64_t bmin = clampi(e - p.degree, 0, bmax);
        const int cnt = (int)(bmax - bmin);
        const int64_t base = (bmax + v - p.tbl_row0) * (W + 2);
        if (Q.active) {
            CB_COOP_EACH(Q, q) {
                if (q == 0) { st[0] = 0u; st[1] = 2u; stm[0] = meta_zero(); stm[1] = meta_zero(); }
            }
            Q.sync();
        }
#pragma unroll 1
        for (int k = 0; k < cnt; k++) {
            Q.phase();
            if (Q.active) {
                CRef x = at<L>(p.res, bmin + k, slot), y = at<L>(p.tbl, base + k, 0);
                Opd a = re_of(x), b = im_of(x);
                Opd c = dyn(re_of_shared(y)), d = dyn(im_of_shared(y));
                const int are = (int)st[0], aim = (int)st[1];   // accumulator columns; the spare ones are their partners
                const int sre = are ^ 1, sim = aim ^ 1;
                auto sr = Q.acol(sre);
                auto si = Q.acol(sim);
                // (the meta records of the two products go to the state words)
                FdotJob<L> jr{a, c, b, d, true, WRef{sr.generic(), decltype(sr)::GSTRIDE}, stm + 2};
                FdotJob<L> ji{a, d, b, c, false, WRef{si.generic(), decltype(si)::GSTRIDE}, stm + 3};
                coop_fdot2_pair<L>(Q, jr, ji, true);
                CB_COOP_EACH(Q, q) {
                    if (q & 1) continue;
                    const int part = q >> 1;
                    const int ac = part ? aim : are, sc = part ? sim : sre;
                    Meta pr = stm[2 + part], tacc = stm[part];
                    if (k == 0) {  // 0 - x = -x exactly: adopt the column, flip the sign
                        if (!is_nan(pr) && !is_zero_mid(pr)) pr.flags ^= F_SIGN;
                        stm[part] = pr;
                        st[part] = (uint32_t)sc;
                    } else {
                        auto A = Q.acol(ac);
                        auto B = Q.acol(sc);
                        decltype(A) where = A;
                        Meta t = arb_addsub_cols<L>(A, tacc, B, pr, true, &where);
                        stm[part] = t;
                        st[part] = (where.a_or_p() == B.a_or_p()) ? (uint32_t)sc : (uint32_t)ac;
                    }
                }
                Q.sync();
            }
        }
        Q.phase();
        if (Q.active) {
            CRef out = at<L>(p.res, bmax, slot);
            const int are = (int)st[0], aim = (int)st[1];
            Opd xa{colref(are), stm[0]}, xb{colref(aim), stm[1]};
            if (p.kind[(bmax + v - p.tbl_row0) * p.kind_stride] == 1) {
                CRef y = at<L>(p.tbl, base + W + 1, 0);
                Opd c = dyn(re_of_shared(y)), d = dyn(im_of_shared(y));
                FdotJob<L> jr{xa, c, xb, d, true, wre(out), out.t}, ji{xa, d, xb, c, false, wim(out), out.t + 1};
                coop_fdot2_pair<L>(Q, jr, ji, true);
            } else {
                CRef y = at<L>(p.tbl, base + cnt, 0);
                CB_COOP_EACH(Q, q) {
                    if (q & 1) continue;
                    const int part = q >> 1;
                    st_meta(out.t + part, arb_div<L>(wpart(out, part), part ? xb : xa, re_of(y), Q.col(q), Q.col(q + 1)));
                }
                Q.sync();
            }
        }
    }
}

}  // namespace cb
