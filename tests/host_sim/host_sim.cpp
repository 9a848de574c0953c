// host_sim.cpp -- TEST INFRASTRUCTURE.  Compiles the per-thread bodies of the CUDA kernels
// (cascade_b200/csrc/cb_solvers.cuh) for the host and runs them one "thread" after another, so
// that the device LOGIC can be checked against the oracle on a machine without a GPU.  It is
// never linked into libcascade_b200.so and nothing in the product loads it.
// Same array layout and arguments as the *_dev entry points of include/cascade_b200.h.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CB_COUNT_PATHS 1  // path statistics (cbsim_path_count): which arithmetic paths a test went through
#include "../../include/cascade_b200.h"
#include "../../cascade_b200/csrc/cb_solvers.cuh"

using namespace cb;

namespace {
struct Tmp {
    std::vector<uint32_t> m;
    std::vector<Meta> t;
    CArr arr(int L, int64_t entries, int64_t S) {
        m.assign((size_t)entries * L * S, 0u);
        t.assign((size_t)entries * S, Meta{0, F_ZERO, 0, 0});
        return CArr{m.data(), t.data(), (size_t)S};
    }
};
template <int L, class F>
void run(int64_t count, F body) {
    std::vector<uint32_t> s0(2 * L + 3), s1(2 * L + 3);
    Scratch<1> S0{s0.data()}, S1{s1.data()};
    for (int64_t i = 0; i < count; i++) body(i, S0, S1);
}
// the kernels that run on hybrid columns at 2048 bits and up (cb_kernels.cuh HybCfg): the same split here, with
// exactly sized parts, so the mapping is checked on the CPU (and under AddressSanitizer)
template <int L, class F>
void run_hyb(int64_t count, F body) {
    if constexpr (L >= 64) {
        constexpr int KS = L + 5, GW = 2 * L + 3 - KS;
        std::vector<uint32_t> s0(KS), s1(KS), g0(GW), g1(GW);
        HybridScratch<1, KS> S0{s0.data(), g0.data(), 1u}, S1{s1.data(), g1.data(), 1u};
        for (int64_t i = 0; i < count; i++) body(i, S0, S1);
    } else {
        run<L>(count, body);
    }
}
#define SIM_DISPATCH(L_, CALL)                                \
    switch (L_) {                                             \
        case 4: { constexpr int L = 4; CALL; } break;         \
        case 32: { constexpr int L = 32; CALL; } break;       \
        case 64: { constexpr int L = 64; CALL; } break;       \
        case 128: { constexpr int L = 128; CALL; } break;     \
        default: return CB200_EINVAL;                         \
    }
}  // namespace
template <int L>
static int sim_fuchs_shared_coop(FuchsSharedParams& sp);

extern "C" {
long long cb_path_counts[8] = {0, 0, 0, 0, 0, 0, 0, 0};
long long cbsim_path_count(int i) { return (i >= 0 && i < 8) ? cb_path_counts[i] : -1; }


int cbsim_fuchs_batch_range(int L_, int order, int degree, int valuation, int64_t n_begin, int64_t n, int64_t batch,
                            const uint32_t* ode_mant, const cb_meta* ode_meta, int shared_ode,
                            uint32_t* res_mant, cb_meta* res_meta);
int cbsim_fuchs_batch(int L_, int order, int degree, int valuation, int64_t n, int64_t batch,
                      const uint32_t* ode_mant, const cb_meta* ode_meta, int shared_ode,
                      uint32_t* res_mant, cb_meta* res_meta) {
    return cbsim_fuchs_batch_range(L_, order, degree, valuation, 0, n, batch, ode_mant, ode_meta, shared_ode, res_mant, res_meta);
}
int cbsim_fuchs_batch_range(int L_, int order, int degree, int valuation, int64_t n_begin, int64_t n, int64_t batch,
                            const uint32_t* ode_mant, const cb_meta* ode_meta, int shared_ode,
                            uint32_t* res_mant, cb_meta* res_meta) {
    if (valuation >= 0) return CB200_EVALUATION;
    if (n_begin < -valuation) n_begin = -valuation;
    if (n_begin > n) return 0;
    CArr odeA{const_cast<uint32_t*>(ode_mant), (Meta*)const_cast<cb_meta*>(ode_meta),
              shared_ode ? (size_t)2 : (size_t)(2 * batch)};
    CArr resA{res_mant, (Meta*)res_meta, (size_t)(2 * batch)};
    const int64_t rows = n + 1 + valuation, W = degree - valuation;
    if (rows <= 0) return 0;
    // columns sized EXACTLY as on the device (cb_kernels.cuh Sizes<L>), each followed by a canary zone, so an
    // out-of-column access is caught here and not only on the GPU
    const bool trunc = (L_ >= 16 && L_ <= 32);
    const size_t aw = trunc ? L_ + 5 : L_ + 2;
    const size_t sizes[7] = {(size_t)(trunc ? 0 : 2 * L_ + 3), (size_t)(trunc ? 0 : 2 * L_ + 3), aw, aw, aw,
                             (size_t)2 * L_ + 3, (size_t)2 * L_ + 3};
    const size_t CAN = 64;
    std::vector<uint32_t> col[7];
    for (int c = 0; c < 7; c++) col[c].assign(sizes[c] + CAN, 0xDEADBEEFu);
    Scratch<1> S0{col[0].data()}, S1{col[1].data()}, A0{col[2].data()}, A1{col[3].data()}, A2{col[4].data()};
    GScratch G0{col[5].data(), 1}, G1{col[6].data(), 1};
    auto canaries_ok = [&]() {
        for (int c = 0; c < 7; c++)
            for (size_t k = sizes[c]; k < sizes[c] + CAN; k++)
                if (col[c][k] != 0xDEADBEEFu) return -1000 - c;
        return 0;
    };
    FuchsSharedParams sp;
    sp.order = order; sp.degree = degree; sp.valuation = valuation; sp.batch = batch;
    sp.res = resA; sp.gcols = nullptr; sp.stats = nullptr; sp.sched = nullptr; sp.range_len = 32;
    { const char* e = getenv("CB200_RR"); sp.no_rr = (e && e[0] == '1') ? 0 : 1; }  // as cb200_fuchs_batch_dev
    FuchsTableParams tp;
    tp.order = order; tp.degree = degree; tp.valuation = valuation; tp.n = n; tp.ode = odeA;
    if (shared_ode) {  // same two-phase path as cb200_fuchs_batch_range_dev
        Tmp tbl, ttmp;
        std::vector<int32_t> kind(rows + 1);
        tp.tbl = tbl.arr(L_, rows * (W + 2), 2);
        tp.tmp = ttmp.arr(L_, 2, 2 * rows);
        tp.kind = kind.data(); tp.row0 = n_begin + valuation; tp.nrows = rows - tp.row0; tp.ninst = 1; tp.tbl_row0 = 0;
        SIM_DISPATCH(L_, run<L>(tp.nrows, [&](int64_t i, Scratch<1> a, Scratch<1> b) { fuchs_table_fast_thread<L>(tp, i, a, b); }));
        sp.n = n; sp.tbl = tp.tbl; sp.kind = kind.data(); sp.n_begin = n_begin; sp.tbl_row0 = 0; sp.kind_stride = 1; sp.per_inst = 0;
        if (getenv("CBSIM_COOP") && L_ >= 64) {
            int rc = L_ == 64 ? sim_fuchs_shared_coop<64>(sp) : sim_fuchs_shared_coop<128>(sp);
            return rc ? rc : canaries_ok();
        }
        if (sp.no_rr) {
            SIM_DISPATCH(L_, for (int64_t i = 0; i < batch; i++) (fuchs_shared_thread<L, false>)(sp, i, S0, S1, A0, A1, A2, G0, G1));
        } else {
            SIM_DISPATCH(L_, for (int64_t i = 0; i < batch; i++) (fuchs_shared_thread<L, true>)(sp, i, S0, S1, A0, A1, A2, G0, G1));
        }
        return canaries_ok();
    }
    // an operator per instance: a table row per instance, a few rows at a time (deliberately small here, so that
    // the range loop of cb200_fuchs_batch_range_dev is exercised)
    const int64_t rt = rows < 7 ? rows : 7;
    Tmp tbl, ttmp;
    std::vector<int32_t> kind(rt * batch + 1);
    tp.tbl = tbl.arr(L_, rt * (W + 2), 2 * batch);
    tp.tmp = ttmp.arr(L_, 2, 2 * rt * batch);
    tp.kind = kind.data(); tp.ninst = batch;
    for (int64_t lo = n_begin; lo <= n; lo += rt) {
        const int64_t hi = lo + rt - 1 < n ? lo + rt - 1 : n;
        tp.row0 = lo + valuation; tp.nrows = hi - lo + 1; tp.tbl_row0 = lo + valuation;
        SIM_DISPATCH(L_, run<L>(tp.nrows * batch, [&](int64_t i, Scratch<1> a, Scratch<1> b) { fuchs_table_fast_thread<L>(tp, i, a, b); }));
        sp.n = hi; sp.tbl = tp.tbl; sp.kind = kind.data(); sp.n_begin = lo; sp.tbl_row0 = lo + valuation; sp.kind_stride = batch;
        sp.per_inst = 1; sp.no_rr = 1;
        SIM_DISPATCH(L_, for (int64_t i = 0; i < batch; i++) (fuchs_shared_thread<L, false, false, true>)(sp, i, S0, S1, A0, A1, A2, G0, G1));
    }
    return canaries_ok();
}

int cbsim_fuchs_intop_batch(int L_, int order, int degree, int valuation, int64_t n, int64_t batch,
                            const int64_t* ode_int, int shared_ode, uint32_t* res_mant, cb_meta* res_meta) {
    if (valuation >= 0) return CB200_EVALUATION;
    FuchsIntParams p;
    p.order = order; p.degree = degree; p.valuation = valuation; p.n = n; p.batch = batch;
    p.ode = ode_int; p.ostride = shared_ode ? 0 : (size_t)batch; p.sched = nullptr; p.range_len = 32;
    p.gs0 = nullptr; p.gs0_stride = 0;
    p.res = CArr{res_mant, (Meta*)res_meta, (size_t)(2 * batch)};
    const size_t sizes[4] = {(size_t)L_ + 3, (size_t)L_ + 2, (size_t)L_ + 2, (size_t)L_ + 2};
    const size_t CAN = 64;
    std::vector<uint32_t> col[4];
    for (int c = 0; c < 4; c++) col[c].assign(sizes[c] + CAN, 0xDEADBEEFu);
    GScratch S0{col[0].data(), 1};  // the product column of the general body is a global-memory column on the device
    Scratch<1> A0{col[1].data()}, A1{col[2].data()}, A2{col[3].data()};
    // the register fast path with the shared-memory body as per-coefficient fallback (what the kernels run at <= 1024
    // bits); CBSIM_INTOP_GENERAL=1 runs the general body alone
    const bool general_only = getenv("CBSIM_INTOP_GENERAL") != nullptr;
    SIM_DISPATCH(L_, for (int64_t i = 0; i < batch; i++) {
        if (L <= 32 && !general_only) { if constexpr (L <= 32) fuchs_intop_fast_thread<L>(p, i, S0, A0, A1, A2); }
        else fuchs_intop_thread<L>(p, i, S0, A0, A1, A2);
    });
    for (int c = 0; c < 4; c++)
        for (size_t k = sizes[c]; k < sizes[c] + CAN; k++)
            if (col[c][k] != 0xDEADBEEFu) return -1000 - c;
    return 0;
}

int cbsim_frobenius1_batch(int L_, int order, int degree, int valuation, int64_t n, int64_t batch,
                           const uint32_t* ode_mant, const cb_meta* ode_meta, int shared_ode,
                           const uint32_t* rho_mant, const cb_meta* rho_meta, int shared_rho,
                           const uint32_t* rhs_mant, const cb_meta* rhs_meta, int64_t rhs_len, int shared_rhs,
                           uint32_t* res_mant, cb_meta* res_meta) {
    FrobParams p;
    p.rhs = CArr{const_cast<uint32_t*>(rhs_mant), (Meta*)const_cast<cb_meta*>(rhs_meta),
                 shared_rhs ? (size_t)2 : (size_t)(2 * batch)};
    p.rhs_len = rhs_len;
    p.order = order; p.degree = degree; p.valuation = valuation; p.n = n; p.batch = batch;
    p.ode = CArr{const_cast<uint32_t*>(ode_mant), (Meta*)const_cast<cb_meta*>(ode_meta),
                 shared_ode ? (size_t)2 : (size_t)(2 * batch)};
    p.rho = CArr{const_cast<uint32_t*>(rho_mant), (Meta*)const_cast<cb_meta*>(rho_meta),
                 shared_rho ? (size_t)2 : (size_t)(2 * batch)};
    p.res = CArr{res_mant, (Meta*)res_meta, (size_t)(2 * batch)};
    Tmp tmp;
    p.tmp = tmp.arr(L_, FROB_NTMP, 2 * batch);
    p.gcols = nullptr;
    p.sched = nullptr;
    p.gcols_words = 0;
    // CBSIM_FROB_RANGE=k: a homogeneous solve in ranges of k coefficients, as the persistent kernel's work units run it
    const char* e = getenv("CBSIM_FROB_RANGE");
    const int64_t range = (e && rhs_len == 0) ? atoi(e) : 0;
    if (range > 0) {
        for (int64_t lo = 0; lo <= n; lo += range) {
            const int64_t hi = lo + range - 1 < n ? lo + range - 1 : n;
            SIM_DISPATCH(L_, run_hyb<L>(batch, [&](int64_t i, auto a, auto b) { frobenius1_thread<L>(p, i, a, b, lo, hi); }));
        }
        return 0;
    }
    SIM_DISPATCH(L_, run_hyb<L>(batch, [&](int64_t i, auto a, auto b) { frobenius1_thread<L>(p, i, a, b); }));
    return 0;
}

int cbsim_frobenius_general_batch(int L_, int order, int degree, int valuation, int64_t n, int64_t batch,
                                  const uint32_t* ode_mant, const cb_meta* ode_meta, int shared_ode,
                                  const uint32_t* rho_mant, const cb_meta* rho_meta, int shared_rho, int mul, int M,
                                  uint32_t* gens_mant, cb_meta* gens_meta) {
    if (degree < 1 || degree > FROBG_MAXD || M < 1 || M > FROBG_MAXM || mul < 1 || mul > M) return CB200_EINVAL;
    FrobGenParams p;
    p.order = order; p.degree = degree; p.valuation = valuation; p.M = M; p.mul = mul; p.n = n; p.batch = batch;
    p.plen = (int64_t)order * n + order + 2;
    p.ode = CArr{const_cast<uint32_t*>(ode_mant), (Meta*)const_cast<cb_meta*>(ode_meta),
                 shared_ode ? (size_t)2 : (size_t)(2 * batch)};
    p.rho = CArr{const_cast<uint32_t*>(rho_mant), (Meta*)const_cast<cb_meta*>(rho_meta),
                 shared_rho ? (size_t)2 : (size_t)(2 * batch)};
    p.gens = CArr{gens_mant, (Meta*)gens_meta, (size_t)(2 * batch)};
    Tmp tmp;
    p.wp = tmp.arr(L_, (int64_t)(degree + 3) * p.plen + FROBG_NTMP + M, 2 * batch);
    // poison the polynomial workspace: reads beyond a polynomial's length must never happen
    std::fill(tmp.m.begin(), tmp.m.end(), 0xDEADBEEFu);
    SIM_DISPATCH(L_, run<L>(batch, [&](int64_t i, Scratch<1> a, Scratch<1> b) { frobenius_general_thread<L>(p, i, a, b); }));
    return 0;
}

int cbsim_solution_op(int which, int L_, int64_t batch, int mul, int M, int64_t cap, int64_t nu,
                      const uint32_t* rho_mant, const cb_meta* rho_meta, uint32_t* gens_mant, cb_meta* gens_meta,
                      uint32_t* f_mant, cb_meta* f_meta, int64_t fcap) {
    if (M < 0 || M > FROBG_MAXM) return CB200_EINVAL;
    SolOpParams p;
    p.which = which; p.M = M; p.mul = mul; p.batch = batch; p.cap = cap; p.fcap = which == 2 ? 0 : fcap; p.nu = nu;
    p.rho = CArr{const_cast<uint32_t*>(rho_mant), (Meta*)const_cast<cb_meta*>(rho_meta), (size_t)(2 * batch)};
    p.gens = CArr{gens_mant, (Meta*)gens_meta, (size_t)(2 * batch)};
    p.f = CArr{f_mant, (Meta*)f_meta, (size_t)(2 * batch)};
    Tmp tmp;
    p.wp = tmp.arr(L_, FROBG_NTMP + M, 2 * batch);
    SIM_DISPATCH(L_, run<L>(batch, [&](int64_t i, Scratch<1> a, Scratch<1> b) { solution_op_thread<L>(p, i, a, b); }));
    return 0;
}

int cbsim_indicial_polynomial(int L_, int order, int degree, int valuation, int64_t batch, const uint32_t* ode_mant,
                              const cb_meta* ode_meta, int shared_ode, int64_t nu, int64_t shift, uint32_t* out_mant,
                              cb_meta* out_meta) {
    IndicialParams p;
    p.order = order; p.degree = degree; p.valuation = valuation; p.batch = batch; p.nu = nu; p.shift = shift;
    p.ode = CArr{const_cast<uint32_t*>(ode_mant), (Meta*)const_cast<cb_meta*>(ode_meta),
                 shared_ode ? (size_t)2 : (size_t)(2 * batch)};
    p.out = CArr{out_mant, (Meta*)out_meta, (size_t)(2 * batch)};
    Tmp tmp;
    p.wp = tmp.arr(L_, FROBG_NTMP, 2 * batch);
    SIM_DISPATCH(L_, run<L>(batch, [&](int64_t i, Scratch<1> a, Scratch<1> b) { indicial_poly_thread<L>(p, i, a, b); }));
    return 0;
}

int cbsim_solution_evaluate(int L_, int M, int64_t cap, int64_t npts, const uint32_t* gens_mant, const cb_meta* gens_meta,
                            const uint32_t* rho_mant, const cb_meta* rho_meta, int pow_mode, uint64_t pow_e,
                            const uint32_t* pts_mant, const cb_meta* pts_meta, uint32_t* out_mant, cb_meta* out_meta) {
    if (M < 1 || M > FROBG_MAXM) return CB200_EINVAL;
    EvalParams p;
    p.M = M; p.pow_mode = pow_mode; p.pow_e = pow_e; p.cap = cap; p.npts = npts;
    p.gens = CArr{const_cast<uint32_t*>(gens_mant), (Meta*)const_cast<cb_meta*>(gens_meta), (size_t)2};
    p.rho = CArr{const_cast<uint32_t*>(rho_mant), (Meta*)const_cast<cb_meta*>(rho_meta), (size_t)2};
    p.pts = CArr{const_cast<uint32_t*>(pts_mant), (Meta*)const_cast<cb_meta*>(pts_meta), (size_t)(2 * npts)};
    p.out = CArr{out_mant, (Meta*)out_meta, (size_t)(2 * npts)};
    Tmp tmp;
    p.wp = tmp.arr(L_, TRANS_NTMP + 4, 2 * npts);
    SIM_DISPATCH(L_, run<L>(npts, [&](int64_t i, Scratch<1> a, Scratch<1> b) { evaluate_thread<L>(p, i, a, b); }));
    return 0;
}

int cbsim_trans(int which, int L_, int64_t n, uint32_t* z_mant, cb_meta* z_meta, const uint32_t* x_mant, const cb_meta* x_meta) {
    TransParams p;
    p.which = which; p.n = n;
    p.z = CArr{z_mant, (Meta*)z_meta, (size_t)(2 * n)};
    p.x = CArr{const_cast<uint32_t*>(x_mant), (Meta*)const_cast<cb_meta*>(x_meta), (size_t)(2 * n)};
    Tmp tmp;
    p.wp = tmp.arr(L_, TRANS_NTMP, 2 * n);
    SIM_DISPATCH(L_, run<L>(n, [&](int64_t i, Scratch<1> a, Scratch<1> b) { trans_thread<L>(p, i, a, b); }));
    return 0;
}

int cbsim_ode_apply(int L_, int order, int degree, int64_t batch, const uint32_t* ode_mant, const cb_meta* ode_meta,
                    int shared_ode, const uint32_t* in_mant, const cb_meta* in_meta, int64_t len_in, uint32_t* out_mant,
                    cb_meta* out_meta, int64_t out_len, int64_t deg, int32_t* solved) {
    ApplyParams p;
    p.order = order; p.degree = degree; p.batch = batch; p.len_in = len_in; p.out_len = out_len; p.deg = deg;
    p.ode = CArr{const_cast<uint32_t*>(ode_mant), (Meta*)const_cast<cb_meta*>(ode_meta),
                 shared_ode ? (size_t)2 : (size_t)(2 * batch)};
    p.in = CArr{const_cast<uint32_t*>(in_mant), (Meta*)const_cast<cb_meta*>(in_meta), (size_t)(2 * batch)};
    p.out = CArr{out_mant, (Meta*)out_meta, (size_t)(2 * batch)};
    Tmp tmp;
    p.wp = tmp.arr(L_, len_in + 1, 2 * batch);
    p.solved = solved;
    SIM_DISPATCH(L_, run<L>(batch, [&](int64_t i, Scratch<1> a, Scratch<1> b) { apply_thread<L>(p, i, a, b); }));
    return 0;
}

int cbsim_horner_batch(int L_, int64_t len, int64_t npts, int nder, const uint32_t* f_mant,
                       const cb_meta* f_meta, int shared_f, const uint32_t* pts_mant,
                       const cb_meta* pts_meta, uint32_t* out_mant, cb_meta* out_meta) {
    if (nder < 1 || nder > HORNER_MAXDER) return CB200_EINVAL;
    HornerParams p;
    p.nder = nder; p.len = len; p.npts = npts;
    p.f = CArr{const_cast<uint32_t*>(f_mant), (Meta*)const_cast<cb_meta*>(f_meta),
               shared_f ? (size_t)2 : (size_t)(2 * npts)};
    p.pts = CArr{const_cast<uint32_t*>(pts_mant), (Meta*)const_cast<cb_meta*>(pts_meta), (size_t)(2 * npts)};
    p.out = CArr{out_mant, (Meta*)out_meta, (size_t)(2 * npts)};
    Tmp tmp;
    p.tmp = tmp.arr(L_, 1, 2 * npts);
    p.gcols = nullptr;
    SIM_DISPATCH(L_, run_hyb<L>(npts, [&](int64_t i, auto a, auto b) { horner_thread<L>(p, i, a, b); }));
    return 0;
}

int cbsim_ew(int op, int L_, int64_t n, uint32_t* z_mant, cb_meta* z_meta, const uint32_t* x_mant,
             const cb_meta* x_meta, const uint32_t* y_mant, const cb_meta* y_meta, const int64_t* k) {
    EwParams p;
    p.op = op; p.n = n;
    p.z = CArr{z_mant, (Meta*)z_meta, (size_t)(2 * n)};
    p.x = CArr{const_cast<uint32_t*>(x_mant), (Meta*)const_cast<cb_meta*>(x_meta), (size_t)(2 * n)};
    p.y = y_mant ? CArr{const_cast<uint32_t*>(y_mant), (Meta*)const_cast<cb_meta*>(y_meta), (size_t)(2 * n)} : p.x;
    p.k = k;
    Tmp tmp;
    p.tmp = tmp.arr(L_, 2, 2 * n);
    SIM_DISPATCH(L_, run<L>(n, [&](int64_t i, Scratch<1> a, Scratch<1> b) { ew_thread<L>(p, i, a, b); }));
    return 0;
}


int cbsim_ode_shift(int L_, int order, int degree, int64_t batch, const uint32_t* ode_mant, const cb_meta* ode_meta,
                    int shared_ode, const uint32_t* a_mant, const cb_meta* a_meta, int shared_a, uint32_t* out_mant,
                    cb_meta* out_meta) {
    ShiftParams p;
    p.order = order; p.degree = degree; p.batch = batch;
    p.in = CArr{const_cast<uint32_t*>(ode_mant), (Meta*)const_cast<cb_meta*>(ode_meta), shared_ode ? (size_t)2 : (size_t)(2 * batch)};
    p.a = CArr{const_cast<uint32_t*>(a_mant), (Meta*)const_cast<cb_meta*>(a_meta), shared_a ? (size_t)2 : (size_t)(2 * batch)};
    p.out = CArr{out_mant, (Meta*)out_meta, (size_t)(2 * batch)};
    Tmp tmp;
    p.tmp = tmp.arr(L_, 1, 2 * batch);
    SIM_DISPATCH(L_, run<L>(batch, [&](int64_t i, Scratch<1> a, Scratch<1> b) { ode_shift_thread<L>(p, i, a, b); }));
    return 0;
}

// the two phases of every sweep of taylor_shift_kernel, the "threads" of a phase one after another
int cbsim_taylor_shift(int L_, int64_t len, int64_t batch, uint32_t* f_mant, cb_meta* f_meta, const uint32_t* a_mant,
                       const cb_meta* a_meta, int shared_a) {
    TShiftParams p;
    p.len = len; p.batch = batch;
    p.f = CArr{f_mant, (Meta*)f_meta, (size_t)(2 * batch)};
    p.a = CArr{const_cast<uint32_t*>(a_mant), (Meta*)const_cast<cb_meta*>(a_meta), shared_a ? (size_t)2 : (size_t)(2 * batch)};
    Tmp tmp;
    p.tmp = tmp.arr(L_, len > 0 ? len : 1, 2 * batch);
    for (int64_t inst = 0; inst < batch; inst++) {
        if (acb_is_exact_zero(at<4>(p.a, 0, shared_a ? 0 : 2 * (size_t)inst))) continue;  // meta only: L irrelevant
        for (int64_t i = len - 2; i >= 0; i--) {
            SIM_DISPATCH(L_, run<L>(len - 1 - i, [&](int64_t k, Scratch<1> a, Scratch<1> b) { taylor_shift_mul<L>(p, inst, i + k, a, b); }));
            SIM_DISPATCH(L_, run<L>(len - 1 - i, [&](int64_t k, Scratch<1> a, Scratch<1> b) { taylor_shift_add<L>(p, inst, i + k, a, b); }));
        }
    }
    return 0;
}

static int sim_valuation_of(const cb_meta* t, int order, int degree) {
    int val = degree;
    for (int i = 0; i <= order; i++) {
        int z = 0;
        for (; z <= degree; z++) {
            const cb_meta* c = t + 2 * (i * (degree + 1) + z);
            bool zero = (c[0].flags & CB_FLAG_ZERO) && c[0].rman == 0 && (c[1].flags & CB_FLAG_ZERO) && c[1].rman == 0;
            if (!zero) break;
        }
        if (z - i < val) val = z - i;
    }
    return val;
}

int cbsim_continuation_series(int L_, int order, int degree, int64_t n, int64_t batch, int64_t path_len,
                              const uint32_t* ode_mant, const cb_meta* ode_meta, const uint32_t* path_mant,
                              const cb_meta* path_meta, uint32_t* y_mant, cb_meta* y_meta, uint32_t* full_mant,
                              cb_meta* full_meta) {
    // same sequence of operations as continuation_impl (abi.cu), on host memory
    const int64_t ne = (int64_t)(order + 1) * (degree + 1), S = 2 * batch;
    std::vector<uint32_t> shm((size_t)ne * L_ * 2), serm((size_t)(n + 1) * L_ * S), tbm((size_t)3 * L_ * 2);
    std::vector<cb_meta> sht((size_t)ne * 2), sert((size_t)(n + 1) * S), tbt(6);
    auto bm = [&](std::vector<uint32_t>& v, int64_t e) { return v.data() + (size_t)e * L_ * 2; };
    auto bt = [&](std::vector<cb_meta>& v, int64_t e) { return v.data() + (size_t)e * 2; };
    for (int64_t t = 0; t + 1 < path_len; t++) {
        const uint32_t* am = path_mant + (size_t)t * L_ * 2;
        const cb_meta* at_ = path_meta + (size_t)t * 2;
        int rc = cbsim_ode_shift(L_, order, degree, 1, ode_mant, ode_meta, 1, am, at_, 1, shm.data(), sht.data());
        if (rc) return rc;
        const int v = sim_valuation_of(sht.data(), order, degree);
        if (v >= 0) return CB200_EVALUATION;
        memcpy(serm.data(), y_mant, (size_t)(-v) * L_ * S * 4);
        memcpy(sert.data(), y_meta, (size_t)(-v) * S * sizeof(cb_meta));
        rc = cbsim_fuchs_batch(L_, order, degree, v, n, batch, shm.data(), sht.data(), 1, serm.data(), sert.data());
        if (rc) return rc;
        cbsim_ew(CB200_OP_SUB, L_, 1, bm(tbm, 0), bt(tbt, 0), path_mant + (size_t)(t + 1) * L_ * 2, path_meta + (size_t)(t + 1) * 2,
                 am, at_, nullptr);
        if (full_mant && t + 2 == path_len) {
            rc = cbsim_taylor_shift(L_, n + 1, batch, serm.data(), sert.data(), bm(tbm, 0), bt(tbt, 0), 1);
            if (rc) return rc;
            memcpy(full_mant, serm.data(), serm.size() * 4);
            memcpy(full_meta, sert.data(), sert.size() * sizeof(cb_meta));
            memcpy(y_mant, serm.data(), (size_t)order * L_ * S * 4);
            memcpy(y_meta, sert.data(), (size_t)order * S * sizeof(cb_meta));
            break;
        }
        HornerParams hp;
        hp.nder = order; hp.len = n + 1; hp.npts = batch;
        hp.f = CArr{serm.data(), (Meta*)sert.data(), (size_t)S};
        hp.pts = CArr{bm(tbm, 0), (Meta*)bt(tbt, 0), (size_t)2};
        hp.out = CArr{y_mant, (Meta*)y_meta, (size_t)S};
        Tmp tmp;
        hp.tmp = tmp.arr(L_, 1, S);
        hp.gcols = nullptr;
        hp.gcols_words = 0;
        SIM_DISPATCH(L_, run_hyb<L>(batch, [&](int64_t i, auto a, auto b) { horner_thread<L>(hp, i, a, b); }));
    }
    return 0;
}
int cbsim_continuation(int L_, int order, int degree, int64_t n, int64_t batch, int64_t path_len,
                       const uint32_t* ode_mant, const cb_meta* ode_meta, const uint32_t* path_mant,
                       const cb_meta* path_meta, uint32_t* y_mant, cb_meta* y_meta) {
    return cbsim_continuation_series(L_, order, degree, n, batch, path_len, ode_mant, ode_meta, path_mant, path_meta, y_mant,
                                     y_meta, nullptr, nullptr);
}

// ---- cooperating lanes (cb_coop.cuh): the four lanes of a quad run their phase one after the other (CB_COOP_EACH),
// each on its own column of exactly the device size (KS shared words + GW global words, followed by canaries)
extern "C++" {
template <int L>
struct SimQuad {
    static constexpr int KS = L + 5, GW = 2 * L + 3 - KS;
    static constexpr size_t CAN = 32;
    std::vector<uint32_t> sh[4], gl[4];
    Quad<1, KS> Q;
    SimQuad() {
        for (int q = 0; q < 4; q++) {
            sh[q].assign(KS + CAN, 0xDEADBEEFu);
            gl[q].assign(GW + CAN, 0xDEADBEEFu);
            Q.p[q] = sh[q].data();
            Q.g[q] = gl[q].data();
        }
    }
    int check() const {
        for (int q = 0; q < 4; q++) {
            for (size_t k = KS; k < KS + CAN; k++) if (sh[q][k] != 0xDEADBEEFu) return -2000 - q;
            for (size_t k = GW; k < GW + CAN; k++) if (gl[q][k] != 0xDEADBEEFu) return -2100 - q;
        }
        return 0;
    }
};
template <int L>
static int sim_fuchs_shared_coop(FuchsSharedParams& sp) {
    if constexpr (L >= 64) {
        SimQuad<L> sq;
        std::vector<uint32_t> acols[4];
        std::vector<uint32_t> state(32, 0u);
        for (int c = 0; c < 4; c++) {
            acols[c].assign(L + 2 + 32, 0xDEADBEEFu);
            sq.Q.p2[c] = acols[c].data();
        }
        sq.Q.st = reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(state.data()) + 15) & ~(uintptr_t)15);
        const int64_t first = sp.n_begin > -sp.valuation ? sp.n_begin : -sp.valuation;
        for (int64_t lo = first; lo <= sp.n; lo += 32)
            for (int64_t i = 0; i < sp.batch; i++) fuchs_shared_coop_quad<L>(sp, i, sq.Q, lo, lo + 31 < sp.n ? lo + 31 : sp.n);
        for (int c = 0; c < 4; c++)
            for (size_t k = L + 2; k < (size_t)L + 2 + 32; k++)
                if (acols[c][k] != 0xDEADBEEFu) return -2200 - c;
        return sq.check();
    } else {
        return CB200_EINVAL;
    }
}
template <int L>
static int sim_frobenius1_coop(FrobParams& p) {
    if constexpr (L >= 64) {
        SimQuad<L> sq;
        // coefficient ranges as the work units of frobenius1_coop_kernel cut them
        for (int64_t lo = 0; lo <= p.n; lo += 8)
            for (int64_t i = 0; i < p.batch; i++) frobenius1_coop_quad<L>(p, i, sq.Q, lo, lo + 7 < p.n ? lo + 7 : p.n);
        return sq.check();
    } else {
        return CB200_EINVAL;
    }
}
template <int L>
static int sim_horner_coop(HornerParams& p) {
    if constexpr (L >= 64) {
        SimQuad<L> sq;
        for (int64_t i = 0; i < p.npts; i++) horner_coop_quad<L>(p, i, sq.Q);
        return sq.check();
    } else {
        return CB200_EINVAL;
    }
}
}  // extern "C++"
int cbsim_frobenius1_coop(int L_, int order, int degree, int valuation, int64_t n, int64_t batch, const uint32_t* ode_mant,
                          const cb_meta* ode_meta, int shared_ode, const uint32_t* rho_mant, const cb_meta* rho_meta,
                          int shared_rho, uint32_t* res_mant, cb_meta* res_meta) {
    FrobParams p;
    p.rhs = CArr{nullptr, nullptr, (size_t)2};
    p.rhs_len = 0;
    p.order = order; p.degree = degree; p.valuation = valuation; p.n = n; p.batch = batch;
    p.ode = CArr{const_cast<uint32_t*>(ode_mant), (Meta*)const_cast<cb_meta*>(ode_meta), shared_ode ? (size_t)2 : (size_t)(2 * batch)};
    p.rho = CArr{const_cast<uint32_t*>(rho_mant), (Meta*)const_cast<cb_meta*>(rho_meta), shared_rho ? (size_t)2 : (size_t)(2 * batch)};
    p.res = CArr{res_mant, (Meta*)res_meta, (size_t)(2 * batch)};
    Tmp tmp;
    p.tmp = tmp.arr(L_, FROB_NTMP, 2 * batch);
    p.gcols = nullptr; p.sched = nullptr; p.gcols_words = 0;
    switch (L_) {
        case 64: return sim_frobenius1_coop<64>(p);
        case 128: return sim_frobenius1_coop<128>(p);
        default: return CB200_EINVAL;
    }
}
int cbsim_horner_coop(int L_, int64_t len, int64_t npts, int nder, const uint32_t* f_mant, const cb_meta* f_meta, int shared_f,
                      const uint32_t* pts_mant, const cb_meta* pts_meta, uint32_t* out_mant, cb_meta* out_meta) {
    HornerParams p;
    p.nder = nder; p.len = len; p.npts = npts;
    p.f = CArr{const_cast<uint32_t*>(f_mant), (Meta*)const_cast<cb_meta*>(f_meta), shared_f ? (size_t)2 : (size_t)(2 * npts)};
    p.pts = CArr{const_cast<uint32_t*>(pts_mant), (Meta*)const_cast<cb_meta*>(pts_meta), (size_t)(2 * npts)};
    p.out = CArr{out_mant, (Meta*)out_meta, (size_t)(2 * npts)};
    Tmp tmp;
    p.tmp = tmp.arr(L_, 1, 2 * npts);
    p.gcols = nullptr; p.gcols_words = 0;
    switch (L_) {
        case 64: return sim_horner_coop<64>(p);
        case 128: return sim_horner_coop<128>(p);
        default: return CB200_EINVAL;
    }
}

int cbsim_radius(int L_, int order, int degree, int64_t batch, const uint32_t* ode_mant, const cb_meta* ode_meta, int shared_ode,
                 int64_t n, uint32_t* out_mant, cb_meta* out_meta, int32_t* n_done) {
    RadiusParams p;
    p.order = order; p.degree = degree; p.batch = batch; p.n = n;
    p.ode = CArr{const_cast<uint32_t*>(ode_mant), (Meta*)const_cast<cb_meta*>(ode_meta), shared_ode ? (size_t)2 : (size_t)(2 * batch)};
    p.out = CArr{out_mant, (Meta*)out_meta, (size_t)(2 * batch)};
    Tmp tmp;
    p.wp = tmp.arr(L_, 2 * (int64_t)(degree + 1) + RADIUS_NTMP + TRANS_NTMP, 2 * batch);
    p.n_done = n_done;
    SIM_DISPATCH(L_, run<L>(batch, [&](int64_t i, Scratch<1> a, Scratch<1> b) { radius_thread<L>(p, i, a, b); }));
    return 0;
}
int cbsim_truncation(int L_, int64_t batch, const uint32_t* eta_mant, const cb_meta* eta_meta, const uint32_t* alpha_mant,
                     const cb_meta* alpha_meta, int64_t bits, uint32_t* n_mant, cb_meta* n_meta, int64_t* n_out) {
    TruncParams p;
    p.batch = batch; p.bits = bits;
    p.eta = CArr{const_cast<uint32_t*>(eta_mant), (Meta*)const_cast<cb_meta*>(eta_meta), (size_t)(2 * batch)};
    p.alpha = CArr{const_cast<uint32_t*>(alpha_mant), (Meta*)const_cast<cb_meta*>(alpha_meta), (size_t)(2 * batch)};
    p.nball = CArr{n_mant, (Meta*)n_meta, (size_t)(2 * batch)};
    Tmp tmp;
    p.wp = tmp.arr(L_, 8 + TRANS_NTMP, 2 * batch);
    p.n_out = n_out;
    SIM_DISPATCH(L_, run<L>(batch, [&](int64_t i, Scratch<1> a, Scratch<1> b) { truncation_thread<L>(p, i, a, b); }));
    return 0;
}

}  // extern "C"
