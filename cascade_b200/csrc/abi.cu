// abi.cu -- the C ABI (include/cascade_b200.h) of libcascade_b200.so.
// sm_100a only; there is no CPU path in this library.
#include <cuda_runtime.h>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "../../include/cascade_b200.h"
#include "cb_kernels.cuh"

using namespace cb;

static_assert(sizeof(cb_meta) == sizeof(Meta), "meta layout");

namespace {

std::atomic<uint64_t> g_launches{0};
unsigned long long* g_rr_stats = nullptr;  // development counters, see CB200_RR_STATS

const LaunchTable* table_for(int L) {
    switch (L) {
        case 4: return &launch_table_L4;
        case 32: return &launch_table_L32;
        case 64: return &launch_table_L64;
        case 128: return &launch_table_L128;
        default: return nullptr;
    }
}

bool supported_L(int L) { return L == 4 || L == 32 || L == 64 || L == 128; }

size_t arr_mant_bytes(int L, int64_t entries, int64_t S) { return (size_t)entries * L * S * 4; }
size_t arr_meta_bytes(int64_t entries, int64_t S) { return (size_t)entries * S * 16; }
size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// carve a ball array out of a workspace
CArr carve(char*& w, int L, int64_t entries, int64_t S) {
    CArr a;
    a.m = (uint32_t*)w;
    w += align256(arr_mant_bytes(L, entries, S));
    a.t = (Meta*)w;
    w += align256(arr_meta_bytes(entries, S));
    a.S = (size_t)S;
    return a;
}
size_t carve_bytes(int L, int64_t entries, int64_t S) {
    return align256(arr_mant_bytes(L, entries, S)) + align256(arr_meta_bytes(entries, S));
}

// does n^(order+1) fit comfortably in int64?  (the integer factors of src/fuchs_solver.c:132,134)
bool factor_fits(int64_t n, int order) {
    long double v = 1;
    for (int i = 0; i <= order; i++) v *= (long double)(n + 1);
    return v < 4.0e18L;
}

// columns of the general-path scratch: the batch padded to whole blocks of the persistent launch (which gives the
// idle lanes of a ragged last block columns of their own)
static_assert(CfgShared<32>::TPB_BIG == 384, "gcols_columns pads to the persistent block size");
size_t gcols_columns(int64_t batch) { return (size_t)(batch + 512); }

// rows of the per-instance table held at a time: about 4 GB of table, at least 8 rows, at most all of them
int64_t pi_table_rows(int L, int64_t W, int64_t rows, int64_t batch) {
    double per_row = (double)(W + 2) * 2.0 * batch * (4.0 * L + 16.0);
    int64_t rt = (int64_t)(4.0e9 / (per_row > 1 ? per_row : 1));
    if (rt < 8) rt = 8;
    if (rt > rows) rt = rows;
    return rt > 0 ? rt : 1;
}

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, n ? n : 1); }
};

}  // namespace

#define CUDA_RET(x)                          \
    do {                                     \
        cudaError_t e_ = (x);                \
        if (e_ != cudaSuccess) return (int)e_; \
    } while (0)

extern "C" {

const char* cb200_version(void) { return "cascade_b200 0.1 (sm_100a)"; }

int cb200_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return -(int)e;
    return n;
}

uint64_t cb200_kernel_launches(void) { return g_launches.load(); }

// development aid: counters of the reduced-radix fused dot product (CB200_RR_STATS=1); out[0] accepted, out[1] rejected
int cb200_debug_rr_stats(uint64_t* out) {
    out[0] = out[1] = 0;
    if (!g_rr_stats) return 0;
    unsigned long long h[4];
    cudaError_t e = cudaMemcpy(h, g_rr_stats, sizeof(h), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return (int)e;
    out[0] = h[0];
    out[1] = h[1];
    return 0;
}

size_t cb200_fuchs_workspace_bytes(int L, int order, int degree, int valuation, int64_t n, int64_t batch,
                                   int shared_ode) {
    (void)order;
    if (valuation >= 0) return 256;  // rejected by the solver (CB200_EVALUATION)
    if (!shared_ode) {
        // a table row per instance, built for one range of coefficient rows at a time
        int64_t W = degree - valuation, rt = pi_table_rows(L, W, n + 1 + valuation, batch);
        return carve_bytes(L, rt * (W + 2), 2 * batch) + carve_bytes(L, 2, 2 * rt * batch) +
               align256(sizeof(int32_t) * (size_t)(rt * batch + 1)) +
               align256((size_t)2 * (2 * L + 3) * gcols_columns(batch) * sizeof(uint32_t));
    }
    int64_t rows = n + 1 + valuation, W = degree - valuation;
    if (rows < 0) rows = 0;
    return carve_bytes(L, rows * (W + 2), 2) + carve_bytes(L, 2, 2 * rows) + align256(sizeof(int32_t) * (rows + 1)) +
           align256((size_t)2 * (2 * L + 3) * gcols_columns(batch) * sizeof(uint32_t)) +
           align256(sizeof(int) * (size_t)(batch / 32 + 2));
}
static size_t hyb_tail_bytes(int L, int64_t count);
static size_t coop_sched_bytes(int64_t count) { return align256(sizeof(int) * (size_t)((count > 0 ? count : 0) / 64 + 2)); }
size_t cb200_frobenius_workspace_bytes(int L, int64_t batch) {
    return carve_bytes(L, FROB_NTMP, 2 * batch) + hyb_tail_bytes(L, batch) + coop_sched_bytes(batch);
}
// global tails of the hybrid scratch columns (2048 bits and up; cb_kernels.cuh HybCfg: two columns per instance) --
// or of the cooperative kernels' columns (CoopCfg: four lanes per instance, the batch padded to whole CTAs)
static size_t hyb_tail_words(int L, int64_t count) {
    if (L < 64) return 0;
    size_t a = (size_t)2 * (size_t)(L - 2) * (size_t)(count > 0 ? count : 0), b = coop_tail_words(L, count);
    return a > b ? a : b;
}
static size_t hyb_tail_bytes(int L, int64_t count) { return align256(hyb_tail_words(L, count) * sizeof(uint32_t)); }
size_t cb200_horner_workspace_bytes(int L, int64_t npts) { return carve_bytes(L, 1, 2 * npts) + hyb_tail_bytes(L, npts); }
size_t cb200_ew_workspace_bytes(int L, int64_t n) { return carve_bytes(L, 2, 2 * n); }

int cb200_fuchs_batch_dev(int L_, int order, int degree, int valuation, int64_t n, int64_t batch,
                          const uint32_t* ode_mant, const cb_meta* ode_meta, int shared_ode,
                          uint32_t* res_mant, cb_meta* res_meta, void* workspace,
                          size_t workspace_bytes, void* stream) {
    return cb200_fuchs_batch_range_dev(L_, order, degree, valuation, 0, n, batch, ode_mant, ode_meta, shared_ode, res_mant,
                                       res_meta, workspace, workspace_bytes, stream);
}

int cb200_fuchs_batch_range_dev(int L_, int order, int degree, int valuation, int64_t n_begin, int64_t n, int64_t batch,
                                const uint32_t* ode_mant, const cb_meta* ode_meta, int shared_ode, uint32_t* res_mant,
                                cb_meta* res_meta, void* workspace, size_t workspace_bytes, void* stream) {
    if (!supported_L(L_) || order < 1 || degree < 0 || batch < 0 || n < 0 || n_begin < 0) return CB200_EINVAL;
    if (valuation >= 0) return CB200_EVALUATION;
    if (n_begin < -valuation) n_begin = -valuation;
    if (n_begin > n) return 0;
    if (!factor_fits(n, order)) return CB200_ERANGE;
    if (workspace_bytes < cb200_fuchs_workspace_bytes(L_, order, degree, valuation, n, batch, shared_ode))
        return CB200_ENOMEM;
    if (n < -valuation) return 0;  // nothing to compute
    if (shared_ode) {
        // one operator for the whole batch: multipliers/divisors once per coefficient index, then
        // the per-instance recurrence with shared-memory resident accumulators
        int64_t rows = n + 1 + valuation, W = degree - valuation;
        char* w = (char*)workspace;
        FuchsTableParams tp;
        tp.order = order;
        tp.degree = degree;
        tp.valuation = valuation;
        tp.n = n;
        tp.ode = CArr{const_cast<uint32_t*>(ode_mant), (Meta*)const_cast<cb_meta*>(ode_meta), (size_t)2};
        tp.tbl = carve(w, L_, rows * (W + 2), 2);
        tp.tmp = carve(w, L_, 2, 2 * rows);
        tp.kind = (int32_t*)w;
        tp.row0 = n_begin + valuation;
        tp.nrows = rows - tp.row0;
        tp.ninst = 1;
        tp.tbl_row0 = 0;
        w += align256(sizeof(int32_t) * (rows + 1));
        g_launches.fetch_add(1);
        int rc = table_for(L_)->fuchs_table(tp, (cudaStream_t)stream);
        if (rc) return rc;
        FuchsSharedParams sp;
        sp.order = order;
        sp.degree = degree;
        sp.valuation = valuation;
        sp.n = n;
        sp.batch = batch;
        sp.tbl = tp.tbl;
        sp.res = CArr{res_mant, (Meta*)res_meta, (size_t)(2 * batch)};
        sp.kind = tp.kind;
        sp.gcols = (uint32_t*)w;
        w += align256((size_t)2 * (2 * L_ + 3) * gcols_columns(batch) * sizeof(uint32_t));
        sp.sched = (int*)w;  // unit counter + one progress word per instance block (persistent launch)
        sp.tbl_row0 = 0;
        sp.kind_stride = 1;
        sp.per_inst = 0;
        sp.range_len = 32;
        if (const char* re = getenv("CB200_RANGE")) { int v = atoi(re); if (v >= 4 && v <= 4096) sp.range_len = v; }  // development
        sp.n_begin = n_begin;
        {
            // the reduced-radix fused dot product (cb_rr.cuh) is bit-identical but measured slower than
            // the carry-chain path inside this kernel: opt-in for experiments
            const char* e = getenv("CB200_RR");
            sp.no_rr = (e && e[0] == '1') ? 0 : 1;
            sp.stats = nullptr;
            const char* st = getenv("CB200_RR_STATS");
            if (st && st[0] == '1') {
                static unsigned long long* dstats = nullptr;
                if (!dstats && cudaMalloc(&dstats, 4 * sizeof(unsigned long long)) == cudaSuccess)
                    cudaMemset(dstats, 0, 4 * sizeof(unsigned long long));
                sp.stats = dstats;
                g_rr_stats = dstats;
            }
        }
        g_launches.fetch_add(batch > 0);
        return table_for(L_)->fuchs_shared(sp, (cudaStream_t)stream);
    }
    // an operator per instance (parameter sweeps): the same two kernels, with a table row per instance built for
    // one range of coefficients at a time (the rows of a range are independent of each other, so the divisions of
    // src/fuchs_solver.c:137 run massively parallel instead of inside the sequential recurrence)
    {
        const int64_t rows = n + 1 + valuation, W = degree - valuation, S = 2 * batch;
        const int64_t rt = pi_table_rows(L_, W, rows, batch);
        char* w = (char*)workspace;
        CArr tbl = carve(w, L_, rt * (W + 2), S);
        CArr ttmp = carve(w, L_, 2, 2 * rt * batch);
        int32_t* kind = (int32_t*)w;
        w += align256(sizeof(int32_t) * (size_t)(rt * batch + 1));
        uint32_t* gcols = (uint32_t*)w;
        for (int64_t lo = n_begin; lo <= n; lo += rt) {
            const int64_t hi = lo + rt - 1 < n ? lo + rt - 1 : n;
            FuchsTableParams tp;
            tp.order = order;
            tp.degree = degree;
            tp.valuation = valuation;
            tp.n = n;
            tp.ode = CArr{const_cast<uint32_t*>(ode_mant), (Meta*)const_cast<cb_meta*>(ode_meta), (size_t)S};
            tp.tbl = tbl;
            tp.tmp = ttmp;
            tp.kind = kind;
            tp.row0 = lo + valuation;
            tp.nrows = hi - lo + 1;
            tp.ninst = batch;
            tp.tbl_row0 = lo + valuation;
            g_launches.fetch_add(batch > 0);
            int rc = table_for(L_)->fuchs_table(tp, (cudaStream_t)stream);
            if (rc) return rc;
            FuchsSharedParams sp;
            sp.order = order;
            sp.degree = degree;
            sp.valuation = valuation;
            sp.n = hi;
            sp.batch = batch;
            sp.tbl = tbl;
            sp.res = CArr{res_mant, (Meta*)res_meta, (size_t)S};
            sp.kind = kind;
            sp.gcols = gcols;
            sp.no_rr = 1;
            sp.stats = nullptr;
            sp.n_begin = lo;
            sp.range_len = 32;
            sp.sched = nullptr;
            sp.tbl_row0 = lo + valuation;
            sp.kind_stride = batch;
            sp.per_inst = 1;
            g_launches.fetch_add(batch > 0);
            rc = table_for(L_)->fuchs_shared(sp, (cudaStream_t)stream);
            if (rc) return rc;
        }
        return 0;
    }
}

int cb200_fuchs_intop_batch_dev(int L_, int order, int degree, int valuation, int64_t n, int64_t batch,
                                const int64_t* ode_int, int shared_ode, uint32_t* res_mant, cb_meta* res_meta,
                                void* stream) {
    if (!supported_L(L_) || order < 1 || degree < 0 || batch < 0 || n < 0) return CB200_EINVAL;
    if ((order + 1) * (degree + 1) > INTOP_MAXENT) return CB200_EINVAL;
    if (valuation >= 0) return CB200_EVALUATION;
    if (!factor_fits(n, order)) return CB200_ERANGE;
    if (n < -valuation) return 0;
    FuchsIntParams p;
    p.order = order;
    p.degree = degree;
    p.valuation = valuation;
    p.n = n;
    p.batch = batch;
    p.ode = ode_int;
    p.ostride = shared_ode ? 0 : (size_t)batch;
    p.res = CArr{res_mant, (Meta*)res_meta, (size_t)(2 * batch)};
    p.range_len = 32;
    p.sched = nullptr;
    if (batch == 0) return 0;
    g_launches.fetch_add(1);
    // Stream-ordered scratch: the global product column of the general body (L + 3 words per lane, lanes padded to
    // whole blocks) and, for the persistent form, the scheduling words.  The persistent form (work units of one
    // instance block x 32 coefficients, one CTA per SM) runs when the batch is more than one wave of blocks with a
    // ragged last wave.
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int tpb = intop_tpb(L_);
    const int64_t nblocks = (batch + tpb - 1) / tpb, waves = (nblocks + sms - 1) / sms;
    const char* e = getenv("CB200_PERSISTENT");
    bool want = nblocks > sms && (double)(waves * sms - nblocks) > 0.04 * (double)nblocks;
    if (e && e[0] == '0') want = false;
    if (e && e[0] == '1') want = true;
    const size_t lanes = (size_t)nblocks * tpb;
    const size_t gbytes = align256((size_t)(L_ + 3) * lanes * sizeof(uint32_t));
    char* scratch = nullptr;
    CUDA_RET(cudaMallocAsync((void**)&scratch, gbytes + sizeof(int) * (size_t)(nblocks + 1), st));
    p.gs0 = (uint32_t*)scratch;
    p.gs0_stride = lanes;
    if (want) {
        p.sched = (int*)(scratch + gbytes);
        cudaError_t me = cudaMemsetAsync(p.sched, 0, sizeof(int) * (size_t)(nblocks + 1), st);
        if (me != cudaSuccess) {
            cudaFreeAsync(scratch, st);
            return (int)me;
        }
    }
    int rc = table_for(L_)->fuchs_intop(p, st);
    cudaFreeAsync(scratch, st);
    return rc;
}

int cb200_frobenius1_rhs_batch_dev(int L_, int order, int degree, int valuation, int64_t n, int64_t batch,
                                   const uint32_t* ode_mant, const cb_meta* ode_meta, int shared_ode,
                                   const uint32_t* rho_mant, const cb_meta* rho_meta, int shared_rho,
                                   const uint32_t* rhs_mant, const cb_meta* rhs_meta, int64_t rhs_len, int shared_rhs,
                                   uint32_t* res_mant, cb_meta* res_meta, void* workspace, size_t workspace_bytes,
                                   void* stream) {
    if (!supported_L(L_) || order < 1 || degree < 1 || batch < 0 || n < 0 || rhs_len < 0) return CB200_EINVAL;
    if (rhs_len > 0 && (!rhs_mant || !rhs_meta)) return CB200_EINVAL;
    if (workspace_bytes < cb200_frobenius_workspace_bytes(L_, batch)) return CB200_ENOMEM;
    FrobParams p;
    p.order = order;
    p.degree = degree;
    p.valuation = valuation;
    p.n = n;
    p.batch = batch;
    p.ode = CArr{const_cast<uint32_t*>(ode_mant), (Meta*)const_cast<cb_meta*>(ode_meta),
                 shared_ode ? (size_t)2 : (size_t)(2 * batch)};
    p.rho = CArr{const_cast<uint32_t*>(rho_mant), (Meta*)const_cast<cb_meta*>(rho_meta),
                 shared_rho ? (size_t)2 : (size_t)(2 * batch)};
    p.res = CArr{res_mant, (Meta*)res_meta, (size_t)(2 * batch)};
    p.rhs = CArr{const_cast<uint32_t*>(rhs_mant), (Meta*)const_cast<cb_meta*>(rhs_meta),
                 shared_rhs ? (size_t)2 : (size_t)(2 * batch)};
    p.rhs_len = rhs_len;
    char* w = (char*)workspace;
    p.tmp = carve(w, L_, FROB_NTMP, 2 * batch);
    p.gcols = (uint32_t*)w;
    p.gcols_words = hyb_tail_words(L_, batch);
    w += hyb_tail_bytes(L_, batch);
    p.sched = (int*)w;
    g_launches.fetch_add(batch > 0);
    return table_for(L_)->frobenius1(p, (cudaStream_t)stream);
}
int cb200_frobenius1_batch_dev(int L_, int order, int degree, int valuation, int64_t n, int64_t batch,
                               const uint32_t* ode_mant, const cb_meta* ode_meta, int shared_ode,
                               const uint32_t* rho_mant, const cb_meta* rho_meta, int shared_rho,
                               uint32_t* res_mant, cb_meta* res_meta, void* workspace,
                               size_t workspace_bytes, void* stream) {
    return cb200_frobenius1_rhs_batch_dev(L_, order, degree, valuation, n, batch, ode_mant, ode_meta, shared_ode, rho_mant,
                                          rho_meta, shared_rho, nullptr, nullptr, 0, 0, res_mant, res_meta, workspace,
                                          workspace_bytes, stream);
}

// ------------------------------------------------------------------ general Frobenius (M > 1)
static int64_t frobg_plen(int order, int64_t n) { return (int64_t)order * n + order + 2; }
size_t cb200_frobenius_general_workspace_bytes(int L, int order, int degree, int M, int64_t n, int64_t batch) {
    return carve_bytes(L, (int64_t)(degree + 3) * frobg_plen(order, n) + FROBG_NTMP + M, 2 * batch);
}
int cb200_frobenius_general_batch_dev(int L_, int order, int degree, int valuation, int64_t n, int64_t batch,
                                      const uint32_t* ode_mant, const cb_meta* ode_meta, int shared_ode,
                                      const uint32_t* rho_mant, const cb_meta* rho_meta, int shared_rho, int mul,
                                      int M, uint32_t* gens_mant, cb_meta* gens_meta, void* workspace,
                                      size_t workspace_bytes, void* stream) {
    if (!supported_L(L_) || order < 1 || degree < 1 || degree > FROBG_MAXD || batch < 0 || n < 0 || M < 1 ||
        M > FROBG_MAXM || mul < 1 || mul > M)
        return CB200_EINVAL;
    if (workspace_bytes < cb200_frobenius_general_workspace_bytes(L_, order, degree, M, n, batch)) return CB200_ENOMEM;
    FrobGenParams p;
    p.order = order;
    p.degree = degree;
    p.valuation = valuation;
    p.M = M;
    p.mul = mul;
    p.n = n;
    p.batch = batch;
    p.plen = frobg_plen(order, n);
    p.ode = CArr{const_cast<uint32_t*>(ode_mant), (Meta*)const_cast<cb_meta*>(ode_meta),
                 shared_ode ? (size_t)2 : (size_t)(2 * batch)};
    p.rho = CArr{const_cast<uint32_t*>(rho_mant), (Meta*)const_cast<cb_meta*>(rho_meta),
                 shared_rho ? (size_t)2 : (size_t)(2 * batch)};
    p.gens = CArr{gens_mant, (Meta*)gens_meta, (size_t)(2 * batch)};
    char* w = (char*)workspace;
    p.wp = carve(w, L_, (int64_t)(degree + 3) * p.plen + FROBG_NTMP + M, 2 * batch);
    g_launches.fetch_add(batch > 0);
    return table_for(L_)->frobenius_general(p, (cudaStream_t)stream);
}

size_t cb200_solution_op_workspace_bytes(int L, int M, int64_t batch) {
    return carve_bytes(L, FROBG_NTMP + (M > 0 ? M : 0), 2 * batch);
}
int cb200_solution_op_dev(int which, int L_, int64_t batch, int mul, int M, int64_t cap, int64_t nu,
                          const uint32_t* rho_mant, const cb_meta* rho_meta, uint32_t* gens_mant, cb_meta* gens_meta,
                          uint32_t* f_mant, cb_meta* f_meta, int64_t fcap, void* workspace, size_t workspace_bytes,
                          void* stream) {
    if (!supported_L(L_) || batch < 0 || M < 0 || M > FROBG_MAXM || cap < 0 || fcap < 0) return CB200_EINVAL;
    if (which != CB200_SOL_EXTEND && which != CB200_SOL_UPDATE && which != CB200_SOL_NORMALIZE) return CB200_EINVAL;
    if (which == CB200_SOL_EXTEND && (nu < 0 || nu >= cap)) return CB200_EINVAL;
    if (which == CB200_SOL_NORMALIZE && (mul < 1 || mul > M)) return CB200_EINVAL;
    if (workspace_bytes < cb200_solution_op_workspace_bytes(L_, M, batch)) return CB200_ENOMEM;
    SolOpParams p;
    p.which = which;
    p.M = M;
    p.mul = mul;
    p.batch = batch;
    p.cap = cap;
    p.fcap = which == CB200_SOL_NORMALIZE ? 0 : fcap;
    p.nu = nu;
    p.rho = CArr{const_cast<uint32_t*>(rho_mant), (Meta*)const_cast<cb_meta*>(rho_meta), (size_t)(2 * batch)};
    p.gens = CArr{gens_mant, (Meta*)gens_meta, (size_t)(2 * batch)};
    p.f = CArr{f_mant, (Meta*)f_meta, (size_t)(2 * batch)};
    char* w = (char*)workspace;
    p.wp = carve(w, L_, FROBG_NTMP + M, 2 * batch);
    g_launches.fetch_add(batch > 0);
    return table_for(L_)->solution_op(p, (cudaStream_t)stream);
}

int cb200_indicial_polynomial_dev(int L_, int order, int degree, int valuation, int64_t batch,
                                  const uint32_t* ode_mant, const cb_meta* ode_meta, int shared_ode, int64_t nu,
                                  int64_t shift, uint32_t* out_mant, cb_meta* out_meta, void* workspace,
                                  size_t workspace_bytes, void* stream) {
    if (!supported_L(L_) || order < 1 || degree < 0 || batch < 0) return CB200_EINVAL;
    if (workspace_bytes < cb200_solution_op_workspace_bytes(L_, 0, batch)) return CB200_ENOMEM;
    IndicialParams p;
    p.order = order;
    p.degree = degree;
    p.valuation = valuation;
    p.batch = batch;
    p.nu = nu;
    p.shift = shift;
    p.ode = CArr{const_cast<uint32_t*>(ode_mant), (Meta*)const_cast<cb_meta*>(ode_meta),
                 shared_ode ? (size_t)2 : (size_t)(2 * batch)};
    p.out = CArr{out_mant, (Meta*)out_meta, (size_t)(2 * batch)};
    char* w = (char*)workspace;
    p.wp = carve(w, L_, FROBG_NTMP, 2 * batch);
    g_launches.fetch_add(batch > 0);
    return table_for(L_)->indicial_poly(p, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ acb_ode_solution_evaluate, acb_exp / acb_log
size_t cb200_evaluate_workspace_bytes(int L, int64_t npts) { return carve_bytes(L, TRANS_NTMP + 4, 2 * npts); }
int cb200_solution_evaluate_dev(int L_, int M, int64_t cap, int64_t npts, const uint32_t* gens_mant,
                                const cb_meta* gens_meta, const uint32_t* rho_mant, const cb_meta* rho_meta, int pow_mode,
                                uint64_t pow_e, const uint32_t* pts_mant, const cb_meta* pts_meta, uint32_t* out_mant,
                                cb_meta* out_meta, void* workspace, size_t workspace_bytes, void* stream) {
    if (!supported_L(L_) || M < 1 || M > FROBG_MAXM || cap < 1 || npts < 0 || pow_mode < 0 || pow_mode > 2 ||
        (pow_mode == 1 && pow_e >= (1ull << 31)))
        return CB200_EINVAL;
    if (workspace_bytes < cb200_evaluate_workspace_bytes(L_, npts)) return CB200_ENOMEM;
    EvalParams p;
    p.M = M;
    p.pow_mode = pow_mode;
    p.pow_e = pow_e;
    p.cap = cap;
    p.npts = npts;
    p.gens = CArr{const_cast<uint32_t*>(gens_mant), (Meta*)const_cast<cb_meta*>(gens_meta), (size_t)2};
    p.rho = CArr{const_cast<uint32_t*>(rho_mant), (Meta*)const_cast<cb_meta*>(rho_meta), (size_t)2};
    p.pts = CArr{const_cast<uint32_t*>(pts_mant), (Meta*)const_cast<cb_meta*>(pts_meta), (size_t)(2 * npts)};
    p.out = CArr{out_mant, (Meta*)out_meta, (size_t)(2 * npts)};
    char* w = (char*)workspace;
    p.wp = carve(w, L_, TRANS_NTMP + 4, 2 * npts);
    g_launches.fetch_add(npts > 0);
    return table_for(L_)->evaluate(p, (cudaStream_t)stream);
}
int cb200_trans_dev(int which, int L_, int64_t n, uint32_t* z_mant, cb_meta* z_meta, const uint32_t* x_mant,
                    const cb_meta* x_meta, void* workspace, size_t workspace_bytes, void* stream) {
    if (!supported_L(L_) || n < 0 || (which != CB200_TRANS_EXP && which != CB200_TRANS_LOG)) return CB200_EINVAL;
    if (workspace_bytes < cb200_evaluate_workspace_bytes(L_, n)) return CB200_ENOMEM;
    TransParams p;
    p.which = which;
    p.n = n;
    p.z = CArr{z_mant, (Meta*)z_meta, (size_t)(2 * n)};
    p.x = CArr{const_cast<uint32_t*>(x_mant), (Meta*)const_cast<cb_meta*>(x_meta), (size_t)(2 * n)};
    char* w = (char*)workspace;
    p.wp = carve(w, L_, TRANS_NTMP, 2 * n);
    g_launches.fetch_add(n > 0);
    return table_for(L_)->trans(p, (cudaStream_t)stream);
}

size_t cb200_ode_apply_workspace_bytes(int L, int64_t len_in, int64_t batch) { return carve_bytes(L, len_in + 1, 2 * batch); }
int cb200_ode_apply_dev(int L_, int order, int degree, int64_t batch, const uint32_t* ode_mant, const cb_meta* ode_meta,
                        int shared_ode, const uint32_t* in_mant, const cb_meta* in_meta, int64_t len_in, uint32_t* out_mant,
                        cb_meta* out_meta, int64_t out_len, int64_t deg, int32_t* solved, void* workspace,
                        size_t workspace_bytes, void* stream) {
    if (!supported_L(L_) || order < 0 || degree < 0 || batch < 0 || len_in < 0 || out_len < 0 || deg < 0) return CB200_EINVAL;
    if (workspace_bytes < cb200_ode_apply_workspace_bytes(L_, len_in, batch)) return CB200_ENOMEM;
    ApplyParams p;
    p.order = order;
    p.degree = degree;
    p.batch = batch;
    p.len_in = len_in;
    p.out_len = out_len;
    p.deg = deg;
    p.ode = CArr{const_cast<uint32_t*>(ode_mant), (Meta*)const_cast<cb_meta*>(ode_meta),
                 shared_ode ? (size_t)2 : (size_t)(2 * batch)};
    p.in = CArr{const_cast<uint32_t*>(in_mant), (Meta*)const_cast<cb_meta*>(in_meta), (size_t)(2 * batch)};
    p.out = CArr{out_mant, (Meta*)out_meta, (size_t)(2 * batch)};
    char* w = (char*)workspace;
    p.wp = carve(w, L_, len_in + 1, 2 * batch);
    p.solved = solved;
    g_launches.fetch_add(batch > 0);
    return table_for(L_)->apply(p, (cudaStream_t)stream);
}

int cb200_horner_batch_dev(int L_, int64_t len, int64_t npts, int nder, const uint32_t* f_mant,
                           const cb_meta* f_meta, int shared_f, const uint32_t* pts_mant,
                           const cb_meta* pts_meta, uint32_t* out_mant, cb_meta* out_meta,
                           void* workspace, size_t workspace_bytes, void* stream) {
    if (!supported_L(L_) || len < 0 || npts < 0 || nder < 1 || nder > HORNER_MAXDER) return CB200_EINVAL;
    if (workspace_bytes < cb200_horner_workspace_bytes(L_, npts)) return CB200_ENOMEM;
    HornerParams p;
    p.nder = nder;
    p.len = len;
    p.npts = npts;
    p.f = CArr{const_cast<uint32_t*>(f_mant), (Meta*)const_cast<cb_meta*>(f_meta),
               shared_f ? (size_t)2 : (size_t)(2 * npts)};
    p.pts = CArr{const_cast<uint32_t*>(pts_mant), (Meta*)const_cast<cb_meta*>(pts_meta), (size_t)(2 * npts)};
    p.out = CArr{out_mant, (Meta*)out_meta, (size_t)(2 * npts)};
    char* w = (char*)workspace;
    p.tmp = carve(w, L_, 1, 2 * npts);
    p.gcols = (uint32_t*)w;
    p.gcols_words = hyb_tail_words(L_, npts);
    g_launches.fetch_add(npts > 0);
    return table_for(L_)->horner(p, (cudaStream_t)stream);
}

int cb200_ew_dev(int op, int L_, int64_t n, uint32_t* z_mant, cb_meta* z_meta, const uint32_t* x_mant,
                 const cb_meta* x_meta, const uint32_t* y_mant, const cb_meta* y_meta, const int64_t* k,
                 void* workspace, size_t workspace_bytes, void* stream) {
    if (!supported_L(L_) || n < 0) return CB200_EINVAL;
    bool binary = op == CB200_OP_ADD || op == CB200_OP_SUB || op == CB200_OP_MUL || op == CB200_OP_DIV;
    bool unary = op == CB200_OP_INV, si = op == CB200_OP_MUL_SI || op == CB200_OP_ADD_SI;
    if (!(binary || unary || si)) return CB200_EINVAL;
    if (si && !k) return CB200_EINVAL;
    if (workspace_bytes < cb200_ew_workspace_bytes(L_, n)) return CB200_ENOMEM;
    EwParams p;
    p.op = op;
    p.n = n;
    p.z = CArr{z_mant, (Meta*)z_meta, (size_t)(2 * n)};
    p.x = CArr{const_cast<uint32_t*>(x_mant), (Meta*)const_cast<cb_meta*>(x_meta), (size_t)(2 * n)};
    p.y = binary ? CArr{const_cast<uint32_t*>(y_mant), (Meta*)const_cast<cb_meta*>(y_meta), (size_t)(2 * n)} : p.x;
    p.k = k;
    char* w = (char*)workspace;
    p.tmp = carve(w, L_, 2, 2 * n);
    g_launches.fetch_add(n > 0);
    return table_for(L_)->ew(p, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ radius_of_convergence, truncation_order
size_t cb200_radius_workspace_bytes(int L, int degree, int64_t batch) {
    return carve_bytes(L, 2 * (int64_t)(degree + 1) + RADIUS_NTMP + TRANS_NTMP, 2 * batch);
}
int cb200_radius_of_convergence_dev(int L_, int order, int degree, int64_t batch, const uint32_t* ode_mant,
                                    const cb_meta* ode_meta, int shared_ode, int64_t n, uint32_t* out_mant, cb_meta* out_meta,
                                    int32_t* n_done, void* workspace, size_t workspace_bytes, void* stream) {
    if (!supported_L(L_) || order < 0 || degree < 0 || batch < 0 || n < 0 || !n_done) return CB200_EINVAL;
    if (workspace_bytes < cb200_radius_workspace_bytes(L_, degree, batch)) return CB200_ENOMEM;
    RadiusParams p;
    p.order = order;
    p.degree = degree;
    p.batch = batch;
    p.n = n;
    p.ode = CArr{const_cast<uint32_t*>(ode_mant), (Meta*)const_cast<cb_meta*>(ode_meta), shared_ode ? (size_t)2 : (size_t)(2 * batch)};
    p.out = CArr{out_mant, (Meta*)out_meta, (size_t)(2 * batch)};
    char* w = (char*)workspace;
    p.wp = carve(w, L_, 2 * (int64_t)(degree + 1) + RADIUS_NTMP + TRANS_NTMP, 2 * batch);
    p.n_done = n_done;
    g_launches.fetch_add(batch > 0);
    return table_for(L_)->radius(p, (cudaStream_t)stream);
}
size_t cb200_truncation_workspace_bytes(int L, int64_t batch) { return carve_bytes(L, 8 + TRANS_NTMP, 2 * batch); }
int cb200_truncation_order_dev(int L_, int64_t batch, const uint32_t* eta_mant, const cb_meta* eta_meta,
                               const uint32_t* alpha_mant, const cb_meta* alpha_meta, int64_t bits, uint32_t* n_mant,
                               cb_meta* n_meta, int64_t* n_out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!supported_L(L_) || batch < 0 || bits < 1 || !n_out) return CB200_EINVAL;
    if (workspace_bytes < cb200_truncation_workspace_bytes(L_, batch)) return CB200_ENOMEM;
    TruncParams p;
    p.batch = batch;
    p.bits = bits;
    p.eta = CArr{const_cast<uint32_t*>(eta_mant), (Meta*)const_cast<cb_meta*>(eta_meta), (size_t)(2 * batch)};
    p.alpha = CArr{const_cast<uint32_t*>(alpha_mant), (Meta*)const_cast<cb_meta*>(alpha_meta), (size_t)(2 * batch)};
    p.nball = CArr{n_mant, (Meta*)n_meta, (size_t)(2 * batch)};
    char* w = (char*)workspace;
    p.wp = carve(w, L_, 8 + TRANS_NTMP, 2 * batch);
    p.n_out = n_out;
    g_launches.fetch_add(batch > 0);
    return table_for(L_)->truncation(p, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ acb_ode_shift, acb_poly_taylor_shift
size_t cb200_ode_shift_workspace_bytes(int L, int64_t batch) { return carve_bytes(L, 1, 2 * batch); }
int cb200_ode_shift_dev(int L_, int order, int degree, int64_t batch, const uint32_t* ode_mant, const cb_meta* ode_meta,
                        int shared_ode, const uint32_t* a_mant, const cb_meta* a_meta, int shared_a, uint32_t* out_mant,
                        cb_meta* out_meta, void* workspace, size_t workspace_bytes, void* stream) {
    if (!supported_L(L_) || order < 0 || degree < 0 || batch < 0) return CB200_EINVAL;
    if (workspace_bytes < cb200_ode_shift_workspace_bytes(L_, batch)) return CB200_ENOMEM;
    ShiftParams p;
    p.order = order;
    p.degree = degree;
    p.batch = batch;
    p.in = CArr{const_cast<uint32_t*>(ode_mant), (Meta*)const_cast<cb_meta*>(ode_meta), shared_ode ? (size_t)2 : (size_t)(2 * batch)};
    p.a = CArr{const_cast<uint32_t*>(a_mant), (Meta*)const_cast<cb_meta*>(a_meta), shared_a ? (size_t)2 : (size_t)(2 * batch)};
    p.out = CArr{out_mant, (Meta*)out_meta, (size_t)(2 * batch)};
    char* w = (char*)workspace;
    p.tmp = carve(w, L_, 1, 2 * batch);
    g_launches.fetch_add(batch > 0);
    return table_for(L_)->ode_shift(p, (cudaStream_t)stream);
}
size_t cb200_taylor_shift_workspace_bytes(int L, int64_t len, int64_t batch) { return carve_bytes(L, len > 0 ? len : 1, 2 * batch); }
int cb200_taylor_shift_dev(int L_, int64_t len, int64_t batch, uint32_t* f_mant, cb_meta* f_meta, const uint32_t* a_mant,
                           const cb_meta* a_meta, int shared_a, void* workspace, size_t workspace_bytes, void* stream) {
    if (!supported_L(L_) || len < 0 || batch < 0) return CB200_EINVAL;
    if (workspace_bytes < cb200_taylor_shift_workspace_bytes(L_, len, batch)) return CB200_ENOMEM;
    TShiftParams p;
    p.len = len;
    p.batch = batch;
    p.f = CArr{f_mant, (Meta*)f_meta, (size_t)(2 * batch)};
    p.a = CArr{const_cast<uint32_t*>(a_mant), (Meta*)const_cast<cb_meta*>(a_meta), shared_a ? (size_t)2 : (size_t)(2 * batch)};
    char* w = (char*)workspace;
    p.tmp = carve(w, L_, len > 0 ? len : 1, 2 * batch);
    g_launches.fetch_add(batch > 0 && len > 1);
    return table_for(L_)->taylor_shift(p, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ analytic continuation
static size_t fuchs_ws_any_valuation(int L, int order, int degree, int64_t n, int64_t batch) {
    size_t m = 0;
    for (int v = -order; v <= -1; v++) {
        size_t b = cb200_fuchs_workspace_bytes(L, order, degree, v, n, batch, 1);
        if (b > m) m = b;
    }
    return m;
}
static size_t continuation_ws(int L, int order, int degree, int64_t n, int64_t batch, bool series) {
    int64_t ne = (int64_t)(order + 1) * (degree + 1);
    return carve_bytes(L, ne, 2)            /* shifted operator */
           + carve_bytes(L, 3, 2)           /* step, and two single-ball temporaries */
           + carve_bytes(L, n + 1, 2 * batch) /* the series of the current step */
           + align256(fuchs_ws_any_valuation(L, order, degree, n, batch)) + align256(cb200_horner_workspace_bytes(L, batch)) +
           align256(cb200_ew_workspace_bytes(L, 1)) + (series ? align256(cb200_taylor_shift_workspace_bytes(L, n + 1, batch)) : 0) +
           1024;
}
size_t cb200_continuation_workspace_bytes(int L, int order, int degree, int64_t n, int64_t batch) {
    return continuation_ws(L, order, degree, n, batch, false);
}
size_t cb200_continuation_series_workspace_bytes(int L, int order, int degree, int64_t n, int64_t batch) {
    return continuation_ws(L, order, degree, n, batch, true);
}

// acb_ode_valuation (reference src/acb_ode.c:159-181) from the meta records of one operator (S = 2)
static int valuation_of(const cb_meta* t, int order, int degree) {
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

static int continuation_impl(int L_, int order, int degree, int64_t n, int64_t batch, int64_t path_len,
                             const uint32_t* ode_mant, const cb_meta* ode_meta, const uint32_t* path_mant,
                             const cb_meta* path_meta, uint32_t* y_mant, cb_meta* y_meta, uint32_t* full_mant,
                             cb_meta* full_meta, void* workspace, size_t workspace_bytes, void* stream) {
    if (!supported_L(L_) || order < 1 || order > HORNER_MAXDER || degree < 0 || batch < 0 || n < order || path_len < 1)
        return CB200_EINVAL;
    const bool series = full_mant != nullptr;
    if (workspace_bytes < continuation_ws(L_, order, degree, n, batch, series)) return CB200_ENOMEM;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t ne = (int64_t)(order + 1) * (degree + 1), S = 2 * batch;
    char* w = (char*)workspace;
    CArr sh = carve(w, L_, ne, 2);
    CArr tb = carve(w, L_, 3, 2);  // [0] step a, [1] temporary of the operator shift
    CArr ser = carve(w, L_, n + 1, S);
    size_t fw = fuchs_ws_any_valuation(L_, order, degree, n, batch);
    void* fws = w;
    w += align256(fw);
    size_t hw = cb200_horner_workspace_bytes(L_, batch);
    void* hws = w;
    w += align256(hw);
    size_t ew = cb200_ew_workspace_bytes(L_, 1);
    void* ews = w;
    w += align256(ew);
    void* tws = w;
    const size_t tw = series ? cb200_taylor_shift_workspace_bytes(L_, n + 1, batch) : 0;
    auto ball_m = [&](const CArr& a, int64_t e) { return a.m + (size_t)e * L_ * a.S; };
    auto ball_t = [&](const CArr& a, int64_t e) { return (cb_meta*)(a.t + (size_t)e * a.S); };
    std::vector<cb_meta> hmeta((size_t)ne * 2);
    for (int64_t t = 0; t + 1 < path_len; t++) {
        // acb_ode_shift(ODE_shift, ODE, path + t): one launch (an exactly zero corner leaves the operator as it is)
        const uint32_t* am = path_mant + (size_t)t * L_ * 2;
        const cb_meta* at_ = path_meta + (size_t)t * 2;
        ShiftParams shp;
        shp.order = order;
        shp.degree = degree;
        shp.batch = 1;
        shp.in = CArr{const_cast<uint32_t*>(ode_mant), (Meta*)const_cast<cb_meta*>(ode_meta), (size_t)2};
        shp.a = CArr{const_cast<uint32_t*>(am), (Meta*)const_cast<cb_meta*>(at_), (size_t)2};
        shp.out = sh;
        shp.tmp = CArr{ball_m(tb, 1), (Meta*)ball_t(tb, 1), (size_t)2};
        g_launches.fetch_add(1);
        int rc = table_for(L_)->ode_shift(shp, st);
        if (rc) return rc;
        // acb_ode_solve_fuchs starts from the valuation of the SHIFTED operator (src/fuchs_solver.c:105): 288 bytes of
        // meta records come back to the host, once per corner (the only synchronisation of the pipeline)
        CUDA_RET(cudaMemcpyAsync(hmeta.data(), sh.t, arr_meta_bytes(ne, 2), cudaMemcpyDeviceToHost, st));
        CUDA_RET(cudaStreamSynchronize(st));
        const int v = valuation_of(hmeta.data(), order, degree);
        if (v >= 0) return CB200_EVALUATION;
        // res[0 .. -v-1] = y (the reference reads only those of res as initial values)
        CUDA_RET(cudaMemcpyAsync(ser.m, y_mant, arr_mant_bytes(L_, -v, S), cudaMemcpyDeviceToDevice, st));
        CUDA_RET(cudaMemcpyAsync(ser.t, y_meta, arr_meta_bytes(-v, S), cudaMemcpyDeviceToDevice, st));
        rc = cb200_fuchs_batch_dev(L_, order, degree, v, n, batch, sh.m, (cb_meta*)sh.t, 1, ser.m, (cb_meta*)ser.t, fws, fw,
                                   stream);
        if (rc) return rc;
        // a = path[t+1] - path[t]
        rc = cb200_ew_dev(CB200_OP_SUB, L_, 1, ball_m(tb, 0), ball_t(tb, 0), path_mant + (size_t)(t + 1) * L_ * 2,
                          path_meta + (size_t)(t + 1) * 2, am, at_, nullptr, ews, ew, stream);
        if (rc) return rc;
        if (series && t + 2 == path_len) {
            // the last corner: the complete acb_poly_taylor_shift, as the reference leaves it in res (:159)
            rc = cb200_taylor_shift_dev(L_, n + 1, batch, ser.m, (cb_meta*)ser.t, ball_m(tb, 0), ball_t(tb, 0), 1, tws, tw, stream);
            if (rc) return rc;
            CUDA_RET(cudaMemcpyAsync(full_mant, ser.m, arr_mant_bytes(L_, n + 1, S), cudaMemcpyDeviceToDevice, st));
            CUDA_RET(cudaMemcpyAsync(full_meta, ser.t, arr_meta_bytes(n + 1, S), cudaMemcpyDeviceToDevice, st));
            CUDA_RET(cudaMemcpyAsync(y_mant, ser.m, arr_mant_bytes(L_, order, S), cudaMemcpyDeviceToDevice, st));
            CUDA_RET(cudaMemcpyAsync(y_meta, ser.t, arr_meta_bytes(order, S), cudaMemcpyDeviceToDevice, st));
            break;
        }
        // y[k] = res^(k)(a)/k! for k < order (all that the next step reads)
        HornerParams hp;
        hp.nder = order;
        hp.len = n + 1;
        hp.npts = batch;
        hp.f = ser;
        hp.pts = CArr{ball_m(tb, 0), (Meta*)ball_t(tb, 0), (size_t)2};
        hp.out = CArr{y_mant, (Meta*)y_meta, (size_t)S};
        char* hwp = (char*)hws;
        hp.tmp = carve(hwp, L_, 1, S);
        hp.gcols = (uint32_t*)hwp;
        hp.gcols_words = hyb_tail_words(L_, batch);
        g_launches.fetch_add(batch > 0);
        rc = table_for(L_)->horner(hp, st);
        if (rc) return rc;
    }
    return 0;
}

int cb200_continuation_dev(int L_, int order, int degree, int64_t n, int64_t batch, int64_t path_len,
                           const uint32_t* ode_mant, const cb_meta* ode_meta, const uint32_t* path_mant,
                           const cb_meta* path_meta, const uint8_t* corner_is_zero, uint32_t* y_mant, cb_meta* y_meta,
                           void* workspace, size_t workspace_bytes, void* stream) {
    (void)corner_is_zero;  // kept for ABI stability: the shift kernel recognises an exactly zero corner itself
    return continuation_impl(L_, order, degree, n, batch, path_len, ode_mant, ode_meta, path_mant, path_meta, y_mant, y_meta,
                             nullptr, nullptr, workspace, workspace_bytes, stream);
}
int cb200_continuation_series_dev(int L_, int order, int degree, int64_t n, int64_t batch, int64_t path_len,
                                  const uint32_t* ode_mant, const cb_meta* ode_meta, const uint32_t* path_mant,
                                  const cb_meta* path_meta, uint32_t* y_mant, cb_meta* y_meta, uint32_t* res_mant,
                                  cb_meta* res_meta, void* workspace, size_t workspace_bytes, void* stream) {
    if (!res_mant || !res_meta || path_len < 2) return CB200_EINVAL;
    return continuation_impl(L_, order, degree, n, batch, path_len, ode_mant, ode_meta, path_mant, path_meta, y_mant, y_meta,
                             res_mant, res_meta, workspace, workspace_bytes, stream);
}


// ------------------------------------------------------------------ host-buffer (end-to-end) entry points
#define CK(x)                                \
    do {                                     \
        cudaError_t e_ = (x);                \
        if (e_ != cudaSuccess) return (int)e_; \
    } while (0)

// ---- pinned host memory for callers without CUDA headers (result buffers of the *_host entry points: copies into
// page-locked memory run at PCIe speed and overlap the kernels; pageable buffers work too, more slowly)
void* cb200_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    return p;
}
void cb200_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

namespace {
// Device buffers, streams and events of the host-buffer solver, kept between calls (a lazily created context:
// allocating and freeing tens of GB per call would cost a few per cent of a BASELINE-size solve).  One per process,
// guarded by a mutex; cb200_host_cache_release() frees it.
struct HostCallCache {
    std::mutex mu;
    int device = -1;
    void* buf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t cap[5] = {0, 0, 0, 0, 0};
    cudaStream_t comp = nullptr, copy = nullptr;
    std::vector<cudaEvent_t> ev;
    cudaError_t reserve(int i, size_t n) {
        if (cap[i] >= n && buf[i]) return cudaSuccess;
        if (buf[i]) cudaFree(buf[i]);
        buf[i] = nullptr;
        cap[i] = 0;
        cudaError_t e = cudaMalloc(&buf[i], n ? n : 1);
        if (e == cudaSuccess) cap[i] = n;
        return e;
    }
    void release() {
        for (int i = 0; i < 5; i++) {
            if (buf[i]) cudaFree(buf[i]);
            buf[i] = nullptr;
            cap[i] = 0;
        }
        for (auto e : ev) cudaEventDestroy(e);
        ev.clear();
        if (comp) cudaStreamDestroy(comp);
        if (copy) cudaStreamDestroy(copy);
        comp = copy = nullptr;
        device = -1;
    }
};
HostCallCache g_hcache;
}  // namespace

void cb200_host_cache_release(void) {
    std::lock_guard<std::mutex> lk(g_hcache.mu);
    if (g_hcache.device >= 0) {
        cudaSetDevice(g_hcache.device);
        g_hcache.release();
    }
}

// acb_ode_solve_fuchs for a batch, host buffers in and out (the call a host program makes; bench.py's `e2e`).
// Only the operator and the -valuation initial rows of res go up; the solve runs in coefficient RANGES on one stream
// (the recurrence resumes from the rows already in HBM) and every finished range -- contiguous rows of the
// [coefficient][limb][slot] layout -- comes down on a second stream, straight into the caller's result arrays, while
// the next range is being computed.  With page-locked result arrays (cb200_host_alloc, cudaHostAlloc, a pinned torch
// tensor) the copies run at PCIe speed; pageable arrays are staged by the driver.
int cb200_fuchs_batch_host(int L, int order, int degree, int valuation, int64_t n, int64_t batch,
                           const uint32_t* ode_mant, const cb_meta* ode_meta, int shared_ode,
                           uint32_t* res_mant, cb_meta* res_meta, int device) {
    if (!supported_L(L) || order < 1 || degree < 0 || batch < 0 || n < 0) return CB200_EINVAL;
    if (valuation >= 0) return CB200_EVALUATION;
    if (!factor_fits(n, order)) return CB200_ERANGE;
    CK(cudaSetDevice(device));
    std::lock_guard<std::mutex> lk(g_hcache.mu);
    HostCallCache& hc = g_hcache;
    if (hc.device != device) {
        if (hc.device >= 0) {
            cudaSetDevice(hc.device);
            hc.release();
            CK(cudaSetDevice(device));
        }
        hc.device = device;
    }
    if (!hc.comp) CK(cudaStreamCreateWithFlags(&hc.comp, cudaStreamNonBlocking));
    if (!hc.copy) CK(cudaStreamCreateWithFlags(&hc.copy, cudaStreamNonBlocking));
    const int64_t So = shared_ode ? 2 : 2 * batch, S = 2 * batch, ne = (int64_t)(order + 1) * (degree + 1);
    const size_t wsb = cb200_fuchs_workspace_bytes(L, order, degree, valuation, n, batch, shared_ode);
    CK(hc.reserve(0, arr_mant_bytes(L, ne, So)));
    CK(hc.reserve(1, arr_meta_bytes(ne, So)));
    CK(hc.reserve(2, arr_mant_bytes(L, n + 1, S)));
    CK(hc.reserve(3, arr_meta_bytes(n + 1, S)));
    CK(hc.reserve(4, wsb));
    uint32_t* d_om = (uint32_t*)hc.buf[0];
    cb_meta* d_ot = (cb_meta*)hc.buf[1];
    uint32_t* d_rm = (uint32_t*)hc.buf[2];
    cb_meta* d_rt = (cb_meta*)hc.buf[3];
    const int64_t ninit = (-valuation < n + 1) ? -valuation : n + 1;
    CK(cudaMemcpyAsync(d_om, ode_mant, arr_mant_bytes(L, ne, So), cudaMemcpyHostToDevice, hc.comp));
    CK(cudaMemcpyAsync(d_ot, ode_meta, arr_meta_bytes(ne, So), cudaMemcpyHostToDevice, hc.comp));
    CK(cudaMemcpyAsync(d_rm, res_mant, arr_mant_bytes(L, ninit, S), cudaMemcpyHostToDevice, hc.comp));
    CK(cudaMemcpyAsync(d_rt, res_meta, arr_meta_bytes(ninit, S), cudaMemcpyHostToDevice, hc.comp));
    if (n < -valuation) return (int)cudaStreamSynchronize(hc.comp);  // nothing to compute
    // coefficient ranges: about 1/8 of the rows each, at least 16 rows, at most 512 MB of result rows
    const size_t row_bytes = (size_t)S * (4 * (size_t)L + 16);
    int64_t rows = (n + 8) / 8;
    if (rows < 16) rows = 16;
    while (rows > 16 && (size_t)rows * row_bytes > ((size_t)512 << 20)) rows = (rows + 1) / 2;
    if (const char* e = getenv("CB200_HOST_ROWS")) { int v = atoi(e); if (v >= 1) rows = v; }  // development
    const int64_t nchunks = (n + 1 + rows - 1) / rows;
    while ((int64_t)hc.ev.size() < nchunks) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        hc.ev.push_back(e);
    }
    // every range is enqueued on the compute stream up front (kernel launches only), so that a blocking copy into
    // pageable memory below never delays the next range
    for (int64_t c = 0; c < nchunks; c++) {
        const int64_t lo = c * rows, hi = (lo + rows - 1 < n) ? lo + rows - 1 : n;
        if (hi >= -valuation) {
            int rc = cb200_fuchs_batch_range_dev(L, order, degree, valuation, lo, hi, batch, d_om, d_ot, shared_ode, d_rm, d_rt,
                                                 hc.buf[4], wsb, hc.comp);
            if (rc) return rc;
        }
        CK(cudaEventRecord(hc.ev[c], hc.comp));
    }
    for (int64_t c = 0; c < nchunks; c++) {
        int64_t lo = c * rows;
        const int64_t hi = (lo + rows - 1 < n) ? lo + rows - 1 : n;
        if (lo < ninit) lo = ninit;  // the initial rows are the caller's own
        if (lo > hi) continue;
        CK(cudaStreamWaitEvent(hc.copy, hc.ev[c], 0));
        CK(cudaMemcpyAsync(res_mant + (size_t)lo * L * S, d_rm + (size_t)lo * L * S, arr_mant_bytes(L, hi - lo + 1, S),
                           cudaMemcpyDeviceToHost, hc.copy));
        CK(cudaMemcpyAsync(res_meta + (size_t)lo * S, d_rt + (size_t)lo * S, arr_meta_bytes(hi - lo + 1, S),
                           cudaMemcpyDeviceToHost, hc.copy));
    }
    CK(cudaStreamSynchronize(hc.comp));
    return (int)cudaStreamSynchronize(hc.copy);
}

int cb200_fuchs_intop_batch_host(int L, int order, int degree, int valuation, int64_t n, int64_t batch,
                                 const int64_t* ode_int, int shared_ode, uint32_t* res_mant, cb_meta* res_meta,
                                 int device) {
    if (!supported_L(L) || order < 1 || degree < 0 || batch < 0 || n < 0) return CB200_EINVAL;
    if (valuation >= 0) return CB200_EVALUATION;
    CK(cudaSetDevice(device));
    int64_t S = 2 * batch, ne = (int64_t)(order + 1) * (degree + 1);
    {   // the int64 multipliers of the kernel must not wrap: (order+1) * max|coefficient| * (n+1)^order < 2^62
        uint64_t pmax = 0;
        for (int64_t i = 0, cnt = ne * (shared_ode ? 1 : batch); i < cnt; i++) {
            uint64_t a = ode_int[i] < 0 ? (uint64_t)0 - (uint64_t)ode_int[i] : (uint64_t)ode_int[i];
            if (a > pmax) pmax = a;
        }
        long double bound = (long double)pmax * (order + 1);
        for (int i = 0; i < order; i++) bound *= (long double)(n + 1);
        if (!(bound < 4.0e18L)) return CB200_ERANGE;
    }
    size_t ob = sizeof(int64_t) * ne * (shared_ode ? 1 : batch);
    DevBuf om, rm, rt;
    CK(om.alloc(ob));
    CK(rm.alloc(arr_mant_bytes(L, n + 1, S)));
    CK(rt.alloc(arr_meta_bytes(n + 1, S)));
    CK(cudaMemcpy(om.p, ode_int, ob, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(rm.p, res_mant, arr_mant_bytes(L, n + 1, S), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(rt.p, res_meta, arr_meta_bytes(n + 1, S), cudaMemcpyHostToDevice));
    int rc = cb200_fuchs_intop_batch_dev(L, order, degree, valuation, n, batch, (const int64_t*)om.p, shared_ode,
                                         (uint32_t*)rm.p, (cb_meta*)rt.p, nullptr);
    if (rc) return rc;
    CK(cudaMemcpy(res_mant, rm.p, arr_mant_bytes(L, n + 1, S), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(res_meta, rt.p, arr_meta_bytes(n + 1, S), cudaMemcpyDeviceToHost));
    return (int)cudaDeviceSynchronize();
}

int cb200_frobenius1_rhs_batch_host(int L, int order, int degree, int valuation, int64_t n, int64_t batch,
                                    const uint32_t* ode_mant, const cb_meta* ode_meta, int shared_ode,
                                    const uint32_t* rho_mant, const cb_meta* rho_meta, int shared_rho,
                                    const uint32_t* rhs_mant, const cb_meta* rhs_meta, int64_t rhs_len, int shared_rhs,
                                    uint32_t* res_mant, cb_meta* res_meta, int device) {
    if (!supported_L(L) || order < 1 || degree < 1 || batch < 0 || n < 0 || rhs_len < 0) return CB200_EINVAL;
    CK(cudaSetDevice(device));
    int64_t So = shared_ode ? 2 : 2 * batch, Sr = shared_rho ? 2 : 2 * batch, Sh = shared_rhs ? 2 : 2 * batch, S = 2 * batch;
    int64_t ne = (int64_t)(order + 1) * (degree + 1);
    DevBuf om, ot, pm, pt, hm, ht, rm, rt, ws;
    size_t wsb = cb200_frobenius_workspace_bytes(L, batch);
    CK(om.alloc(arr_mant_bytes(L, ne, So)));
    CK(ot.alloc(arr_meta_bytes(ne, So)));
    CK(pm.alloc(arr_mant_bytes(L, 1, Sr)));
    CK(pt.alloc(arr_meta_bytes(1, Sr)));
    CK(hm.alloc(arr_mant_bytes(L, rhs_len, Sh)));
    CK(ht.alloc(arr_meta_bytes(rhs_len, Sh)));
    CK(rm.alloc(arr_mant_bytes(L, n + 1, S)));
    CK(rt.alloc(arr_meta_bytes(n + 1, S)));
    CK(ws.alloc(wsb));
    CK(cudaMemcpy(om.p, ode_mant, arr_mant_bytes(L, ne, So), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ot.p, ode_meta, arr_meta_bytes(ne, So), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(pm.p, rho_mant, arr_mant_bytes(L, 1, Sr), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(pt.p, rho_meta, arr_meta_bytes(1, Sr), cudaMemcpyHostToDevice));
    if (rhs_len > 0) {
        CK(cudaMemcpy(hm.p, rhs_mant, arr_mant_bytes(L, rhs_len, Sh), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ht.p, rhs_meta, arr_meta_bytes(rhs_len, Sh), cudaMemcpyHostToDevice));
    }
    int rc = cb200_frobenius1_rhs_batch_dev(L, order, degree, valuation, n, batch, (uint32_t*)om.p, (cb_meta*)ot.p,
                                            shared_ode, (uint32_t*)pm.p, (cb_meta*)pt.p, shared_rho, (uint32_t*)hm.p,
                                            (cb_meta*)ht.p, rhs_len, shared_rhs, (uint32_t*)rm.p, (cb_meta*)rt.p, ws.p, wsb,
                                            nullptr);
    if (rc) return rc;
    CK(cudaMemcpy(res_mant, rm.p, arr_mant_bytes(L, n + 1, S), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(res_meta, rt.p, arr_meta_bytes(n + 1, S), cudaMemcpyDeviceToHost));
    return (int)cudaDeviceSynchronize();
}
int cb200_frobenius1_batch_host(int L, int order, int degree, int valuation, int64_t n, int64_t batch,
                                const uint32_t* ode_mant, const cb_meta* ode_meta, int shared_ode,
                                const uint32_t* rho_mant, const cb_meta* rho_meta, int shared_rho,
                                uint32_t* res_mant, cb_meta* res_meta, int device) {
    return cb200_frobenius1_rhs_batch_host(L, order, degree, valuation, n, batch, ode_mant, ode_meta, shared_ode, rho_mant,
                                           rho_meta, shared_rho, nullptr, nullptr, 0, 0, res_mant, res_meta, device);
}

int cb200_frobenius_general_batch_host(int L, int order, int degree, int valuation, int64_t n, int64_t batch,
                                       const uint32_t* ode_mant, const cb_meta* ode_meta, int shared_ode,
                                       const uint32_t* rho_mant, const cb_meta* rho_meta, int shared_rho, int mul,
                                       int M, uint32_t* gens_mant, cb_meta* gens_meta, int device) {
    if (!supported_L(L) || order < 1 || degree < 1 || degree > FROBG_MAXD || batch < 0 || n < 0 || M < 1 ||
        M > FROBG_MAXM || mul < 1 || mul > M)
        return CB200_EINVAL;
    CK(cudaSetDevice(device));
    int64_t So = shared_ode ? 2 : 2 * batch, Sr = shared_rho ? 2 : 2 * batch, S = 2 * batch;
    int64_t ne = (int64_t)(order + 1) * (degree + 1), ng = (int64_t)M * (n + 1);
    DevBuf om, ot, pm, pt, gm, gt, ws;
    size_t wsb = cb200_frobenius_general_workspace_bytes(L, order, degree, M, n, batch);
    CK(om.alloc(arr_mant_bytes(L, ne, So)));
    CK(ot.alloc(arr_meta_bytes(ne, So)));
    CK(pm.alloc(arr_mant_bytes(L, 1, Sr)));
    CK(pt.alloc(arr_meta_bytes(1, Sr)));
    CK(gm.alloc(arr_mant_bytes(L, ng, S)));
    CK(gt.alloc(arr_meta_bytes(ng, S)));
    CK(ws.alloc(wsb));
    CK(cudaMemcpy(om.p, ode_mant, arr_mant_bytes(L, ne, So), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ot.p, ode_meta, arr_meta_bytes(ne, So), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(pm.p, rho_mant, arr_mant_bytes(L, 1, Sr), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(pt.p, rho_meta, arr_meta_bytes(1, Sr), cudaMemcpyHostToDevice));
    int rc = cb200_frobenius_general_batch_dev(L, order, degree, valuation, n, batch, (uint32_t*)om.p, (cb_meta*)ot.p,
                                               shared_ode, (uint32_t*)pm.p, (cb_meta*)pt.p, shared_rho, mul, M,
                                               (uint32_t*)gm.p, (cb_meta*)gt.p, ws.p, wsb, nullptr);
    if (rc) return rc;
    CK(cudaMemcpy(gens_mant, gm.p, arr_mant_bytes(L, ng, S), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(gens_meta, gt.p, arr_meta_bytes(ng, S), cudaMemcpyDeviceToHost));
    return (int)cudaDeviceSynchronize();
}

int cb200_solution_op_host(int which, int L, int64_t batch, int mul, int M, int64_t cap, int64_t nu,
                           const uint32_t* rho_mant, const cb_meta* rho_meta, uint32_t* gens_mant, cb_meta* gens_meta,
                           uint32_t* f_mant, cb_meta* f_meta, int64_t fcap, int device) {
    if (!supported_L(L) || batch < 0 || M < 0 || M > FROBG_MAXM || cap < 0 || fcap < 0) return CB200_EINVAL;
    CK(cudaSetDevice(device));
    int64_t S = 2 * batch, ng = (int64_t)M * cap;
    bool hasf = which != CB200_SOL_NORMALIZE;
    DevBuf pm, pt, gm, gt, fm, ft, ws;
    size_t wsb = cb200_solution_op_workspace_bytes(L, M, batch);
    CK(pm.alloc(arr_mant_bytes(L, 1, S)));
    CK(pt.alloc(arr_meta_bytes(1, S)));
    CK(gm.alloc(arr_mant_bytes(L, ng, S)));
    CK(gt.alloc(arr_meta_bytes(ng, S)));
    CK(fm.alloc(arr_mant_bytes(L, hasf ? fcap : 0, S)));
    CK(ft.alloc(arr_meta_bytes(hasf ? fcap : 0, S)));
    CK(ws.alloc(wsb));
    CK(cudaMemcpy(pm.p, rho_mant, arr_mant_bytes(L, 1, S), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(pt.p, rho_meta, arr_meta_bytes(1, S), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(gm.p, gens_mant, arr_mant_bytes(L, ng, S), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(gt.p, gens_meta, arr_meta_bytes(ng, S), cudaMemcpyHostToDevice));
    if (hasf) {
        CK(cudaMemcpy(fm.p, f_mant, arr_mant_bytes(L, fcap, S), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ft.p, f_meta, arr_meta_bytes(fcap, S), cudaMemcpyHostToDevice));
    }
    int rc = cb200_solution_op_dev(which, L, batch, mul, M, cap, nu, (uint32_t*)pm.p, (cb_meta*)pt.p, (uint32_t*)gm.p,
                                   (cb_meta*)gt.p, (uint32_t*)fm.p, (cb_meta*)ft.p, fcap, ws.p, wsb, nullptr);
    if (rc) return rc;
    CK(cudaMemcpy(gens_mant, gm.p, arr_mant_bytes(L, ng, S), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(gens_meta, gt.p, arr_meta_bytes(ng, S), cudaMemcpyDeviceToHost));
    if (hasf) {
        CK(cudaMemcpy(f_mant, fm.p, arr_mant_bytes(L, fcap, S), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(f_meta, ft.p, arr_meta_bytes(fcap, S), cudaMemcpyDeviceToHost));
    }
    return (int)cudaDeviceSynchronize();
}

int cb200_indicial_polynomial_host(int L, int order, int degree, int valuation, int64_t batch,
                                   const uint32_t* ode_mant, const cb_meta* ode_meta, int shared_ode, int64_t nu,
                                   int64_t shift, uint32_t* out_mant, cb_meta* out_meta, int device) {
    if (!supported_L(L) || order < 1 || degree < 0 || batch < 0) return CB200_EINVAL;
    CK(cudaSetDevice(device));
    int64_t So = shared_ode ? 2 : 2 * batch, S = 2 * batch, ne = (int64_t)(order + 1) * (degree + 1);
    DevBuf om, ot, rm, rt, ws;
    size_t wsb = cb200_solution_op_workspace_bytes(L, 0, batch);
    CK(om.alloc(arr_mant_bytes(L, ne, So)));
    CK(ot.alloc(arr_meta_bytes(ne, So)));
    CK(rm.alloc(arr_mant_bytes(L, order + 2, S)));
    CK(rt.alloc(arr_meta_bytes(order + 2, S)));
    CK(ws.alloc(wsb));
    CK(cudaMemcpy(om.p, ode_mant, arr_mant_bytes(L, ne, So), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ot.p, ode_meta, arr_meta_bytes(ne, So), cudaMemcpyHostToDevice));
    int rc = cb200_indicial_polynomial_dev(L, order, degree, valuation, batch, (uint32_t*)om.p, (cb_meta*)ot.p,
                                           shared_ode, nu, shift, (uint32_t*)rm.p, (cb_meta*)rt.p, ws.p, wsb, nullptr);
    if (rc) return rc;
    CK(cudaMemcpy(out_mant, rm.p, arr_mant_bytes(L, order + 2, S), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out_meta, rt.p, arr_meta_bytes(order + 2, S), cudaMemcpyDeviceToHost));
    return (int)cudaDeviceSynchronize();
}

int cb200_solution_evaluate_host(int L, int M, int64_t cap, int64_t npts, const uint32_t* gens_mant,
                                 const cb_meta* gens_meta, const uint32_t* rho_mant, const cb_meta* rho_meta, int pow_mode,
                                 uint64_t pow_e, const uint32_t* pts_mant, const cb_meta* pts_meta, uint32_t* out_mant,
                                 cb_meta* out_meta, int device) {
    if (!supported_L(L) || M < 1 || M > FROBG_MAXM || cap < 1 || npts < 0) return CB200_EINVAL;
    CK(cudaSetDevice(device));
    int64_t S = 2 * npts, ng = (int64_t)M * cap;
    DevBuf gm, gt, rm, rt, pm, pt, om, ot, ws;
    size_t wsb = cb200_evaluate_workspace_bytes(L, npts);
    CK(gm.alloc(arr_mant_bytes(L, ng, 2)));
    CK(gt.alloc(arr_meta_bytes(ng, 2)));
    CK(rm.alloc(arr_mant_bytes(L, 1, 2)));
    CK(rt.alloc(arr_meta_bytes(1, 2)));
    CK(pm.alloc(arr_mant_bytes(L, 1, S)));
    CK(pt.alloc(arr_meta_bytes(1, S)));
    CK(om.alloc(arr_mant_bytes(L, 1, S)));
    CK(ot.alloc(arr_meta_bytes(1, S)));
    CK(ws.alloc(wsb));
    CK(cudaMemcpy(gm.p, gens_mant, arr_mant_bytes(L, ng, 2), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(gt.p, gens_meta, arr_meta_bytes(ng, 2), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(rm.p, rho_mant, arr_mant_bytes(L, 1, 2), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(rt.p, rho_meta, arr_meta_bytes(1, 2), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(pm.p, pts_mant, arr_mant_bytes(L, 1, S), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(pt.p, pts_meta, arr_meta_bytes(1, S), cudaMemcpyHostToDevice));
    int rc = cb200_solution_evaluate_dev(L, M, cap, npts, (uint32_t*)gm.p, (cb_meta*)gt.p, (uint32_t*)rm.p, (cb_meta*)rt.p,
                                         pow_mode, pow_e, (uint32_t*)pm.p, (cb_meta*)pt.p, (uint32_t*)om.p, (cb_meta*)ot.p,
                                         ws.p, wsb, nullptr);
    if (rc) return rc;
    CK(cudaMemcpy(out_mant, om.p, arr_mant_bytes(L, 1, S), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out_meta, ot.p, arr_meta_bytes(1, S), cudaMemcpyDeviceToHost));
    return (int)cudaDeviceSynchronize();
}
int cb200_trans_host(int which, int L, int64_t n, uint32_t* z_mant, cb_meta* z_meta, const uint32_t* x_mant,
                     const cb_meta* x_meta, int device) {
    if (!supported_L(L) || n < 0) return CB200_EINVAL;
    CK(cudaSetDevice(device));
    int64_t S = 2 * n;
    DevBuf zm, zt, xm, xt, ws;
    size_t wsb = cb200_evaluate_workspace_bytes(L, n);
    CK(zm.alloc(arr_mant_bytes(L, 1, S)));
    CK(zt.alloc(arr_meta_bytes(1, S)));
    CK(xm.alloc(arr_mant_bytes(L, 1, S)));
    CK(xt.alloc(arr_meta_bytes(1, S)));
    CK(ws.alloc(wsb));
    CK(cudaMemcpy(xm.p, x_mant, arr_mant_bytes(L, 1, S), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(xt.p, x_meta, arr_meta_bytes(1, S), cudaMemcpyHostToDevice));
    int rc = cb200_trans_dev(which, L, n, (uint32_t*)zm.p, (cb_meta*)zt.p, (uint32_t*)xm.p, (cb_meta*)xt.p, ws.p, wsb, nullptr);
    if (rc) return rc;
    CK(cudaMemcpy(z_mant, zm.p, arr_mant_bytes(L, 1, S), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(z_meta, zt.p, arr_meta_bytes(1, S), cudaMemcpyDeviceToHost));
    return (int)cudaDeviceSynchronize();
}

int cb200_ode_apply_host(int L, int order, int degree, int64_t batch, const uint32_t* ode_mant, const cb_meta* ode_meta,
                         int shared_ode, const uint32_t* in_mant, const cb_meta* in_meta, int64_t len_in, uint32_t* out_mant,
                         cb_meta* out_meta, int64_t out_len, int64_t deg, int32_t* solved, int device) {
    if (!supported_L(L) || order < 0 || degree < 0 || batch < 0 || len_in < 0 || out_len < 0) return CB200_EINVAL;
    CK(cudaSetDevice(device));
    int64_t So = shared_ode ? 2 : 2 * batch, S = 2 * batch, ne = (int64_t)(order + 1) * (degree + 1);
    DevBuf om, ot, im, it, rm, rt, sv, ws;
    size_t wsb = cb200_ode_apply_workspace_bytes(L, len_in, batch);
    CK(om.alloc(arr_mant_bytes(L, ne, So)));
    CK(ot.alloc(arr_meta_bytes(ne, So)));
    CK(im.alloc(arr_mant_bytes(L, len_in, S)));
    CK(it.alloc(arr_meta_bytes(len_in, S)));
    CK(rm.alloc(arr_mant_bytes(L, out_len, S)));
    CK(rt.alloc(arr_meta_bytes(out_len, S)));
    CK(sv.alloc(sizeof(int32_t) * batch));
    CK(ws.alloc(wsb));
    CK(cudaMemcpy(om.p, ode_mant, arr_mant_bytes(L, ne, So), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ot.p, ode_meta, arr_meta_bytes(ne, So), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(im.p, in_mant, arr_mant_bytes(L, len_in, S), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(it.p, in_meta, arr_meta_bytes(len_in, S), cudaMemcpyHostToDevice));
    int rc = cb200_ode_apply_dev(L, order, degree, batch, (uint32_t*)om.p, (cb_meta*)ot.p, shared_ode, (uint32_t*)im.p,
                                 (cb_meta*)it.p, len_in, (uint32_t*)rm.p, (cb_meta*)rt.p, out_len, deg,
                                 solved ? (int32_t*)sv.p : nullptr, ws.p, wsb, nullptr);
    if (rc) return rc;
    CK(cudaMemcpy(out_mant, rm.p, arr_mant_bytes(L, out_len, S), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out_meta, rt.p, arr_meta_bytes(out_len, S), cudaMemcpyDeviceToHost));
    if (solved) CK(cudaMemcpy(solved, sv.p, sizeof(int32_t) * batch, cudaMemcpyDeviceToHost));
    return (int)cudaDeviceSynchronize();
}

int cb200_horner_batch_host(int L, int64_t len, int64_t npts, int nder, const uint32_t* f_mant,
                            const cb_meta* f_meta, int shared_f, const uint32_t* pts_mant,
                            const cb_meta* pts_meta, uint32_t* out_mant, cb_meta* out_meta, int device) {
    if (!supported_L(L) || len < 0 || npts < 0 || nder < 1 || nder > HORNER_MAXDER) return CB200_EINVAL;
    CK(cudaSetDevice(device));
    int64_t Sf = shared_f ? 2 : 2 * npts, S = 2 * npts;
    DevBuf fm, ft, pm, pt, om, ot, ws;
    size_t wsb = cb200_horner_workspace_bytes(L, npts);
    CK(fm.alloc(arr_mant_bytes(L, len, Sf)));
    CK(ft.alloc(arr_meta_bytes(len, Sf)));
    CK(pm.alloc(arr_mant_bytes(L, 1, S)));
    CK(pt.alloc(arr_meta_bytes(1, S)));
    CK(om.alloc(arr_mant_bytes(L, nder, S)));
    CK(ot.alloc(arr_meta_bytes(nder, S)));
    CK(ws.alloc(wsb));
    CK(cudaMemcpy(fm.p, f_mant, arr_mant_bytes(L, len, Sf), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ft.p, f_meta, arr_meta_bytes(len, Sf), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(pm.p, pts_mant, arr_mant_bytes(L, 1, S), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(pt.p, pts_meta, arr_meta_bytes(1, S), cudaMemcpyHostToDevice));
    int rc = cb200_horner_batch_dev(L, len, npts, nder, (uint32_t*)fm.p, (cb_meta*)ft.p, shared_f,
                                    (uint32_t*)pm.p, (cb_meta*)pt.p, (uint32_t*)om.p, (cb_meta*)ot.p, ws.p,
                                    wsb, nullptr);
    if (rc) return rc;
    CK(cudaMemcpy(out_mant, om.p, arr_mant_bytes(L, nder, S), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out_meta, ot.p, arr_meta_bytes(nder, S), cudaMemcpyDeviceToHost));
    return (int)cudaDeviceSynchronize();
}

int cb200_ew_host(int op, int L, int64_t n, uint32_t* z_mant, cb_meta* z_meta, const uint32_t* x_mant,
                  const cb_meta* x_meta, const uint32_t* y_mant, const cb_meta* y_meta, const int64_t* k,
                  int device) {
    if (!supported_L(L) || n < 0) return CB200_EINVAL;
    CK(cudaSetDevice(device));
    int64_t S = 2 * n;
    DevBuf zm, zt, xm, xt, ym, yt, kk, ws;
    size_t wsb = cb200_ew_workspace_bytes(L, n);
    CK(zm.alloc(arr_mant_bytes(L, 1, S)));
    CK(zt.alloc(arr_meta_bytes(1, S)));
    CK(xm.alloc(arr_mant_bytes(L, 1, S)));
    CK(xt.alloc(arr_meta_bytes(1, S)));
    CK(ym.alloc(arr_mant_bytes(L, 1, S)));
    CK(yt.alloc(arr_meta_bytes(1, S)));
    CK(kk.alloc(sizeof(int64_t) * n));
    CK(ws.alloc(wsb));
    CK(cudaMemcpy(xm.p, x_mant, arr_mant_bytes(L, 1, S), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(xt.p, x_meta, arr_meta_bytes(1, S), cudaMemcpyHostToDevice));
    if (y_mant && y_meta) {
        CK(cudaMemcpy(ym.p, y_mant, arr_mant_bytes(L, 1, S), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(yt.p, y_meta, arr_meta_bytes(1, S), cudaMemcpyHostToDevice));
    }
    if (k) CK(cudaMemcpy(kk.p, k, sizeof(int64_t) * n, cudaMemcpyHostToDevice));
    int rc = cb200_ew_dev(op, L, n, (uint32_t*)zm.p, (cb_meta*)zt.p, (uint32_t*)xm.p, (cb_meta*)xt.p,
                          (uint32_t*)ym.p, (cb_meta*)yt.p, k ? (int64_t*)kk.p : nullptr, ws.p, wsb, nullptr);
    if (rc) return rc;
    CK(cudaMemcpy(z_mant, zm.p, arr_mant_bytes(L, 1, S), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(z_meta, zt.p, arr_meta_bytes(1, S), cudaMemcpyDeviceToHost));
    return (int)cudaDeviceSynchronize();
}

static int continuation_host_impl(int L, int order, int degree, int64_t n, int64_t batch, int64_t path_len,
                                  const uint32_t* ode_mant, const cb_meta* ode_meta, const uint32_t* path_mant,
                                  const cb_meta* path_meta, uint32_t* y_mant, cb_meta* y_meta, uint32_t* res_mant,
                                  cb_meta* res_meta, int device) {
    if (!supported_L(L) || order < 1 || degree < 0 || batch < 0 || n < order || path_len < 1) return CB200_EINVAL;
    CK(cudaSetDevice(device));
    const bool series = res_mant != nullptr;
    int64_t ne = (int64_t)(order + 1) * (degree + 1), S = 2 * batch;
    DevBuf om, ot, pm, pt, ym, yt, rm, rt, ws;
    size_t wsb = continuation_ws(L, order, degree, n, batch, series);
    CK(om.alloc(arr_mant_bytes(L, ne, 2)));
    CK(ot.alloc(arr_meta_bytes(ne, 2)));
    CK(pm.alloc(arr_mant_bytes(L, path_len, 2)));
    CK(pt.alloc(arr_meta_bytes(path_len, 2)));
    CK(ym.alloc(arr_mant_bytes(L, order, S)));
    CK(yt.alloc(arr_meta_bytes(order, S)));
    if (series) {
        CK(rm.alloc(arr_mant_bytes(L, n + 1, S)));
        CK(rt.alloc(arr_meta_bytes(n + 1, S)));
    }
    CK(ws.alloc(wsb));
    CK(cudaMemcpy(om.p, ode_mant, arr_mant_bytes(L, ne, 2), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ot.p, ode_meta, arr_meta_bytes(ne, 2), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(pm.p, path_mant, arr_mant_bytes(L, path_len, 2), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(pt.p, path_meta, arr_meta_bytes(path_len, 2), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ym.p, y_mant, arr_mant_bytes(L, order, S), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(yt.p, y_meta, arr_meta_bytes(order, S), cudaMemcpyHostToDevice));
    int rc = continuation_impl(L, order, degree, n, batch, path_len, (uint32_t*)om.p, (cb_meta*)ot.p, (uint32_t*)pm.p,
                               (cb_meta*)pt.p, (uint32_t*)ym.p, (cb_meta*)yt.p, series ? (uint32_t*)rm.p : nullptr,
                               series ? (cb_meta*)rt.p : nullptr, ws.p, wsb, nullptr);
    if (rc) return rc;
    CK(cudaMemcpy(y_mant, ym.p, arr_mant_bytes(L, order, S), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(y_meta, yt.p, arr_meta_bytes(order, S), cudaMemcpyDeviceToHost));
    if (series) {
        CK(cudaMemcpy(res_mant, rm.p, arr_mant_bytes(L, n + 1, S), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(res_meta, rt.p, arr_meta_bytes(n + 1, S), cudaMemcpyDeviceToHost));
    }
    return (int)cudaDeviceSynchronize();
}
int cb200_continuation_host(int L, int order, int degree, int64_t n, int64_t batch, int64_t path_len,
                            const uint32_t* ode_mant, const cb_meta* ode_meta, const uint32_t* path_mant,
                            const cb_meta* path_meta, uint32_t* y_mant, cb_meta* y_meta, int device) {
    return continuation_host_impl(L, order, degree, n, batch, path_len, ode_mant, ode_meta, path_mant, path_meta, y_mant,
                                  y_meta, nullptr, nullptr, device);
}
int cb200_continuation_series_host(int L, int order, int degree, int64_t n, int64_t batch, int64_t path_len,
                                   const uint32_t* ode_mant, const cb_meta* ode_meta, const uint32_t* path_mant,
                                   const cb_meta* path_meta, uint32_t* y_mant, cb_meta* y_meta, uint32_t* res_mant,
                                   cb_meta* res_meta, int device) {
    if (!res_mant || !res_meta || path_len < 2) return CB200_EINVAL;
    return continuation_host_impl(L, order, degree, n, batch, path_len, ode_mant, ode_meta, path_mant, path_meta, y_mant,
                                  y_meta, res_mant, res_meta, device);
}

int cb200_ode_shift_host(int L, int order, int degree, int64_t batch, const uint32_t* ode_mant, const cb_meta* ode_meta,
                         int shared_ode, const uint32_t* a_mant, const cb_meta* a_meta, int shared_a, uint32_t* out_mant,
                         cb_meta* out_meta, int device) {
    if (!supported_L(L) || order < 0 || degree < 0 || batch < 0) return CB200_EINVAL;
    CK(cudaSetDevice(device));
    int64_t ne = (int64_t)(order + 1) * (degree + 1), S = 2 * batch, So = shared_ode ? 2 : S, Sa = shared_a ? 2 : S;
    DevBuf om, ot, am, at_, rm, rt, ws;
    size_t wsb = cb200_ode_shift_workspace_bytes(L, batch);
    CK(om.alloc(arr_mant_bytes(L, ne, So)));
    CK(ot.alloc(arr_meta_bytes(ne, So)));
    CK(am.alloc(arr_mant_bytes(L, 1, Sa)));
    CK(at_.alloc(arr_meta_bytes(1, Sa)));
    CK(rm.alloc(arr_mant_bytes(L, ne, S)));
    CK(rt.alloc(arr_meta_bytes(ne, S)));
    CK(ws.alloc(wsb));
    CK(cudaMemcpy(om.p, ode_mant, arr_mant_bytes(L, ne, So), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ot.p, ode_meta, arr_meta_bytes(ne, So), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(am.p, a_mant, arr_mant_bytes(L, 1, Sa), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(at_.p, a_meta, arr_meta_bytes(1, Sa), cudaMemcpyHostToDevice));
    int rc = cb200_ode_shift_dev(L, order, degree, batch, (uint32_t*)om.p, (cb_meta*)ot.p, shared_ode, (uint32_t*)am.p,
                                 (cb_meta*)at_.p, shared_a, (uint32_t*)rm.p, (cb_meta*)rt.p, ws.p, wsb, nullptr);
    if (rc) return rc;
    CK(cudaMemcpy(out_mant, rm.p, arr_mant_bytes(L, ne, S), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out_meta, rt.p, arr_meta_bytes(ne, S), cudaMemcpyDeviceToHost));
    return (int)cudaDeviceSynchronize();
}

int cb200_taylor_shift_host(int L, int64_t len, int64_t batch, uint32_t* f_mant, cb_meta* f_meta, const uint32_t* a_mant,
                            const cb_meta* a_meta, int shared_a, int device) {
    if (!supported_L(L) || len < 0 || batch < 0) return CB200_EINVAL;
    CK(cudaSetDevice(device));
    int64_t S = 2 * batch, Sa = shared_a ? 2 : S;
    DevBuf fm, ft, am, at_, ws;
    size_t wsb = cb200_taylor_shift_workspace_bytes(L, len, batch);
    CK(fm.alloc(arr_mant_bytes(L, len, S)));
    CK(ft.alloc(arr_meta_bytes(len, S)));
    CK(am.alloc(arr_mant_bytes(L, 1, Sa)));
    CK(at_.alloc(arr_meta_bytes(1, Sa)));
    CK(ws.alloc(wsb));
    CK(cudaMemcpy(fm.p, f_mant, arr_mant_bytes(L, len, S), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ft.p, f_meta, arr_meta_bytes(len, S), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(am.p, a_mant, arr_mant_bytes(L, 1, Sa), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(at_.p, a_meta, arr_meta_bytes(1, Sa), cudaMemcpyHostToDevice));
    int rc = cb200_taylor_shift_dev(L, len, batch, (uint32_t*)fm.p, (cb_meta*)ft.p, (uint32_t*)am.p, (cb_meta*)at_.p,
                                    shared_a, ws.p, wsb, nullptr);
    if (rc) return rc;
    CK(cudaMemcpy(f_mant, fm.p, arr_mant_bytes(L, len, S), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(f_meta, ft.p, arr_meta_bytes(len, S), cudaMemcpyDeviceToHost));
    return (int)cudaDeviceSynchronize();
}

int cb200_radius_of_convergence_host(int L, int order, int degree, int64_t batch, const uint32_t* ode_mant,
                                     const cb_meta* ode_meta, int shared_ode, int64_t n, uint32_t* out_mant, cb_meta* out_meta,
                                     int32_t* n_done, int device) {
    if (!supported_L(L) || order < 0 || degree < 0 || batch < 0 || n < 0 || !n_done) return CB200_EINVAL;
    CK(cudaSetDevice(device));
    int64_t ne = (int64_t)(order + 1) * (degree + 1), S = 2 * batch, So = shared_ode ? 2 : S;
    DevBuf om, ot, rm, rt, nd, ws;
    size_t wsb = cb200_radius_workspace_bytes(L, degree, batch);
    CK(om.alloc(arr_mant_bytes(L, ne, So)));
    CK(ot.alloc(arr_meta_bytes(ne, So)));
    CK(rm.alloc(arr_mant_bytes(L, 1, S)));
    CK(rt.alloc(arr_meta_bytes(1, S)));
    CK(nd.alloc(sizeof(int32_t) * batch));
    CK(ws.alloc(wsb));
    CK(cudaMemcpy(om.p, ode_mant, arr_mant_bytes(L, ne, So), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ot.p, ode_meta, arr_meta_bytes(ne, So), cudaMemcpyHostToDevice));
    int rc = cb200_radius_of_convergence_dev(L, order, degree, batch, (uint32_t*)om.p, (cb_meta*)ot.p, shared_ode, n,
                                             (uint32_t*)rm.p, (cb_meta*)rt.p, (int32_t*)nd.p, ws.p, wsb, nullptr);
    if (rc) return rc;
    CK(cudaMemcpy(out_mant, rm.p, arr_mant_bytes(L, 1, S), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out_meta, rt.p, arr_meta_bytes(1, S), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(n_done, nd.p, sizeof(int32_t) * batch, cudaMemcpyDeviceToHost));
    return (int)cudaDeviceSynchronize();
}
int cb200_truncation_order_host(int L, int64_t batch, const uint32_t* eta_mant, const cb_meta* eta_meta,
                                const uint32_t* alpha_mant, const cb_meta* alpha_meta, int64_t bits, uint32_t* n_mant,
                                cb_meta* n_meta, int64_t* n_out, int device) {
    if (!supported_L(L) || batch < 0 || bits < 1 || !n_out) return CB200_EINVAL;
    CK(cudaSetDevice(device));
    int64_t S = 2 * batch;
    DevBuf em, et, am, at_, nm, nt, no, ws;
    size_t wsb = cb200_truncation_workspace_bytes(L, batch);
    CK(em.alloc(arr_mant_bytes(L, 1, S)));
    CK(et.alloc(arr_meta_bytes(1, S)));
    CK(am.alloc(arr_mant_bytes(L, 1, S)));
    CK(at_.alloc(arr_meta_bytes(1, S)));
    CK(nm.alloc(arr_mant_bytes(L, 1, S)));
    CK(nt.alloc(arr_meta_bytes(1, S)));
    CK(no.alloc(sizeof(int64_t) * batch));
    CK(ws.alloc(wsb));
    CK(cudaMemcpy(em.p, eta_mant, arr_mant_bytes(L, 1, S), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(et.p, eta_meta, arr_meta_bytes(1, S), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(am.p, alpha_mant, arr_mant_bytes(L, 1, S), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(at_.p, alpha_meta, arr_meta_bytes(1, S), cudaMemcpyHostToDevice));
    int rc = cb200_truncation_order_dev(L, batch, (uint32_t*)em.p, (cb_meta*)et.p, (uint32_t*)am.p, (cb_meta*)at_.p, bits,
                                        (uint32_t*)nm.p, (cb_meta*)nt.p, (int64_t*)no.p, ws.p, wsb, nullptr);
    if (rc) return rc;
    CK(cudaMemcpy(n_mant, nm.p, arr_mant_bytes(L, 1, S), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(n_meta, nt.p, arr_meta_bytes(1, S), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(n_out, no.p, sizeof(int64_t) * batch, cudaMemcpyDeviceToHost));
    return (int)cudaDeviceSynchronize();
}

}  // extern "C"
