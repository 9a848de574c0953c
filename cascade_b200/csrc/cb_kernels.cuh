// cb_kernels.cuh -- __global__ wrappers around the per-thread solver bodies, and their launchers.
// Included by one translation unit per precision (kern_L*.cu) so the precisions compile in parallel.
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>
#include "cb_solvers.cuh"

namespace cb {

// threads per block for each precision: bounded by the shared-memory scratch
// (2 columns of 2L+3 words per thread) and by 255 registers per thread.
template <int L> struct Cfg;
template <> struct Cfg<4> { static constexpr int TPB = 128; };
template <> struct Cfg<32> { static constexpr int TPB = 224; };
template <> struct Cfg<64> { static constexpr int TPB = 192; };
template <> struct Cfg<128> { static constexpr int TPB = 112; };

// threads per block of the shared-operator recurrence kernel (two product columns + three
// accumulator columns per thread)
template <int L> struct CfgShared;
template <> struct CfgShared<4> { static constexpr int TPB = 128; static constexpr int TPB_BIG = 128; static constexpr int TPB_HUGE = 128; };
// 1024 bits: two block sizes, picked per launch to minimise the idle tail of the last wave
// (7 warps at <= 255 registers, or 12 warps at 168 registers = 3 warps per scheduler)
template <> struct CfgShared<32> { static constexpr int TPB = 224; static constexpr int TPB_BIG = 384; static constexpr int TPB_HUGE = 448; };
template <> struct CfgShared<64> { static constexpr int TPB = 96; static constexpr int TPB_BIG = 96; static constexpr int TPB_HUGE = 96; };
template <> struct CfgShared<128> { static constexpr int TPB = 32; static constexpr int TPB_BIG = 32; static constexpr int TPB_HUGE = 32; };

template <int L> struct Sizes {
    static constexpr int WORDS = 2 * L + 3;  // words per product scratch column
    // words per accumulator column; with truncated products the spare one also parks a high
    // product (L+4 words), so all three are L+5
    static constexpr int AWORDS = use_trunc<L>() ? L + 5 : L + 2;
    static constexpr size_t SMEM = (size_t)2 * WORDS * Cfg<L>::TPB * sizeof(uint32_t);
    // recurrence kernel: with truncated products S0 only parks a high product (L+5 words) and the
    // second full column lives in global memory (general path only)
    static constexpr int S0WORDS = use_trunc<L>() ? 0 : WORDS;
    static constexpr int S1WORDS = use_trunc<L>() ? 0 : WORDS;
    static constexpr size_t SMEM_SHARED_PER_THREAD = (size_t)(S0WORDS + S1WORDS + 3 * AWORDS) * sizeof(uint32_t);
};

#define CB_KERNEL_PROLOGUE(COUNT)                                                              \
    extern __shared__ uint32_t cb_smem[];                                                      \
    constexpr int TPB = Cfg<L>::TPB;                                                           \
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(cb_smem) + 4u * threadIdx.x;     \
    Scratch<TPB> S0{sbase};                                                                    \
    Scratch<TPB> S1{sbase + 4u * (uint32_t)(Sizes<L>::WORDS * TPB)};                           \
    const int64_t idx = (int64_t)blockIdx.x * TPB + threadIdx.x;                               \
    if (idx >= (COUNT)) return;

// 2048 / 4096 bits: hybrid columns (L + 5 words in shared memory, the tail of the rare exact paths in global
// memory), so that the block is bounded by the 255 registers per thread, not by shared memory
#ifndef CB_HYB64_TPB
#define CB_HYB64_TPB 256
#endif
template <int L> struct HybCfg {
    static constexpr bool ON = L >= 64;
    static constexpr int KS = L + 5;                  // shared words per column
    static constexpr int GW = 2 * L + 3 - KS;         // global words per column
    static constexpr int TPB = L == 64 ? 256 : (L == 128 ? 192 : Cfg<L>::TPB);
    static constexpr size_t SMEM = (size_t)2 * KS * TPB * sizeof(uint32_t);
};
// the Horner kernels on hybrid columns have their own block size (2048 bits: more warps at fewer registers)
template <int L> struct HornerHyb {
    static constexpr int TPB = L == 64 ? CB_HYB64_TPB : HybCfg<L>::TPB;
    static constexpr size_t SMEM = (size_t)2 * HybCfg<L>::KS * TPB * sizeof(uint32_t);
};
#define CB_KERNEL_PROLOGUE_HYB(COUNT, GCOLS, TPBV)                                             \
    extern __shared__ uint32_t cb_smem[];                                                      \
    constexpr int TPB = TPBV;                                                                  \
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(cb_smem) + 4u * threadIdx.x;     \
    const int64_t idx = (int64_t)blockIdx.x * TPB + threadIdx.x;                               \
    if (idx >= (COUNT)) return;                                                                \
    HybridScratch<TPB, HybCfg<L>::KS> S0{sbase, (GCOLS) + idx, (uint32_t)(COUNT)};             \
    HybridScratch<TPB, HybCfg<L>::KS> S1{sbase + 4u * (uint32_t)(HybCfg<L>::KS * TPB),         \
                                         (GCOLS) + (size_t)HybCfg<L>::GW * (size_t)(COUNT) + idx, (uint32_t)(COUNT)};

template <int L>
__global__ void __launch_bounds__(Cfg<L>::TPB, 1) frobenius1_kernel(FrobParams p) {
    CB_KERNEL_PROLOGUE(p.batch)
    frobenius1_thread<L, true>(p, idx, S0, S1);
}
template <int L>
__global__ void __launch_bounds__(HybCfg<L>::TPB, 1) frobenius1_hyb_kernel(FrobParams p) {
    if constexpr (HybCfg<L>::ON) {
        CB_KERNEL_PROLOGUE_HYB(p.batch, p.gcols, HybCfg<L>::TPB)
        frobenius1_thread<L, true>(p, idx, S0, S1);
    }
}
// ---- the 4096-bit Frobenius recurrence for batches of more than one wave: persistent CTAs of 208 threads on hybrid
// columns (6.5 warps per SM against the 3.5 of the all-shared kernel), work units of (208 instances, 8 coefficients)
// handed out by an atomic counter as in the recurrence kernel -- no wave quantisation, any batch size
template <int L> struct FrobPersCfg {
    static constexpr bool ON = L == 128;
    static constexpr int TPB = 208, KS = HybCfg<L>::KS, RANGE = 8;
    static constexpr size_t SMEM = (size_t)2 * KS * TPB * sizeof(uint32_t);
};
template <int L>
__global__ void __launch_bounds__(FrobPersCfg<L>::TPB, 1) frobenius1_pers_kernel(FrobParams p) {
    if constexpr (FrobPersCfg<L>::ON) {
        extern __shared__ uint32_t cb_smem[];
        __shared__ int unit_s;
        constexpr int TPB = FrobPersCfg<L>::TPB, KS = FrobPersCfg<L>::KS;
        const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(cb_smem) + 4u * threadIdx.x;
        const int64_t nblocks = (p.batch + TPB - 1) / TPB;
        const int64_t nranges = (p.n + FrobPersCfg<L>::RANGE) / FrobPersCfg<L>::RANGE;  // coefficients 0 .. n
        const int64_t total = nblocks * nranges;
        const uint32_t gs = (uint32_t)(nblocks * TPB);
        for (;;) {
            if (threadIdx.x == 0) {
                int u = atomicAdd(p.sched, 1);
                if (u < total) {
                    int64_t b = u % nblocks, r = u / nblocks;
                    while (r > 0 && atomicAdd(p.sched + 1 + b, 0) < (int)r) __nanosleep(200);
                    __threadfence();
                }
                unit_s = u;
            }
            __syncthreads();
            const int64_t u = unit_s;
            if (u >= total) break;
            const int64_t b = u % nblocks, r = u / nblocks;
            const int64_t inst = b * TPB + threadIdx.x;
            const int64_t lo = r * FrobPersCfg<L>::RANGE;
            const int64_t hi = lo + FrobPersCfg<L>::RANGE - 1 < p.n ? lo + FrobPersCfg<L>::RANGE - 1 : p.n;
            HybridScratch<TPB, KS> S0{sbase, p.gcols + inst, gs};
            HybridScratch<TPB, KS> S1{sbase + 4u * (uint32_t)(KS * TPB), p.gcols + (size_t)HybCfg<L>::GW * gs + inst, gs};
            // every thread walks the loops (phase barriers keep the 6.5 warps on the same instruction cache lines:
            // without them this kernel runs at 0.68 of the speed); threads beyond the batch touch no memory
            const bool live = inst < p.batch;
            frobenius1_thread<L, true>(p, live ? inst : 0, S0, S1, lo, hi, live);
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) atomicExch(p.sched + 1 + b, (int)(r + 1));
        }
    }
}
// ---- cooperating lanes (cb_coop.cuh): four lanes per instance, one hybrid column per lane, 64 instances per CTA
template <int L> struct CoopCfg {
    static constexpr bool ON = L >= 64;
    static constexpr int TPB = 256, G = TPB / 4;
    static constexpr int KS = L + 5, GW = 2 * L + 3 - KS;
    static constexpr size_t SMEM = (size_t)KS * TPB * sizeof(uint32_t);
    static constexpr int RANGE = 8;  // coefficients per work unit of the Frobenius recurrence
};
// two lanes per instance: G instances per CTA, four columns each, laid out [word][role][instance]; G = 8 mod 16 keeps the
// columns the two lanes of an instance touch at the same time (roles r and r + 2) sixteen banks apart
template <int L> struct Coop2Cfg {
    static constexpr int G = L >= 128 ? 104 : 120, TPB = 2 * G, STRIDE = 4 * G;
    static constexpr int KS = L + 5;
    static constexpr size_t SMEM = (size_t)KS * STRIDE * sizeof(uint32_t);
    static_assert(G % 16 == 8, "bank layout");
};
// the instance of a thread, and its Quad, for both forms
template <int L, int LANES>
struct CoopGeom {
    static constexpr int G = LANES == 4 ? CoopCfg<L>::G : Coop2Cfg<L>::G;
    static constexpr int TPB = LANES * G, STRIDE = 4 * G, KS = CoopCfg<L>::KS;
    static constexpr size_t SMEM = (size_t)KS * STRIDE * sizeof(uint32_t);
    using Q = Quad<STRIDE, KS>;
    static __device__ __forceinline__ int local_instance() { return (int)threadIdx.x / LANES; }
    // sbase: shared-space address of the column region; gcols: global tails; blk, nblocks: this CTA's instance block
    static __device__ __forceinline__ Q quad(uint32_t sbase, uint32_t* gcols, int64_t blk, int64_t nblocks) {
        const int i = local_instance(), r = (int)threadIdx.x % LANES;
        Q q;
        q.q0 = LANES == 4 ? r : 2 * r;
        q.per = 4 / LANES;
        q.qmask = LANES == 4 ? (0xFu << (threadIdx.x & 28)) : (0x3u << (threadIdx.x & 30));
        const uint32_t col0 = LANES == 4 ? 4u * (uint32_t)i : (uint32_t)i;
        q.cstep = LANES == 4 ? 1u : (uint32_t)G;
        q.a0 = sbase + 4u * col0;
        q.g0 = gcols + (size_t)blk * STRIDE + col0;
        q.gs = (uint32_t)(nblocks * STRIDE);
        return q;
    }
};
// Work units = (block of 64 instances, range of RANGE coefficients), handed out by an atomic counter in
// coefficient-major order; unit (b, r) waits for (b, r - 1) (taken a full round earlier: all CTAs are resident and
// units are taken in order, so the wait cannot deadlock).  Any batch size fills the SMs without a last-wave tail.
template <int L, int LANES>
__global__ void __launch_bounds__(CoopGeom<L, LANES>::TPB, 1) frobenius1_coop_kernel(FrobParams p) {
    if constexpr (CoopCfg<L>::ON) {
        extern __shared__ uint32_t cb_smem[];
        __shared__ int unit_s;
        using Geo = CoopGeom<L, LANES>;
        constexpr int G = Geo::G;
        const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(cb_smem);
        const int64_t nblocks = (p.batch + G - 1) / G;
        const int64_t nranges = (p.n + CoopCfg<L>::RANGE) / CoopCfg<L>::RANGE;  // coefficients 0 .. n
        const int64_t total = nblocks * nranges;
        for (;;) {
            if (threadIdx.x == 0) {
                int u = atomicAdd(p.sched, 1);
                if (u < total) {
                    int64_t b = u % nblocks, r = u / nblocks;
                    while (r > 0 && atomicAdd(p.sched + 1 + b, 0) < (int)r) __nanosleep(200);
                    __threadfence();
                }
                unit_s = u;
            }
            __syncthreads();
            const int64_t u = unit_s;
            if (u >= total) break;
            const int64_t b = u % nblocks, r = u / nblocks;
            const int64_t inst = b * G + Geo::local_instance();
            const int64_t lo = r * CoopCfg<L>::RANGE;
            const int64_t hi = lo + CoopCfg<L>::RANGE - 1 < p.n ? lo + CoopCfg<L>::RANGE - 1 : p.n;
            typename Geo::Q Q = Geo::quad(sbase, p.gcols, b, nblocks);
            // every lane walks the loops and meets the CTA-wide phase barriers (instruction cache); lanes without an
            // instance touch no memory
            Q.active = inst < p.batch;
            Q.cta_sync = true;
            frobenius1_coop_quad<L>(p, Q.active ? inst : 0, Q, lo, hi);
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) atomicExch(p.sched + 1 + b, (int)(r + 1));
        }
    }
}
// the shared-operator Fuchs recurrence on four lanes per instance (few solutions of one operator: analytic
// continuation, monodromy), work units of G instances x RANGE coefficients as above
#ifndef CB_COOPF_SYNC
#define CB_COOPF_SYNC 0  // (phase barriers in the cooperative Fuchs kernel: measured 3-5 % slower, its code is small)
#endif
#ifndef CB_COOPF128_TPB
#define CB_COOPF128_TPB 128  // (192 measured: 5.3 T against 6.9 T limb-MAC/s at 65 536 instances)
#endif
template <int L> struct CoopFuchsCfg {
    static constexpr int TPB = L <= 64 ? 256 : CB_COOPF128_TPB, G = TPB / 4, RANGE = 32, STATE = 24;
    static constexpr size_t SMEM = (size_t)(CoopCfg<L>::KS + L + 2) * TPB * sizeof(uint32_t) + (size_t)STATE * G * sizeof(uint32_t);
};
template <int L>
__global__ void __launch_bounds__(CoopFuchsCfg<L>::TPB, 1) fuchs_shared_coop_kernel(FuchsSharedParams p) {
    if constexpr (CoopCfg<L>::ON) {
        extern __shared__ __align__(16) uint32_t cb_smem_cf[];
        __shared__ int unit_s;
        constexpr int TPB = CoopFuchsCfg<L>::TPB, G = CoopFuchsCfg<L>::G, KS = CoopCfg<L>::KS;
        const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(cb_smem_cf);
        uint32_t* state = cb_smem_cf + (size_t)(KS + L + 2) * TPB;
        const int64_t nblocks = (p.batch + G - 1) / G;
        const int64_t first = p.n_begin > -p.valuation ? p.n_begin : -p.valuation;
        const int64_t nranges = (p.n - first + CoopFuchsCfg<L>::RANGE) / CoopFuchsCfg<L>::RANGE;
        const int64_t total = nblocks * nranges;
        for (;;) {
            if (threadIdx.x == 0) {
                int u = atomicAdd(p.sched, 1);
                if (u < total) {
                    int64_t b = u % nblocks, r = u / nblocks;
                    while (r > 0 && atomicAdd(p.sched + 1 + b, 0) < (int)r) __nanosleep(200);
                    __threadfence();
                }
                unit_s = u;
            }
            __syncthreads();
            const int64_t u = unit_s;
            if (u >= total) break;
            const int64_t b = u % nblocks, r = u / nblocks;
            const int64_t inst = b * G + (threadIdx.x >> 2);
            const int64_t lo = first + r * CoopFuchsCfg<L>::RANGE;
            const int64_t hi = lo + CoopFuchsCfg<L>::RANGE - 1 < p.n ? lo + CoopFuchsCfg<L>::RANGE - 1 : p.n;
            Quad<TPB, KS> Q;
            Q.q0 = (int)(threadIdx.x & 3);
            Q.per = 1;
            Q.qmask = 0xFu << (threadIdx.x & 28);
            Q.cstep = 1u;
            Q.a0 = sbase + 4u * (threadIdx.x & ~3u);
            Q.g0 = p.gcols + (size_t)b * TPB + (threadIdx.x & ~3u);
            Q.gs = (uint32_t)(nblocks * TPB);
            Q.a20 = sbase + 4u * (uint32_t)(KS * TPB) + 4u * (threadIdx.x & ~3u);
            Q.st = state + (size_t)CoopFuchsCfg<L>::STATE * (threadIdx.x >> 2);
            Q.active = inst < p.batch;  // (lanes without an instance walk the loops for the phase barriers only)
            Q.cta_sync = CB_COOPF_SYNC != 0;
            fuchs_shared_coop_quad<L>(p, Q.active ? inst : 0, Q, lo, hi);
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) atomicExch(p.sched + 1 + b, (int)(r + 1));
        }
    }
}
// Horner on four (two) lanes per point: one CTA per block of 64 (120 / 104) points
template <int L, int LANES>
__global__ void __launch_bounds__(CoopGeom<L, LANES>::TPB, 1) horner_coop_kernel(HornerParams p) {
    if constexpr (CoopCfg<L>::ON) {
        extern __shared__ uint32_t cb_smem[];
        using Geo = CoopGeom<L, LANES>;
        const int64_t pt = (int64_t)blockIdx.x * Geo::G + Geo::local_instance();
        typename Geo::Q Q = Geo::quad((uint32_t)__cvta_generic_to_shared(cb_smem), p.gcols, blockIdx.x, gridDim.x);
        Q.active = pt < p.npts;  // (lanes without a point walk the loops for the phase barriers only)
        Q.cta_sync = true;
        horner_coop_quad<L>(p, Q.active ? pt : 0, Q);
    }
}

template <int L>
__global__ void __launch_bounds__(Cfg<L>::TPB, 1) frobenius_general_kernel(FrobGenParams p) {
    CB_KERNEL_PROLOGUE(p.batch)
    frobenius_general_thread<L>(p, idx, S0, S1);
}
template <int L>
__global__ void __launch_bounds__(Cfg<L>::TPB, 1) solution_op_kernel(SolOpParams p) {
    CB_KERNEL_PROLOGUE(p.batch)
    solution_op_thread<L>(p, idx, S0, S1);
}
template <int L>
__global__ void __launch_bounds__(Cfg<L>::TPB, 1) indicial_poly_kernel(IndicialParams p) {
    CB_KERNEL_PROLOGUE(p.batch)
    indicial_poly_thread<L>(p, idx, S0, S1);
}
template <int L>
__global__ void __launch_bounds__(Cfg<L>::TPB, 1) evaluate_kernel(EvalParams p) {
    CB_KERNEL_PROLOGUE(p.npts)
    evaluate_thread<L>(p, idx, S0, S1);
}
template <int L>
__global__ void __launch_bounds__(Cfg<L>::TPB, 1) apply_kernel(ApplyParams p) {
    CB_KERNEL_PROLOGUE(p.batch)
    apply_thread<L>(p, idx, S0, S1);
}
template <int L>
__global__ void __launch_bounds__(Cfg<L>::TPB, 1) trans_kernel(TransParams p) {
    CB_KERNEL_PROLOGUE(p.n)
    trans_thread<L>(p, idx, S0, S1);
}
template <int L>
__global__ void __launch_bounds__(Cfg<L>::TPB, 1) radius_kernel(RadiusParams p) {
    CB_KERNEL_PROLOGUE(p.batch)
    radius_thread<L>(p, idx, S0, S1);
}
template <int L>
__global__ void __launch_bounds__(Cfg<L>::TPB, 1) truncation_kernel(TruncParams p) {
    CB_KERNEL_PROLOGUE(p.batch)
    truncation_thread<L>(p, idx, S0, S1);
}
#ifndef CB_HORNER_SYNC
#define CB_HORNER_SYNC 1
#endif
constexpr bool HORNER_SYNC = CB_HORNER_SYNC != 0;
template <int L>
__global__ void __launch_bounds__(Cfg<L>::TPB, 1) horner_kernel(HornerParams p) {
    CB_KERNEL_PROLOGUE(p.npts)
    horner_thread<L, HORNER_SYNC>(p, idx, S0, S1);
}
template <int L>
__global__ void __launch_bounds__(HornerHyb<L>::TPB, 1) horner_hyb_kernel(HornerParams p) {
    if constexpr (HybCfg<L>::ON) {
        CB_KERNEL_PROLOGUE_HYB(p.npts, p.gcols, HornerHyb<L>::TPB)
        horner_thread<L, HORNER_SYNC>(p, idx, S0, S1);
    }
}
template <int L>
__global__ void __launch_bounds__(Cfg<L>::TPB, 1) ode_shift_kernel(ShiftParams p) {
    CB_KERNEL_PROLOGUE(p.batch)
    ode_shift_thread<L>(p, idx, S0, S1);
}
// one CTA per series (blockIdx.x); see TShiftParams
template <int L>
__global__ void __launch_bounds__(Cfg<L>::TPB, 1) taylor_shift_kernel(TShiftParams p) {
    extern __shared__ uint32_t cb_smem[];
    constexpr int TPB = Cfg<L>::TPB;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(cb_smem) + 4u * threadIdx.x;
    Scratch<TPB> S0{sbase};
    Scratch<TPB> S1{sbase + 4u * (uint32_t)(Sizes<L>::WORDS * TPB)};
    const int64_t inst = blockIdx.x;
    if (acb_is_exact_zero(at<L>(p.a, 0, (p.a.S == 2) ? 0 : 2 * (size_t)inst))) return;  // uniform over the CTA
    for (int64_t i = p.len - 2; i >= 0; i--) {
        for (int64_t j = i + threadIdx.x; j <= p.len - 2; j += TPB) taylor_shift_mul<L>(p, inst, j, S0, S1);
        __syncthreads();
        for (int64_t j = i + threadIdx.x; j <= p.len - 2; j += TPB) taylor_shift_add<L>(p, inst, j, S0, S1);
        __syncthreads();
    }
}
// ---- Horner with the series staged by TMA: every point of the launch reads the SAME series, one coefficient per
// step (reference src/acb_ode_solution.c:40: acb_poly_evaluate of one stored series at many points).  Instead of every
// thread loading that coefficient through L1 / L2 on its critical path, one elected thread streams the series into
// shared memory in tiles of T coefficients with cp.async.bulk (1-D bulk copies: an entry of the [entry][limb][2]
// layout is contiguous, 8 L + 32 bytes), double buffered, completion on an mbarrier; the tile for the next T steps
// is in flight while the current one is consumed.  Same arithmetic as horner_thread: the coefficient is merely read
// from shared memory.
template <int L> struct HornerTile {
    static constexpr int T = (L <= 64 && HornerHyb<L>::TPB <= 256) ? 16 : 8;
    static constexpr uint32_t MANT = (uint32_t)T * L * 2 * 4, META = (uint32_t)T * 2 * 16;
    static constexpr size_t BYTES = (size_t)2 * (MANT + META);
};
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void tma_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void tma_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
        "r"(parity)
        : "memory");
}
#endif
template <int L>
__global__ void __launch_bounds__(HornerHyb<L>::TPB, 1) horner_tma_kernel(HornerParams p) {
    if constexpr (HybCfg<L>::ON) {
#if defined(__CUDA_ARCH__)
        extern __shared__ __align__(128) uint32_t cb_smem_tma[];
        __shared__ __align__(8) uint64_t mbar[2];
        constexpr int TPB = HornerHyb<L>::TPB, T = HornerTile<L>::T;
        constexpr bool SYNC = HORNER_SYNC;
        const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(cb_smem_tma) + 4u * threadIdx.x;
        char* tiles = reinterpret_cast<char*>(cb_smem_tma) + HornerHyb<L>::SMEM;
        auto tile_m = [&](int b) { return reinterpret_cast<uint32_t*>(tiles + (size_t)b * (HornerTile<L>::MANT + HornerTile<L>::META)); };
        auto tile_t = [&](int b) {
            return reinterpret_cast<Meta*>(tiles + (size_t)b * (HornerTile<L>::MANT + HornerTile<L>::META) + HornerTile<L>::MANT);
        };
        const int64_t idx = (int64_t)blockIdx.x * TPB + threadIdx.x;
        if (threadIdx.x == 0) {
            tma_mbar_init(&mbar[0], 1);
            tma_mbar_init(&mbar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        const int64_t ntiles = (p.len + T - 1) / T;
        // tile k holds entries j_lo(k) .. j_hi(k), j_hi(k) = len - 1 - k T, walking down the series as Horner does
        auto issue = [&](int64_t k) {
            const int64_t j_hi = p.len - 1 - k * T, j_lo = j_hi - T + 1 > 0 ? j_hi - T + 1 : 0;
            const uint32_t c = (uint32_t)(j_hi - j_lo + 1);
            const int b = (int)(k & 1);
            tma_expect(&mbar[b], c * (uint32_t)(L * 2 * 4 + 2 * 16));
            tma_bulk_g2s(tile_m(b), p.f.m + (size_t)j_lo * L * 2, c * (uint32_t)(L * 2 * 4), &mbar[b]);
            tma_bulk_g2s(tile_t(b), p.f.t + (size_t)j_lo * 2, c * (uint32_t)(2 * 16), &mbar[b]);
        };
        if (threadIdx.x == 0 && ntiles > 0) issue(0);
        if (idx >= p.npts) return;  // (thread 0 of an existing block is always a live point)
        HybridScratch<TPB, HybCfg<L>::KS> S0{sbase, p.gcols + idx, (uint32_t)p.npts};
        HybridScratch<TPB, HybCfg<L>::KS> S1{sbase + 4u * (uint32_t)(HybCfg<L>::KS * TPB),
                                             p.gcols + (size_t)HybCfg<L>::GW * (size_t)p.npts + idx, (uint32_t)p.npts};
        const size_t slot = 2 * (size_t)idx;
        CRef a = at<L>(p.pts, 0, (p.pts.S == 2) ? 0 : slot), tmp = at<L>(p.tmp, 0, slot);
        for (int k = 0; k < p.nder; k++) acb_zero<L>(at<L>(p.out, k, slot));
#pragma unroll 1
        for (int64_t k = 0; k < ntiles; k++) {
            const int b = (int)(k & 1);
            const int64_t j_hi = p.len - 1 - k * T, j_lo = j_hi - T + 1 > 0 ? j_hi - T + 1 : 0;
            // the buffer of tile k + 1 was last read during tile k - 1: every live thread has passed the barrier at the
            // end of that tile
            if (threadIdx.x == 0 && k + 1 < ntiles) issue(k + 1);
            tma_wait(&mbar[b], (uint32_t)((k >> 1) & 1));
#pragma unroll 1
            for (int64_t j = j_hi; j >= j_lo; j--) {
                CRef fj{tile_m(b) + (size_t)(j - j_lo) * L * 2, (size_t)2, tile_t(b) + (size_t)(j - j_lo) * 2};
                if (j == p.len - 1) {
                    acb_set<L>(at<L>(p.out, 0, slot), fj);
                    continue;
                }
#pragma unroll 1
                for (int d = p.nder - 1; d >= 0; d--) {
                    CRef o = at<L>(p.out, d, slot);
                    CB_PHASE_SYNC();
                    acb_mul<L>(tmp, o, a, S0, S1);
                    CB_PHASE_SYNC();
                    if (d > 0)
                        acb_addsub<L>(o, tmp, at<L>(p.out, d - 1, slot), false, S0, S1);
                    else
                        acb_addsub<L>(o, tmp, fj, false, S0, S1);
                }
            }
            __syncthreads();
        }
#endif
    }
}

template <int L>
__global__ void __launch_bounds__(Cfg<L>::TPB, 1) ew_kernel(EwParams p) {
    CB_KERNEL_PROLOGUE(p.n)
    ew_thread<L>(p, idx, S0, S1);
}

template <int L>
__global__ void __launch_bounds__(Cfg<L>::TPB, 1) fuchs_table_kernel(FuchsTableParams p) {
    CB_KERNEL_PROLOGUE(p.nrows * p.ninst)
    fuchs_table_fast_thread<L>(p, idx, S0, S1);
}
template <int L, int TPB, bool RR = false, bool SYNC = false, bool PI = false>
__global__ void __launch_bounds__(TPB, 1) fuchs_shared_kernel(FuchsSharedParams p) {
    extern __shared__ uint32_t cb_smem[];
    constexpr uint32_t W0 = Sizes<L>::S0WORDS, W1 = Sizes<L>::S1WORDS, AW = Sizes<L>::AWORDS;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(cb_smem) + 4u * threadIdx.x;
    Scratch<TPB> S0{sbase}, S1{sbase + 4u * W0 * TPB};
    Scratch<TPB> A0{sbase + 4u * (W0 + W1) * TPB}, A1{sbase + 4u * (W0 + W1 + AW) * TPB},
        A2{sbase + 4u * (W0 + W1 + 2 * AW) * TPB};
    const int64_t idx = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (idx >= p.batch) return;
    GScratch G0{p.gcols + idx, (size_t)p.batch};
    GScratch G1{p.gcols + (size_t)Sizes<L>::WORDS * p.batch + idx, (size_t)p.batch};
    fuchs_shared_thread<L, RR, SYNC, PI>(p, idx, S0, S1, A0, A1, A2, G0, G1);
}

// Persistent form of the recurrence kernel.  Every coefficient is formed from coefficients read back from res, so
// the work of an instance block can be cut anywhere along the coefficient index: a work unit is (block of TPB
// instances, range of range_len coefficients).  One CTA per SM takes units from a counter in coefficient-major
// order; unit (b, r) waits for (b, r - 1), which was handed out one full round earlier.  This removes the
// last-wave tail for ANY batch size, so the 12-warp variant (with phase barriers: 1.19x per SM) can serve batches
// that are not a multiple of 148 * 384 instances -- BASELINE's 65 536 included.
// Idle lanes of a ragged last block recompute the last valid instance (identical stores), so that every thread
// of the CTA reaches every barrier.
template <int L, int TPB>
__global__ void __launch_bounds__(TPB, 1) fuchs_shared_persistent_kernel(FuchsSharedParams p) {
    extern __shared__ uint32_t cb_smem[];
    __shared__ int unit_s;
    constexpr uint32_t W0 = Sizes<L>::S0WORDS, W1 = Sizes<L>::S1WORDS, AW = Sizes<L>::AWORDS;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(cb_smem) + 4u * threadIdx.x;
    Scratch<TPB> S0{sbase}, S1{sbase + 4u * W0 * TPB};
    Scratch<TPB> A0{sbase + 4u * (W0 + W1) * TPB}, A1{sbase + 4u * (W0 + W1 + AW) * TPB},
        A2{sbase + 4u * (W0 + W1 + 2 * AW) * TPB};
    const int64_t nblocks = (p.batch + TPB - 1) / TPB;
    const int64_t first = p.n_begin > -p.valuation ? p.n_begin : -p.valuation;
    const int64_t nranges = (p.n - first + p.range_len) / p.range_len;
    const int64_t total = nblocks * nranges;
    for (;;) {
        if (threadIdx.x == 0) {
            int u = atomicAdd(p.sched, 1);
            if (u < total) {
                int64_t b = u % nblocks, r = u / nblocks;
                while (r > 0 && atomicAdd(p.sched + 1 + b, 0) < (int)r) __nanosleep(200);
                __threadfence();
            }
            unit_s = u;
        }
        __syncthreads();
        const int64_t u = unit_s;
        if (u >= total) break;
        const int64_t b = u % nblocks, r = u / nblocks;
        int64_t inst = b * TPB + threadIdx.x;
        if (inst >= p.batch) inst = p.batch - 1;
        const int64_t lo = first + r * p.range_len;
        const int64_t hi = lo + p.range_len - 1 < p.n ? lo + p.range_len - 1 : p.n;
        // every thread -- idle lanes included -- owns a column of the (padded) general-path scratch
        const size_t gstride = (size_t)nblocks * TPB, gi = (size_t)b * TPB + threadIdx.x;
        GScratch G0{p.gcols + gi, gstride};
        GScratch G1{p.gcols + (size_t)Sizes<L>::WORDS * gstride + gi, gstride};
        fuchs_shared_thread<L, false, true>(p, inst, S0, S1, A0, A1, A2, G0, G1, lo, hi);
        __threadfence();  // this thread's coefficients are visible device-wide before the block reports the range
        __syncthreads();
        if (threadIdx.x == 0) atomicExch(p.sched + 1 + b, (int)(r + 1));
    }
}

// integer-operator kernel: three accumulator columns of L + 2 words per thread in shared memory (the general body's
// product column is in global memory: FuchsIntParams::gs0).  1024 bits: 12 warps at 168 registers -- 16 warps at 128
// registers (CB_INTOP_TPB32=512) were measured 4 % SLOWER: 420 bytes of spills put the gain back into local-memory waits
#ifndef CB_INTOP_TPB32
#define CB_INTOP_TPB32 384
#endif
template <int L> struct IntopCfg {
    static constexpr int TPB = L <= 4 ? 384 : (L <= 32 ? CB_INTOP_TPB32 : (L == 64 ? 192 : 96));
    static constexpr size_t SMEM = (size_t)3 * (L + 2) * TPB * sizeof(uint32_t);
};
inline int intop_tpb(int L) { return L <= 4 ? IntopCfg<4>::TPB : (L <= 32 ? IntopCfg<32>::TPB : (L == 64 ? IntopCfg<64>::TPB : IntopCfg<128>::TPB)); }
template <int L>
__global__ void __launch_bounds__(IntopCfg<L>::TPB, 1) fuchs_intop_kernel(FuchsIntParams p) {
    extern __shared__ uint32_t cb_smem[];
    constexpr int TPB = IntopCfg<L>::TPB;
    constexpr uint32_t AW = L + 2;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(cb_smem) + 4u * threadIdx.x;
    Scratch<TPB> A0{sbase}, A1{sbase + 4u * AW * TPB}, A2{sbase + 4u * 2 * AW * TPB};
    const int64_t idx = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (idx >= p.batch) return;
    GScratch S0{p.gs0 + idx, p.gs0_stride};
    if constexpr (L <= 32) fuchs_intop_fast_thread<L>(p, idx, S0, A0, A1, A2);
    else fuchs_intop_thread<L>(p, idx, S0, A0, A1, A2);
}

// persistent form of the integer-operator kernel (same work-unit scheme as fuchs_shared_persistent_kernel; the
// thread body has no barriers, so idle lanes of a ragged last block simply skip the work)
template <int L>
__global__ void __launch_bounds__(IntopCfg<L>::TPB, 1) fuchs_intop_persistent_kernel(FuchsIntParams p) {
    extern __shared__ uint32_t cb_smem[];
    __shared__ int unit_s;
    constexpr int TPB = IntopCfg<L>::TPB;
    constexpr uint32_t AW = L + 2;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(cb_smem) + 4u * threadIdx.x;
    Scratch<TPB> A0{sbase}, A1{sbase + 4u * AW * TPB}, A2{sbase + 4u * 2 * AW * TPB};
    const int64_t nblocks = (p.batch + TPB - 1) / TPB;
    const int64_t first = -p.valuation;
    const int64_t nranges = (p.n - first + p.range_len) / p.range_len;
    const int64_t total = nblocks * nranges;
    for (;;) {
        if (threadIdx.x == 0) {
            int u = atomicAdd(p.sched, 1);
            if (u < total) {
                int64_t b = u % nblocks, r = u / nblocks;
                while (r > 0 && atomicAdd(p.sched + 1 + b, 0) < (int)r) __nanosleep(200);
                __threadfence();
            }
            unit_s = u;
        }
        __syncthreads();
        const int64_t u = unit_s;
        if (u >= total) break;
        const int64_t b = u % nblocks, r = u / nblocks;
        const int64_t inst = b * TPB + threadIdx.x;
        const int64_t lo = first + r * p.range_len;
        const int64_t hi = lo + p.range_len - 1 < p.n ? lo + p.range_len - 1 : p.n;
        GScratch S0{p.gs0 + (size_t)b * TPB + threadIdx.x, p.gs0_stride};
        if (inst < p.batch) {
            if constexpr (L <= 32) fuchs_intop_fast_thread<L>(p, inst, S0, A0, A1, A2, lo, hi);
            else fuchs_intop_thread<L>(p, inst, S0, A0, A1, A2, lo, hi);
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicExch(p.sched + 1 + b, (int)(r + 1));
    }
}

// time of a launch is proportional to (number of waves) x (instances per SM per wave) / throughput;
// relative per-SM throughputs measured on B200 (7, 12 and 14 warps per SM)
// (measured, 98-coefficient runs: 224 threads 21.7 instances/ms/SM, 384 threads 23.4 -- 25.9 with the phase
// barriers it is built with --, 448 threads -- 128 registers, 532 B of spills -- 21.4: never chosen; persistent
// launches at the full BASELINE size: 384 threads 317 ms, 352 340 ms, 416 400 ms, 448 382 ms)
constexpr double THR_BIG = 1.19, THR_HUGE = 0.98, THR_PERS64 = 1.23;  // 2048 bits, measured: 5.3-5.45 T against 4.3 T whole-wave
// 0: TPB, 1: TPB_BIG, 2: TPB_HUGE.  CB200_TPB_CHOICE=0/1/2 in the environment overrides (development)
template <int L>
int choose_block_size(int64_t batch, int sms) {
    const char* e = getenv("CB200_TPB_CHOICE");
    if (e && e[0] >= '0' && e[0] <= '2') return e[0] - '0';
    auto cost = [&](int tpb, double thr) {
        int64_t blocks = (batch + tpb - 1) / tpb;
        int64_t waves = (blocks + sms - 1) / sms;
        return (double)waves * tpb / thr;
    };
    int best = 0;
    double c = cost(CfgShared<L>::TPB, 1.0);
    if (CfgShared<L>::TPB_BIG != CfgShared<L>::TPB && cost(CfgShared<L>::TPB_BIG, THR_BIG) < c) {
        best = 1;
        c = cost(CfgShared<L>::TPB_BIG, THR_BIG);
    }
    if (CfgShared<L>::TPB_HUGE != CfgShared<L>::TPB && cost(CfgShared<L>::TPB_HUGE, THR_HUGE) < c) best = 2;
    return best;
}

template <int L, int PT>
int launch_persistent(const FuchsSharedParams& p, int sms, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(fuchs_shared_persistent_kernel<L, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(Sizes<L>::SMEM_SHARED_PER_THREAD * PT));
    if (e != cudaSuccess) return (int)e;
    int64_t nblocks = (p.batch + PT - 1) / PT;
    e = cudaMemsetAsync(p.sched, 0, sizeof(int) * (size_t)(nblocks + 1), s);
    if (e != cudaSuccess) return (int)e;
    int per_sm = 1;  // resident CTAs per SM (2048 bits: two 96-thread CTAs of 81 KB)
    if (L == 64) per_sm = 2;
    if (const char* e = getenv("CB200_PERS_CTAS")) { int v = atoi(e); if (v >= 1 && v <= 4) per_sm = v; }  // development
    int64_t cap = (int64_t)sms * per_sm;
    int grid = (int)(nblocks < cap ? nblocks : cap);
    fuchs_shared_persistent_kernel<L, PT><<<grid, PT, Sizes<L>::SMEM_SHARED_PER_THREAD * PT, s>>>(p);
    return (int)cudaGetLastError();
}

// persistent launch: worthwhile when the batch is more than one wave of its blocks and the fractional-wave cost
// of the barrier-synchronised variant beats the best whole-wave choice.  CB200_PERSISTENT=0/1 in the environment
// overrides.  1024 bits: 12 warps (TPB_BIG); 2048 bits: the same 3 warps as the plain kernel, so the gain is the
// missing last-wave tail plus what the phase barriers give.
template <int L> struct PersCfg { static constexpr int TPB = 0; static constexpr double THR = 1.0; };
template <> struct PersCfg<32> { static constexpr int TPB = CfgShared<32>::TPB_BIG; static constexpr double THR = THR_BIG; };
template <> struct PersCfg<64> { static constexpr int TPB = CfgShared<64>::TPB; static constexpr double THR = THR_PERS64; };
template <int L>
bool use_persistent(int64_t batch, int sms) {
    if (PersCfg<L>::TPB == 0) return false;
    const char* e = getenv("CB200_PERSISTENT");
    if (e && e[0] == '0') return false;
    if (e && e[0] == '1') return true;
    auto cost = [&](int tpb, double thr) {
        int64_t blocks = (batch + tpb - 1) / tpb;
        int64_t waves = (blocks + sms - 1) / sms;
        return (double)waves * tpb / thr;
    };
    double best = cost(CfgShared<L>::TPB, 1.0);
    if (CfgShared<L>::TPB_BIG != CfgShared<L>::TPB) {
        double big = cost(CfgShared<L>::TPB_BIG, THR_BIG);
        if (big < best) best = big;
    }
    constexpr int PT = PersCfg<L>::TPB > 0 ? PersCfg<L>::TPB : 1;
    int64_t blocks = (batch + PT - 1) / PT;
    if (blocks <= sms) return L == 64 && batch >= PT;  // 2048 bits: same block size, the barriers alone give 1.23x
    double pers = (double)blocks * PT / sms / PersCfg<L>::THR * 1.02;  // 2 % for the ragged last round
    return pers < best;
}

inline int sm_count() {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}
template <int L, class P>
int launch(void (*kern)(P), const P& p, int64_t count, cudaStream_t st, int tpb = Cfg<L>::TPB,
           size_t smem = Sizes<L>::SMEM) {
    if (count <= 0) return 0;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    unsigned grid = (unsigned)((count + tpb - 1) / tpb);
    kern<<<grid, tpb, smem, st>>>(p);
    return (int)cudaGetLastError();
}

// hybrid-column kernel or the all-shared one?  A wave of the hybrid kernel carries more threads but takes
// WAVE_RATIO times as long (measured, full waves); CB200_HYBRID=0/1 in the environment overrides (development).
template <int L>
bool use_hybrid(int64_t count, int sms, double wave_ratio) {
    if (!HybCfg<L>::ON) return false;
    const char* e = getenv("CB200_HYBRID");
    if (e && e[0] == '0') return false;
    if (e && e[0] == '1') return true;
    auto waves = [&](int tpb) { return (double)(((count + tpb - 1) / tpb + sms - 1) / sms); };
    return waves(HybCfg<L>::TPB) * wave_ratio < waves(Cfg<L>::TPB);
}
// time of a full wave of the hybrid kernel / time of a full wave of the all-shared kernel
template <int L> struct HybRatio { static constexpr double HORNER = 1.0, FROB = 1.0; };
// (B200, whole waves of both: Horner 2048 bits 5.45 -> 6.99 T limb-MAC/s, 4096 bits 4.79 -> 6.45; Frobenius M = 1
// 2048 bits 2.30 -> 2.66, 4096 bits 2.21 -> 2.80 -- since the long division keeps its window in a ring of L + 2 words;
// with the window sliding through the global tails the 4096-bit Frobenius kernel lost 7 %)
template <> struct HybRatio<64> { static constexpr double HORNER = 1.04, FROB = 1.16; };
template <> struct HybRatio<128> { static constexpr double HORNER = 1.27, FROB = 1.35; };

// ---- cooperative kernels: when, and how they are launched
// columns (lanes) the cooperative kernels need tails for: the batch padded to whole CTAs, four lanes per instance
inline size_t coop_tail_words(int L, int64_t count) {
    if (L < 64) return 0;
    const int64_t c = count > 0 ? count : 0;
    const int64_t g2 = L >= 128 ? Coop2Cfg<128>::G : Coop2Cfg<64>::G;  // the two-lane form pads to its own CTA size
    const int64_t cols4 = ((c + 63) / 64) * 256, cols2 = ((c + g2 - 1) / g2) * 4 * g2;
    return (size_t)(L - 2) * (size_t)(cols4 > cols2 ? cols4 : cols2);
}
// CB200_COOP=0/1 in the environment overrides (development; 2 / 4 force that many lanes for the Frobenius recurrence).
// Default, from measurements on B200 (tools/quick_coop.py, profiles/coop_r02.txt): four lanes per instance win while
// the batch is at most ONE wave of cooperative CTAs (64 instances per SM) -- Horner 2.0-3.4x at 2 .. 4096 points (2048
// and 4096 bits; with the CTA-wide phase barriers it stays ahead up to three waves at 2048 bits and at every size at
// 4096 bits: 1.5-1.8x), the Frobenius recurrence 1.9-2.3x -- and lose beyond about two waves (the lanes that only form
// products idle in the tails and the long divisions).  Beyond one wave the Frobenius recurrence runs on TWO lanes per
// instance (every lane busy in every phase, twice the warps of the one-thread kernel; with CTA-wide phase barriers,
// without which seven warps at different places of the code stall on instruction fetch): 1.5x at 2048 bits, 1.09x at
// 4096 bits up to one wave of the one-thread kernel; larger 4096-bit batches go to the persistent hybrid kernel.
template <int L>
bool coop_wanted(int64_t count, int sms, bool frobenius) {
    if (!CoopCfg<L>::ON) return false;
    const char* e = getenv("CB200_COOP");
    if (e && e[0] == '0') return false;
    if (e && e[0] == '1') return true;
    // Horner at 4096 bits: four lanes at every size (65 536 points: 54 ms against 96 ms on the one-thread hybrid kernel)
    if (!frobenius && L >= 128) return true;
    // Horner at 2048 bits: 1.8x at 16 384 points, level at 65 536, behind at 262 144
    if (!frobenius) return count <= (int64_t)3 * sms * CoopCfg<L>::G;
    return count <= (int64_t)sms * CoopCfg<L>::G;
}
// lanes per instance for the Frobenius recurrence (0: the one-thread kernels)
template <int L>
int coop_frob_lanes(const FrobParams& p) {
    if (!CoopCfg<L>::ON || p.rhs_len != 0 || p.sched == nullptr || p.gcols_words < coop_tail_words(L, p.batch)) return 0;
    const char* e = getenv("CB200_COOP");
    if (e && e[0] == '2') return 2;
    if (e && e[0] == '4') return 4;
    if (e && e[0] == '0') return 0;
    if (coop_wanted<L>(p.batch, sm_count(), true)) return 4;
    if (L == 64) return 2;
    // 4096 bits: two lanes up to one wave of the all-shared kernel (1.09x); beyond that the persistent hybrid kernel
    return p.batch <= (int64_t)sm_count() * Cfg<L>::TPB ? 2 : 0;
}
// the persistent hybrid kernel: homogeneous solves of more than one wave of the all-shared kernel (CB200_FROBPERS=0/1
// in the environment overrides)
template <int L>
bool use_frob_pers(const FrobParams& p) {
    if (!FrobPersCfg<L>::ON || p.rhs_len != 0 || p.sched == nullptr) return false;
    const int64_t padded = (p.batch + FrobPersCfg<L>::TPB - 1) / FrobPersCfg<L>::TPB * FrobPersCfg<L>::TPB;
    if (p.gcols_words < (size_t)2 * HybCfg<L>::GW * (size_t)padded) return false;
    const char* e = getenv("CB200_FROBPERS");
    if (e && e[0] == '0') return false;
    if (e && e[0] == '1') return true;
    return p.batch > (int64_t)sm_count() * Cfg<L>::TPB;
}
template <int L>
int launch_frob_pers(const FrobParams& p, cudaStream_t s) {
    if (p.batch <= 0) return 0;
    cudaError_t e = cudaFuncSetAttribute(frobenius1_pers_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FrobPersCfg<L>::SMEM);
    if (e != cudaSuccess) return (int)e;
    const int64_t nblocks = (p.batch + FrobPersCfg<L>::TPB - 1) / FrobPersCfg<L>::TPB;
    e = cudaMemsetAsync(p.sched, 0, sizeof(int) * (size_t)(nblocks + 1), s);
    if (e != cudaSuccess) return (int)e;
    const int sms = sm_count();
    const int grid = (int)(nblocks < sms ? nblocks : sms);
    frobenius1_pers_kernel<L><<<grid, FrobPersCfg<L>::TPB, FrobPersCfg<L>::SMEM, s>>>(p);
    return (int)cudaGetLastError();
}
template <int L, int LANES>
int launch_coop_frob(const FrobParams& p, cudaStream_t s) {
    if (p.batch <= 0) return 0;
    using Geo = CoopGeom<L, LANES>;
    cudaError_t e = cudaFuncSetAttribute(frobenius1_coop_kernel<L, LANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo::SMEM);
    if (e != cudaSuccess) return (int)e;
    const int64_t nblocks = (p.batch + Geo::G - 1) / Geo::G;
    e = cudaMemsetAsync(p.sched, 0, sizeof(int) * (size_t)(nblocks + 1), s);
    if (e != cudaSuccess) return (int)e;
    const int sms = sm_count();
    const int grid = (int)(nblocks < sms ? nblocks : sms);
    frobenius1_coop_kernel<L, LANES><<<grid, Geo::TPB, Geo::SMEM, s>>>(p);
    return (int)cudaGetLastError();
}
// the Fuchs recurrence with a shared operator on four lanes per instance: at 2048 / 4096 bits at EVERY batch size.
// Measured on B200 (profiles/coop_fuchs_r02.txt, 60 coefficients): 2048 bits 2.4x at 256 .. 4 096 instances, 1.9x at
// 16 384, 1.65x at 262 144 (8.0 T against 4.8 T limb-MAC/s); 4096 bits 3.2-3.7x everywhere (6.9 T against 2.0 T) --
// the one-thread kernels run 3 warps (2048 bits) or ONE warp (4096 bits) per SM, the cooperative kernel eight / four.
template <int L>
bool use_coop_fuchs(const FuchsSharedParams& p, int sms) {
    (void)sms;
    if (!CoopCfg<L>::ON || p.per_inst || !p.sched) return false;
    const char* e = getenv("CB200_COOP");
    if (e && e[0] == '0') return false;
    return true;
}
template <int L>
int launch_coop_fuchs(const FuchsSharedParams& p, int sms, cudaStream_t s) {
    if (p.batch <= 0) return 0;
    cudaError_t e = cudaFuncSetAttribute(fuchs_shared_coop_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)CoopFuchsCfg<L>::SMEM);
    if (e != cudaSuccess) return (int)e;
    const int64_t nblocks = (p.batch + CoopFuchsCfg<L>::G - 1) / CoopFuchsCfg<L>::G;
    e = cudaMemsetAsync(p.sched, 0, sizeof(int) * (size_t)(nblocks + 1), s);
    if (e != cudaSuccess) return (int)e;
    const int grid = (int)(nblocks < sms ? nblocks : sms);
    fuchs_shared_coop_kernel<L><<<grid, CoopFuchsCfg<L>::TPB, CoopFuchsCfg<L>::SMEM, s>>>(p);
    return (int)cudaGetLastError();
}
template <int L>
bool use_coop_horner(const HornerParams& p) {
    if (!CoopCfg<L>::ON || p.gcols_words < coop_tail_words(L, p.npts)) return false;
    const char* e = getenv("CB200_COOP");
    if (e && e[0] == '2') return true;
    return coop_wanted<L>(p.npts, sm_count(), false);
}
template <int L, int LANES>
int launch_coop_horner_lanes(const HornerParams& p, cudaStream_t s) {
    using Geo = CoopGeom<L, LANES>;
    cudaError_t e = cudaFuncSetAttribute(horner_coop_kernel<L, LANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo::SMEM);
    if (e != cudaSuccess) return (int)e;
    const unsigned grid = (unsigned)((p.npts + Geo::G - 1) / Geo::G);
    horner_coop_kernel<L, LANES><<<grid, Geo::TPB, Geo::SMEM, s>>>(p);
    return (int)cudaGetLastError();
}
// CB200_COOP=2 forces two lanes per point (development)
template <int L>
int launch_coop_horner(const HornerParams& p, cudaStream_t s) {
    if (p.npts <= 0) return 0;
    const char* e = getenv("CB200_COOP");
    if (e && e[0] == '2') return launch_coop_horner_lanes<L, 2>(p, s);
    return launch_coop_horner_lanes<L, 4>(p, s);
}

// the TMA-staged Horner kernel: one series for all points (S = 2), long enough for a few tiles, 16-byte aligned
// CB200_TMA=0/1 in the environment overrides (development)
template <int L>
bool use_horner_tma(const HornerParams& p) {
    if (!HybCfg<L>::ON) return false;
    if (p.f.S != 2 || ((uintptr_t)p.f.m & 15) || ((uintptr_t)p.f.t & 15)) return false;
    const char* e = getenv("CB200_TMA");
    if (e && e[0] == '0') return false;
    if (e && e[0] == '1') return true;
    return p.len >= 4 * HornerTile<L>::T;
}

// per-precision launch table, filled by kern_L*.cu
struct LaunchTable {
    int (*frobenius1)(const FrobParams&, cudaStream_t);
    int (*horner)(const HornerParams&, cudaStream_t);
    int (*ew)(const EwParams&, cudaStream_t);
    int (*fuchs_table)(const FuchsTableParams&, cudaStream_t);
    int (*fuchs_shared)(const FuchsSharedParams&, cudaStream_t);
    int (*fuchs_intop)(const FuchsIntParams&, cudaStream_t);
    int (*frobenius_general)(const FrobGenParams&, cudaStream_t);
    int (*solution_op)(const SolOpParams&, cudaStream_t);
    int (*indicial_poly)(const IndicialParams&, cudaStream_t);
    int (*evaluate)(const EvalParams&, cudaStream_t);
    int (*trans)(const TransParams&, cudaStream_t);
    int (*apply)(const ApplyParams&, cudaStream_t);
    int (*ode_shift)(const ShiftParams&, cudaStream_t);
    int (*taylor_shift)(const TShiftParams&, cudaStream_t);
    int (*radius)(const RadiusParams&, cudaStream_t);
    int (*truncation)(const TruncParams&, cudaStream_t);
};

// The launch functions of one precision are spread over four translation units (kern_L<LL>.cu: Frobenius M = 1,
// Horner, elementwise + the table itself; kern_L<LL>f.cu: the Fuchs recurrences; kern_L<LL>t.cu: evaluation, exp / log, radius / truncation; kern_L<LL>x.cu: the
// rest), so
// that the build parallelises: each kernel is compiled on its own and the 2048 / 4096-bit units took minutes each.
#define CB_DECLARE_LAUNCHERS(LL)                                                                           \
    int ftab_##LL(const FuchsTableParams&, cudaStream_t);                                                  \
    int fsh_##LL(const FuchsSharedParams&, cudaStream_t);                                                  \
    int fint_##LL(const FuchsIntParams&, cudaStream_t);                                                    \
    int frobg_##LL(const FrobGenParams&, cudaStream_t);                                                    \
    int solop_##LL(const SolOpParams&, cudaStream_t);                                                      \
    int indp_##LL(const IndicialParams&, cudaStream_t);                                                    \
    int eval_##LL(const EvalParams&, cudaStream_t);                                                        \
    int trans_##LL(const TransParams&, cudaStream_t);                                                      \
    int apply_##LL(const ApplyParams&, cudaStream_t);                                                      \
    int radius_##LL(const RadiusParams&, cudaStream_t);                                                    \
    int trunc_##LL(const TruncParams&, cudaStream_t);                                                      \
    int oshift_##LL(const ShiftParams&, cudaStream_t);                                                     \
    int tshift_##LL(const TShiftParams&, cudaStream_t);

#define CB_DEFINE_LAUNCH_FUCHS(LL)                                                                         \
    namespace cb {                                                                                         \
    CB_DECLARE_LAUNCHERS(LL)                                                                               \
    int ftab_##LL(const FuchsTableParams& p, cudaStream_t s) {                                      \
        return launch<LL>(fuchs_table_kernel<LL>, p, p.nrows * p.ninst, s);                            \
    }                                                                                                      \
    int fsh_##LL(const FuchsSharedParams& p, cudaStream_t s) {                                      \
        int dev = 0, sms = 148;                                                                            \
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); \
        if (use_coop_fuchs<LL>(p, sms)) return launch_coop_fuchs<LL>(p, sms, s);                           \
        if (p.per_inst) /* a table row per instance (operators differ between instances) */                \
            return launch<LL>(fuchs_shared_kernel<LL, CfgShared<LL>::TPB, false, false, true>, p, p.batch, s,       \
                              CfgShared<LL>::TPB, Sizes<LL>::SMEM_SHARED_PER_THREAD * CfgShared<LL>::TPB);  \
        if (use_rr<LL>() && !p.no_rr) /* opt-in experiment (CB200_RR=1): its own instantiation */          \
            return launch<LL>(fuchs_shared_kernel<LL, CfgShared<LL>::TPB, true>, p, p.batch, s, CfgShared<LL>::TPB, \
                              Sizes<LL>::SMEM_SHARED_PER_THREAD * CfgShared<LL>::TPB);                     \
        if (p.sched && use_persistent<LL>(p.batch, sms)) {                                                 \
            return launch_persistent<LL, (PersCfg<LL>::TPB > 0 ? PersCfg<LL>::TPB : CfgShared<LL>::TPB)>(p, sms, s); \
        }                                                                                                  \
        int choice = choose_block_size<LL>(p.batch, sms);                                                  \
        if (choice == 2)                                                                                   \
            return launch<LL>(fuchs_shared_kernel<LL, CfgShared<LL>::TPB_HUGE>, p, p.batch, s, CfgShared<LL>::TPB_HUGE, \
                              Sizes<LL>::SMEM_SHARED_PER_THREAD * CfgShared<LL>::TPB_HUGE);                \
        if (choice == 1)                                                                                   \
            return launch<LL>(fuchs_shared_kernel<LL, CfgShared<LL>::TPB_BIG, false, true>, p, p.batch, s, CfgShared<LL>::TPB_BIG, \
                              Sizes<LL>::SMEM_SHARED_PER_THREAD * CfgShared<LL>::TPB_BIG);                 \
        return launch<LL>(fuchs_shared_kernel<LL, CfgShared<LL>::TPB>, p, p.batch, s, CfgShared<LL>::TPB,  \
                          Sizes<LL>::SMEM_SHARED_PER_THREAD * CfgShared<LL>::TPB);                         \
    }                                                                                                      \
    int fint_##LL(const FuchsIntParams& p, cudaStream_t s) {                                        \
        const size_t smem = IntopCfg<LL>::SMEM;                                                            \
        if (p.sched) { /* persistent launch: no last-wave tail */                                          \
            int dev = 0, sms = 148;                                                                        \
            if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); \
            cudaError_t e = cudaFuncSetAttribute(fuchs_intop_persistent_kernel<LL>,                        \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
            if (e != cudaSuccess) return (int)e;                                                           \
            int64_t nblocks = (p.batch + IntopCfg<LL>::TPB - 1) / IntopCfg<LL>::TPB;                       \
            int grid = (int)(nblocks < sms ? nblocks : sms);                                               \
            fuchs_intop_persistent_kernel<LL><<<grid, IntopCfg<LL>::TPB, smem, s>>>(p);                    \
            return (int)cudaGetLastError();                                                                \
        }                                                                                                  \
        return launch<LL>(fuchs_intop_kernel<LL>, p, p.batch, s, IntopCfg<LL>::TPB, smem);                 \
    }                                                                                                      \
    }

#define CB_DEFINE_LAUNCH_OTHER(LL)                                                                         \
    namespace cb {                                                                                         \
    CB_DECLARE_LAUNCHERS(LL)                                                                               \
    int frobg_##LL(const FrobGenParams& p, cudaStream_t s) { return launch<LL>(frobenius_general_kernel<LL>, p, p.batch, s); } \
    int solop_##LL(const SolOpParams& p, cudaStream_t s) { return launch<LL>(solution_op_kernel<LL>, p, p.batch, s); } \
    int indp_##LL(const IndicialParams& p, cudaStream_t s) { return launch<LL>(indicial_poly_kernel<LL>, p, p.batch, s); } \
    int apply_##LL(const ApplyParams& p, cudaStream_t s) { return launch<LL>(apply_kernel<LL>, p, p.batch, s); } \
    int oshift_##LL(const ShiftParams& p, cudaStream_t s) { return launch<LL>(ode_shift_kernel<LL>, p, p.batch, s); } \
    int tshift_##LL(const TShiftParams& p, cudaStream_t s) {                                        \
        if (p.batch <= 0 || p.len <= 1) return 0;                                                          \
        cudaError_t e = cudaFuncSetAttribute(taylor_shift_kernel<LL>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)Sizes<LL>::SMEM);                                        \
        if (e != cudaSuccess) return (int)e;                                                               \
        taylor_shift_kernel<LL><<<(unsigned)p.batch, Cfg<LL>::TPB, Sizes<LL>::SMEM, s>>>(p);               \
        return (int)cudaGetLastError();                                                                    \
    }                                                                                                      \
    }

#define CB_DEFINE_LAUNCH_TRANS(LL)                                                                         \
    namespace cb {                                                                                         \
    CB_DECLARE_LAUNCHERS(LL)                                                                               \
    int eval_##LL(const EvalParams& p, cudaStream_t s) { return launch<LL>(evaluate_kernel<LL>, p, p.npts, s); } \
    int trans_##LL(const TransParams& p, cudaStream_t s) { return launch<LL>(trans_kernel<LL>, p, p.n, s); } \
    int radius_##LL(const RadiusParams& p, cudaStream_t s) { return launch<LL>(radius_kernel<LL>, p, p.batch, s); } \
    int trunc_##LL(const TruncParams& p, cudaStream_t s) { return launch<LL>(truncation_kernel<LL>, p, p.batch, s); } \
    }

#define CB_DEFINE_LAUNCH_TABLE(LL)                                                                         \
    namespace cb {                                                                                         \
    CB_DECLARE_LAUNCHERS(LL)                                                                               \
    static int frob_##LL(const FrobParams& p, cudaStream_t s) {                                            \
        if (int lanes = coop_frob_lanes<LL>(p)) return lanes == 2 ? launch_coop_frob<LL, 2>(p, s) : launch_coop_frob<LL, 4>(p, s);                                       \
        if (use_frob_pers<LL>(p)) return launch_frob_pers<LL>(p, s);                                       \
        if (use_hybrid<LL>(p.batch, sm_count(), HybRatio<LL>::FROB))                                       \
            return launch<LL>(frobenius1_hyb_kernel<LL>, p, p.batch, s, HybCfg<LL>::TPB, HybCfg<LL>::SMEM); \
        return launch<LL>(frobenius1_kernel<LL>, p, p.batch, s);                                           \
    }                                                                                                      \
    static int horner_##LL(const HornerParams& p, cudaStream_t s) {                                        \
        if (use_coop_horner<LL>(p)) return launch_coop_horner<LL>(p, s);                                   \
        if (use_hybrid<LL>(p.npts, sm_count(), HybRatio<LL>::HORNER)) {                                    \
            if (use_horner_tma<LL>(p))                                                                     \
                return launch<LL>(horner_tma_kernel<LL>, p, p.npts, s, HornerHyb<LL>::TPB,                  \
                                  HornerHyb<LL>::SMEM + HornerTile<LL>::BYTES);                            \
            return launch<LL>(horner_hyb_kernel<LL>, p, p.npts, s, HornerHyb<LL>::TPB, HornerHyb<LL>::SMEM); \
        }                                                                                                  \
        return launch<LL>(horner_kernel<LL>, p, p.npts, s);                                                \
    }                                                                                                      \
    static int ew_##LL(const EwParams& p, cudaStream_t s) { return launch<LL>(ew_kernel<LL>, p, p.n, s); }                \
    extern const LaunchTable launch_table_L##LL = {frob_##LL, horner_##LL, ew_##LL, ftab_##LL, fsh_##LL, fint_##LL, \
                                                   frobg_##LL, solop_##LL, indp_##LL, eval_##LL, trans_##LL, apply_##LL, \
                                                   oshift_##LL, tshift_##LL, radius_##LL, trunc_##LL}; \
    }

extern const LaunchTable launch_table_L4, launch_table_L32, launch_table_L64, launch_table_L128;

}  // namespace cb
