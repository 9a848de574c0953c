// cb_mul.cuh -- limb products held entirely in registers (sm_100a): mul_full (exact L x L -> 2L),
// mul_high (the columns >= K only, with a stated error bound) and add_chain (PTX carry chain).
//
// Replaces the mpn_mul_n calls that Arb's arf_mul / arf_complex_mul make underneath
// acb_mul at reference src/fuchs_solver.c:118 and src/frobenius_solver.c:59,107.
//
// Schoolbook rows.  Row i multiplies all limbs of `a` by b[i]; products that land on
// an even column pair go to the `ev` accumulator, those on an odd column pair go to
// `od` (od[k] is column k+1).  Each (row, parity) is one carry chain of L/2
// mad.lo.cc/madc.hi.cc pairs; ptxas fuses every pair into a single
// IMAD.WIDE.U32.X Rd, Pout, Ra, Rb, Rc, Pin, so a product is L*L IMAD.WIDE plus
// 2L-1 IADD3.X for the final ev + (od << 32).  Accumulators start as compile-time
// zeros, which ptxas turns into RZ operands (no MOVs).
#pragma once
#include <stdint.h>

namespace cb {

#if defined(__CUDA_ARCH__)

template <int L, int K0, int J0, bool CARRY, int N>
__device__ __forceinline__ void mul_chain(uint32_t (&acc)[N], const uint32_t (&a)[L], uint32_t b) {
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;"
                 : "=r"(acc[K0]), "=r"(acc[K0 + 1])
                 : "r"(a[J0]), "r"(b), "r"(acc[K0]), "r"(acc[K0 + 1]));
#pragma unroll
    for (int j = 2; j < L; j += 2)
        asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;"
                     : "=r"(acc[K0 + j]), "=r"(acc[K0 + j + 1])
                     : "r"(a[J0 + j]), "r"(b), "r"(acc[K0 + j]), "r"(acc[K0 + j + 1]));
    if (CARRY) asm volatile("addc.u32 %0, %1, 0;" : "=r"(acc[K0 + L]) : "r"(acc[K0 + L]));
}

template <int L, int I>
struct MulRows {
    template <int N>
    static __device__ __forceinline__ void run(uint32_t (&ev)[N], uint32_t (&od)[N],
                                               const uint32_t (&a)[L], const uint32_t (&b)[L]) {
        if constexpr (I < L) {
            if constexpr ((I & 1) == 0) {
                mul_chain<L, I, 0, true>(ev, a, b[I]);
                mul_chain<L, I, 1, true>(od, a, b[I]);
            } else {
                mul_chain<L, I - 1, 0, true>(od, a, b[I]);
                mul_chain<L, I + 1, 1, (I + 1 + L < 2 * L)>(ev, a, b[I]);
            }
            MulRows<L, I + 1>::run(ev, od, a, b);
        }
    }
};

// r[0..2L) = a[0..L) * b[0..L), little-endian limbs, exact.
template <int L>
__device__ __forceinline__ void mul_full(uint32_t (&r)[2 * L], const uint32_t (&a)[L],
                                         const uint32_t (&b)[L]) {
    static_assert(L % 2 == 0, "even limb count");
    uint32_t ev[2 * L], od[2 * L];
#pragma unroll
    for (int k = 0; k < 2 * L; k++) { ev[k] = 0; od[k] = 0; }
    MulRows<L, 0>::run(ev, od, a, b);
    r[0] = ev[0];
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r[1]) : "r"(ev[1]), "r"(od[0]));
#pragma unroll
    for (int k = 2; k < 2 * L; k++)
        asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r[k]) : "r"(ev[k]), "r"(od[k - 1]));
}


// o[0..N) += r[0..N): one add.cc/addc.cc chain; returns the carry out.  (Written in PTX on purpose:
// a C-level 64-bit carry chain right after mul_full makes ptxas interleave it with the product's
// chains and spill their carry predicates -- 4x the instructions, measured.)
template <int N>
__device__ __forceinline__ uint32_t add_chain(uint32_t (&o)[N], const uint32_t (&r)[N]) {
    asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(o[0]) : "r"(r[0]));
#pragma unroll
    for (int k = 1; k < N; k++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(o[k]) : "r"(r[k]));
    uint32_t c;
    asm volatile("addc.u32 %0, 0, 0;" : "=r"(c));
    return c;
}

// ---------------------------------------------------------------- truncated ("high") product
// r[0 .. 2L-K) = ( sum over limb products with i + j >= K of a[j] b[i] 2^(32 (i+j)) ) >> 32 K.
// Whole limb products whose low word falls in a column below K are dropped (including the part of
// their high word that would land in column K), so with P = a*b:
//     0 <= P - (r << 32 K) < (K + 1) * 2^(32 (K + 1)).
// Same even/odd carry-chain structure as mul_full; K must be even.
template <int L, int K, int KIDX, int J0, int JS, int N>
__device__ __forceinline__ void mulh_chain(uint32_t (&acc)[N], const uint32_t (&a)[L], uint32_t b) {
    // products a[J0 + j] * b for even j in [JS, L), accumulated at acc[KIDX + j], acc[KIDX + j + 1]
    if constexpr (JS < L) {
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;"
                     : "=r"(acc[KIDX + JS]), "=r"(acc[KIDX + JS + 1])
                     : "r"(a[J0 + JS]), "r"(b), "r"(acc[KIDX + JS]), "r"(acc[KIDX + JS + 1]));
#pragma unroll
        for (int j = JS + 2; j < L; j += 2)
            asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;"
                         : "=r"(acc[KIDX + j]), "=r"(acc[KIDX + j + 1])
                         : "r"(a[J0 + j]), "r"(b), "r"(acc[KIDX + j]), "r"(acc[KIDX + j + 1]));
        if constexpr (KIDX + L < N) asm volatile("addc.u32 %0, %1, 0;" : "=r"(acc[KIDX + L]) : "r"(acc[KIDX + L]));
    }
}
template <int L, int K, int I>
struct MulHighRows {
    template <int N>
    static __device__ __forceinline__ void run(uint32_t (&ev)[N], uint32_t (&od)[N], const uint32_t (&a)[L],
                                               const uint32_t (&b)[L]) {
        if constexpr (I < L) {
            if constexpr ((I & 1) == 0) {
                constexpr int JS = (K - I) > 0 ? (K - I) : 0;
                mulh_chain<L, K, I - K, 0, JS>(ev, a, b[I]);  // column I + j      -> ev[I + j - K]
                mulh_chain<L, K, I - K, 1, JS>(od, a, b[I]);  // column I + 1 + j  -> od[I + j - K]
            } else {
                constexpr int JSO = (K + 1 - I) > 0 ? (K + 1 - I) : 0;
                constexpr int JSE = (K - 1 - I) > 0 ? (K - 1 - I) : 0;
                mulh_chain<L, K, I - 1 - K, 0, JSO>(od, a, b[I]);  // column I + j (odd) -> od[I + j - 1 - K]
                mulh_chain<L, K, I + 1 - K, 1, JSE>(ev, a, b[I]);  // column I + 1 + j   -> ev[I + 1 + j - K]
            }
            MulHighRows<L, K, I + 1>::run(ev, od, a, b);
        }
    }
};
template <int L, int K>
__device__ __forceinline__ void mul_high(uint32_t (&r)[2 * L - K], const uint32_t (&a)[L], const uint32_t (&b)[L]) {
    static_assert(L % 2 == 0 && K % 2 == 0 && K >= 0 && K < L, "even limb counts");
    constexpr int N = 2 * L - K + 2;
    uint32_t ev[N], od[N];
#pragma unroll
    for (int k = 0; k < N; k++) { ev[k] = 0; od[k] = 0; }
    MulHighRows<L, K, 0>::run(ev, od, a, b);
    r[0] = ev[0];
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r[1]) : "r"(ev[1]), "r"(od[0]));
#pragma unroll
    for (int k = 2; k < 2 * L - K; k++)
        asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r[k]) : "r"(ev[k]), "r"(od[k - 1]));
}

#else  // host build of the same interface (tests/host_sim only; never shipped in the library)

template <int L>
inline void mul_full(uint32_t (&r)[2 * L], const uint32_t (&a)[L], const uint32_t (&b)[L]) {
    for (int k = 0; k < 2 * L; k++) r[k] = 0;
    for (int i = 0; i < L; i++) {
        uint64_t c = 0;
        for (int j = 0; j < L; j++) {
            uint64_t t = (uint64_t)a[j] * b[i] + r[i + j] + c;
            r[i + j] = (uint32_t)t;
            c = t >> 32;
        }
        r[i + L] = (uint32_t)c;
    }
}


template <int N>
inline uint32_t add_chain(uint32_t (&o)[N], const uint32_t (&r)[N]) {
    uint32_t c = 0;
    for (int k = 0; k < N; k++) {
        uint64_t t = (uint64_t)o[k] + r[k] + c;
        o[k] = (uint32_t)t;
        c = (uint32_t)(t >> 32);
    }
    return c;
}

template <int L, int K>
inline void mul_high(uint32_t (&r)[2 * L - K], const uint32_t (&a)[L], const uint32_t (&b)[L]) {
    uint32_t t[2 * L + 2];
    for (int k = 0; k < 2 * L + 2; k++) t[k] = 0;
    for (int i = 0; i < L; i++)
        for (int j = 0; j < L; j++) {
            if (i + j < K) continue;
            uint64_t p = (uint64_t)a[j] * b[i];
            uint64_t c = (uint32_t)p;
            int k = i + j;
            uint64_t s = (uint64_t)t[k] + c;
            t[k] = (uint32_t)s;
            uint64_t carry = (s >> 32) + (p >> 32);
            for (k = k + 1; carry; k++) {
                s = (uint64_t)t[k] + (uint32_t)carry;
                t[k] = (uint32_t)s;
                carry = (carry >> 32) + (s >> 32);
            }
        }
    for (int k = 0; k < 2 * L - K; k++) r[k] = t[K + k];
}

#endif

}  // namespace cb
