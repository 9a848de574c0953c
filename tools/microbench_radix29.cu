// Reduced-radix product microbenchmark for B200 (design study for the next recurrence kernel).
// The shipped 1024-bit product (cb::mul_full<32>) is 1024 carry-chained IMAD.WIDE.U32.X, a form that
// issues at half the plain IMAD.WIDE rate and serialises on the carry.  With 29-bit digits a 1024-bit
// mantissa is 36 digits, a digit product is < 2^58, and a whole column (<= 36 products) fits a 64-bit
// accumulator: the product is 1296 PLAIN mad.wide.u32 with no carry chain and column-level ILP.
// Measures products/s of: scan36 (digits in registers), scan36_conv (32-bit limbs in, 32-bit limbs
// out: conversion both ways inside the loop), scan36_high (only the columns a truncated product
// needs), against mul_full<32>.  Prints one JSON object.  Usage: microbench_radix29 [iters]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../cascade_b200/csrc/cb_mul.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int N = 36;             // 29-bit digits
constexpr uint32_t M29 = (1u << 29) - 1;

// columns KLO .. 2N-2 of a * b (digits), carry out in r[2N-1]; two accumulators per column
template <int KLO>
__device__ __forceinline__ void mul_scan(uint32_t (&r)[2 * N], const uint32_t (&a)[N], const uint32_t (&b)[N]) {
    uint64_t carry = 0;
#pragma unroll
    for (int k = KLO; k < 2 * N - 1; k++) {
        uint64_t acc0 = carry, acc1 = 0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            int j = k - i;
            if (j < 0 || j >= N) continue;
            if (i & 1) acc1 += (uint64_t)a[i] * b[j];
            else acc0 += (uint64_t)a[i] * b[j];
        }
        uint64_t acc = acc0 + acc1;
        r[k] = (uint32_t)acc & M29;
        carry = acc >> 29;
    }
    r[2 * N - 1] = (uint32_t)carry;
}

// 32 x 32-bit limbs -> 36 x 29-bit digits (static funnel shifts)
__device__ __forceinline__ void to29(uint32_t (&d)[N], const uint32_t (&w)[32]) {
#pragma unroll
    for (int i = 0; i < N; i++) {
        int bit = 29 * i, word = bit >> 5, off = bit & 31;
        uint32_t lo = w[word], hi = (word + 1 < 32) ? w[word + 1] : 0u;
        d[i] = __funnelshift_r(lo, hi, off) & M29;
    }
}
// digits KD.. of a 72-digit product -> the 32 top 32-bit words (bits 1024..2047 of the 2048-bit product)
__device__ __forceinline__ void from29_high(uint32_t (&w)[32], const uint32_t (&r)[2 * N]) {
#pragma unroll
    for (int j = 0; j < 32; j++) {
        int bit = 1024 + 32 * j, d = bit / 29, o = bit % 29;
        uint64_t v = (uint64_t)r[d] >> o;
        v |= (uint64_t)r[d + 1] << (29 - o);
        if (d + 2 < 2 * N) v |= (uint64_t)r[d + 2] << (58 - o);
        w[j] = (uint32_t)v;
    }
}

template <int TPB, int KLO>
__global__ void __launch_bounds__(TPB, 1) k_scan(uint32_t* out, const uint32_t* in, int stride, int iters) {
    int t = blockIdx.x * TPB + threadIdx.x;
    uint32_t a[N], b[N], r[2 * N];
#pragma unroll
    for (int i = 0; i < N; i++) { a[i] = in[i * stride + t] & M29; b[i] = in[(N + i) * stride + t] & M29; }
    for (int it = 0; it < iters; it++) {
        mul_scan<KLO>(r, a, b);
#pragma unroll
        for (int i = 0; i < N; i++) a[i] ^= r[N + i] & M29;
        b[0] = (b[0] + r[2 * N - 2]) & M29;
    }
#pragma unroll
    for (int i = 0; i < N; i++) out[i * stride + t] = a[i];
}

template <int TPB, int KLO>
__global__ void __launch_bounds__(TPB, 1) k_scan_conv(uint32_t* out, const uint32_t* in, int stride, int iters) {
    int t = blockIdx.x * TPB + threadIdx.x;
    uint32_t aw[32], bw[32];
#pragma unroll
    for (int i = 0; i < 32; i++) { aw[i] = in[i * stride + t]; bw[i] = in[(32 + i) * stride + t]; }
    for (int it = 0; it < iters; it++) {
        uint32_t a[N], b[N], r[2 * N], hw[32];
        to29(a, aw);
        to29(b, bw);
        mul_scan<KLO>(r, a, b);
        from29_high(hw, r);
#pragma unroll
        for (int i = 0; i < 32; i++) aw[i] ^= hw[i];
        bw[0] += hw[0];
    }
#pragma unroll
    for (int i = 0; i < 32; i++) out[i * stride + t] = aw[i];
}

template <int TPB>
__global__ void __launch_bounds__(TPB, 1) k_mul32(uint32_t* out, const uint32_t* in, int stride, int iters) {
    int t = blockIdx.x * TPB + threadIdx.x;
    uint32_t a[32], b[32], r[64];
#pragma unroll
    for (int i = 0; i < 32; i++) { a[i] = in[i * stride + t]; b[i] = in[(32 + i) * stride + t]; }
    for (int it = 0; it < iters; it++) {
        cb::mul_full<32>(r, a, b);
#pragma unroll
        for (int i = 0; i < 32; i++) a[i] ^= r[32 + i];
        b[0] += r[0];
    }
#pragma unroll
    for (int i = 0; i < 32; i++) out[i * stride + t] = a[i];
}

// operand scanning (row-wise) with one 64-bit accumulator per column: a[i] is the same register for a
// whole row (operand-reuse cache), accumulators are independent.  R digits per operand.
template <int TPB, int R>
__global__ void __launch_bounds__(TPB, 1) k_rows(uint32_t* out, const uint32_t* in, int stride, int iters) {
    int t = blockIdx.x * TPB + threadIdx.x;
    uint32_t a[R], b[R];
#pragma unroll
    for (int i = 0; i < R; i++) { a[i] = in[i * stride + t] & M29; b[i] = in[(R + i) * stride + t] & M29; }
    for (int it = 0; it < iters; it++) {
        uint64_t acc[2 * R];
#pragma unroll
        for (int k = 0; k < 2 * R; k++) acc[k] = 0;
#pragma unroll
        for (int i = 0; i < R; i++)
#pragma unroll
            for (int j = 0; j < R; j++) acc[i + j] += (uint64_t)a[i] * b[j];
#pragma unroll
        for (int i = 0; i < R; i++) a[i] ^= (uint32_t)(acc[R + i - 1] >> 7) & M29;
        b[0] = (b[0] + (uint32_t)acc[0]) & M29;
    }
#pragma unroll
    for (int i = 0; i < R; i++) out[i * stride + t] = a[i];
}

template <typename F>
static float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

template <int TPB, int KLO, bool CONV>
static void run(const char* name, void* buf, int sms, int iters, int nprod) {
    int blocks = sms, stride = blocks * TPB;
    float ms = time_ms([&] {
        if (CONV) k_scan_conv<TPB, KLO><<<blocks, TPB>>>((uint32_t*)buf + (32u << 20), (const uint32_t*)buf, stride, iters);
        else k_scan<TPB, KLO><<<blocks, TPB>>>((uint32_t*)buf + (32u << 20), (const uint32_t*)buf, stride, iters);
    }, 5);
    double prods = (double)blocks * TPB * iters;
    printf(", \"%s\": {\"Gprod_s\": %.3f, \"imad_Tops\": %.3f}", name, prods / ms / 1e6, prods * nprod / ms / 1e9);
}

int main(int argc, char** argv) {
    int iters = argc > 1 ? atoi(argv[1]) : 200;
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    void* buf; CK(cudaMalloc(&buf, 256u << 20)); CK(cudaMemset(buf, 0x5a, 256u << 20));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"iters\": %d", p.name, sms, iters);
    {
        const int TPB = 256; int blocks = sms, stride = blocks * TPB;
        float ms = time_ms([&] { k_mul32<TPB><<<blocks, TPB>>>((uint32_t*)buf + (32u << 20), (const uint32_t*)buf, stride, iters); }, 5);
        double prods = (double)blocks * TPB * iters;
        printf(", \"mul_full32_tpb256\": {\"Gprod_s\": %.3f, \"imad_Tops\": %.3f}", prods / ms / 1e6, prods * 1024 / ms / 1e9);
    }
    // number of digit products with columns >= KLO: sum over k of count(k)
    auto count = [](int klo) { int c = 0; for (int k = klo; k < 2 * N - 1; k++) for (int i = 0; i < N; i++) { int j = k - i; if (j >= 0 && j < N) c++; } return c; };
    run<128, 0, false>("scan36_tpb128", buf, sms, iters, count(0));
    run<256, 0, false>("scan36_tpb256", buf, sms, iters, count(0));
    run<384, 0, false>("scan36_tpb384", buf, sms, iters, count(0));
    run<256, 0, true>("scan36_conv_tpb256", buf, sms, iters, count(0));
    run<384, 0, true>("scan36_conv_tpb384", buf, sms, iters, count(0));
    run<256, 31, false>("scan36_high31_tpb256", buf, sms, iters, count(31));
    run<256, 31, true>("scan36_high31_conv_tpb256", buf, sms, iters, count(31));
    run<384, 31, true>("scan36_high31_conv_tpb384", buf, sms, iters, count(31));
    run<512, 31, true>("scan36_high31_conv_tpb512", buf, sms, iters, count(31));
    {
        const int TPB = 256, R = 16; int blocks = sms, stride = blocks * TPB;
        float ms = time_ms([&] { k_rows<TPB, R><<<blocks, TPB>>>((uint32_t*)buf + (32u << 20), (const uint32_t*)buf, stride, iters * 4); }, 5);
        printf(", \"rows16_tpb256_imad_Tops\": %.3f", (double)blocks * TPB * iters * 4 * R * R / ms / 1e9);
    }
    {
        const int TPB = 512, R = 16; int blocks = sms, stride = blocks * TPB;
        float ms = time_ms([&] { k_rows<TPB, R><<<blocks, TPB>>>((uint32_t*)buf + (32u << 20), (const uint32_t*)buf, stride, iters * 4); }, 5);
        printf(", \"rows16_tpb512_imad_Tops\": %.3f", (double)blocks * TPB * iters * 4 * R * R / ms / 1e9);
    }
    {
        const int TPB = 256, R = 24; int blocks = sms, stride = blocks * TPB;
        float ms = time_ms([&] { k_rows<TPB, R><<<blocks, TPB>>>((uint32_t*)buf + (32u << 20), (const uint32_t*)buf, stride, iters * 2); }, 5);
        printf(", \"rows24_tpb256_imad_Tops\": %.3f", (double)blocks * TPB * iters * 2 * R * R / ms / 1e9);
    }
    printf("}\n");
    return 0;
}
