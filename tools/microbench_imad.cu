// Integer-pipe peak microbenchmark for B200.
// Measures, with CUDA events on the launching stream:
//   wide_indep : independent mad.wide.u32 chains (IMAD.WIDE.U32 issue rate)
//   lo_indep   : independent mad.lo.u32 chains (IMAD issue rate)
//   mul32      : the real in-register 32x32-limb product cb::mul_full<32> (carry chains)
//   shfl       : __shfl_sync rate, for the design note on warp-cooperative limbs
// Prints one JSON object.  Usage: microbench_imad [iters]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../cascade_b200/csrc/cb_mul.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void k_wide(uint64_t* out, uint32_t x, uint32_t y, int iters) {
    uint64_t acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = threadIdx.x + i;
    uint32_t a = x + threadIdx.x, b = y + blockIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a), "r"(b));
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_lo(uint32_t* out, uint32_t x, uint32_t y, int iters) {
    uint32_t acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = threadIdx.x + i;
    uint32_t a = x + threadIdx.x, b = y + blockIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++)
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b));
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_shfl(uint32_t* out, int iters) {
    uint32_t v[4];
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] = threadIdx.x * 7 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int i = 0; i < 4; i++) v[i] = __shfl_sync(0xffffffffu, v[i], (u + i + 1) & 31);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = v[0] ^ v[1] ^ v[2] ^ v[3];
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
        for (int i = 0; i < 32; i++) a[i] ^= r[32 + i];   // feedback so the loop cannot be hoisted
        b[0] += r[0];
    }
#pragma unroll
    for (int i = 0; i < 32; i++) out[i * stride + t] = a[i];
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

int main(int argc, char** argv) {
    int iters = argc > 1 ? atoi(argv[1]) : 2000;
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    void* buf; CK(cudaMalloc(&buf, 256u << 20)); CK(cudaMemset(buf, 0x5a, 256u << 20));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"iters\": %d", p.name, sms, iters);
    {   // 1024 threads per SM, 64 IMAD.WIDE per iter per thread
        int blocks = sms * 4, tpb = 256;
        float ms = time_ms([&] { k_wide<<<blocks, tpb>>>((uint64_t*)buf, 3, 5, iters); }, 5);
        double ops = (double)blocks * tpb * iters * 64.0;
        printf(", \"wide_indep_Tops\": %.4f", ops / ms / 1e9);
    }
    {
        int blocks = sms * 4, tpb = 256;
        float ms = time_ms([&] { k_lo<<<blocks, tpb>>>((uint32_t*)buf, 3, 5, iters); }, 5);
        double ops = (double)blocks * tpb * iters * 64.0;
        printf(", \"lo_indep_Tops\": %.4f", ops / ms / 1e9);
    }
    {
        int blocks = sms * 4, tpb = 256;
        float ms = time_ms([&] { k_shfl<<<blocks, tpb>>>((uint32_t*)buf, iters); }, 5);
        double ops = (double)blocks * tpb * iters * 64.0;
        printf(", \"shfl_Tlane_ops\": %.4f", ops / ms / 1e9);
    }
    int mit = iters / 20 > 0 ? iters / 20 : 1;
    {
        const int TPB = 256; int blocks = sms; int stride = blocks * TPB;
        float ms = time_ms([&] { k_mul32<TPB><<<blocks, TPB>>>((uint32_t*)buf + (32u << 20), (const uint32_t*)buf, stride, mit); }, 5);
        double ops = (double)blocks * TPB * mit * 1024.0;
        printf(", \"mul32_tpb256_Tops\": %.4f", ops / ms / 1e9);
    }
    {
        const int TPB = 320; int blocks = sms; int stride = blocks * TPB;
        float ms = time_ms([&] { k_mul32<TPB><<<blocks, TPB>>>((uint32_t*)buf + (32u << 20), (const uint32_t*)buf, stride, mit); }, 5);
        double ops = (double)blocks * TPB * mit * 1024.0;
        printf(", \"mul32_tpb320_Tops\": %.4f", ops / ms / 1e9);
    }
    {
        const int TPB = 384; int blocks = sms; int stride = blocks * TPB;
        float ms = time_ms([&] { k_mul32<TPB><<<blocks, TPB>>>((uint32_t*)buf + (32u << 20), (const uint32_t*)buf, stride, mit); }, 5);
        double ops = (double)blocks * TPB * mit * 1024.0;
        printf(", \"mul32_tpb384_Tops\": %.4f", ops / ms / 1e9);
    }
    {
        const int TPB = 128; int blocks = sms; int stride = blocks * TPB;
        float ms = time_ms([&] { k_mul32<TPB><<<blocks, TPB>>>((uint32_t*)buf + (32u << 20), (const uint32_t*)buf, stride, mit); }, 5);
        double ops = (double)blocks * TPB * mit * 1024.0;
        printf(", \"mul32_tpb128_Tops\": %.4f", ops / ms / 1e9);
    }
    printf("}\n");
    return 0;
}
