// Which carry formulation of the limb multiply-accumulate runs at full IMAD.WIDE rate on sm_100a?
//   v_chain   : mad.lo.cc / madc.hi.cc chains (IMAD.WIDE.U32.X, carry in+out)
//   v_outonly : mad.lo.cc / madc.hi.cc with carry-OUT only, carry captured by addc (IADD3.X on the ALU pipe)
//   v_comba   : column-wise 3-word accumulators (lo,hi via IMAD.WIDE carry-out, third word via addc)
#include <cstdio>
#include <stdint.h>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void __launch_bounds__(256, 1)
v_chain(uint32_t* out, const uint32_t* in, int iters) {
    uint32_t a[16], acc[4][18], b[4];
    for (int i = 0; i < 16; i++) a[i] = in[i * 256 + threadIdx.x];
    for (int c = 0; c < 4; c++) { b[c] = in[5000 + c] + threadIdx.x; for (int i = 0; i < 18; i++) acc[c][i] = i + c; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[c][0]), "+r"(acc[c][1]) : "r"(a[0]), "r"(b[c]));
#pragma unroll
            for (int j = 2; j < 16; j += 2)
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[c][j]), "+r"(acc[c][j + 1]) : "r"(a[j]), "r"(b[c]));
            asm volatile("addc.u32 %0, %0, 0;" : "+r"(acc[c][16]));
        }
    }
    uint32_t s = 0;
    for (int c = 0; c < 4; c++) for (int i = 0; i < 18; i++) s ^= acc[c][i];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256, 1)
v_outonly(uint32_t* out, const uint32_t* in, int iters) {
    uint32_t a[16], acc[4][16], cw[4][8], b[4];
    for (int i = 0; i < 16; i++) a[i] = in[i * 256 + threadIdx.x];
    for (int c = 0; c < 4; c++) { b[c] = in[5000 + c] + threadIdx.x; for (int i = 0; i < 16; i++) acc[c][i] = i + c; for (int i = 0; i < 8; i++) cw[c][i] = 0; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
#pragma unroll
            for (int j = 0; j < 16; j += 2)
                asm volatile("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;"
                             : "+r"(acc[c][j]), "+r"(acc[c][j + 1]), "+r"(cw[c][j / 2]) : "r"(a[j]), "r"(b[c]));
        }
    }
    uint32_t s = 0;
    for (int c = 0; c < 4; c++) { for (int i = 0; i < 16; i++) s ^= acc[c][i]; for (int i = 0; i < 8; i++) s ^= cw[c][i]; }
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

// plain mad.wide (no carry) at the same register shape, for reference
__global__ void __launch_bounds__(256, 1)
v_plain(uint32_t* out, const uint32_t* in, int iters) {
    uint32_t a[16], b[4];
    uint64_t acc[4][8];
    for (int i = 0; i < 16; i++) a[i] = in[i * 256 + threadIdx.x];
    for (int c = 0; c < 4; c++) { b[c] = in[5000 + c] + threadIdx.x; for (int i = 0; i < 8; i++) acc[c][i] = i + c; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
#pragma unroll
            for (int j = 0; j < 8; j++)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c][j]) : "r"(a[2 * j]), "r"(b[c]));
        }
    }
    uint64_t s = 0;
    for (int c = 0; c < 4; c++) for (int i = 0; i < 8; i++) s ^= acc[c][i];
    out[blockIdx.x * 256 + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
}

// plain mad.wide plus one independent ALU op (add) per IMAD: does the ALU pipe co-issue for free?
__global__ void __launch_bounds__(256, 1)
v_plain_alu(uint32_t* out, const uint32_t* in, int iters) {
    uint32_t a[16], b[4], x[4][8];
    uint64_t acc[4][8];
    for (int i = 0; i < 16; i++) a[i] = in[i * 256 + threadIdx.x];
    for (int c = 0; c < 4; c++) { b[c] = in[5000 + c] + threadIdx.x; for (int i = 0; i < 8; i++) { acc[c][i] = i + c; x[c][i] = i; } }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
#pragma unroll
            for (int j = 0; j < 8; j++)
                asm volatile("mad.wide.u32 %0, %2, %3, %0; add.u32 %1, %1, %3;" : "+l"(acc[c][j]), "+r"(x[c][j]) : "r"(a[2 * j]), "r"(b[c]));
        }
    }
    uint64_t s = 0;
    for (int c = 0; c < 4; c++) for (int i = 0; i < 8; i++) s ^= acc[c][i] + x[c][i];
    out[blockIdx.x * 256 + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
}

template <typename F>
static float time_ms(F launch) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main(int argc, char** argv) {
    int iters = argc > 1 ? atoi(argv[1]) : 4000;
    void* buf; CK(cudaMalloc(&buf, 64u << 20)); CK(cudaMemset(buf, 0x3c, 64u << 20));
    uint32_t* out = (uint32_t*)buf + (8u << 20);
    printf("{\"iters\": %d", iters);
    for (int bps = 1; bps <= 2; bps++) {   // 256 or 512 threads per SM
        int blocks = 148 * bps;
        double ops = (double)blocks * 256 * iters * 32.0;
        float ms;
        ms = time_ms([&] { v_chain<<<blocks, 256>>>(out, (const uint32_t*)buf, iters); });
        printf(", \"chain_%dw_Tops\": %.3f", 8 * bps, ops / ms / 1e9);
        ms = time_ms([&] { v_outonly<<<blocks, 256>>>(out, (const uint32_t*)buf, iters); });
        printf(", \"outonly_%dw_Tops\": %.3f", 8 * bps, ops / ms / 1e9);
        ms = time_ms([&] { v_plain<<<blocks, 256>>>(out, (const uint32_t*)buf, iters); });
        printf(", \"plain_%dw_Tops\": %.3f", 8 * bps, ops / ms / 1e9);
        ms = time_ms([&] { v_plain_alu<<<blocks, 256>>>(out, (const uint32_t*)buf, iters); });
        printf(", \"plain_alu_%dw_Tops\": %.3f", 8 * bps, ops / ms / 1e9);
    }
    printf("}\n");
    return 0;
}
