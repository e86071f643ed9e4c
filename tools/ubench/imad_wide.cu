// Microbenchmark: IMAD.WIDE.U32 (u32 x u32 + u64) vs IMAD (32-bit) vs DFMA vs DADD issue throughput on sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(unsigned long long* out, int iters, uint32_t a, uint32_t b) {
    unsigned long long acc[8];
    double d[8];
    uint32_t s[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { acc[i] = threadIdx.x + i; d[i] = threadIdx.x * 0.5 + i; s[i] = threadIdx.x + i; }
    const double da = 1.0000001, db = 0.5;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) acc[i] += (unsigned long long)a * (uint32_t)(s[i] + it);      // IMAD.WIDE.U32 with 64-bit addend
            else if (MODE == 1) s[i] = s[i] * a + b;                                      // IMAD
            else if (MODE == 2) d[i] = __fma_rn(d[i], da, db);                            // DFMA
            else d[i] = __dadd_rn(d[i], db);                                              // DADD
        }
    }
    unsigned long long r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r += acc[i] + (unsigned long long)d[i] + s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
    unsigned long long* d; cudaMalloc(&d, 148 * 8 * 256 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    const char* names[4] = {"IMAD.WIDE.U32 (+u64)", "IMAD (32-bit)", "DFMA", "DADD"};
    for (int mode = 0; mode < 4; mode++)
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(d, iters, 12345u, 7u);
            if (mode == 1) k<1><<<148 * 8, 256>>>(d, iters, 12345u, 7u);
            if (mode == 2) k<2><<<148 * 8, 256>>>(d, iters, 12345u, 7u);
            if (mode == 3) k<3><<<148 * 8, 256>>>(d, iters, 12345u, 7u);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double ops = 148.0 * 8 * 256 * (double)iters * 8;
            printf("%-24s %.3f ms  %.1f ops per clk per SM @1.965GHz\n", names[mode], ms, ops / (ms * 1e-3) / 148 / 1.965e9);
        }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
