// Microbenchmark: FFMA vs FFMA2 (fma.rn.f32x2) issue throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
    float2 x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    const float2 A = make_float2(a, a), B = make_float2(b, b);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { x[i].x = __fmaf_rn(x[i].x, a, b); x[i].y = __fmaf_rn(x[i].y, a, b); }
            else if (MODE == 1) x[i] = __ffma2_rn(x[i], A, B);
            else { x[i] = __fadd2_rn(x[i], A); }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 3; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(d, iters, 1.0001f, 0.5f);
            if (mode == 1) k<1><<<148 * 8, 256>>>(d, iters, 1.0001f, 0.5f);
            if (mode == 2) k<2><<<148 * 8, 256>>>(d, iters, 1.0001f, 0.5f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double lanes = 148.0 * 8 * 256 * (double)iters * 16;  // scalar f32 ops
            printf("mode %d (%s): %.3f ms  %.2f T lane-ops/s  (%.1f per clk per SM @1.965GHz)\n", mode,
                   mode == 0 ? "FFMA" : mode == 1 ? "FFMA2" : "FADD2", ms, lanes / ms / 1e9, lanes / (ms * 1e-3) / 148 / 1.965e9);
        }
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
