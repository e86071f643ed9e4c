// Shared plumbing for libyamb200: context, error reporting, launch bookkeeping, device helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "yamb200.h"

#define YAM_NUM_SMS_FALLBACK 148

struct yam_ctx {
    int device;
    int num_sms;
    cudaStream_t stream;      // stream work is enqueued on
    cudaStream_t own_stream;  // created with the context
    void* scratch;            // grow-only device scratch
    size_t scratch_bytes;
    void* scratch2;           // second arena: helpers called by operators that already hold `scratch`
    size_t scratch2_bytes;
    void* pinned;             // small pinned host buffer for scalar/histogram read-back
    size_t pinned_bytes;
    int64_t launches;
};

void yam_set_error(const char* fmt, ...);

#define YAM_CUDA(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            yam_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                          __LINE__);                                                        \
            return YAM_ECUDA;                                                               \
        }                                                                                   \
    } while (0)

#define YAM_REQUIRE(cond, ...)        \
    do {                              \
        if (!(cond)) {                \
            yam_set_error(__VA_ARGS__); \
            return YAM_EINVAL;        \
        }                             \
    } while (0)

// Check the launch itself (configuration errors); execution errors surface at the next sync.
#define YAM_LAUNCHED(ctx)                                                               \
    do {                                                                                \
        (ctx)->launches++;                                                              \
        cudaError_t _e = cudaGetLastError();                                            \
        if (_e != cudaSuccess) {                                                        \
            yam_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),   \
                          __FILE__, __LINE__);                                          \
            return YAM_ECUDA;                                                           \
        }                                                                               \
    } while (0)

// Device scratch: returns a pointer valid until the next yam_scratch call on this context.
int yam_scratch(yam_ctx* ctx, size_t bytes, void** out);
// Second arena with the same rule (histogram slabs under operators whose own buffers live in the first).
int yam_scratch2(yam_ctx* ctx, size_t bytes, void** out);
int yam_pinned(yam_ctx* ctx, size_t bytes, void** out);
int yam_enter(yam_ctx* ctx);  // cudaSetDevice + sanity

static inline size_t yam_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int yam_dtype_size(int dtype) {
    switch (dtype) {
        case YAM_U8: return 1;
        case YAM_U16: return 2;
        case YAM_F32: return 4;
        case YAM_I32: return 4;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// device helpers
#ifdef __CUDACC__

__device__ __forceinline__ int yam_border(int i, int n, int mode) {
    // cv2 borderInterpolate for REFLECT_101 / REPLICATE (index may be far outside for tiny n)
    if ((unsigned)i < (unsigned)n) return i;
    if (mode == YAM_BORDER_REPLICATE) return i < 0 ? 0 : n - 1;
    if (n == 1) return 0;
    do {
        if (i < 0) i = -i;
        else i = 2 * (n - 1) - i;
    } while ((unsigned)i >= (unsigned)n);
    return i;
}

template <typename T>
__device__ __forceinline__ T yam_warp_min(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other < v ? other : v;
    }
    return v;
}
template <typename T>
__device__ __forceinline__ T yam_warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other > v ? other : v;
    }
    return v;
}
template <typename T>
__device__ __forceinline__ T yam_warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 128-bit accesses: data touched once should not displace L1 contents (LUTs, taps)
__device__ __forceinline__ uint4 yam_ld_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void yam_st_stream(uint4* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

// cv::saturate_cast<T>(float) == clamp(rint(x)); rint = round-half-even (cvRound)
__device__ __forceinline__ int yam_rint_sat(float x, int hi) {
    // clamp in float first so the conversion cannot overflow
    x = fminf(fmaxf(x, 0.0f), (float)hi);
    return __float2int_rn(x);
}

#endif  // __CUDACC__
