// K1 / K2 / threshold: elementwise passes, 128-bit coalesced, one read + one write per pixel.
//
// Every kernel here is HBM-bound: algorithmic bytes per pixel are listed in DESIGN.md §4.
// Layout: grid.y = frame, grid.x = a multiple of the SM count with a grid-stride loop over
// 16-byte groups; the scalar tail (count % group) is handled by the last threads.
#include "yam_common.cuh"

namespace {

constexpr int kThreads = 256;

inline dim3 grid_for(const yam_ctx* ctx, int64_t groups, int64_t frames) {
    int64_t bx = (groups + kThreads - 1) / kThreads;
    int64_t cap = (int64_t)ctx->num_sms * 8;
    if (frames > 1) cap = (cap + frames - 1) / frames < ctx->num_sms ? ctx->num_sms : (cap + frames - 1) / frames;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    return dim3((unsigned)bx, (unsigned)frames, 1);
}

// ------------------------------------------------------------------------------- bgr2gray
// cv2 BGR2GRAY fixed point (15 bit): (B*3735 + G*19235 + R*9798 + 2^14) >> 15
__device__ __forceinline__ uint32_t gray_fix(uint32_t b, uint32_t g, uint32_t r) {
    return (b * 3735u + g * 19235u + r * 9798u + (1u << 14)) >> 15;
}

__global__ void __launch_bounds__(kThreads) bgr2gray_u16_kernel(const uint16_t* __restrict__ src,
                                                                uint16_t* __restrict__ dst,
                                                                int64_t count) {
    // 8 pixels per thread: 3 x 16 B in, 16 B out
    const int64_t groups = count / 8;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
        const uint4* p = reinterpret_cast<const uint4*>(src + g * 24);
        uint4 a = yam_ld_stream(p), b = yam_ld_stream(p + 1), c = yam_ld_stream(p + 2);
        uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            // element e = 3*i + ch ; word = e/2 ; half = e&1
            uint32_t e0 = 3 * i, e1 = 3 * i + 1, e2 = 3 * i + 2;
            uint32_t bb = (w[e0 >> 1] >> ((e0 & 1) * 16)) & 0xffffu;
            uint32_t gg = (w[e1 >> 1] >> ((e1 & 1) * 16)) & 0xffffu;
            uint32_t rr = (w[e2 >> 1] >> ((e2 & 1) * 16)) & 0xffffu;
            o[i] = gray_fix(bb, gg, rr);
        }
        uint4 out = make_uint4(o[0] | (o[1] << 16), o[2] | (o[3] << 16), o[4] | (o[5] << 16),
                               o[6] | (o[7] << 16));
        yam_st_stream(reinterpret_cast<uint4*>(dst + g * 8), out);
    }
    // tail
    const int64_t done = groups * 8;
    for (int64_t i = done + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        dst[i] = (uint16_t)gray_fix(src[3 * i], src[3 * i + 1], src[3 * i + 2]);
}

__global__ void __launch_bounds__(kThreads) bgr2gray_u8_kernel(const uint8_t* __restrict__ src,
                                                               uint8_t* __restrict__ dst,
                                                               int64_t count) {
    // 16 pixels per thread: 3 x 16 B in, 16 B out
    const int64_t groups = count / 16;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
        const uint4* p = reinterpret_cast<const uint4*>(src + g * 48);
        uint4 a = yam_ld_stream(p), b = yam_ld_stream(p + 1), c = yam_ld_stream(p + 2);
        uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
        uint32_t o[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 16; i++) {
            uint32_t e0 = 3 * i, e1 = 3 * i + 1, e2 = 3 * i + 2;
            uint32_t bb = (w[e0 >> 2] >> ((e0 & 3) * 8)) & 0xffu;
            uint32_t gg = (w[e1 >> 2] >> ((e1 & 3) * 8)) & 0xffu;
            uint32_t rr = (w[e2 >> 2] >> ((e2 & 3) * 8)) & 0xffu;
            o[i >> 2] |= gray_fix(bb, gg, rr) << ((i & 3) * 8);
        }
        yam_st_stream(reinterpret_cast<uint4*>(dst + g * 16), make_uint4(o[0], o[1], o[2], o[3]));
    }
    const int64_t done = groups * 16;
    for (int64_t i = done + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        dst[i] = (uint8_t)gray_fix(src[3 * i], src[3 * i + 1], src[3 * i + 2]);
}

__global__ void __launch_bounds__(kThreads) bgr2gray_f32_kernel(const float* __restrict__ src,
                                                                float* __restrict__ dst,
                                                                int64_t count) {
    // cv2 4.13 float path: b*0.114f + g*0.587f + r*0.299f, left to right, no contraction
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        float b = src[3 * i], g = src[3 * i + 1], r = src[3 * i + 2];
        float s = __fadd_rn(__fmul_rn(b, 0.114f), __fmul_rn(g, 0.587f));
        dst[i] = __fadd_rn(s, __fmul_rn(r, 0.299f));
    }
}

// ------------------------------------------------------------------------------- min / max
__device__ __forceinline__ uint32_t f32_ordered(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_unordered(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

template <typename T>
__device__ __forceinline__ uint32_t key_of(T v);
template <>
__device__ __forceinline__ uint32_t key_of<uint8_t>(uint8_t v) { return v; }
template <>
__device__ __forceinline__ uint32_t key_of<uint16_t>(uint16_t v) { return v; }
template <>
__device__ __forceinline__ uint32_t key_of<float>(float v) { return f32_ordered(v); }

// mm[2*frame] = min key, mm[2*frame+1] = max key
template <typename T>
__global__ void __launch_bounds__(kThreads) minmax_kernel(const T* __restrict__ src,
                                                          int64_t frame_px,
                                                          uint32_t* __restrict__ mm) {
    constexpr int VEC = 16 / sizeof(T);
    const T* base = src + (int64_t)blockIdx.y * frame_px;
    uint32_t lo = 0xffffffffu, hi = 0u;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool aligned = (reinterpret_cast<uintptr_t>(base) & 15) == 0;
    int64_t done = 0;
    if (aligned) {
        const int64_t groups = frame_px / VEC;
        for (int64_t g = tid; g < groups; g += stride) {
            uint4 v = yam_ld_stream(reinterpret_cast<const uint4*>(base) + g);
            const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
            for (int i = 0; i < VEC; i++) {
                uint32_t k = key_of<T>(e[i]);
                lo = min(lo, k);
                hi = max(hi, k);
            }
        }
        done = groups * VEC;
    }
    for (int64_t i = done + tid; i < frame_px; i += stride) {
        uint32_t k = key_of<T>(base[i]);
        lo = min(lo, k);
        hi = max(hi, k);
    }
    lo = yam_warp_min(lo);
    hi = yam_warp_max(hi);
    __shared__ uint32_t s_lo[kThreads / 32], s_hi[kThreads / 32];
    if ((threadIdx.x & 31) == 0) {
        s_lo[threadIdx.x >> 5] = lo;
        s_hi[threadIdx.x >> 5] = hi;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        lo = threadIdx.x < kThreads / 32 ? s_lo[threadIdx.x] : 0xffffffffu;
        hi = threadIdx.x < kThreads / 32 ? s_hi[threadIdx.x] : 0u;
        lo = yam_warp_min(lo);
        hi = yam_warp_max(hi);
        if (threadIdx.x == 0) {
            atomicMin(&mm[2 * blockIdx.y], lo);
            atomicMax(&mm[2 * blockIdx.y + 1], hi);
        }
    }
}

__global__ void minmax_init_kernel(uint32_t* mm, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        mm[2 * i] = 0xffffffffu;
        mm[2 * i + 1] = 0u;
    }
}

__global__ void minmax_finish_kernel(const uint32_t* mm, double* out, int64_t n, int is_f32) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        if (is_f32) {
            out[2 * i] = (double)f32_unordered(mm[2 * i]);
            out[2 * i + 1] = (double)f32_unordered(mm[2 * i + 1]);
        } else {
            out[2 * i] = (double)mm[2 * i];
            out[2 * i + 1] = (double)mm[2 * i + 1];
        }
    }
}

// ------------------------------------------------------------------------------- normalize
// cv2.normalize NORM_MINMAX: scale/shift in fp64 (no contraction), one f32 FMA per pixel.
template <typename T>
__device__ __forceinline__ T scale_store(float x);
template <>
__device__ __forceinline__ uint8_t scale_store<uint8_t>(float x) { return (uint8_t)yam_rint_sat(x, 255); }
template <>
__device__ __forceinline__ uint16_t scale_store<uint16_t>(float x) { return (uint16_t)yam_rint_sat(x, 65535); }
template <>
__device__ __forceinline__ float scale_store<float>(float x) { return x; }

template <typename T>
__global__ void __launch_bounds__(kThreads) normalize_kernel(const T* __restrict__ src,
                                                             T* __restrict__ dst, int64_t frame_px,
                                                             const double* __restrict__ mm,
                                                             double lo, double hi) {
    constexpr int VEC = 16 / sizeof(T);
    const double mn = mm[2 * blockIdx.y], mx = mm[2 * blockIdx.y + 1];
    const double range = __dsub_rn(mx, mn);
    const double sc_d = range > 2.220446049250313e-16 ? __dmul_rn(__dsub_rn(hi, lo), __ddiv_rn(1.0, range)) : 0.0;
    const double sh_d = __dsub_rn(lo, __dmul_rn(mn, sc_d));
    const float sc = __double2float_rn(sc_d), sh = __double2float_rn(sh_d);
    const T* s = src + (int64_t)blockIdx.y * frame_px;
    T* d = dst + (int64_t)blockIdx.y * frame_px;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool aligned = ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 15) == 0;
    int64_t done = 0;
    if (aligned) {
        const int64_t groups = frame_px / VEC;
        for (int64_t g = tid; g < groups; g += stride) {
            uint4 v = yam_ld_stream(reinterpret_cast<const uint4*>(s) + g);
            const T* e = reinterpret_cast<const T*>(&v);
            uint4 o;
            T* oe = reinterpret_cast<T*>(&o);
#pragma unroll
            for (int i = 0; i < VEC; i++) oe[i] = scale_store<T>(__fmaf_rn((float)e[i], sc, sh));
            yam_st_stream(reinterpret_cast<uint4*>(d) + g, o);
        }
        done = groups * VEC;
    }
    for (int64_t i = done + tid; i < frame_px; i += stride)
        d[i] = scale_store<T>(__fmaf_rn((float)s[i], sc, sh));
}

// ------------------------------------------------------------------------------- convertScaleAbs
template <typename T>
__global__ void __launch_bounds__(kThreads) scale_abs_kernel(const T* __restrict__ src,
                                                             uint8_t* __restrict__ dst,
                                                             int64_t count, float a, float b) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // 16 outputs per thread: one 16 B store; inputs 16*sizeof(T) bytes
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
    int64_t done = 0;
    if (aligned) {
        const int64_t groups = count / 16;
        for (int64_t g = tid; g < groups; g += stride) {
            uint32_t o[4] = {0, 0, 0, 0};
            constexpr int NV = sizeof(T);  // number of 16 B input vectors per 16 pixels
#pragma unroll
            for (int v = 0; v < NV; v++) {
                uint4 in = yam_ld_stream(reinterpret_cast<const uint4*>(src) + g * NV + v);
                const T* e = reinterpret_cast<const T*>(&in);
                constexpr int PER = 16 / sizeof(T);
#pragma unroll
                for (int i = 0; i < PER; i++) {
                    int px = v * PER + i;
                    uint32_t r = (uint32_t)yam_rint_sat(fabsf(__fmaf_rn((float)e[i], a, b)), 255);
                    o[px >> 2] |= r << ((px & 3) * 8);
                }
            }
            yam_st_stream(reinterpret_cast<uint4*>(dst) + g, make_uint4(o[0], o[1], o[2], o[3]));
        }
        done = groups * 16;
    }
    for (int64_t i = done + tid; i < count; i += stride)
        dst[i] = (uint8_t)yam_rint_sat(fabsf(__fmaf_rn((float)src[i], a, b)), 255);
}

// ------------------------------------------------------------------------------- LUT (u8)
struct Lut256 {
    uint8_t t[256];
};

__global__ void __launch_bounds__(kThreads) lut_u8_kernel(const uint8_t* __restrict__ src,
                                                          uint8_t* __restrict__ dst, int64_t count,
                                                          Lut256 lut) {
    __shared__ uint8_t s_lut[256];
    if (threadIdx.x < 256) s_lut[threadIdx.x] = lut.t[threadIdx.x];
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
    int64_t done = 0;
    if (aligned) {
        const int64_t groups = count / 16;
        for (int64_t g = tid; g < groups; g += stride) {
            uint4 v = yam_ld_stream(reinterpret_cast<const uint4*>(src) + g);
            const uint8_t* e = reinterpret_cast<const uint8_t*>(&v);
            uint4 o;
            uint8_t* oe = reinterpret_cast<uint8_t*>(&o);
#pragma unroll
            for (int i = 0; i < 16; i++) oe[i] = s_lut[e[i]];
            yam_st_stream(reinterpret_cast<uint4*>(dst) + g, o);
        }
        done = groups * 16;
    }
    for (int64_t i = done + tid; i < count; i += stride) dst[i] = s_lut[src[i]];
}

// ------------------------------------------------------------------------------- threshold
// dst = src > t ? maxval : 0.  t is a constant or read per frame from device memory (Otsu).
template <typename T>
__global__ void __launch_bounds__(kThreads) threshold_kernel(const T* __restrict__ src,
                                                             T* __restrict__ dst, int64_t frame_px,
                                                             const int32_t* __restrict__ t_dev,
                                                             float t_const, T maxval) {
    constexpr int VEC = 16 / sizeof(T);
    const float t = t_dev ? (float)t_dev[blockIdx.y] : t_const;
    const T* s = src + (int64_t)blockIdx.y * frame_px;
    T* d = dst + (int64_t)blockIdx.y * frame_px;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool aligned = ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 15) == 0;
    int64_t done = 0;
    if (aligned) {
        const int64_t groups = frame_px / VEC;
        for (int64_t g = tid; g < groups; g += stride) {
            uint4 v = yam_ld_stream(reinterpret_cast<const uint4*>(s) + g);
            const T* e = reinterpret_cast<const T*>(&v);
            uint4 o;
            T* oe = reinterpret_cast<T*>(&o);
#pragma unroll
            for (int i = 0; i < VEC; i++) oe[i] = (float)e[i] > t ? maxval : (T)0;
            yam_st_stream(reinterpret_cast<uint4*>(d) + g, o);
        }
        done = groups * VEC;
    }
    for (int64_t i = done + tid; i < frame_px; i += stride) d[i] = (float)s[i] > t ? maxval : (T)0;
}

}  // namespace

// internal entry used by the Otsu path (yam_hist.cu): threshold with per-frame device thresholds
int yam_threshold_dev(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t frame_px, int dtype,
                      const int32_t* t_dev, double maxval) {
    dim3 grid = grid_for(ctx, frame_px / (16 / yam_dtype_size(dtype)), n);
    if (dtype == YAM_U8) {
        double m = rint(maxval);
        uint8_t mv = (uint8_t)(m < 0 ? 0 : m > 255 ? 255 : m);
        threshold_kernel<uint8_t><<<grid, kThreads, 0, ctx->stream>>>((const uint8_t*)src, (uint8_t*)dst,
                                                                      frame_px, t_dev, 0.f, mv);
    } else if (dtype == YAM_U16) {
        double m = rint(maxval);
        uint16_t mv = (uint16_t)(m < 0 ? 0 : m > 65535 ? 65535 : m);
        threshold_kernel<uint16_t><<<grid, kThreads, 0, ctx->stream>>>(
            (const uint16_t*)src, (uint16_t*)dst, frame_px, t_dev, 0.f, mv);
    } else {
        yam_set_error("threshold: unsupported dtype %d", dtype);
        return YAM_EINVAL;
    }
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

// ---- interleaved colour <-> planes (neighbourhood filters on BGR input run per channel, like cv2) ----
template <typename T>
__global__ void __launch_bounds__(kThreads) split_channels_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t px, int ch) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < px * ch; i += stride) {
        const int64_t p = i / ch;
        const int c = (int)(i - p * ch);
        dst[(int64_t)c * px + p] = src[i];          // coalesced read, strided write
    }
}
template <typename T>
__global__ void __launch_bounds__(kThreads) merge_channels_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t px, int ch) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < px * ch; i += stride) {
        const int64_t p = i / ch;
        const int c = (int)(i - p * ch);
        dst[i] = src[(int64_t)c * px + p];          // strided read, coalesced write
    }
}

// ---- order-independent 64-bit content checksum (parity evidence for sharded runs) -------------------
// sum over elements of mix64((index_base + i) * GOLDEN + value) mod 2^64: the sum does not depend on
// how the elements are split over launches, strips or ranks, so a sharded run can add its parts up
// (all-reduce) and compare with the dense run.  mix64 = splitmix64's finalizer.
__device__ __forceinline__ unsigned long long yam_mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

template <typename T>
__global__ void __launch_bounds__(kThreads) checksum_kernel(const T* __restrict__ src, int64_t count,
                                                            unsigned long long base,
                                                            unsigned long long* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long acc = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        acc += yam_mix64((base + (unsigned long long)i) * 0x9E3779B97F4A7C15ull + (unsigned long long)(uint32_t)src[i]);
    acc = yam_warp_sum(acc);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

extern "C" {

int yam_bgr2gray(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w, int dtype) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(src && dst && n > 0 && h > 0 && w > 0, "bgr2gray: bad shape");
    const int64_t count = n * h * w;
    if (dtype == YAM_U16) {
        dim3 grid = grid_for(ctx, count / 8, 1);
        bgr2gray_u16_kernel<<<grid, kThreads, 0, ctx->stream>>>((const uint16_t*)src, (uint16_t*)dst, count);
    } else if (dtype == YAM_U8) {
        dim3 grid = grid_for(ctx, count / 16, 1);
        bgr2gray_u8_kernel<<<grid, kThreads, 0, ctx->stream>>>((const uint8_t*)src, (uint8_t*)dst, count);
    } else if (dtype == YAM_F32) {
        dim3 grid = grid_for(ctx, count, 1);
        bgr2gray_f32_kernel<<<grid, kThreads, 0, ctx->stream>>>((const float*)src, (float*)dst, count);
    } else {
        YAM_REQUIRE(false, "bgr2gray: unsupported dtype %d", dtype);
    }
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

int yam_minmax(yam_ctx* ctx, const void* src, int64_t n, int64_t h, int64_t w, int dtype,
               double* out_dev, double* out_host) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(src && n > 0 && h > 0 && w > 0 && n <= 65535, "minmax: bad shape");
    YAM_REQUIRE(dtype == YAM_U8 || dtype == YAM_U16 || dtype == YAM_F32, "minmax: unsupported dtype %d", dtype);
    void* scratch = nullptr;
    const size_t mm_bytes = yam_align_up(sizeof(uint32_t) * 2 * n, 256);
    if (int rc = yam_scratch(ctx, mm_bytes + sizeof(double) * 2 * n, &scratch)) return rc;
    uint32_t* mm = (uint32_t*)scratch;
    double* out = out_dev ? out_dev : (double*)((char*)scratch + mm_bytes);
    const int64_t frame_px = h * w;
    minmax_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(mm, n);
    YAM_LAUNCHED(ctx);
    dim3 grid = grid_for(ctx, frame_px / (16 / yam_dtype_size(dtype)), n);
    if (dtype == YAM_U8)
        minmax_kernel<uint8_t><<<grid, kThreads, 0, ctx->stream>>>((const uint8_t*)src, frame_px, mm);
    else if (dtype == YAM_U16)
        minmax_kernel<uint16_t><<<grid, kThreads, 0, ctx->stream>>>((const uint16_t*)src, frame_px, mm);
    else
        minmax_kernel<float><<<grid, kThreads, 0, ctx->stream>>>((const float*)src, frame_px, mm);
    YAM_LAUNCHED(ctx);
    minmax_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(mm, out, n, dtype == YAM_F32);
    YAM_LAUNCHED(ctx);
    if (out_host) {
        YAM_CUDA(cudaMemcpyAsync(out_host, out, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, ctx->stream));
        YAM_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return YAM_OK;
}

int yam_normalize_minmax(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w,
                         int dtype, double alpha, double beta) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(src && dst && n > 0 && h > 0 && w > 0 && n <= 65535, "normalize: bad shape");
    YAM_REQUIRE(dtype == YAM_U8 || dtype == YAM_U16 || dtype == YAM_F32, "normalize: unsupported dtype %d", dtype);
    // per-frame min/max into the scratch tail (yam_minmax lays out [mm keys | doubles])
    if (int rc = yam_minmax(ctx, src, n, h, w, dtype, nullptr, nullptr)) return rc;
    const size_t mm_bytes = yam_align_up(sizeof(uint32_t) * 2 * n, 256);
    const double* mm = (const double*)((char*)ctx->scratch + mm_bytes);
    const double lo = alpha < beta ? alpha : beta, hi = alpha < beta ? beta : alpha;
    const int64_t frame_px = h * w;
    dim3 grid = grid_for(ctx, frame_px / (16 / yam_dtype_size(dtype)), n);
    if (dtype == YAM_U8)
        normalize_kernel<uint8_t><<<grid, kThreads, 0, ctx->stream>>>((const uint8_t*)src, (uint8_t*)dst, frame_px, mm, lo, hi);
    else if (dtype == YAM_U16)
        normalize_kernel<uint16_t><<<grid, kThreads, 0, ctx->stream>>>((const uint16_t*)src, (uint16_t*)dst, frame_px, mm, lo, hi);
    else
        normalize_kernel<float><<<grid, kThreads, 0, ctx->stream>>>((const float*)src, (float*)dst, frame_px, mm, lo, hi);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

int yam_convert_scale_abs(yam_ctx* ctx, const void* src, void* dst, int64_t count, int dtype,
                          double alpha, double beta) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(src && dst && count > 0, "convert_scale_abs: bad shape");
    dim3 grid = grid_for(ctx, count / 16, 1);
    const float a = (float)alpha, b = (float)beta;
    if (dtype == YAM_U8)
        scale_abs_kernel<uint8_t><<<grid, kThreads, 0, ctx->stream>>>((const uint8_t*)src, (uint8_t*)dst, count, a, b);
    else if (dtype == YAM_U16)
        scale_abs_kernel<uint16_t><<<grid, kThreads, 0, ctx->stream>>>((const uint16_t*)src, (uint8_t*)dst, count, a, b);
    else if (dtype == YAM_F32)
        scale_abs_kernel<float><<<grid, kThreads, 0, ctx->stream>>>((const float*)src, (uint8_t*)dst, count, a, b);
    else
        YAM_REQUIRE(false, "convert_scale_abs: unsupported dtype %d", dtype);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

int yam_lut_u8(yam_ctx* ctx, const void* src, void* dst, int64_t count, const uint8_t* table_host) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(src && dst && table_host && count > 0, "lut_u8: bad arguments");
    Lut256 lut;
    memcpy(lut.t, table_host, 256);
    dim3 grid = grid_for(ctx, count / 16, 1);
    lut_u8_kernel<<<grid, kThreads, 0, ctx->stream>>>((const uint8_t*)src, (uint8_t*)dst, count, lut);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

int yam_threshold(yam_ctx* ctx, const void* src, void* dst, int64_t count, int dtype, double thresh,
                  double maxval) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(src && dst && count > 0, "threshold: bad shape");
    dim3 grid = grid_for(ctx, count / (16 / (yam_dtype_size(dtype) ? yam_dtype_size(dtype) : 1)), 1);
    // cv2: integer images compare against floor(thresh); maxval is rounded and saturated
    const float t = (float)floor(thresh);
    if (dtype == YAM_U8) {
        double m = rint(maxval);
        uint8_t mv = (uint8_t)(m < 0 ? 0 : m > 255 ? 255 : m);
        threshold_kernel<uint8_t><<<grid, kThreads, 0, ctx->stream>>>((const uint8_t*)src, (uint8_t*)dst, count, nullptr, t, mv);
    } else if (dtype == YAM_U16) {
        double m = rint(maxval);
        uint16_t mv = (uint16_t)(m < 0 ? 0 : m > 65535 ? 65535 : m);
        threshold_kernel<uint16_t><<<grid, kThreads, 0, ctx->stream>>>((const uint16_t*)src, (uint16_t*)dst, count, nullptr, t, mv);
    } else if (dtype == YAM_F32) {
        threshold_kernel<float><<<grid, kThreads, 0, ctx->stream>>>((const float*)src, (float*)dst, count, nullptr, (float)thresh, (float)maxval);
    } else {
        YAM_REQUIRE(false, "threshold: unsupported dtype %d", dtype);
    }
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

int yam_checksum64(yam_ctx* ctx, const void* src, int64_t count, int dtype, int64_t index_base, uint64_t* sum_dev) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(src && sum_dev && count > 0 && index_base >= 0, "checksum64: bad arguments");
    int64_t bx = (count + kThreads * 8 - 1) / (kThreads * 8);
    const int64_t cap = (int64_t)ctx->num_sms * 16;
    if (bx > cap) bx = cap;
    unsigned long long* out = (unsigned long long*)sum_dev;
    if (dtype == YAM_U8) checksum_kernel<uint8_t><<<(unsigned)bx, kThreads, 0, ctx->stream>>>((const uint8_t*)src, count, (unsigned long long)index_base, out);
    else if (dtype == YAM_U16) checksum_kernel<uint16_t><<<(unsigned)bx, kThreads, 0, ctx->stream>>>((const uint16_t*)src, count, (unsigned long long)index_base, out);
    else if (dtype == YAM_I32) checksum_kernel<int32_t><<<(unsigned)bx, kThreads, 0, ctx->stream>>>((const int32_t*)src, count, (unsigned long long)index_base, out);
    else YAM_REQUIRE(false, "checksum64: unsupported dtype %d", dtype);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

int yam_split_channels(yam_ctx* ctx, const void* src, void* dst, int64_t px, int channels, int dtype) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(src && dst && src != dst && px > 0 && channels >= 1 && channels <= 4, "split_channels: bad arguments");
    int64_t bx = (px * channels + kThreads * 4 - 1) / (kThreads * 4);
    const int64_t cap = (int64_t)ctx->num_sms * 16;
    if (bx > cap) bx = cap;
    if (dtype == YAM_U8) split_channels_kernel<uint8_t><<<(unsigned)bx, kThreads, 0, ctx->stream>>>((const uint8_t*)src, (uint8_t*)dst, px, channels);
    else if (dtype == YAM_U16) split_channels_kernel<uint16_t><<<(unsigned)bx, kThreads, 0, ctx->stream>>>((const uint16_t*)src, (uint16_t*)dst, px, channels);
    else if (dtype == YAM_F32) split_channels_kernel<float><<<(unsigned)bx, kThreads, 0, ctx->stream>>>((const float*)src, (float*)dst, px, channels);
    else YAM_REQUIRE(false, "split_channels: unsupported dtype %d", dtype);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

int yam_merge_channels(yam_ctx* ctx, const void* src, void* dst, int64_t px, int channels, int dtype) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(src && dst && src != dst && px > 0 && channels >= 1 && channels <= 4, "merge_channels: bad arguments");
    int64_t bx = (px * channels + kThreads * 4 - 1) / (kThreads * 4);
    const int64_t cap = (int64_t)ctx->num_sms * 16;
    if (bx > cap) bx = cap;
    if (dtype == YAM_U8) merge_channels_kernel<uint8_t><<<(unsigned)bx, kThreads, 0, ctx->stream>>>((const uint8_t*)src, (uint8_t*)dst, px, channels);
    else if (dtype == YAM_U16) merge_channels_kernel<uint16_t><<<(unsigned)bx, kThreads, 0, ctx->stream>>>((const uint16_t*)src, (uint16_t*)dst, px, channels);
    else if (dtype == YAM_F32) merge_channels_kernel<float><<<(unsigned)bx, kThreads, 0, ctx->stream>>>((const float*)src, (float*)dst, px, channels);
    else YAM_REQUIRE(false, "merge_channels: unsupported dtype %d", dtype);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

}  // extern "C"


