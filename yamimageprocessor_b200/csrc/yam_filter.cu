// K3 / K5 / K9 / K4: separable Gaussian (cv2 fixed-point and float32 orders), box filter,
// Gaussian adaptive threshold and small-window median — all as ONE fused pass per op:
// a (TH+2r) x (TW+2r) halo tile is staged in shared memory with 128-bit loads, the horizontal
// pass writes an intermediate tile to shared memory, the vertical pass reads it back register-
// tiled and writes the output with coalesced stores.  HBM traffic = 1 read + 1 write per pixel
// (halo re-reads hit L2).  Arithmetic follows oracle/np_oracle.py bit for bit.
#include <math.h>

#include <type_traits>

#include "yam_common.cuh"
#include "yam_host.h"

// yam_pointwise.cu: dst = src > t_dev[frame] ? maxval : 0 per frame
int yam_threshold_dev(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t frame_px, int dtype,
                      const int32_t* t_dev, double maxval);
#include "yam_median_net.h"

namespace {

constexpr int kThreads = 256;
constexpr int TW = 64;   // output tile width  (2 columns per lane in the vertical pass)
constexpr int TH = 64;   // output tile height (8 warps x 8 rows)
constexpr int RB = 8;    // rows per warp in the vertical pass

struct TapsU {
    uint32_t v[YAM_MAX_TAPS + 1];
};
struct TapsF {
    float v[YAM_MAX_TAPS + 1];
};

enum { EPI_SHIFT = 0, EPI_BOX = 1 };
enum { FEPI_STORE = 0, FEPI_ADAPTIVE = 1, FEPI_ADAPTIVE_BITS = 2 };

template <typename T>
struct FixedTraits;
template <>
struct FixedTraits<uint8_t> {
    typedef uint32_t Acc;
    static constexpr int bits = 8;
};
template <>
struct FixedTraits<uint16_t> {
    typedef unsigned long long Acc;
    static constexpr int bits = 16;
};

__host__ __device__ constexpr int round_up_c(int v, int a) { return (v + a - 1) / a * a; }

template <typename Tout>
__device__ __forceinline__ void store_tile(const Tout* __restrict__ s_out, Tout* __restrict__ dst, int h, int w,
                                           int x0, int y0);

// ----------------------------------------------------------------------------------------------
// tile loader: s_in[rows][SW] <- src[(y0-r .. y0+TH+r) x (x0-ra .. x0+TW+ra)] with border mapping.
// TS = smem element type (T itself, or float when Tin is converted on load).
template <typename T, typename TS>
__device__ __forceinline__ void load_tile(const T* __restrict__ src, TS* __restrict__ s_in, int h, int w,
                                          int x0, int y0, int r, int ra, int SW, int rows, int border,
                                          int pitch = 0) {
    constexpr int VEC = 16 / sizeof(T);
    if (pitch == 0) pitch = SW;  // SW = loaded width; pitch = row stride in shared memory
    const int vec_per_row = SW / VEC;
    const int total = rows * vec_per_row;
    const bool row_aligned = ((w % VEC) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    for (int v = threadIdx.x; v < total; v += kThreads) {
        const int ry = v / vec_per_row;
        const int vx = v - ry * vec_per_row;
        const int gy = yam_border(y0 - r + ry, h, border);
        const int gx = x0 - ra + vx * VEC;
        TS* d = s_in + ry * pitch + vx * VEC;
        const T* row = src + (int64_t)gy * w;
        if (row_aligned && gx >= 0 && gx + VEC <= w) {
            uint4 q = *reinterpret_cast<const uint4*>(row + gx);
            const T* e = reinterpret_cast<const T*>(&q);
            if (sizeof(TS) == sizeof(T)) {
                *reinterpret_cast<uint4*>(d) = q;
            } else {
#pragma unroll
                for (int i = 0; i < VEC; i++) d[i] = (TS)e[i];
            }
        } else {
#pragma unroll 4
            for (int i = 0; i < VEC; i++) d[i] = (TS)row[yam_border(gx + i, w, border)];
        }
    }
}

// ----------------------------------------------------------------------------------------------
// fixed-point separable filter, compile-time K (register-tiled)
template <typename T, int KS, int EPI>
__global__ void __launch_bounds__(kThreads) sep_fixed_tiled(const T* __restrict__ src, T* __restrict__ dst,
                                                            int h, int w, TapsU taps, int border,
                                                            uint32_t box_div) {
    typedef typename FixedTraits<T>::Acc Acc;
    constexpr int bits = FixedTraits<T>::bits;
    constexpr int VEC = 16 / sizeof(T);
    constexpr int R = KS / 2;
    constexpr int RA = round_up_c(R, VEC);
    constexpr int SW0 = TW + 2 * RA;
    // u16: pitches padded so that a quarter-warp of 4 column groups x 2 rows is bank-conflict free
    // (input pitch 96 px = 12 x 16 B == 4 mod 8, intermediate pitch 68 words == 1 mod 8 groups)
    constexpr bool PAIR = (sizeof(T) == 2) && (SW0 == 80);
    constexpr int SW = PAIR ? 96 : SW0;
    constexpr int TWP = PAIR ? TW + 4 : TW;
    constexpr int ROWS = TH + 2 * R;
    constexpr int NV = 1 + 2 * RA / VEC;  // aligned vectors covering VEC outputs + halo
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s_in = reinterpret_cast<T*>(smem_raw);
    uint32_t* s_t = reinterpret_cast<uint32_t*>(smem_raw + round_up_c(ROWS * SW * (int)sizeof(T), 16));

    const int64_t frame = blockIdx.z;
    src += frame * (int64_t)h * w;
    dst += frame * (int64_t)h * w;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;

    load_tile<T, T>(src, s_in, h, w, x0, y0, R, RA, SW0, ROWS, border, SW);
    __syncthreads();

    // horizontal pass: VEC outputs per item, exact integer, symmetric pairs share one multiply
    constexpr int GROUPS = TW / VEC;
    constexpr int ITEMS = PAIR ? ((ROWS + 1) / 2) * 16 : ROWS * GROUPS;
    for (int item = threadIdx.x; item < ITEMS; item += kThreads) {
        int ry, cg;
        if (PAIR) {
            const int l = item & 7, q = item >> 3;
            cg = (q & 1) * 4 + (l & 3);
            ry = 2 * (q >> 1) + (l >> 2);
            if (ry >= ROWS) continue;
        } else {
            ry = item / GROUPS;
            cg = item - ry * GROUPS;
        }
        const T* p = s_in + ry * SW + cg * VEC;
        uint32_t e[NV * VEC];
#pragma unroll
        for (int v = 0; v < NV; v++) {
            uint4 q = *reinterpret_cast<const uint4*>(p + v * VEC);
            const T* qe = reinterpret_cast<const T*>(&q);
#pragma unroll
            for (int i = 0; i < VEC; i++) e[v * VEC + i] = qe[i];
        }
        uint32_t out[VEC];
#pragma unroll
        for (int j = 0; j < VEC; j++) {
            const int c = RA + j;  // centre element
            uint32_t acc = taps.v[R] * e[c];
#pragma unroll
            for (int i = 1; i <= R; i++) acc += taps.v[R + i] * (e[c - i] + e[c + i]);
            out[j] = acc;
        }
        uint32_t* t = s_t + ry * TWP + cg * VEC;
#pragma unroll
        for (int j = 0; j < VEC; j += 4)
            *reinterpret_cast<uint4*>(t + j) = make_uint4(out[j], out[j + 1], out[j + 2], out[j + 3]);
    }
    __syncthreads();

    // vertical pass: lane -> column pair, warp -> RB output rows; sliding window in registers
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cx = 2 * lane;
    const int r0 = warp * RB;
    T* s_out = reinterpret_cast<T*>(s_t + ROWS * TWP);
    if constexpr (sizeof(T) == 2 && EPI == EPI_SHIFT) {
        // 16-bit Gaussian: sum of 16-bit taps x 32-bit row sums needs 48 bits.  IMAD.WIDE runs at ~36 / clk / SM
        // (tools/ubench/imad_wide.cu); fp64 FMA at 63 / clk / SM holds every partial sum exactly (< 2^48 < 2^53),
        // leaves the integer pipe to the horizontal passes of the other resident blocks, and needs no
        // conversion instruction: (0x43300000 : t) is the double 2^52 + t, and the rounded result
        // (acc + 2^31) >> 32 is read from the mantissa of acc + (2^52 + 2^31).
        double kd[KS];
#pragma unroll
        for (int k = 0; k < KS; k++) kd[k] = (double)taps.v[k];
        double a0[RB], a1[RB];
#pragma unroll
        for (int j = 0; j < RB; j++) a0[j] = a1[j] = 0.0;
#pragma unroll
        for (int i = 0; i < RB + 2 * R; i++) {
            const uint2 tv = *reinterpret_cast<const uint2*>(s_t + (r0 + i) * TWP + cx);
            const double t0 = __dsub_rn(__hiloint2double(0x43300000, (int)tv.x), 4503599627370496.0);
            const double t1 = __dsub_rn(__hiloint2double(0x43300000, (int)tv.y), 4503599627370496.0);
#pragma unroll
            for (int j = 0; j < RB; j++) {
                const int k = i - j;  // tap index for output row j
                if (k >= 0 && k < KS) {
                    a0[j] = __fma_rn(kd[k], t0, a0[j]);
                    a1[j] = __fma_rn(kd[k], t1, a1[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < RB; j++) {
            const uint32_t o0 = (uint32_t)__double2hiint(__dadd_rn(a0[j], 4503599627370496.0 + 2147483648.0)) & 0xffffu;
            const uint32_t o1 = (uint32_t)__double2hiint(__dadd_rn(a1[j], 4503599627370496.0 + 2147483648.0)) & 0xffffu;
            *reinterpret_cast<uint32_t*>(s_out + (r0 + j) * TW + cx) = o0 | (o1 << 16);
        }
    } else {
        // box sums of 16-bit pixels stay below 2^32 (k <= 31), so only the 16-bit Gaussian needs wide sums
        typedef typename std::conditional<EPI == EPI_BOX, uint32_t, Acc>::type VAcc;
        VAcc acc0[RB], acc1[RB];
#pragma unroll
        for (int j = 0; j < RB; j++) acc0[j] = acc1[j] = 0;
#pragma unroll
        for (int i = 0; i < RB + 2 * R; i++) {
            const uint2 tv = *reinterpret_cast<const uint2*>(s_t + (r0 + i) * TWP + cx);
#pragma unroll
            for (int j = 0; j < RB; j++) {
                const int k = i - j;  // tap index for output row j
                if (k >= 0 && k < KS) {
                    acc0[j] += (VAcc)taps.v[k] * tv.x;
                    acc1[j] += (VAcc)taps.v[k] * tv.y;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < RB; j++) {
            uint32_t o0, o1;
            if (EPI == EPI_SHIFT) {
                o0 = (uint32_t)((acc0[j] + ((VAcc)1 << (2 * bits - 1))) >> (2 * bits));
                o1 = (uint32_t)((acc1[j] + ((VAcc)1 << (2 * bits - 1))) >> (2 * bits));
            } else {
                // rint(sum / k^2), k^2 odd: (2 sum + k^2) / (2 k^2)
                o0 = (uint32_t)((2 * (uint32_t)acc0[j] + box_div) / (2 * box_div));
                o1 = (uint32_t)((2 * (uint32_t)acc1[j] + box_div) / (2 * box_div));
            }
            T* d = s_out + (r0 + j) * TW + cx;
            if (sizeof(T) == 2)
                *reinterpret_cast<uint32_t*>(d) = o0 | (o1 << 16);
            else
                *reinterpret_cast<uint16_t*>(d) = (uint16_t)(o0 | (o1 << 8));
        }
    }
    __syncthreads();
    store_tile<T>(s_out, dst, h, w, x0, y0);
}

// fixed-point separable filter, runtime K (any odd K <= YAM_MAX_TAPS that fits shared memory)
template <typename T, int EPI>
__global__ void __launch_bounds__(kThreads) sep_fixed_generic(const T* __restrict__ src, T* __restrict__ dst,
                                                              int h, int w, TapsU taps, int ks, int border,
                                                              uint32_t box_div) {
    typedef typename FixedTraits<T>::Acc Acc;
    constexpr int bits = FixedTraits<T>::bits;
    constexpr int VEC = 16 / sizeof(T);
    const int R = ks / 2;
    const int RA = round_up_c(R, VEC);
    const int SW = TW + 2 * RA;
    const int ROWS = TH + 2 * R;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s_in = reinterpret_cast<T*>(smem_raw);
    uint32_t* s_t = reinterpret_cast<uint32_t*>(smem_raw + round_up_c(ROWS * SW * (int)sizeof(T), 16));
    const int64_t frame = blockIdx.z;
    src += frame * (int64_t)h * w;
    dst += frame * (int64_t)h * w;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    load_tile<T, T>(src, s_in, h, w, x0, y0, R, RA, SW, ROWS, border);
    __syncthreads();
    for (int item = threadIdx.x; item < ROWS * TW; item += kThreads) {
        const int ry = item / TW, cx = item - ry * TW;
        const T* p = s_in + ry * SW + cx + RA - R;
        uint32_t acc = 0;
        for (int i = 0; i < ks; i++) acc += taps.v[i] * (uint32_t)p[i];
        s_t[item] = acc;
    }
    __syncthreads();
    for (int item = threadIdx.x; item < TH * TW; item += kThreads) {
        const int ry = item / TW, cx = item - ry * TW;
        const int gx = x0 + cx, gy = y0 + ry;
        if (gx >= w || gy >= h) continue;
        Acc acc = 0;
        for (int i = 0; i < ks; i++) acc += (Acc)taps.v[i] * s_t[(ry + i) * TW + cx];
        uint32_t o;
        if (EPI == EPI_SHIFT)
            o = (uint32_t)((acc + ((Acc)1 << (2 * bits - 1))) >> (2 * bits));
        else
            o = (uint32_t)((2 * (uint32_t)acc + box_div) / (2 * box_div));
        dst[(int64_t)gy * w + gx] = (T)o;
    }
}

// ----------------------------------------------------------------------------------------------
// float32 separable Gaussian in cv2 4.13's summation order (see oracle gaussian_f32):
//   rows  K>=7: s = k0*x0; s = fma(x_i, k_i, s)            (sequential)
//         K==5: s = (x-1 + x+1)*k1; s = fma(x0,kc,s); s = fma(x-2 + x+2, k2, s)
//         K==3: s = fma(x-1 + x+1, k1, x0*kc)
//   cols       t = kc*y0; t = fma(y+j + y-j, k_{c+j}, t)    (centre, then symmetric pairs)
template <int KSC>
__device__ __forceinline__ float row_dot(const float* __restrict__ x, const TapsF& taps, int ks) {
    // x points at the left-most tap
    const int K = KSC > 0 ? KSC : ks;
    if (K == 3) return __fmaf_rn(__fadd_rn(x[0], x[2]), taps.v[2], __fmul_rn(x[1], taps.v[1]));
    if (K == 5) {
        float s = __fmul_rn(__fadd_rn(x[1], x[3]), taps.v[3]);
        s = __fmaf_rn(x[2], taps.v[2], s);
        return __fmaf_rn(__fadd_rn(x[0], x[4]), taps.v[4], s);
    }
    float s = __fmul_rn(taps.v[0], x[0]);
    if (KSC > 0) {
#pragma unroll
        for (int i = 1; i < (KSC > 0 ? KSC : 1); i++) s = __fmaf_rn(x[i], taps.v[i], s);
    } else {
        for (int i = 1; i < K; i++) s = __fmaf_rn(x[i], taps.v[i], s);
    }
    return s;
}

template <typename Tin, typename Tout, int KSC, int FEPI>
__global__ void __launch_bounds__(kThreads) sep_f32_kernel(const Tin* __restrict__ src, Tout* __restrict__ dst,
                                                           int h, int w, TapsF taps, int ks, int border,
                                                           int sat_hi, int idelta) {
    constexpr int VEC = 16 / sizeof(Tin);
    const int K = KSC > 0 ? KSC : ks;
    const int R = K / 2;
    const int RA = round_up_c(R, VEC);
    const int SW = TW + 2 * RA;
    const int ROWS = TH + 2 * R;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_in = reinterpret_cast<float*>(smem_raw);
    float* s_t = s_in + ROWS * SW;
    const int64_t frame = blockIdx.z;
    src += frame * (int64_t)h * w;
    dst += frame * (int64_t)h * w;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    load_tile<Tin, float>(src, s_in, h, w, x0, y0, R, RA, SW, ROWS, border);
    __syncthreads();

    // horizontal pass, 4 outputs per item
    constexpr int PX = 4;
    constexpr int GROUPS = TW / PX;
    for (int item = threadIdx.x; item < ROWS * GROUPS; item += kThreads) {
        const int ry = item / GROUPS, cg = item - ry * GROUPS;
        const float* p = s_in + ry * SW + cg * PX + RA - R;
        float o[PX];
        if (KSC > 0) {
            float e[PX + (KSC > 0 ? KSC : 1) - 1];
#pragma unroll
            for (int i = 0; i < PX + (KSC > 0 ? KSC : 1) - 1; i++) e[i] = p[i];
#pragma unroll
            for (int j = 0; j < PX; j++) o[j] = row_dot<KSC>(e + j, taps, ks);
        } else {
#pragma unroll
            for (int j = 0; j < PX; j++) o[j] = row_dot<0>(p + j, taps, ks);
        }
        *reinterpret_cast<float4*>(s_t + ry * TW + cg * PX) = make_float4(o[0], o[1], o[2], o[3]);
    }
    __syncthreads();

    // vertical pass: lane -> column pair, warp -> RB rows
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cx = 2 * lane;
    const int r0 = warp * RB;
    const int gx = x0 + cx;
    float out0[RB], out1[RB];
    if (KSC > 0) {
        constexpr int KK = KSC > 0 ? KSC : 1;
        constexpr int RR = KK / 2;
        float c0[RB + KK - 1], c1[RB + KK - 1];
#pragma unroll
        for (int i = 0; i < RB + KK - 1; i++) {
            const float2 v = *reinterpret_cast<const float2*>(s_t + (r0 + i) * TW + cx);
            c0[i] = v.x;
            c1[i] = v.y;
        }
#pragma unroll
        for (int j = 0; j < RB; j++) {
            float a = __fmul_rn(taps.v[RR], c0[j + RR]);
            float b = __fmul_rn(taps.v[RR], c1[j + RR]);
#pragma unroll
            for (int q = 1; q <= RR; q++) {
                a = __fmaf_rn(__fadd_rn(c0[j + RR + q], c0[j + RR - q]), taps.v[RR + q], a);
                b = __fmaf_rn(__fadd_rn(c1[j + RR + q], c1[j + RR - q]), taps.v[RR + q], b);
            }
            out0[j] = a;
            out1[j] = b;
        }
    } else {
#pragma unroll 1
        for (int j = 0; j < RB; j++) {
            const float* col = s_t + (r0 + j + R) * TW + cx;
            float a = __fmul_rn(taps.v[R], col[0]);
            float b = __fmul_rn(taps.v[R], col[1]);
            for (int q = 1; q <= R; q++) {
                a = __fmaf_rn(__fadd_rn(col[q * TW], col[-q * TW]), taps.v[R + q], a);
                b = __fmaf_rn(__fadd_rn(col[q * TW + 1], col[-q * TW + 1]), taps.v[R + q], b);
            }
            out0[j] = a;
            out1[j] = b;
        }
    }
#pragma unroll
    for (int j = 0; j < RB; j++) {
        const int gy = y0 + r0 + j;
        if (gy >= h || gx >= w) continue;
        Tout* d = dst + (int64_t)gy * w + gx;
        if (FEPI == FEPI_ADAPTIVE) {
            // mean = saturate(rint(blur)); dst = (src - mean > -idelta) ? 255 : 0
            const float* sp = s_in + (r0 + j + R) * SW + RA + cx;
            const int m0 = yam_rint_sat(out0[j], sat_hi), m1 = yam_rint_sat(out1[j], sat_hi);
            const int s0 = (int)sp[0], s1 = (int)sp[1];
            const uint32_t o0 = (s0 - m0 > -idelta) ? 255u : 0u, o1 = (s1 - m1 > -idelta) ? 255u : 0u;
            if (gx + 1 < w && ((reinterpret_cast<uintptr_t>(d) & 1) == 0))
                *reinterpret_cast<uint16_t*>(d) = (uint16_t)(o0 | (o1 << 8));
            else {
                d[0] = (Tout)o0;
                if (gx + 1 < w) d[1] = (Tout)o1;
            }
        } else {
            d[0] = (Tout)out0[j];
            if (gx + 1 < w) d[1] = (Tout)out1[j];
        }
    }
}

// ----------------------------------------------------------------------------------------------
// float32 separable Gaussian, compile-time K, throughput version (same arithmetic order as above).
//   load   integer pixels become floats with PRMT + FADD (0x4B000000 | v) - 2^23 instead of I2F
//   H pass 8 outputs per item from aligned 128-bit shared loads
//   V pass sliding window in registers; adaptive epilogue rounds with the 1.5*2^23 magic add
//   store  the output tile is staged in shared memory and written with 16-byte coalesced stores
__device__ __forceinline__ float u16lo_to_f32(uint32_t wd) {
    return __fadd_rn(__uint_as_float(__byte_perm(wd, 0x4B000000u, 0x7410)), -8388608.0f);
}
__device__ __forceinline__ float u16hi_to_f32(uint32_t wd) {
    return __fadd_rn(__uint_as_float(__byte_perm(wd, 0x4B000000u, 0x7432)), -8388608.0f);
}
template <int K>
__device__ __forceinline__ float u8_to_f32(uint32_t wd) {
    return __fadd_rn(__uint_as_float(__byte_perm(wd, 0x4B000000u, 0x7440 | K)), -8388608.0f);
}

template <typename T>
__device__ __forceinline__ void vec_to_f32(const uint4& q, float* d);
template <>
__device__ __forceinline__ void vec_to_f32<uint16_t>(const uint4& q, float* d) {
    const uint32_t wd[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        d[2 * i] = u16lo_to_f32(wd[i]);
        d[2 * i + 1] = u16hi_to_f32(wd[i]);
    }
}
template <>
__device__ __forceinline__ void vec_to_f32<uint8_t>(const uint4& q, float* d) {
    const uint32_t wd[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        d[4 * i] = u8_to_f32<0>(wd[i]);
        d[4 * i + 1] = u8_to_f32<1>(wd[i]);
        d[4 * i + 2] = u8_to_f32<2>(wd[i]);
        d[4 * i + 3] = u8_to_f32<3>(wd[i]);
    }
}
template <>
__device__ __forceinline__ void vec_to_f32<float>(const uint4& q, float* d) {
    d[0] = __uint_as_float(q.x); d[1] = __uint_as_float(q.y); d[2] = __uint_as_float(q.z); d[3] = __uint_as_float(q.w);
}

// cooperative store of a TH x TW staged tile (row pitch TW elements) with 16-byte vectors
template <typename Tout>
__device__ __forceinline__ void store_tile(const Tout* __restrict__ s_out, Tout* __restrict__ dst, int h, int w,
                                           int x0, int y0) {
    constexpr int VEC = 16 / sizeof(Tout);
    constexpr int VPR = TW / VEC;
    const bool row_aligned = ((w % VEC) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    for (int v = threadIdx.x; v < TH * VPR; v += kThreads) {
        const int ty = v / VPR, vx = v - ty * VPR;
        const int gy = y0 + ty, gx = x0 + vx * VEC;
        if (gy >= h || gx >= w) continue;
        const Tout* sp = s_out + ty * TW + vx * VEC;
        Tout* d = dst + (int64_t)gy * w + gx;
        if (row_aligned && gx + VEC <= w) {
            *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(sp);
        } else {
            for (int i = 0; i < VEC && gx + i < w; i++) d[i] = sp[i];
        }
    }
}

template <typename Tin, typename Tout, int KS, int FEPI>
__global__ void __launch_bounds__(kThreads) sep_f32_tiled(const Tin* __restrict__ src, Tout* __restrict__ dst,
                                                          int h, int w, TapsF taps, int border, int idelta,
                                                          uint32_t* __restrict__ bits_out, int wpr) {
    constexpr int VEC = 16 / sizeof(Tin);
    constexpr int R = KS / 2;
    constexpr int RA = round_up_c(R, VEC > 4 ? VEC : 4);
    constexpr int SW = TW + 2 * RA;
    constexpr int ROWS = TH + 2 * R;
    // Row pitches are padded so that a quarter-warp made of 4 column groups x 2 consecutive rows hits
    // 8 distinct 16-byte bank groups: input pitch == 4 (mod 8) float4-groups... i.e. SWP/4 odd,
    // intermediate pitch TWP/4 odd.  ncu (profiles/r01_ncu_c2_v1.txt): 45 % of the shared wavefronts
    // of the unpadded layout were bank conflicts and the LSU data pipe was the limiter (78 %).
    constexpr int SWP = ((SW / 4) % 2 == 0) ? SW + 4 : SW;
    constexpr int TWP = TW + 4;
    constexpr int OFF = (RA - R) & 3;                 // misalignment of the first tap inside a float4
    constexpr int NV = (OFF + 8 + 2 * R + 3) / 4;     // float4 loads per 8 outputs
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_in = reinterpret_cast<float*>(smem_raw);
    float* s_t = s_in + ROWS * SWP + 8;               // +8 floats slack: the last item may read past its row
    // Byte outputs (adaptive threshold): the 4 KiB staging tile aliases the start of s_t, which every
    // warp has copied into registers before the first output is written (barrier below); 45 KB per
    // block instead of 49 KB lets five blocks share an SM.  Float outputs keep their own tile.
    constexpr bool ALIAS_OUT = sizeof(Tout) == 1;
    static_assert(!ALIAS_OUT || TH * TW <= ROWS * TWP * 4, "aliased output tile must fit inside s_t");
    Tout* s_out = ALIAS_OUT ? reinterpret_cast<Tout*>(s_t) : reinterpret_cast<Tout*>(s_t + ROWS * TWP);
    src += (int64_t)blockIdx.z * h * w;
    if (FEPI != FEPI_ADAPTIVE_BITS) dst += (int64_t)blockIdx.z * h * w;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;

    // ---- load + convert
    {
        constexpr int VPR = SW / VEC;
        const bool row_aligned = ((w % VEC) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
        for (int v = threadIdx.x; v < ROWS * VPR; v += kThreads) {
            const int ry = v / VPR, vx = v - ry * VPR;
            const int gy = yam_border(y0 - R + ry, h, border);
            const int gx = x0 - RA + vx * VEC;
            const Tin* row = src + (int64_t)gy * w;
            float f[VEC];
            if (row_aligned && gx >= 0 && gx + VEC <= w) {
                const uint4 q = *reinterpret_cast<const uint4*>(row + gx);
                vec_to_f32<Tin>(q, f);
            } else {
#pragma unroll 4
                for (int i = 0; i < VEC; i++) f[i] = (float)row[yam_border(gx + i, w, border)];
            }
            float* d = s_in + ry * SWP + vx * VEC;
#pragma unroll
            for (int i = 0; i < VEC; i += 4) *reinterpret_cast<float4*>(d + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
        }
    }
    __syncthreads();

    // ---- horizontal pass: 8 outputs per item
    {
        static_assert(TW / 8 == 8, "quarter-warp mapping below assumes 8 column groups per row");
        constexpr int ROW_PAIRS = (ROWS + 1) / 2;
        for (int item = threadIdx.x; item < ROW_PAIRS * 16; item += kThreads) {
            // 8 consecutive items = 4 column groups x 2 rows (conflict-free 128-bit accesses)
            const int l = item & 7, q = item >> 3;
            const int cg = (q & 1) * 4 + (l & 3);
            const int ry = 2 * (q >> 1) + (l >> 2);
            if (ry >= ROWS) continue;
            const float* p = s_in + ry * SWP + cg * 8 + (RA - R) - OFF;
            float e[NV * 4];
#pragma unroll
            for (int v = 0; v < NV; v++) {
                const float4 q = *reinterpret_cast<const float4*>(p + 4 * v);
                e[4 * v] = q.x; e[4 * v + 1] = q.y; e[4 * v + 2] = q.z; e[4 * v + 3] = q.w;
            }
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; j++) o[j] = row_dot<KS>(e + OFF + j, taps, KS);
            float* t = s_t + ry * TWP + cg * 8;
            *reinterpret_cast<float4*>(t) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(t + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
    }
    __syncthreads();

    // ---- vertical pass: lane -> column pair, warp -> RB rows.  The two columns of a lane form one
    // packed fp32 operand (FFMA2 / FADD2 / FMUL2: each half is an IEEE fp32 operation, so results
    // are bit-identical to scalar code at half the issue slots).
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int cx = 2 * lane;
        const int r0 = warp * RB;
        float2 c[RB + KS - 1];
#pragma unroll
        for (int i = 0; i < RB + KS - 1; i++) c[i] = *reinterpret_cast<const float2*>(s_t + (r0 + i) * TWP + cx);
        if (ALIAS_OUT) __syncthreads();  // all of s_t is in registers: its storage may now take the outputs
        const float2 kc = make_float2(taps.v[R], taps.v[R]);
#pragma unroll
        for (int j = 0; j < RB; j++) {
            float2 a = __fmul2_rn(kc, c[j + R]);
#pragma unroll
            for (int q = 1; q <= R; q++)
                a = __ffma2_rn(__fadd2_rn(c[j + R + q], c[j + R - q]), make_float2(taps.v[R + q], taps.v[R + q]), a);
            if (FEPI == FEPI_ADAPTIVE || FEPI == FEPI_ADAPTIVE_BITS) {
                // mean = rint(blur) (round-half-even via the 1.5*2^23 add; blur is a convex
                // combination of pixel values so saturation can never trigger);
                // dst = (src - mean > -idelta) ? 255 : 0, all values exact integers in float
                const float2 sp = *reinterpret_cast<const float2*>(s_in + (r0 + j + R) * SWP + RA + cx);
                const float2 big = make_float2(12582912.0f, 12582912.0f), nbig = make_float2(-12582912.0f, -12582912.0f);
                const float2 m = __fadd2_rn(__fadd2_rn(a, big), nbig);
                // sp - m, exact: one rounding of an exactly representable difference of integers
                const float2 d = __ffma2_rn(m, make_float2(-1.0f, -1.0f), sp);
                const float nd = -(float)idelta;
                const uint32_t o0 = (d.x > nd) ? 255u : 0u;
                const uint32_t o1 = (d.y > nd) ? 255u : 0u;
                *reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(s_out) + (r0 + j) * TW + cx) = (uint16_t)(o0 | (o1 << 8));
            } else {
                *reinterpret_cast<float2*>(reinterpret_cast<float*>(s_out) + (r0 + j) * TW + cx) = a;
            }
        }
    }
    __syncthreads();
    if (FEPI == FEPI_ADAPTIVE_BITS) {
        // pack the 64x64 byte tile to 2 words per row: 1 bit per pixel, bits beyond the image are 0
        static_assert(TW == 64, "bit packing assumes two words per tile row");
        uint32_t* bits = bits_out + (int64_t)blockIdx.z * h * wpr;
        const uint8_t* s8 = reinterpret_cast<const uint8_t*>(s_out);
        for (int t = threadIdx.x; t < TH * 2; t += kThreads) {
            const int row = t >> 1, half = t & 1;
            const int gy = y0 + row, gw = (x0 >> 5) + half;
            if (gy >= h || gw >= wpr) continue;
            const uint4 q0 = *reinterpret_cast<const uint4*>(s8 + row * TW + half * 32);
            const uint4 q1 = *reinterpret_cast<const uint4*>(s8 + row * TW + half * 32 + 16);
            const uint32_t wd[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
            uint32_t word = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) word |= (((wd[i] & 0x01010101u) * 0x01020408u) >> 24) << (4 * i);
            const int valid = w - gw * 32;
            if (valid < 32) word &= (1u << valid) - 1u;
            bits[(int64_t)gy * wpr + gw] = word;
        }
    } else {
        store_tile<Tout>(s_out, dst, h, w, x0, y0);
    }
}

// ----------------------------------------------------------------------------------------------
// median 3x3 / 5x5 (cv2.medianBlur, BORDER_REPLICATE): exact rank filter via min/max exchanges
template <typename T, int KS>
__global__ void __launch_bounds__(kThreads) median_kernel(const T* __restrict__ src, T* __restrict__ dst,
                                                          int h, int w) {
    constexpr int VEC = 16 / sizeof(T);
    constexpr int R = KS / 2;
    constexpr int RA = round_up_c(R, VEC);
    constexpr int SW = TW + 2 * RA;
    constexpr int ROWS = TH + 2 * R;
    __shared__ __align__(16) T s_in[ROWS * SW];
    const int64_t frame = blockIdx.z;
    src += frame * (int64_t)h * w;
    dst += frame * (int64_t)h * w;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    load_tile<T, T>(src, s_in, h, w, x0, y0, R, RA, SW, ROWS, YAM_BORDER_REPLICATE);
    __syncthreads();
    for (int item = threadIdx.x; item < TH * TW; item += kThreads) {
        const int ry = item / TW, cx = item - ry * TW;
        const int gx = x0 + cx, gy = y0 + ry;
        if (gx >= w || gy >= h) continue;
        uint32_t p[KS * KS];
#pragma unroll
        for (int dy = 0; dy < KS; dy++)
#pragma unroll
            for (int dx = 0; dx < KS; dx++) p[dy * KS + dx] = s_in[(ry + dy) * SW + cx + RA - R + dx];
        dst[(int64_t)gy * w + gx] = (T)(KS == 3 ? median9(p) : median25(p));
    }
}

// ----------------------------------------------------------------------------------------------
// launch helpers

inline dim3 tile_grid(int64_t n, int64_t h, int64_t w) {
    return dim3((unsigned)((w + TW - 1) / TW), (unsigned)((h + TH - 1) / TH), (unsigned)n);
}

template <typename T>
size_t fixed_smem(int ks) {
    const int VEC = 16 / sizeof(T);
    const int R = ks / 2, RA = round_up_c(R, VEC), SW = TW + 2 * RA, ROWS = TH + 2 * R;
    // upper bound covering the padded pitches of the tiled u16 kernel (96 px, 68 words)
    const int SWP = SW < 96 ? 96 : SW, TWP = TW + 4;
    return (size_t)round_up_c(ROWS * SWP * (int)sizeof(T), 16) + (size_t)ROWS * TWP * 4 + (size_t)TH * TW * sizeof(T);
}

template <typename Tin>
size_t f32_smem(int ks) {
    const int VEC = 16 / sizeof(Tin);
    const int R = ks / 2, RA = round_up_c(R, VEC), SW = TW + 2 * RA, ROWS = TH + 2 * R;
    return (size_t)ROWS * SW * 4 + (size_t)ROWS * TW * 4;
}

template <typename Tin, typename Tout>
size_t f32_tiled_smem(int ks) {
    const int VEC = 16 / sizeof(Tin);
    const int R = ks / 2, RA = round_up_c(R, VEC > 4 ? VEC : 4), SW = TW + 2 * RA, ROWS = TH + 2 * R;
    const int SWP = ((SW / 4) % 2 == 0) ? SW + 4 : SW, TWP = TW + 4;
    return ((size_t)ROWS * SWP + 8 + (size_t)ROWS * TWP) * 4 + (sizeof(Tout) == 1 ? 0 : (size_t)TH * TW * sizeof(Tout));
}

// median for 8-bit images and windows of 7x7 .. 15x15 (cv2.medianBlur accepts ksize > 5 for CV_8U only,
// modules/preprocessing.py:147, range ui/control_metadata.py:210-218): the rank (k*k + 1) / 2 element is
// found by an 8-step binary search on the value, counting window pixels <= mid with the SIMD video
// compare (four pixels per instruction).  Thread = one output pixel; the (64 + 2r) x (64 + 2r) byte tile
// lives in shared memory.  Exact: a rank filter has no arithmetic to round.
__global__ void __launch_bounds__(kThreads) median_u8_search_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                                    int h, int w, int ks) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int R = ks / 2;
    const int RA = round_up_c(R, 16);
    const int SW = TW + 2 * RA;
    const int ROWS = TH + 2 * R;
    uint8_t* s_in = smem_raw;
    const int64_t frame = blockIdx.z;
    src += frame * (int64_t)h * w;
    dst += frame * (int64_t)h * w;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    load_tile<uint8_t, uint8_t>(src, s_in, h, w, x0, y0, R, RA, SW, ROWS, YAM_BORDER_REPLICATE);
    __syncthreads();
    const int need = (ks * ks + 1) / 2;
    for (int item = threadIdx.x; item < TH * TW; item += kThreads) {
        const int ry = item / TW, cx = item - ry * TW;
        const int gx = x0 + cx, gy = y0 + ry;
        if (gx >= w || gy >= h) continue;
        const uint8_t* win = s_in + ry * SW + (RA - R) + cx;   // top-left pixel of the window
        int lo = 0, hi = 255;                                   // smallest v with count(window <= v) >= need
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const uint32_t m4 = (uint32_t)mid * 0x01010101u;
            int count = 0;
            for (int dy = 0; dy < ks; dy++) {
                const uint8_t* row = win + dy * SW;
                int dx = 0;
                for (; dx + 4 <= ks; dx += 4) {
                    const uint32_t v = (uint32_t)row[dx] | ((uint32_t)row[dx + 1] << 8) | ((uint32_t)row[dx + 2] << 16) |
                                       ((uint32_t)row[dx + 3] << 24);
                    count += __popc(__vcmpleu4(v, m4)) >> 3;     // 0xff per byte that is <= mid
                }
                for (; dx < ks; dx++) count += row[dx] <= mid;
            }
            if (count >= need) hi = mid;
            else lo = mid + 1;
        }
        dst[(int64_t)gy * w + gx] = (uint8_t)lo;
    }
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        if (bytes > 227 * 1024) {
            yam_set_error("filter window too large for shared memory (%zu bytes)", bytes);
            return YAM_EINVAL;
        }
        YAM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    }
    return YAM_OK;
}

template <typename T, int EPI>
int launch_fixed(yam_ctx* ctx, const T* src, T* dst, int64_t n, int64_t h, int64_t w, const TapsU& taps,
                 int ks, int border, uint32_t box_div) {
    dim3 grid = tile_grid(n, h, w);
    const size_t smem = fixed_smem<T>(ks);
#define YAM_FIXED_CASE(K)                                                                         \
    case K: {                                                                                     \
        if (int rc = set_smem(sep_fixed_tiled<T, K, EPI>, smem)) return rc;                       \
        sep_fixed_tiled<T, K, EPI><<<grid, kThreads, smem, ctx->stream>>>(src, dst, (int)h, (int)w, \
                                                                          taps, border, box_div); \
        break;                                                                                    \
    }
    switch (ks) {
        YAM_FIXED_CASE(3)
        YAM_FIXED_CASE(5)
        YAM_FIXED_CASE(7)
        YAM_FIXED_CASE(9)
        YAM_FIXED_CASE(11)
        YAM_FIXED_CASE(13)
        YAM_FIXED_CASE(15)
        YAM_FIXED_CASE(17)
        default: {
            if (int rc = set_smem(sep_fixed_generic<T, EPI>, smem)) return rc;
            sep_fixed_generic<T, EPI><<<grid, kThreads, smem, ctx->stream>>>(src, dst, (int)h, (int)w, taps,
                                                                            ks, border, box_div);
        }
    }
#undef YAM_FIXED_CASE
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

template <typename Tin, typename Tout, int FEPI>
int launch_f32(yam_ctx* ctx, const Tin* src, Tout* dst, int64_t n, int64_t h, int64_t w, const TapsF& taps,
               int ks, int border, int sat_hi, int idelta, uint32_t* bits_out = nullptr, int wpr = 0) {
    dim3 grid = tile_grid(n, h, w);
    const size_t smem = f32_smem<Tin>(ks);
#define YAM_F32_CASE(K)                                                                              \
    case K: {                                                                                        \
        const size_t tsmem = f32_tiled_smem<Tin, Tout>(K);                                           \
        if (int rc = set_smem(sep_f32_tiled<Tin, Tout, K, FEPI>, tsmem)) return rc;                  \
        sep_f32_tiled<Tin, Tout, K, FEPI><<<grid, kThreads, tsmem, ctx->stream>>>(                   \
            src, dst, (int)h, (int)w, taps, border, idelta, bits_out, wpr);                          \
        break;                                                                                       \
    }
    switch (ks) {
        YAM_F32_CASE(3)
        YAM_F32_CASE(5)
        YAM_F32_CASE(7)
        YAM_F32_CASE(11)
        YAM_F32_CASE(15)
        default: {
            if (FEPI == FEPI_ADAPTIVE_BITS) {
                yam_set_error("adaptive_threshold_bits: block size %d has no fused kernel (use 3, 5, 7, 11 or 15)", ks);
                return YAM_EINVAL;
            }
            constexpr int GEPI = FEPI == FEPI_ADAPTIVE_BITS ? FEPI_ADAPTIVE : FEPI;
            if (int rc = set_smem(sep_f32_kernel<Tin, Tout, 0, GEPI>, smem)) return rc;
            sep_f32_kernel<Tin, Tout, 0, GEPI><<<grid, kThreads, smem, ctx->stream>>>(
                src, dst, (int)h, (int)w, taps, ks, border, sat_hi, idelta);
        }
    }
#undef YAM_F32_CASE
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

int check_shape(const void* src, const void* dst, int64_t n, int64_t h, int64_t w, const char* what) {
    YAM_REQUIRE(src && dst, "%s: NULL image pointer", what);
    YAM_REQUIRE(n > 0 && h > 0 && w > 0, "%s: empty image (%lld,%lld,%lld)", what, (long long)n, (long long)h, (long long)w);
    YAM_REQUIRE(n <= 65535, "%s: at most 65535 frames per call", what);
    YAM_REQUIRE(h < (1 << 30) && w < (1 << 30), "%s: image side too large", what);
    return YAM_OK;
}

}  // namespace

extern "C" {

int yam_gaussian(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w, int dtype,
                 int ksize, double sigma, int border) {
    if (int rc = yam_enter(ctx)) return rc;
    if (int rc = check_shape(src, dst, n, h, w, "gaussian")) return rc;
    YAM_REQUIRE(src != dst, "gaussian: in-place operation is not supported");
    YAM_REQUIRE(border == YAM_BORDER_REFLECT101 || border == YAM_BORDER_REPLICATE, "gaussian: unknown border %d", border);
    if (ksize <= 0) {
        // cv2 createGaussianKernels: ksize from sigma
        YAM_REQUIRE(sigma > 0, "gaussian: ksize and sigma cannot both be <= 0");
        ksize = (int)lrint(sigma * (dtype == YAM_U8 ? 3 : 4) * 2 + 1) | 1;
    }
    YAM_REQUIRE((ksize & 1) && ksize <= YAM_MAX_TAPS, "gaussian: ksize must be odd and <= %d, got %d", YAM_MAX_TAPS, ksize);
    double kf[YAM_MAX_TAPS];
    yam_host_gaussian_taps(ksize, sigma, kf);
    if (dtype == YAM_U8 || dtype == YAM_U16) {
        int64_t kq[YAM_MAX_TAPS];
        yam_host_fixed_taps(kf, ksize, dtype == YAM_U8 ? 8 : 16, kq);
        TapsU taps;
        for (int i = 0; i < ksize; i++) {
            YAM_REQUIRE(kq[i] >= 0, "gaussian: negative fixed-point tap (sigma too small for ksize)");
            taps.v[i] = (uint32_t)kq[i];
        }
        if (dtype == YAM_U8)
            return launch_fixed<uint8_t, EPI_SHIFT>(ctx, (const uint8_t*)src, (uint8_t*)dst, n, h, w, taps, ksize, border, 0);
        {
            int handled = 0;  // TMA-staged persistent kernel (yam_gauss_tma.cu) for the shapes a tensor map can describe
            if (int rc = yam_gauss16_tma(ctx, src, dst, n, h, w, ksize, taps.v, border, &handled)) return rc;
            if (handled) return YAM_OK;
        }
        return launch_fixed<uint16_t, EPI_SHIFT>(ctx, (const uint16_t*)src, (uint16_t*)dst, n, h, w, taps, ksize, border, 0);
    }
    YAM_REQUIRE(dtype == YAM_F32, "gaussian: unsupported dtype %d", dtype);
    TapsF taps;
    for (int i = 0; i < ksize; i++) taps.v[i] = (float)kf[i];
    return launch_f32<float, float, FEPI_STORE>(ctx, (const float*)src, (float*)dst, n, h, w, taps, ksize, border, 0, 0);
}

int yam_box(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w, int dtype, int ksize) {
    if (int rc = yam_enter(ctx)) return rc;
    if (int rc = check_shape(src, dst, n, h, w, "box")) return rc;
    YAM_REQUIRE(src != dst, "box: in-place operation is not supported");
    YAM_REQUIRE((ksize & 1) && ksize >= 1 && ksize <= 31, "box: ksize must be odd in [1,31], got %d", ksize);
    TapsU taps;
    for (int i = 0; i < ksize; i++) taps.v[i] = 1;
    const uint32_t kk = (uint32_t)(ksize * ksize);
    if (dtype == YAM_U8)
        return launch_fixed<uint8_t, EPI_BOX>(ctx, (const uint8_t*)src, (uint8_t*)dst, n, h, w, taps, ksize, YAM_BORDER_REFLECT101, kk);
    YAM_REQUIRE(dtype == YAM_U16, "box: unsupported dtype %d", dtype);
    return launch_fixed<uint16_t, EPI_BOX>(ctx, (const uint16_t*)src, (uint16_t*)dst, n, h, w, taps, ksize, YAM_BORDER_REFLECT101, kk);
}

int yam_adaptive_threshold(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w,
                           int dtype, int block_size, double C) {
    if (int rc = yam_enter(ctx)) return rc;
    if (int rc = check_shape(src, dst, n, h, w, "adaptive_threshold")) return rc;
    YAM_REQUIRE(src != dst, "adaptive_threshold: in-place operation is not supported");
    YAM_REQUIRE((block_size & 1) && block_size >= 3 && block_size <= YAM_MAX_TAPS,
                "adaptive_threshold: block_size must be odd in [3,%d], got %d", YAM_MAX_TAPS, block_size);
    double kf[YAM_MAX_TAPS];
    yam_host_gaussian_taps(block_size, 0.0, kf);
    TapsF taps;
    for (int i = 0; i < block_size; i++) taps.v[i] = (float)kf[i];
    const int idelta = (int)ceil(C);
    if (dtype == YAM_U8)
        return launch_f32<uint8_t, uint8_t, FEPI_ADAPTIVE>(ctx, (const uint8_t*)src, (uint8_t*)dst, n, h, w, taps,
                                                           block_size, YAM_BORDER_REPLICATE, 255, idelta);
    YAM_REQUIRE(dtype == YAM_U16, "adaptive_threshold: unsupported dtype %d", dtype);
    return launch_f32<uint16_t, uint8_t, FEPI_ADAPTIVE>(ctx, (const uint16_t*)src, (uint8_t*)dst, n, h, w, taps,
                                                        block_size, YAM_BORDER_REPLICATE, 65535, idelta);
}

static int adaptive_threshold_bits_impl(yam_ctx* ctx, const void* src, uint32_t* bits_out, int64_t n, int64_t h, int64_t w,
                                        int dtype, int block_size, double C, const int32_t* t_dev, void* mask_out, double maxval);

int yam_adaptive_threshold_bits(yam_ctx* ctx, const void* src, uint32_t* bits_out, int64_t n, int64_t h, int64_t w,
                                int dtype, int block_size, double C) {
    if (int rc = yam_enter(ctx)) return rc;
    return adaptive_threshold_bits_impl(ctx, src, bits_out, n, h, w, dtype, block_size, C, nullptr, nullptr, 0.0);
}

int yam_adaptive_threshold_bits_mask(yam_ctx* ctx, const void* src, uint32_t* bits_out, int64_t n, int64_t h, int64_t w,
                                     int dtype, int block_size, double C, const int32_t* thresh_dev, void* mask_out,
                                     double maxval) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(thresh_dev && mask_out && mask_out != src && n <= 65535, "adaptive_threshold_bits_mask: bad arguments");
    return adaptive_threshold_bits_impl(ctx, src, bits_out, n, h, w, dtype, block_size, C, thresh_dev, mask_out, maxval);
}

static int adaptive_threshold_bits_impl(yam_ctx* ctx, const void* src, uint32_t* bits_out, int64_t n, int64_t h, int64_t w,
                                        int dtype, int block_size, double C, const int32_t* t_dev, void* mask_out, double maxval) {
    if (int rc = check_shape(src, bits_out, n, h, w, "adaptive_threshold_bits")) return rc;
    YAM_REQUIRE((block_size & 1) && block_size >= 3 && block_size <= YAM_MAX_TAPS,
                "adaptive_threshold_bits: block_size must be odd in [3,%d], got %d", YAM_MAX_TAPS, block_size);
    double kf[YAM_MAX_TAPS];
    yam_host_gaussian_taps(block_size, 0.0, kf);
    TapsF taps;
    for (int i = 0; i < block_size; i++) taps.v[i] = (float)kf[i];
    const int idelta = (int)ceil(C);
    const int wpr = (int)((w + 31) / 32);
    YAM_REQUIRE(dtype == YAM_U8 || dtype == YAM_U16, "adaptive_threshold_bits: unsupported dtype %d", dtype);
    {
        int handled = 0;  // TMA-staged kernel (yam_adaptive.cu) for the shapes a tensor map can describe
        // the fused mask exists for 16-bit input; 8-bit input (and shapes the TMA kernel does not take) get the
        // mask from the plain threshold kernel
        const bool fuse_mask = mask_out && dtype == YAM_U16;
        if (int rc = yam_adaptive_bits_tma(ctx, src, bits_out, n, h, w, dtype, block_size, taps.v, idelta, &handled,
                                           fuse_mask ? t_dev : nullptr, fuse_mask ? mask_out : nullptr, maxval))
            return rc;
        if (handled && mask_out && !fuse_mask)
            if (int rc = yam_threshold_dev(ctx, src, mask_out, n, h * w, dtype, t_dev, maxval)) return rc;
        if (handled) return YAM_OK;
    }
    if (mask_out)
        if (int rc = yam_threshold_dev(ctx, src, mask_out, n, h * w, dtype, t_dev, maxval)) return rc;
    if (dtype == YAM_U8)
        return launch_f32<uint8_t, uint8_t, FEPI_ADAPTIVE_BITS>(ctx, (const uint8_t*)src, (uint8_t*)nullptr, n, h, w, taps,
                                                                block_size, YAM_BORDER_REPLICATE, 255, idelta, bits_out, wpr);
    YAM_REQUIRE(dtype == YAM_U16, "adaptive_threshold_bits: unsupported dtype %d", dtype);
    return launch_f32<uint16_t, uint8_t, FEPI_ADAPTIVE_BITS>(ctx, (const uint16_t*)src, (uint8_t*)nullptr, n, h, w, taps,
                                                             block_size, YAM_BORDER_REPLICATE, 65535, idelta, bits_out, wpr);
}

int yam_median(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w, int dtype, int ksize) {
    if (int rc = yam_enter(ctx)) return rc;
    if (int rc = check_shape(src, dst, n, h, w, "median")) return rc;
    YAM_REQUIRE(src != dst, "median: in-place operation is not supported");
    YAM_REQUIRE(dtype == YAM_U8 || dtype == YAM_U16, "median: unsupported dtype %d", dtype);
    dim3 grid = tile_grid(n, h, w);
    if (dtype == YAM_U8 && ksize >= 7 && ksize <= 15 && (ksize & 1)) {
        const int R = ksize / 2, RA = round_up_c(R, 16);
        const size_t smem = (size_t)(TH + 2 * R) * (TW + 2 * RA);
        median_u8_search_kernel<<<grid, kThreads, smem, ctx->stream>>>((const uint8_t*)src, (uint8_t*)dst, (int)h, (int)w, ksize);
        YAM_LAUNCHED(ctx);
        return YAM_OK;
    }
    YAM_REQUIRE(ksize == 3 || ksize == 5, "median: ksize must be 3 or 5, or 7..15 (odd) for uint8 as in cv2 (got %d)", ksize);
    if (dtype == YAM_U8) {
        if (ksize == 3) median_kernel<uint8_t, 3><<<grid, kThreads, 0, ctx->stream>>>((const uint8_t*)src, (uint8_t*)dst, (int)h, (int)w);
        else median_kernel<uint8_t, 5><<<grid, kThreads, 0, ctx->stream>>>((const uint8_t*)src, (uint8_t*)dst, (int)h, (int)w);
    } else {
        if (ksize == 3) median_kernel<uint16_t, 3><<<grid, kThreads, 0, ctx->stream>>>((const uint16_t*)src, (uint16_t*)dst, (int)h, (int)w);
        else median_kernel<uint16_t, 5><<<grid, kThreads, 0, ctx->stream>>>((const uint16_t*)src, (uint16_t*)dst, (int)h, (int)w);
    }
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

}  // extern "C"
