// the remaining cheap per-pixel steps of the reference's menus, so that a whole
// UI pipeline can stay on the device.
//   add_weighted   cv2.addWeighted (SharpenModule, modules/preprocessing.py:167-171)
//   select_channel SelectChannelModule (modules/preprocessing.py:188-209)
//   border_clear   remove_border_regions (core/segmentation.py:316-325)
//   edge_filter    sobel_operator / prewitt_operator / laplacian_operator (core/segmentation.py:150-169)
// All integer stencils are exact (int32 accumulators: |sum| <= 65535 * 64 * 20 for the 7-tap Sobel);
// the square roots are IEEE correctly rounded (__dsqrt_rn / __fsqrt_rn).
#include "yam_common.cuh"

namespace {

constexpr int kThreads = 256;

// ---- addWeighted: r = fmaf(a, alpha, fmaf(b, beta, gamma)), saturate(rint(r)) ------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) add_weighted_kernel(const T* __restrict__ a, const T* __restrict__ b,
                                                                T* __restrict__ dst, int64_t count, float alpha,
                                                                float beta, float gamma) {
    constexpr int VEC = 16 / sizeof(T);
    constexpr int HI = sizeof(T) == 1 ? 255 : 65535;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool aligned = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
    int64_t done = 0;
    if (aligned) {
        const int64_t groups = count / VEC;
        for (int64_t g = tid; g < groups; g += stride) {
            const uint4 qa = yam_ld_stream(reinterpret_cast<const uint4*>(a) + g);
            const uint4 qb = yam_ld_stream(reinterpret_cast<const uint4*>(b) + g);
            const T* pa = reinterpret_cast<const T*>(&qa);
            const T* pb = reinterpret_cast<const T*>(&qb);
            uint4 qo;
            T* po = reinterpret_cast<T*>(&qo);
#pragma unroll
            for (int i = 0; i < VEC; i++)
                po[i] = (T)yam_rint_sat(__fmaf_rn((float)pa[i], alpha, __fmaf_rn((float)pb[i], beta, gamma)), HI);
            yam_st_stream(reinterpret_cast<uint4*>(dst) + g, qo);
        }
        done = groups * VEC;
    }
    for (int64_t i = done + tid; i < count; i += stride)
        dst[i] = (T)yam_rint_sat(__fmaf_rn((float)a[i], alpha, __fmaf_rn((float)b[i], beta, gamma)), HI);
}

// ---- channel selection from interleaved BGR ----------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) select_channel_kernel(const T* __restrict__ bgr, T* __restrict__ dst,
                                                                  int64_t px, int c0, int c1) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < px; i += stride) {
        const T* p = bgr + 3 * i;
        // pair modes: np.uint8((x.astype(f32) + y.astype(f32)) / 2) == (x + y) >> 1 for uint8 inputs
        dst[i] = c1 < 0 ? p[c0] : (T)(((uint32_t)p[c0] + (uint32_t)p[c1]) >> 1);
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) gray2bgr_kernel(const T* __restrict__ gray, T* __restrict__ bgr, int64_t px) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < px; i += stride) {
        const T v = gray[i];
        bgr[3 * i] = v;
        bgr[3 * i + 1] = v;
        bgr[3 * i + 2] = v;
    }
}

// ---- border clearing: element (y, x, ch) survives iff d <= y < h-d and d <= x < w-d ---------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) border_clear_kernel(const T* __restrict__ src, T* __restrict__ dst, int h,
                                                                int64_t row_elems, int ch, int d, int64_t frames) {
    // one thread per element of a row; rows/frames in the grid's y/z
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= row_elems) return;
    const int x = (int)(e / ch);
    const int w = (int)(row_elems / ch);
    const bool keep_x = x >= d && x < w - d;
    for (int64_t f = blockIdx.z; f < frames; f += gridDim.z)
        for (int y = blockIdx.y; y < h; y += gridDim.y) {
            const int64_t idx = (f * h + y) * row_elems + e;
            dst[idx] = (keep_x && y >= d && y < h - d) ? src[idx] : (T)0;
        }
}

// ---- 3x3 .. 7x7 derivative stencils ------------------------------------------------------------------------------
struct EdgeTaps {
    int k;          // taps per axis (1, 3, 5 or 7; Laplacian ksize 1 uses 3 taps with smooth = delta)
    int smooth[7];  // smoothing kernel
    int d1[7];      // first derivative
    int d2[7];      // second derivative
};
enum { EDGE_SOBEL = 0, EDGE_PREWITT = 1, EDGE_LAPLACIAN = 2, EDGE_LAPLACIAN4 = 3 };

constexpr int ETW = 32, ETH = 32;

template <typename T, int KIND>
__global__ void __launch_bounds__(kThreads) edge_kernel(const T* __restrict__ src, uint8_t* __restrict__ dst, int h,
                                                        int w, EdgeTaps taps) {
    constexpr int RMAX = 3;
    __shared__ int s_in[(ETH + 2 * RMAX) * (ETW + 2 * RMAX)];
    const int r = taps.k / 2;
    const int SW = ETW + 2 * r, ROWS = ETH + 2 * r;
    src += (int64_t)blockIdx.z * h * w;
    dst += (int64_t)blockIdx.z * h * w;
    const int x0 = blockIdx.x * ETW, y0 = blockIdx.y * ETH;
    for (int i = threadIdx.x; i < ROWS * SW; i += kThreads) {
        const int ry = i / SW, rx = i - ry * SW;
        const int gy = yam_border(y0 - r + ry, h, YAM_BORDER_REFLECT101);
        const int gx = yam_border(x0 - r + rx, w, YAM_BORDER_REFLECT101);
        s_in[i] = (int)src[(int64_t)gy * w + gx];
    }
    __syncthreads();
    constexpr int HI = sizeof(T) == 1 ? 255 : 65535;
    for (int o = threadIdx.x; o < ETH * ETW; o += kThreads) {
        const int ty = o / ETW, tx = o - ty * ETW;
        const int gy = y0 + ty, gx = x0 + tx;
        if (gy >= h || gx >= w) continue;
        int ax = 0, ay = 0;  // Sobel/Prewitt: d/dx, d/dy; Laplacian: the two second derivatives summed in ax
        for (int i = 0; i < taps.k; i++) {
            const int* row = s_in + (ty + i) * SW + tx;
            int sd = 0, ss = 0;
            for (int j = 0; j < taps.k; j++) {
                const int v = row[j];
                if (KIND == EDGE_LAPLACIAN || KIND == EDGE_LAPLACIAN4) sd += taps.d2[j] * v;
                else sd += taps.d1[j] * v;
                ss += taps.smooth[j] * v;
            }
            if (KIND == EDGE_LAPLACIAN) {
                ax += taps.smooth[i] * sd + taps.d2[i] * ss;
            } else if (KIND == EDGE_LAPLACIAN4) {
                // ksize 1: [1,-2,1] along x on the centre row + [1,-2,1] along y on the centre column
                if (i == 1) ax += sd;
                ax += taps.d2[i] * row[1];
            } else {
                ax += taps.smooth[i] * sd;
                ay += taps.d1[i] * ss;
            }
        }
        int out;
        if (KIND == EDGE_SOBEL) {
            // CV_64F gradients (exact), cv2.magnitude in double, clip, truncate
            const double dx = (double)ax, dy = (double)ay;
            const double m = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
            out = m >= 255.0 ? 255 : (int)m;
        } else if (KIND == EDGE_PREWITT) {
            // filter2D(ddepth=-1) saturates each gradient to the input dtype; float32 magnitude
            const float fx = (float)min(max(ax, 0), HI), fy = (float)min(max(ay, 0), HI);
            const float m = __fsqrt_rn(__fmaf_rn(fx, fx, __fmul_rn(fy, fy)));
            out = m >= 255.0f ? 255 : (int)m;
        } else {
            const int a = ax < 0 ? -ax : ax;
            out = a > 255 ? 255 : a;
        }
        dst[(int64_t)gy * w + gx] = (uint8_t)out;
    }
}

// ---- per-row power sums of the set pixels of a mask: out[y] = (count, sum x, sum x^2, sum x^3) ------------
// (exact int64; the host combines the rows with powers of y in arbitrary precision -> cv2.moments)
template <typename T>
__global__ void __launch_bounds__(kThreads) mask_row_moments_kernel(const T* __restrict__ mask, int h, int w,
                                                                    long long* __restrict__ out) {
    const int warp = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (warp >= h) return;
    const T* row = mask + (int64_t)warp * w;
    long long s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int x = lane; x < w; x += 32) {
        if (row[x]) {
            const long long xx = x;
            s0 += 1;
            s1 += xx;
            s2 += xx * xx;
            s3 += xx * xx * xx;
        }
    }
    s0 = yam_warp_sum(s0);
    s1 = yam_warp_sum(s1);
    s2 = yam_warp_sum(s2);
    s3 = yam_warp_sum(s3);
    if (lane == 0) {
        long long* o = out + (int64_t)warp * 4;
        o[0] = s0; o[1] = s1; o[2] = s2; o[3] = s3;
    }
}

unsigned stream_blocks(yam_ctx* ctx, int64_t items) {
    int64_t b = (items + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)ctx->num_sms * 8;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" {

int yam_add_weighted(yam_ctx* ctx, const void* a, const void* b, void* dst, int64_t count, int dtype, double alpha,
                     double beta, double gamma) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(a && b && dst && count > 0, "add_weighted: bad arguments");
    YAM_REQUIRE(dtype == YAM_U8 || dtype == YAM_U16, "add_weighted: unsupported dtype %d", dtype);
    const unsigned blocks = stream_blocks(ctx, count / (dtype == YAM_U8 ? 16 : 8) + 1);
    if (dtype == YAM_U8)
        add_weighted_kernel<uint8_t><<<blocks, kThreads, 0, ctx->stream>>>((const uint8_t*)a, (const uint8_t*)b, (uint8_t*)dst,
                                                                           count, (float)alpha, (float)beta, (float)gamma);
    else
        add_weighted_kernel<uint16_t><<<blocks, kThreads, 0, ctx->stream>>>((const uint16_t*)a, (const uint16_t*)b,
                                                                            (uint16_t*)dst, count, (float)alpha, (float)beta,
                                                                            (float)gamma);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

int yam_select_channel(yam_ctx* ctx, const void* bgr, void* dst, int64_t px, int dtype, int mode) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(bgr && dst && px > 0, "select_channel: bad arguments");
    YAM_REQUIRE(dtype == YAM_U8 || dtype == YAM_U16, "select_channel: unsupported dtype %d", dtype);
    YAM_REQUIRE(mode >= 0 && mode <= YAM_CHANNEL_BR, "select_channel: unknown mode %d", mode);
    YAM_REQUIRE(mode <= YAM_CHANNEL_R || dtype == YAM_U8, "select_channel: two-channel means are uint8 only");
    // interleaved order is B, G, R
    static const int first[] = {0, 1, 2, 2, 1, 0}, second[] = {-1, -1, -1, 1, 0, 2};
    const unsigned blocks = stream_blocks(ctx, px);
    if (dtype == YAM_U8)
        select_channel_kernel<uint8_t><<<blocks, kThreads, 0, ctx->stream>>>((const uint8_t*)bgr, (uint8_t*)dst, px, first[mode], second[mode]);
    else
        select_channel_kernel<uint16_t><<<blocks, kThreads, 0, ctx->stream>>>((const uint16_t*)bgr, (uint16_t*)dst, px, first[mode], second[mode]);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

int yam_gray2bgr(yam_ctx* ctx, const void* gray, void* bgr, int64_t px, int dtype) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(gray && bgr && px > 0, "gray2bgr: bad arguments");
    YAM_REQUIRE(dtype == YAM_U8 || dtype == YAM_U16, "gray2bgr: unsupported dtype %d", dtype);
    const unsigned blocks = stream_blocks(ctx, px);
    if (dtype == YAM_U8) gray2bgr_kernel<uint8_t><<<blocks, kThreads, 0, ctx->stream>>>((const uint8_t*)gray, (uint8_t*)bgr, px);
    else gray2bgr_kernel<uint16_t><<<blocks, kThreads, 0, ctx->stream>>>((const uint16_t*)gray, (uint16_t*)bgr, px);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

int yam_border_clear(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w, int channels, int dtype,
                     int border_distance) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(src && dst && n > 0 && h > 0 && w > 0 && channels >= 1 && channels <= 4, "border_clear: bad arguments");
    YAM_REQUIRE(h < (1 << 30) && w < (1 << 30), "border_clear: image side too large");
    YAM_REQUIRE(dtype == YAM_U8 || dtype == YAM_U16, "border_clear: unsupported dtype %d", dtype);
    // the reference slices [d:-d]: d == 0 ("[0:-0]" is empty) or 2d >= side clears everything
    int d = border_distance;
    if (d <= 0 || 2ll * d >= h || 2ll * d >= w) d = (int)((h > w ? h : w) + 1);
    const int64_t row_elems = w * channels;
    dim3 grid((unsigned)((row_elems + kThreads - 1) / kThreads), (unsigned)(h < 1024 ? h : 1024), (unsigned)(n < 64 ? n : 64));
    if (dtype == YAM_U8)
        border_clear_kernel<uint8_t><<<grid, kThreads, 0, ctx->stream>>>((const uint8_t*)src, (uint8_t*)dst, (int)h, row_elems, channels, d, n);
    else
        border_clear_kernel<uint16_t><<<grid, kThreads, 0, ctx->stream>>>((const uint16_t*)src, (uint16_t*)dst, (int)h, row_elems, channels, d, n);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

int yam_mask_row_moments(yam_ctx* ctx, const void* mask, int64_t h, int64_t w, int dtype, int64_t* out_dev) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(mask && out_dev && h > 0 && w > 0, "mask_row_moments: bad arguments");
    YAM_REQUIRE(h < (1 << 30) && w < (1 << 20), "mask_row_moments: image side too large for exact int64 row sums");
    YAM_REQUIRE(dtype == YAM_U8 || dtype == YAM_U16, "mask_row_moments: unsupported dtype %d", dtype);
    const unsigned blocks = (unsigned)((h * 32 + kThreads - 1) / kThreads);
    if (dtype == YAM_U8)
        mask_row_moments_kernel<uint8_t><<<blocks, kThreads, 0, ctx->stream>>>((const uint8_t*)mask, (int)h, (int)w, (long long*)out_dev);
    else
        mask_row_moments_kernel<uint16_t><<<blocks, kThreads, 0, ctx->stream>>>((const uint16_t*)mask, (int)h, (int)w, (long long*)out_dev);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

int yam_edge_filter(yam_ctx* ctx, const void* src, void* dst_u8, int64_t n, int64_t h, int64_t w, int dtype, int kind,
                    int ksize) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(src && dst_u8 && n > 0 && n <= 65535 && h > 0 && w > 0, "edge_filter: bad arguments");
    YAM_REQUIRE(h < (1 << 30) && w < (1 << 30), "edge_filter: image side too large");
    YAM_REQUIRE(dtype == YAM_U8 || dtype == YAM_U16, "edge_filter: unsupported dtype %d", dtype);
    YAM_REQUIRE(kind == YAM_EDGE_SOBEL || kind == YAM_EDGE_PREWITT || kind == YAM_EDGE_LAPLACIAN, "edge_filter: unknown kind %d", kind);
    if (kind == YAM_EDGE_PREWITT) ksize = 3;
    YAM_REQUIRE(ksize == 1 || ksize == 3 || ksize == 5 || ksize == 7,
                "edge_filter: ksize %d is outside the exact integer range of the device kernel (1, 3, 5, 7)", ksize);
    // cv2.getDerivKernels(normalize=False)
    static const int S[4][7] = {{1}, {1, 2, 1}, {1, 4, 6, 4, 1}, {1, 6, 15, 20, 15, 6, 1}};
    static const int D1[4][7] = {{-1, 0, 1}, {-1, 0, 1}, {-1, -2, 0, 2, 1}, {-1, -4, -5, 0, 5, 4, 1}};
    static const int D2[4][7] = {{1, -2, 1}, {1, -2, 1}, {1, 0, -2, 0, 1}, {1, 2, -1, -4, -1, 2, 1}};
    EdgeTaps t;
    memset(&t, 0, sizeof(t));
    const int idx = ksize / 2;
    t.k = ksize == 1 ? 3 : ksize;
    for (int i = 0; i < t.k; i++) {
        t.d1[i] = D1[idx][i];
        t.d2[i] = D2[idx][i];
        t.smooth[i] = ksize == 1 ? (i == 1 ? 1 : 0) : S[idx][i];
    }
    int k = kind;
    if (kind == YAM_EDGE_PREWITT) {
        // filter2D correlates with [[1,0,-1]]*3 and its transpose (core/segmentation.py:159-160)
        const int pd[3] = {1, 0, -1};
        for (int i = 0; i < 3; i++) {
            t.d1[i] = pd[i];
            t.smooth[i] = 1;
        }
    } else if (kind == YAM_EDGE_LAPLACIAN && ksize == 1) {
        k = EDGE_LAPLACIAN4;
    }
    dim3 grid((unsigned)((w + ETW - 1) / ETW), (unsigned)((h + ETH - 1) / ETH), (unsigned)n);
#define YAM_EDGE_LAUNCH(T, KIND) edge_kernel<T, KIND><<<grid, kThreads, 0, ctx->stream>>>((const T*)src, (uint8_t*)dst_u8, (int)h, (int)w, t)
    if (dtype == YAM_U8) {
        if (k == EDGE_SOBEL) YAM_EDGE_LAUNCH(uint8_t, EDGE_SOBEL);
        else if (k == EDGE_PREWITT) YAM_EDGE_LAUNCH(uint8_t, EDGE_PREWITT);
        else if (k == EDGE_LAPLACIAN) YAM_EDGE_LAUNCH(uint8_t, EDGE_LAPLACIAN);
        else YAM_EDGE_LAUNCH(uint8_t, EDGE_LAPLACIAN4);
    } else {
        if (k == EDGE_SOBEL) YAM_EDGE_LAUNCH(uint16_t, EDGE_SOBEL);
        else if (k == EDGE_PREWITT) YAM_EDGE_LAUNCH(uint16_t, EDGE_PREWITT);
        else if (k == EDGE_LAPLACIAN) YAM_EDGE_LAUNCH(uint16_t, EDGE_LAPLACIAN);
        else YAM_EDGE_LAUNCH(uint16_t, EDGE_LAPLACIAN4);
    }
#undef YAM_EDGE_LAUNCH
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

}  // extern "C"
