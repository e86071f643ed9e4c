// Luma access for colour images: Preprocessor.histogram_equalization on a BGR image
// (core/preprocessing.py:74-79) is  cvtColor(BGR2YCrCb) -> equalizeHist(Y) -> cvtColor(YCrCb2BGR).
// The Cr / Cb planes never need to exist in memory: yam_bgr_luma_ycrcb writes the Y plane (the input of
// the ordinary uint8 equalisation), yam_bgr_replace_luma_ycrcb recomputes (Y, Cr, Cb) of every pixel from
// the source, swaps in the new Y and converts back -- cv2's 14-bit fixed point, exact over all 2^24
// colours (oracle/np_oracle.py: bgr2ycrcb_u8, ycrcb2bgr_u8, pinned against cv2 4.13.0).
// Four pixels (12 bytes = three 32-bit words) per thread step; a scalar loop takes the tail and
// pointers that are not 4-byte aligned.
#include "yam_common.cuh"

namespace {

constexpr int kShift = 14;
constexpr int kHalf = 1 << (kShift - 1);
constexpr int kDelta = 128 << kShift;

__device__ __forceinline__ int sat8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

__device__ __forceinline__ int luma_of(int b, int g, int r) { return (r * 4899 + g * 9617 + b * 1868 + kHalf) >> kShift; }

// (b, g, r) -> the pixel with its luma replaced: what YCrCb2BGR returns for (y_new, Cr, Cb) of the source
__device__ __forceinline__ void replace_luma(int b, int g, int r, int y_new, int& ob, int& og, int& orr) {
    const int y = luma_of(b, g, r);                                           // <= 255 by construction
    const int cr = sat8(((r - y) * 11682 + kDelta + kHalf) >> kShift) - 128;  // stored as uint8, read back
    const int cb = sat8(((b - y) * 9241 + kDelta + kHalf) >> kShift) - 128;
    ob = sat8(y_new + ((cb * 29049 + kHalf) >> kShift));
    og = sat8(y_new + ((cb * -5636 + cr * -11698 + kHalf) >> kShift));
    orr = sat8(y_new + ((cr * 22987 + kHalf) >> kShift));
}

__device__ __forceinline__ int byte_of(const uint32_t (&w)[3], int i) { return (int)((w[i >> 2] >> (8 * (i & 3))) & 255u); }

__global__ void __launch_bounds__(256) bgr_luma_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ y, int64_t px,
                                                       int64_t quads) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t* s32 = (const uint32_t*)src;
    uint32_t* y32 = (uint32_t*)y;
    for (int64_t q = t0; q < quads; q += stride) {
        uint32_t w[3] = {s32[3 * q], s32[3 * q + 1], s32[3 * q + 2]};
        uint32_t out = 0;
#pragma unroll
        for (int p = 0; p < 4; p++) out |= (uint32_t)luma_of(byte_of(w, 3 * p), byte_of(w, 3 * p + 1), byte_of(w, 3 * p + 2)) << (8 * p);
        y32[q] = out;
    }
    for (int64_t i = 4 * quads + t0; i < px; i += stride) y[i] = (uint8_t)luma_of(src[3 * i], src[3 * i + 1], src[3 * i + 2]);
}

__global__ void __launch_bounds__(256) bgr_replace_luma_kernel(const uint8_t* __restrict__ src, const uint8_t* __restrict__ y_new,
                                                               uint8_t* __restrict__ dst, int64_t px, int64_t quads) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t* s32 = (const uint32_t*)src;
    const uint32_t* y32 = (const uint32_t*)y_new;
    uint32_t* d32 = (uint32_t*)dst;
    for (int64_t q = t0; q < quads; q += stride) {
        uint32_t w[3] = {s32[3 * q], s32[3 * q + 1], s32[3 * q + 2]};
        const uint32_t yn = y32[q];
        uint32_t o[3] = {0, 0, 0};
#pragma unroll
        for (int p = 0; p < 4; p++) {
            int c[3];
            replace_luma(byte_of(w, 3 * p), byte_of(w, 3 * p + 1), byte_of(w, 3 * p + 2), (int)((yn >> (8 * p)) & 255u), c[0], c[1], c[2]);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const int i = 3 * p + k;
                o[i >> 2] |= (uint32_t)c[k] << (8 * (i & 3));
            }
        }
        d32[3 * q] = o[0];
        d32[3 * q + 1] = o[1];
        d32[3 * q + 2] = o[2];
    }
    for (int64_t i = 4 * quads + t0; i < px; i += stride) {
        int b, g, r;
        replace_luma(src[3 * i], src[3 * i + 1], src[3 * i + 2], y_new[i], b, g, r);
        dst[3 * i] = (uint8_t)b;
        dst[3 * i + 1] = (uint8_t)g;
        dst[3 * i + 2] = (uint8_t)r;
    }
}

}  // namespace

static inline bool aligned4(const void* p) { return ((uintptr_t)p & 3u) == 0; }

static inline unsigned color_grid(const yam_ctx* ctx, int64_t px) {
    int64_t bx = (px / 4 + 255) / 256;
    const int64_t cap = (int64_t)ctx->num_sms * 8;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    return (unsigned)bx;
}

extern "C" {

int yam_bgr_luma_ycrcb(yam_ctx* ctx, const uint8_t* bgr, uint8_t* y, int64_t px) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(bgr && y && px > 0, "bgr_luma_ycrcb: bad arguments");
    const int64_t quads = (aligned4(bgr) && aligned4(y)) ? px / 4 : 0;
    bgr_luma_kernel<<<color_grid(ctx, px), 256, 0, ctx->stream>>>(bgr, y, px, quads);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

int yam_bgr_replace_luma_ycrcb(yam_ctx* ctx, const uint8_t* bgr, const uint8_t* y_new, uint8_t* dst, int64_t px) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(bgr && y_new && dst && px > 0, "bgr_replace_luma_ycrcb: bad arguments");
    const int64_t quads = (aligned4(bgr) && aligned4(y_new) && aligned4(dst)) ? px / 4 : 0;
    bgr_replace_luma_kernel<<<color_grid(ctx, px), 256, 0, ctx->stream>>>(bgr, y_new, dst, px, quads);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

}  // extern "C"
