// SURVEY.md 8(f) N4 -- the watershed front half (core/segmentation.py:97-111):
//   thresh  = cv2.threshold(gray, 0, 255, THRESH_BINARY_INV + THRESH_OTSU)      yam_threshold_inv
//   opening = morphologyEx(thresh, OPEN, ones(k,k), iterations)                  yam_morph (existing)
//   sure_bg = dilate(opening, ones(k,k), iterations)                             yam_morph (existing)
//   dist    = cv2.distanceTransform(opening, DIST_L2, 5)                         yam_distance_transform
//   sure_fg = uint8(threshold(dist, factor * dist.max(), 255, BINARY))           yam_minmax + yam_threshold + yam_convert_scale_abs
//   markers = connectedComponents(sure_fg) + 1; markers[sure_bg - sure_fg == 255] = 0      yam_ccl_label + yam_watershed_combine
//
// cv2 4.13's DIST_L2 / 5x5 transform is the float32 two-pass chamfer scan with weights a = 1, b = 1.4,
// c = 2.1969 (verified bit for bit against the oracle's restatement): a sequential raster recurrence.
// Its result is the shortest path length from the nearest zero pixel in the chamfer graph, every path
// length being a float32 sum accumulated from the zero pixel outwards.  The kernel below computes the
// LEAST FIXED POINT of d(p) = min(d(p), d(q) + w(p - q)) by tile-wise relaxation in shared memory until
// nothing changes: the minimum over ALL paths.  In exact arithmetic both are the chamfer metric; in
// float32 the two-pass scan fixes one accumulation order per path family, so the fixed point can be
// one ulp lower at isolated pixels.  Parity is therefore stated as a tolerance (1e-5 relative, the
// north_star's float tolerance), not bit-exactness; tests/test_gpu_n4.py reports the exact-match rate.
#include "yam_common.cuh"
#include "yam_host.h"

namespace {

constexpr int kT = 32;          // tile side (one pixel per thread)
constexpr int kH = 2;           // halo: the 5x5 mask reaches two pixels
constexpr int kS = kT + 2 * kH;
constexpr float kBig = 1.0e30f;
constexpr float kA = 1.0f, kB = 1.4f, kC = 2.1969f;

__global__ void __launch_bounds__(256) dt_init_kernel(const uint8_t* __restrict__ mask, float* __restrict__ d, int64_t count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) d[i] = mask[i] ? kBig : 0.0f;
}

// one launch = up to `iters` Jacobi sweeps of every 32 x 32 tile against a halo read at launch time;
// *changed is raised when any pixel got smaller, the host relaunches until a launch changes nothing
__global__ void __launch_bounds__(kT * kT) dt_relax_kernel(float* __restrict__ d, int h, int w, int iters,
                                                         int* __restrict__ changed) {
    __shared__ float s[kS][kS + 1];
    float* frame = d + (int64_t)blockIdx.z * h * w;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int x0 = blockIdx.x * kT, y0 = blockIdx.y * kT;
    for (int i = ty * kT + tx; i < kS * kS; i += kT * kT) {
        const int ry = i / kS, rx = i - ry * kS;
        const int gy = y0 - kH + ry, gx = x0 - kH + rx;
        s[ry][rx] = (gy >= 0 && gy < h && gx >= 0 && gx < w) ? frame[(int64_t)gy * w + gx] : kBig;
    }
    __syncthreads();
    const int cy = ty + kH, cx = tx + kH;
    float v = s[cy][cx];
    bool any = false;
    for (int it = 0; it < iters; it++) {
        float nv = v;
        if (v > 0.0f) {
            const float a4 = fminf(fminf(s[cy - 1][cx], s[cy + 1][cx]), fminf(s[cy][cx - 1], s[cy][cx + 1]));
            const float b4 = fminf(fminf(s[cy - 1][cx - 1], s[cy - 1][cx + 1]), fminf(s[cy + 1][cx - 1], s[cy + 1][cx + 1]));
            const float c8 = fminf(fminf(fminf(s[cy - 2][cx - 1], s[cy - 2][cx + 1]), fminf(s[cy + 2][cx - 1], s[cy + 2][cx + 1])),
                                   fminf(fminf(s[cy - 1][cx - 2], s[cy - 1][cx + 2]), fminf(s[cy + 1][cx - 2], s[cy + 1][cx + 2])));
            // x -> x + w is monotone under round-to-nearest, so min(neighbours) + w == min(neighbour + w)
            nv = fminf(fminf(v, __fadd_rn(a4, kA)), fminf(__fadd_rn(b4, kB), __fadd_rn(c8, kC)));
        }
        const bool lower = nv < v;
        __syncthreads();             // every thread has read the old neighbourhood
        if (lower) s[cy][cx] = v = nv;
        const int moved = __syncthreads_or(lower ? 1 : 0);
        if (!moved) break;
        any = any || lower;
    }
    const int gy = y0 + ty, gx = x0 + tx;
    if (any && gy < h && gx < w) {
        frame[(int64_t)gy * w + gx] = v;
        *changed = 1;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) threshold_inv_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t frame_px,
                                                            const int32_t* __restrict__ t_dev, float t_host, T maxval) {
    const float t = t_dev ? (float)t_dev[blockIdx.z] : t_host;
    const T* s = src + (int64_t)blockIdx.z * frame_px;
    T* d = dst + (int64_t)blockIdx.z * frame_px;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < frame_px; i += stride)
        d[i] = ((float)s[i] > t) ? (T)0 : maxval;
}

__global__ void __launch_bounds__(256) watershed_combine_kernel(const int32_t* __restrict__ labels, const uint8_t* __restrict__ sure_bg,
                                                                const uint8_t* __restrict__ sure_fg, int32_t* __restrict__ markers,
                                                                int64_t count) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        // unknown = cv2.subtract(sure_bg, sure_fg) == 255  <=>  sure_bg == 255 and sure_fg == 0
        const bool unknown = sure_bg[i] == 255 && sure_fg[i] == 0;
        markers[i] = unknown ? 0 : labels[i] + 1;
    }
}

// second-order raw moments per label: thread = 8 consecutive pixels of a row, runs of one label are
// summed in registers and flushed with three 64-bit atomics (regions are a few runs per row)
__global__ void __launch_bounds__(256) region_moments_kernel(const int32_t* __restrict__ labels, int h, int w, int64_t n_labels,
                                                             unsigned long long* __restrict__ out) {
    const int groups = (w + 7) / 8;
    const int64_t total = (int64_t)h * groups;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
        const int y = (int)(g / groups), x0 = (int)(g - (int64_t)y * groups) * 8;
        const int32_t* row = labels + (int64_t)y * w;
        int cur = 0;
        unsigned long long cnt = 0, sx = 0, sxx = 0;
        auto flush = [&]() {
            if (cur > 0 && cur <= n_labels) {
                unsigned long long* o = out + (int64_t)(cur - 1) * 3;
                atomicAdd(o + 0, cnt * (unsigned long long)y * (unsigned long long)y);   // sum r^2
                atomicAdd(o + 1, sxx);                                                    // sum c^2
                atomicAdd(o + 2, sx * (unsigned long long)y);                             // sum r*c
            }
        };
        for (int i = 0; i < 8 && x0 + i < w; i++) {
            const int l = row[x0 + i];
            if (l != cur) {
                flush();
                cur = l;
                cnt = sx = sxx = 0;
            }
            const unsigned long long x = (unsigned long long)(x0 + i);
            cnt++;
            sx += x;
            sxx += x * x;
        }
        flush();
    }
}

}  // namespace

extern "C" {

int yam_region_moments(yam_ctx* ctx, const int32_t* labels, int64_t h, int64_t w, int64_t n_labels, int64_t* moments_dev) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(labels && moments_dev && h > 0 && w > 0 && n_labels >= 0, "region_moments: bad arguments");
    YAM_REQUIRE(h < (1 << 24) && w < (1 << 24), "region_moments: image side must be below 2^24");
    if (n_labels == 0) return YAM_OK;
    YAM_CUDA(cudaMemsetAsync(moments_dev, 0, (size_t)n_labels * 3 * sizeof(int64_t), ctx->stream));
    const int64_t total = h * ((w + 7) / 8);
    int64_t bx = (total + 255) / 256;
    const int64_t cap = (int64_t)ctx->num_sms * 16;
    if (bx > cap) bx = cap;
    region_moments_kernel<<<(unsigned)bx, 256, 0, ctx->stream>>>(labels, (int)h, (int)w, n_labels, (unsigned long long*)moments_dev);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

int yam_distance_transform(yam_ctx* ctx, const uint8_t* mask, float* dist, int64_t n, int64_t h, int64_t w,
                           int* launches_out) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(mask && dist && n > 0 && h > 0 && w > 0 && n <= 65535, "distance_transform: bad arguments");
    YAM_REQUIRE(h < (1 << 30) && w < (1 << 30), "distance_transform: image side too large");
    const int64_t count = n * h * w;
    dt_init_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(mask, dist, count);
    YAM_LAUNCHED(ctx);
    void* pin = nullptr;
    if (int rc = yam_pinned(ctx, 64, &pin)) return rc;
    int* flag_host = (int*)pin;
    void* scratch = nullptr;
    if (int rc = yam_scratch(ctx, 256, &scratch)) return rc;
    int* flag_dev = (int*)scratch;
    dim3 grid((unsigned)((w + kT - 1) / kT), (unsigned)((h + kT - 1) / kT), (unsigned)n);
    YAM_REQUIRE(grid.y <= 65535, "distance_transform: more than 65535 tile rows");
    int launches = 0;
    // a launch settles everything within ~32 px of its information; distances grow by at least 1 per
    // pixel, so h + w launches bound even a single zero pixel in a full frame
    const int64_t limit = (h + w) / kT + 4;
    while (true) {
        YAM_CUDA(cudaMemsetAsync(flag_dev, 0, sizeof(int), ctx->stream));
        dt_relax_kernel<<<grid, dim3(kT, kT), 0, ctx->stream>>>(dist, (int)h, (int)w, 2 * kT, flag_dev);
        YAM_LAUNCHED(ctx);
        launches++;
        YAM_CUDA(cudaMemcpyAsync(flag_host, flag_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        YAM_CUDA(cudaStreamSynchronize(ctx->stream));
        if (!*flag_host) break;
        YAM_REQUIRE(launches <= limit, "distance_transform: relaxation did not converge (%d launches)", launches);
    }
    if (launches_out) *launches_out = launches;
    return YAM_OK;
}

int yam_threshold_inv(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w, int dtype,
                      const int32_t* thresh_dev, double thresh, double maxval) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(src && dst && n > 0 && h > 0 && w > 0 && n <= 65535, "threshold_inv: bad arguments");
    const int64_t frame_px = h * w;
    int64_t bx = (frame_px + 256 * 8 - 1) / (256 * 8);
    const int64_t cap = (int64_t)ctx->num_sms * 8;
    if (bx > cap) bx = cap;
    dim3 grid((unsigned)bx, 1, (unsigned)n);
    const double m = rint(maxval);
    const float t = (float)floor(thresh);  // cv2: integer images compare against floor(thresh)
    if (dtype == YAM_U8)
        threshold_inv_kernel<uint8_t><<<grid, 256, 0, ctx->stream>>>((const uint8_t*)src, (uint8_t*)dst, frame_px, thresh_dev, t,
                                                                       (uint8_t)(m < 0 ? 0 : m > 255 ? 255 : m));
    else if (dtype == YAM_U16)
        threshold_inv_kernel<uint16_t><<<grid, 256, 0, ctx->stream>>>((const uint16_t*)src, (uint16_t*)dst, frame_px, thresh_dev, t,
                                                                         (uint16_t)(m < 0 ? 0 : m > 65535 ? 65535 : m));
    else
        YAM_REQUIRE(false, "threshold_inv: unsupported dtype %d", dtype);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

int yam_watershed_combine(yam_ctx* ctx, const int32_t* labels, const uint8_t* sure_bg, const uint8_t* sure_fg,
                          int32_t* markers, int64_t count) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(labels && sure_bg && sure_fg && markers && count > 0, "watershed_combine: bad arguments");
    int64_t bx = (count + 256 * 8 - 1) / (256 * 8);
    const int64_t cap = (int64_t)ctx->num_sms * 8;
    if (bx > cap) bx = cap;
    watershed_combine_kernel<<<(unsigned)bx, 256, 0, ctx->stream>>>(labels, sure_bg, sure_fg, markers, count);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

}  // extern "C"
