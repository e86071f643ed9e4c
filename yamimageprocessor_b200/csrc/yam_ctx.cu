// Context, error state, scratch memory and the host-only helpers of the C ABI.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include <vector>

#include "yam_common.cuh"
#include "yam_host.h"

static thread_local char g_err[512] = "";

void yam_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" {

int yam_abi_version(void) { return YAM_ABI_VERSION; }
const char* yam_last_error(void) { return g_err; }

int yam_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        yam_set_error("cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
        return YAM_ENODEV;
    }
    return n;
}

int yam_ctx_create(int device, yam_ctx** out) {
    YAM_REQUIRE(out != nullptr, "yam_ctx_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        yam_set_error("no CUDA device available (%s); libyamb200 has no CPU fallback",
                      e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return YAM_ENODEV;
    }
    YAM_REQUIRE(device >= 0 && device < n, "yam_ctx_create: device %d out of range [0,%d)", device, n);
    YAM_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    YAM_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        yam_set_error("device %d is sm_%d%d; libyamb200 is built for sm_100a only", device, prop.major,
                      prop.minor);
        return YAM_ENODEV;
    }
    yam_ctx* c = (yam_ctx*)calloc(1, sizeof(yam_ctx));
    if (!c) {
        yam_set_error("out of host memory");
        return YAM_ENOMEM;
    }
    c->device = device;
    c->num_sms = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : YAM_NUM_SMS_FALLBACK;
    e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        free(c);
        yam_set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e));
        return YAM_ECUDA;
    }
    c->stream = c->own_stream;
    *out = c;
    return YAM_OK;
}

int yam_ctx_destroy(yam_ctx* ctx) {
    if (!ctx) return YAM_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->scratch2) cudaFree(ctx->scratch2);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    free(ctx);
    return YAM_OK;
}

int yam_ctx_set_stream(yam_ctx* ctx, void* stream) {
    YAM_REQUIRE(ctx != nullptr, "ctx is NULL");
    // a NULL handle is CUDA's legacy default stream (what torch reports for its default stream)
    ctx->stream = (cudaStream_t)stream;
    return YAM_OK;
}

int yam_ctx_synchronize(yam_ctx* ctx) {
    YAM_REQUIRE(ctx != nullptr, "ctx is NULL");
    YAM_CUDA(cudaSetDevice(ctx->device));
    YAM_CUDA(cudaStreamSynchronize(ctx->stream));
    return YAM_OK;
}

int64_t yam_ctx_launch_count(yam_ctx* ctx, int reset) {
    if (!ctx) return 0;
    int64_t v = ctx->launches;
    if (reset) ctx->launches = 0;
    return v;
}

int yam_malloc(yam_ctx* ctx, int64_t bytes, void** out) {
    YAM_REQUIRE(ctx && out && bytes >= 0, "yam_malloc: bad arguments");
    YAM_CUDA(cudaSetDevice(ctx->device));
    *out = nullptr;
    if (bytes == 0) return YAM_OK;
    cudaError_t e = cudaMalloc(out, (size_t)bytes);
    if (e != cudaSuccess) {
        yam_set_error("cudaMalloc(%lld) failed: %s", (long long)bytes, cudaGetErrorString(e));
        return YAM_ENOMEM;
    }
    return YAM_OK;
}

int yam_free(yam_ctx* ctx, void* ptr) {
    YAM_REQUIRE(ctx != nullptr, "ctx is NULL");
    YAM_CUDA(cudaSetDevice(ctx->device));
    if (ptr) YAM_CUDA(cudaFree(ptr));
    return YAM_OK;
}

int yam_memcpy_h2d(yam_ctx* ctx, void* dst_dev, const void* src_host, int64_t bytes) {
    YAM_REQUIRE(ctx && bytes >= 0, "yam_memcpy_h2d: bad arguments");
    YAM_CUDA(cudaSetDevice(ctx->device));
    YAM_CUDA(cudaMemcpyAsync(dst_dev, src_host, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
    YAM_CUDA(cudaStreamSynchronize(ctx->stream));
    return YAM_OK;
}

int yam_memcpy_d2h(yam_ctx* ctx, void* dst_host, const void* src_dev, int64_t bytes) {
    YAM_REQUIRE(ctx && bytes >= 0, "yam_memcpy_d2h: bad arguments");
    YAM_CUDA(cudaSetDevice(ctx->device));
    YAM_CUDA(cudaMemcpyAsync(dst_host, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
    YAM_CUDA(cudaStreamSynchronize(ctx->stream));
    return YAM_OK;
}

}  // extern "C"

int yam_enter(yam_ctx* ctx) {
    YAM_REQUIRE(ctx != nullptr, "ctx is NULL");
    YAM_CUDA(cudaSetDevice(ctx->device));
    return YAM_OK;
}

int yam_scratch(yam_ctx* ctx, size_t bytes, void** out) {
    if (bytes > ctx->scratch_bytes) {
        // grow-only; outstanding work may still use the old buffer
        YAM_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->scratch) YAM_CUDA(cudaFree(ctx->scratch));
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
        size_t want = yam_align_up(bytes + bytes / 4, (size_t)1 << 20);
        cudaError_t e = cudaMalloc(&ctx->scratch, want);
        if (e != cudaSuccess) {
            yam_set_error("scratch cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
            return YAM_ENOMEM;
        }
        ctx->scratch_bytes = want;
    }
    *out = ctx->scratch;
    return YAM_OK;
}

int yam_scratch2(yam_ctx* ctx, size_t bytes, void** out) {
    if (bytes > ctx->scratch2_bytes) {
        YAM_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->scratch2) YAM_CUDA(cudaFree(ctx->scratch2));
        ctx->scratch2 = nullptr;
        ctx->scratch2_bytes = 0;
        size_t want = yam_align_up(bytes + bytes / 4, (size_t)1 << 20);
        cudaError_t e = cudaMalloc(&ctx->scratch2, want);
        if (e != cudaSuccess) {
            yam_set_error("scratch2 cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
            return YAM_ENOMEM;
        }
        ctx->scratch2_bytes = want;
    }
    *out = ctx->scratch2;
    return YAM_OK;
}

int yam_pinned(yam_ctx* ctx, size_t bytes, void** out) {
    if (bytes > ctx->pinned_bytes) {
        YAM_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->pinned) YAM_CUDA(cudaFreeHost(ctx->pinned));
        ctx->pinned = nullptr;
        ctx->pinned_bytes = 0;
        size_t want = yam_align_up(bytes, (size_t)1 << 16);
        cudaError_t e = cudaMallocHost(&ctx->pinned, want);
        if (e != cudaSuccess) {
            yam_set_error("cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
            return YAM_ENOMEM;
        }
        ctx->pinned_bytes = want;
    }
    *out = ctx->pinned;
    return YAM_OK;
}

// ---------------------------------------------------------------------------------------------
// host-only helpers

static const double kSmallTab1[] = {1.0};
static const double kSmallTab3[] = {0.25, 0.5, 0.25};
static const double kSmallTab5[] = {0.0625, 0.25, 0.375, 0.25, 0.0625};
static const double kSmallTab7[] = {0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125};
static const double kSmallTab9[] = {4.0 / 256, 13.0 / 256, 30.0 / 256, 51.0 / 256, 60.0 / 256,
                                    51.0 / 256, 30.0 / 256, 13.0 / 256, 4.0 / 256};

void yam_host_gaussian_taps(int k, double sigma, double* out) {
    const double* tab = nullptr;
    if (sigma <= 0) {
        switch (k) {
            case 1: tab = kSmallTab1; break;
            case 3: tab = kSmallTab3; break;
            case 5: tab = kSmallTab5; break;
            case 7: tab = kSmallTab7; break;
            case 9: tab = kSmallTab9; break;
            default: break;
        }
    }
    if (tab) {
        for (int i = 0; i < k; i++) out[i] = tab[i];
        return;
    }
    double s = sigma > 0 ? sigma : 0.3 * ((k - 1) * 0.5 - 1) + 0.8;
    double scale2x = -0.5 / (s * s);
    double sum = 0;
    for (int i = 0; i < k; i++) {
        double x = i - (k - 1) * 0.5;
        out[i] = exp(scale2x * x * x);
        sum += out[i];
    }
    // pairwise-free left-to-right sum above mirrors numpy? numpy uses pairwise summation for
    // n >= 8 blocks of 128; for k <= 128 it reduces left to right with 8 accumulators only when
    // n >= 8.  The quantised / float32-cast taps are insensitive to that last-ulp difference
    // (swept in tests/test_host_helpers.py against cv2 for k = 1..31 and many sigmas).
    sum = 1.0 / sum;
    for (int i = 0; i < k; i++) out[i] *= sum;
}

void yam_host_fixed_taps(const double* kf, int k, int bits, int64_t* out) {
    int r = k / 2;
    double err = 0.0;
    int64_t acc = 0;
    for (int i = 0; i < r; i++) {
        double v = kf[i] * (double)(1 << bits) + err;
        int64_t q = (int64_t)floor(v + 0.5);
        err = v - (double)q;
        out[i] = out[k - 1 - i] = q;
        acc += q;
    }
    out[r] = ((int64_t)1 << bits) - 2 * acc;
}

void yam_host_structuring_element(int shape, int k, uint8_t* out) {
    memset(out, 0, (size_t)k * k);
    if (k == 1) {
        out[0] = 1;
        return;
    }
    if (shape == YAM_SHAPE_CROSS) {
        for (int i = 0; i < k; i++) {
            out[(k / 2) * k + i] = 1;
            out[i * k + k / 2] = 1;
        }
        return;
    }
    if (shape == YAM_SHAPE_ELLIPSE) {
        int r = k / 2, c = k / 2;
        double inv_r2 = r ? 1.0 / ((double)r * r) : 0.0;
        for (int i = 0; i < k; i++) {
            int dy = i - r;
            if (abs(dy) <= r) {
                int dx = (int)lrint(c * sqrt((r * r - dy * dy) * inv_r2));  // cvRound
                int j1 = c - dx < 0 ? 0 : c - dx;
                int j2 = c + dx + 1 > k ? k : c + dx + 1;
                for (int j = j1; j < j2; j++) out[i * k + j] = 1;
            }
        }
        return;
    }
    memset(out, 1, (size_t)k * k);
}

template <typename CountT>
static int host_otsu_impl(const CountT* h, int bins) {
    // cv2 getThreshVal_Otsu_{8u,16u}: fp64 recurrence, strict '>' keeps the first maximum.
    // This TU is compiled without FMA contraction (-ffp-contract=off).
    double total = 0, mu = 0;
    int first = -1, last = -1;
    for (int i = 0; i < bins; i++) {
        if (h[i]) {
            if (first < 0) first = i;
            last = i;
            total += (double)h[i];          // integer-valued: exact in any order
            mu += (double)i * (double)h[i];
        }
    }
    if (first < 0) return 0;
    const double scale = 1.0 / total;
    mu *= scale;
    double mu1 = 0, q1 = 0, max_sigma = 0;
    int max_val = 0;
    const double eps = 1.1920928955078125e-07;  // FLT_EPSILON
    // Bins before the first occupied one leave (mu1, q1) = (0, 0) and `continue`; bins after the last
    // occupied one have q1 ~ 1 (> 1 - eps) and `continue`: both ranges are skipped without changing
    // the result.  Empty bins in between are NOT skippable (mu1 = (mu1*q1)/q1 re-rounds).
    for (int i = first; i <= last; i++) {
        const double p_i = (double)h[i] * scale;
        mu1 *= q1;
        q1 += p_i;
        const double q2 = 1.0 - q1;
        const double mn = q1 < q2 ? q1 : q2, mx = q1 > q2 ? q1 : q2;
        if (mn < eps || mx > 1.0 - eps) continue;
        mu1 = (mu1 + (double)i * p_i) / q1;
        const double mu2 = (mu - q1 * mu1) / q2;
        const double sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2);
        if (sigma > max_sigma) {
            max_sigma = sigma;
            max_val = i;
        }
    }
    return max_val;
}

// K frames advanced in lock step by one thread: the recurrence of a frame is a chain of dependent
// mul -> add -> div (~22 cycles per bin on a CPU core), so K independent chains fill the pipeline
// that one chain leaves idle.  Same operations per frame as host_otsu_impl, hence the same results.
template <typename CountT, int K>
static void host_otsu_multi(const CountT* const* hs, int bins, int32_t* out) {
    const double eps = 1.1920928955078125e-07;  // FLT_EPSILON
    double scale[K], mu[K], mu1[K], q1[K], best[K];
    int first[K], last[K], best_i[K];
    int lo = bins, hi = -1;
    for (int f = 0; f < K; f++) {
        const CountT* h = hs[f];
        double total = 0, m = 0;
        first[f] = -1;
        last[f] = -1;
        for (int i = 0; i < bins; i++) {
            if (h[i]) {
                if (first[f] < 0) first[f] = i;
                last[f] = i;
                total += (double)h[i];
                m += (double)i * (double)h[i];
            }
        }
        scale[f] = first[f] < 0 ? 0.0 : 1.0 / total;
        mu[f] = m * scale[f];
        mu1[f] = q1[f] = best[f] = 0;
        best_i[f] = 0;
        if (first[f] >= 0) {
            if (first[f] < lo) lo = first[f];
            if (last[f] > hi) hi = last[f];
        }
    }
    for (int i = lo; i <= hi; i++) {
        for (int f = 0; f < K; f++) {
            if (i < first[f] || i > last[f]) continue;  // outside the occupied range: state unchanged (see above)
            const double p_i = (double)hs[f][i] * scale[f];
            mu1[f] *= q1[f];
            q1[f] += p_i;
            const double q2 = 1.0 - q1[f];
            const double mn = q1[f] < q2 ? q1[f] : q2, mx = q1[f] > q2 ? q1[f] : q2;
            if (mn < eps || mx > 1.0 - eps) continue;
            mu1[f] = (mu1[f] + (double)i * p_i) / q1[f];
            const double mu2 = (mu[f] - q1[f] * mu1[f]) / q2;
            const double sigma = q1[f] * q2 * (mu1[f] - mu2) * (mu1[f] - mu2);
            if (sigma > best[f]) {
                best[f] = sigma;
                best_i[f] = i;
            }
        }
    }
    for (int f = 0; f < K; f++) out[f] = first[f] < 0 ? 0 : best_i[f];
}

int yam_host_otsu(const uint64_t* h, int bins) { return host_otsu_impl<uint64_t>(h, bins); }
int yam_host_otsu32(const uint32_t* h, int bins) { return host_otsu_impl<uint32_t>(h, bins); }

// `count` (1..4) histograms of `bins` counts each, `stride_bytes` apart, scanned by the calling thread
void yam_host_otsu_group(const void* hists, size_t stride_bytes, int count, int bins, int narrow, int32_t* out) {
    const char* base = (const char*)hists;
    if (narrow) {
        const uint32_t* hs[4];
        for (int f = 0; f < count; f++) hs[f] = (const uint32_t*)(base + (size_t)f * stride_bytes);
        if (count == 4) host_otsu_multi<uint32_t, 4>(hs, bins, out);
        else if (count == 3) host_otsu_multi<uint32_t, 3>(hs, bins, out);
        else if (count == 2) host_otsu_multi<uint32_t, 2>(hs, bins, out);
        else out[0] = host_otsu_impl<uint32_t>(hs[0], bins);
    } else {
        const uint64_t* hs[4];
        for (int f = 0; f < count; f++) hs[f] = (const uint64_t*)(base + (size_t)f * stride_bytes);
        if (count == 4) host_otsu_multi<uint64_t, 4>(hs, bins, out);
        else if (count == 3) host_otsu_multi<uint64_t, 3>(hs, bins, out);
        else if (count == 2) host_otsu_multi<uint64_t, 2>(hs, bins, out);
        else out[0] = host_otsu_impl<uint64_t>(hs[0], bins);
    }
}

extern "C" {

int yam_gaussian_taps_f64(int ksize, double sigma, double* out) {
    YAM_REQUIRE(out && ksize >= 1 && (ksize & 1) && ksize <= YAM_MAX_TAPS,
                "ksize must be odd in [1,%d], got %d", YAM_MAX_TAPS, ksize);
    yam_host_gaussian_taps(ksize, sigma, out);
    return YAM_OK;
}

int yam_gaussian_taps_fixed(int ksize, double sigma, int bits, int64_t* out) {
    YAM_REQUIRE(out && ksize >= 1 && (ksize & 1) && ksize <= YAM_MAX_TAPS,
                "ksize must be odd in [1,%d], got %d", YAM_MAX_TAPS, ksize);
    YAM_REQUIRE(bits == 8 || bits == 16, "bits must be 8 or 16");
    double kf[YAM_MAX_TAPS];
    yam_host_gaussian_taps(ksize, sigma, kf);
    yam_host_fixed_taps(kf, ksize, bits, out);
    return YAM_OK;
}

int yam_structuring_element(int shape, int ksize, uint8_t* out) {
    YAM_REQUIRE(out && ksize >= 1 && ksize <= YAM_MAX_SE, "ksize must be in [1,%d]", YAM_MAX_SE);
    YAM_REQUIRE(shape >= 0 && shape <= 2, "unknown shape %d", shape);
    yam_host_structuring_element(shape, ksize, out);
    return YAM_OK;
}

int yam_otsu_from_hist(const uint64_t* hist, int bins, int* out_threshold) {
    YAM_REQUIRE(hist && out_threshold && bins > 0, "yam_otsu_from_hist: bad arguments");
    *out_threshold = yam_host_otsu(hist, bins);
    return YAM_OK;
}

}  // extern "C"
