// K3 fast path: cv2.GaussianBlur on uint16 (fixed-point taps, bit-exact; oracle gaussian_fixed) for the
// shapes a TMA tensor map can describe.  Same skeleton as yam_adaptive.cu:
//   persistent CTA of 16 warps per SM, tiles of 240 x 80 outputs handed out round-robin, the raw tile of
//   the next-but-one tile in flight (two TMA buffers, one mbarrier each);
//   stage   raw uint16 tile + halo (256 px x (80 + 2r) rows) by ONE cp.async.bulk.tensor.3d; image-border
//           tiles rewrite the zero-filled out-of-range pixels with the BORDER_REFLECT_101 mirror;
//   H pass  warp = row pair, lane = 8 pixels of both rows in registers, r-pixel halos from the
//           neighbour lanes by SHFL, exact 32-bit integer sums (symmetric taps share one IMAD);
//   V pass  warp = 64 columns x 10 rows, lane = two adjacent columns; 48-bit sums on the fp64 pipe
//           (DFMA, exact below 2^53; (0x43300000 : t) is the double 2^52 + t, the rounded result is read
//           from the mantissa of acc + 2^52 + 2^31), results stored straight to global memory as
//           packed uint16 pairs (128 contiguous bytes per warp and row).
// Compared with sep_fixed_tiled (yam_filter.cu) the raw tile is never re-staged by the SM, the H pass
// reads each pixel from shared memory once, and there is no output staging tile.
#include <cuda.h>

#include "yam_common.cuh"
#include "yam_host.h"

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kBoxW = 256;   // staged pixels per row: 32 lanes x 8 px
constexpr int kOutW = 240;   // outputs per tile row: lanes 1..30 (lanes 0 / 31 only feed halos)
constexpr int kMargin = 8;   // left margin in pixels (16 bytes: TMA box origin alignment)
constexpr int kTP = 240;     // pitch of the intermediate tile in 32-bit words
constexpr int kTH = 80;      // output rows per tile
constexpr int kRB = 10;      // rows per vertical-pass item: 8 row blocks x 4 column groups = 2 items per warp

struct GaussTaps {
    uint32_t q[16];   // fixed-point taps (sum = 2^16)
    double d[16];     // the same values as doubles for the fp64 vertical pass
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    unsigned long long spins = 0;
    while (!done) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!done && ++spins > (1ull << 26)) __trap();  // a copy that never lands must not hang the device
    }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int KS>
__global__ void __launch_bounds__(kThreads, 1)
gauss16_tma_kernel(const __grid_constant__ CUtensorMap tmap, uint16_t* __restrict__ dst, int h, int w, GaussTaps taps,
                   int border, int tiles_x, int tiles_y, int total_tiles) {
    constexpr int R = KS / 2, ROWS = kTH + 2 * R, RP = ROWS / 2;
    constexpr int ROUNDS = (RP + kWarps - 1) / kWarps;
    static_assert(R <= 8 && ROWS % 2 == 0 && ROUNDS <= 3, "tile geometry");
    constexpr size_t RAW_BYTES = (size_t)ROWS * kBoxW * sizeof(uint16_t);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t* s_t = reinterpret_cast<uint32_t*>(smem_raw + 2 * RAW_BYTES);   // [ROWS][240] row sums (< 2^32)
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + 2 * RAW_BYTES + (size_t)ROWS * kTP * 4);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    auto issue = [&](int k) {
        const int tile = blockIdx.x + k * gridDim.x;
        if (tile >= total_tiles) return;
        const int tx = tile % tiles_x, rest = tile / tiles_x;
        const int ty = rest % tiles_y, fr = rest / tiles_y;
        mbar_expect_tx(bar + (k & 1), (uint32_t)RAW_BYTES);
        tma_load_3d(smem_raw + (k & 1) * RAW_BYTES, &tmap, bar + (k & 1), tx * kOutW - kMargin, ty * kTH - R, fr);
    };
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
    }
    __syncthreads();
    if (tid == 0) {
        issue(0);
        issue(1);
    }

    int tx = (int)(blockIdx.x % tiles_x), ty = (int)((blockIdx.x / tiles_x) % tiles_y), frame = (int)(blockIdx.x / (tiles_x * tiles_y));
    for (int k = 0;; k++) {
        const int tile = blockIdx.x + k * gridDim.x;
        if (tile >= total_tiles) break;
        const int x0 = tx * kOutW, y0 = ty * kTH;
        const int bx0 = x0 - kMargin, by0 = y0 - R;
        uint16_t* s_raw = reinterpret_cast<uint16_t*>(smem_raw + (k & 1) * RAW_BYTES);   // [ROWS][256]
        mbar_wait(bar + (k & 1), (k >> 1) & 1);

        // ---- image border: TMA zero-fills out-of-range pixels; rewrite them with cv2's border mapping.
        // Sources are always in-range pixels of this box (frames are at least one box wide and high, so a
        // mirror never leaves it): columns of the in-range rows first, then whole out-of-range rows.
        if (bx0 < 0 || by0 < 0 || bx0 + kBoxW > w || by0 + ROWS > h) {
            const int cl = max(0, -bx0), cr = min(kBoxW, w - bx0);
            const int rt = max(0, -by0), rbm = min(ROWS, h - by0);
            const int ncol = cl + (kBoxW - cr);
            for (int i = tid; i < (rbm - rt) * ncol; i += kThreads) {
                const int ry = rt + i / ncol, kk = i - (i / ncol) * ncol;
                const int cx = kk < cl ? kk : cr + (kk - cl);
                // only the R pixels next to the image feed valid outputs; mirrors of farther ones may fall outside
                // the box and are clamped to it (their values are never used)
                const int sx = min(max(yam_border(bx0 + cx, w, border) - bx0, cl), cr - 1);
                s_raw[ry * kBoxW + cx] = s_raw[ry * kBoxW + sx];
            }
            __syncthreads();
            const int nrow = rt + (ROWS - rbm);
            for (int i = tid; i < nrow * kBoxW; i += kThreads) {
                const int kk = i / kBoxW, cx = i - kk * kBoxW;
                const int ry = kk < rt ? kk : rbm + (kk - rt);
                const int sy = min(max(yam_border(by0 + ry, h, border) - by0, rt), rbm - 1);
                s_raw[ry * kBoxW + cx] = s_raw[sy * kBoxW + cx];
            }
            __syncthreads();
        }

        // ---- horizontal pass (exact, modulo 2^32; the true sums are < 2^32)
        {
            const int m = lane - 1;
            const int sw = (m >> 2) & 1;   // pair-swapped 16-byte chunks in every second 128-byte group (bank conflicts)
            const int off0 = ((2 * m) ^ sw) * 4, off1 = ((2 * m + 1) ^ sw) * 4;
            const bool stores = m >= 0 && m < kOutW / 8;
#pragma unroll
            for (int it = 0; it < ROUNDS; it++) {
                const int rp = min(warp + it * kWarps, RP - 1);
                const uint16_t* ra = s_raw + (2 * rp) * kBoxW + 8 * lane;
                const uint4 A = *reinterpret_cast<const uint4*>(ra), B = *reinterpret_cast<const uint4*>(ra + kBoxW);
                const uint32_t aw[4] = {A.x, A.y, A.z, A.w}, bw[4] = {B.x, B.y, B.z, B.w};
                uint32_t ea[8 + 2 * R], eb[8 + 2 * R];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    ea[R + i] = (i & 1) ? (aw[i >> 1] >> 16) : (aw[i >> 1] & 0xffffu);
                    eb[R + i] = (i & 1) ? (bw[i >> 1] >> 16) : (bw[i >> 1] & 0xffffu);
                }
#pragma unroll
                for (int i = 0; i < R; i++) {
                    ea[i] = __shfl_up_sync(0xffffffffu, ea[R + 8 - R + i], 1);
                    eb[i] = __shfl_up_sync(0xffffffffu, eb[R + 8 - R + i], 1);
                    ea[R + 8 + i] = __shfl_down_sync(0xffffffffu, ea[R + i], 1);
                    eb[R + 8 + i] = __shfl_down_sync(0xffffffffu, eb[R + i], 1);
                }
                uint32_t oa[8], ob[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    uint32_t sa = taps.q[R] * ea[R + j], sb = taps.q[R] * eb[R + j];
#pragma unroll
                    for (int i = 1; i <= R; i++) {
                        sa += taps.q[R + i] * (ea[R + j - i] + ea[R + j + i]);
                        sb += taps.q[R + i] * (eb[R + j - i] + eb[R + j + i]);
                    }
                    oa[j] = sa;
                    ob[j] = sb;
                }
                if (stores) {
                    uint32_t* ta = s_t + (2 * rp) * kTP;
                    uint32_t* tb = ta + kTP;
                    *reinterpret_cast<uint4*>(ta + off0) = make_uint4(oa[0], oa[1], oa[2], oa[3]);
                    *reinterpret_cast<uint4*>(ta + off1) = make_uint4(oa[4], oa[5], oa[6], oa[7]);
                    *reinterpret_cast<uint4*>(tb + off0) = make_uint4(ob[0], ob[1], ob[2], ob[3]);
                    *reinterpret_cast<uint4*>(tb + off1) = make_uint4(ob[4], ob[5], ob[6], ob[7]);
                }
            }
        }
        __syncthreads();

        // ---- vertical pass on the fp64 pipe: 8 row blocks x 4 column groups, two items per warp
        static_assert((kTH / kRB) * 4 == 2 * kWarps, "two V-pass items per warp");
        uint16_t* out = dst + (int64_t)frame * h * w;
#pragma unroll 1
        for (int it = 0; it < 2; it++) {
            const int item = warp + it * kWarps;
            const int rbk = item >> 2, cg = item & 3;
            const int r0 = rbk * kRB;
            const int col = 64 * cg + 2 * lane;                       // first of this lane's two columns
            const bool live = col < kOutW;
            const int ccol = live ? col : 0;
            const int phys = (((ccol >> 2) ^ ((ccol >> 5) & 1)) << 2) | (ccol & 3);
            const uint32_t* tp = s_t + r0 * kTP + phys;
            double a0[kRB], a1[kRB];
#pragma unroll
            for (int j = 0; j < kRB; j++) a0[j] = a1[j] = 0.0;
#pragma unroll
            for (int i = 0; i < kRB + 2 * R; i++) {
                const uint2 tv = *reinterpret_cast<const uint2*>(tp + i * kTP);
                const double t0 = __dsub_rn(__hiloint2double(0x43300000, (int)tv.x), 4503599627370496.0);
                const double t1 = __dsub_rn(__hiloint2double(0x43300000, (int)tv.y), 4503599627370496.0);
#pragma unroll
                for (int j = 0; j < kRB; j++) {
                    const int kk = i - j;
                    if (kk >= 0 && kk < KS) {
                        a0[j] = __fma_rn(taps.d[kk], t0, a0[j]);
                        a1[j] = __fma_rn(taps.d[kk], t1, a1[j]);
                    }
                }
            }
            const int gx = x0 + col;
            if (live && gx < w) {
                uint32_t* op = reinterpret_cast<uint32_t*>(out + (int64_t)(y0 + r0) * w + gx);
                const int rows = min(kRB, h - (y0 + r0));
#pragma unroll
                for (int j = 0; j < kRB; j++) {
                    const uint32_t o0 = (uint32_t)__double2hiint(__dadd_rn(a0[j], 4503599627370496.0 + 2147483648.0)) & 0xffffu;
                    const uint32_t o1 = (uint32_t)__double2hiint(__dadd_rn(a1[j], 4503599627370496.0 + 2147483648.0)) & 0xffffu;
                    if (j < rows) op[(int64_t)j * (w / 2)] = o0 | (o1 << 16);
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            fence_proxy_async();   // the buffer was touched through the generic proxy (reads, border fix-up)
            issue(k + 2);
        }
        tx += (int)gridDim.x;
        while (tx >= tiles_x) {
            tx -= tiles_x;
            if (++ty == tiles_y) {
                ty = 0;
                frame++;
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            (void)cudaGetLastError();
    }
    return fn;
}

template <int KS>
int launch(yam_ctx* ctx, const uint16_t* src, uint16_t* dst, int64_t n, int64_t h, int64_t w, const GaussTaps& taps, int border) {
    constexpr int ROWS = kTH + 2 * (KS / 2);
    CUtensorMap map;
    const cuuint64_t gdim[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    const cuuint64_t gstride[2] = {(cuuint64_t)w * 2, (cuuint64_t)w * h * 2};
    const cuuint32_t box[3] = {kBoxW, (cuuint32_t)ROWS, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult rc = encode_tiled()(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<uint16_t*>(src), gdim, gstride, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        yam_set_error("gaussian: cuTensorMapEncodeTiled failed (%d)", (int)rc);
        return YAM_ECUDA;
    }
    constexpr size_t smem = 2 * (size_t)ROWS * kBoxW * 2 + (size_t)ROWS * kTP * 4 + 16;
    YAM_CUDA(cudaFuncSetAttribute(gauss16_tma_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles_x = (int)((w + kOutW - 1) / kOutW), tiles_y = (int)((h + kTH - 1) / kTH);
    const int64_t total = (int64_t)tiles_x * tiles_y * n;
    if (total >= (1ll << 31)) {
        yam_set_error("gaussian: too many tiles");
        return YAM_EINVAL;
    }
    const unsigned grid = (unsigned)(total < ctx->num_sms ? total : ctx->num_sms);
    gauss16_tma_kernel<KS><<<grid, kThreads, smem, ctx->stream>>>(map, dst, (int)h, (int)w, taps, border, tiles_x, tiles_y, (int)total);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

}  // namespace

// *handled = 1 when the TMA kernel took the call (uint16, odd ksize 3..15, rows of 16-byte multiples from a
// 16-byte aligned base, frames at least one box wide and high); otherwise the caller uses sep_fixed_tiled.
int yam_gauss16_tma(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w, int ksize,
                    const uint32_t* taps_q, int border, int* handled) {
    *handled = 0;
    if (ksize < 3 || ksize > 15 || !(ksize & 1)) return YAM_OK;
    if ((w % 8) || ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) || w < kBoxW || h < 96 || n > 65535)
        return YAM_OK;
    const char* legacy = getenv("YAM_GAUSS_LEGACY");
    if (legacy && legacy[0] == '1') return YAM_OK;
    if (!encode_tiled()) return YAM_OK;
    GaussTaps taps;
    for (int i = 0; i < 16; i++) {
        taps.q[i] = i < ksize ? taps_q[i] : 0u;
        taps.d[i] = (double)taps.q[i];
    }
    int rc = YAM_EINVAL;
    const uint16_t* s = (const uint16_t*)src;
    uint16_t* d = (uint16_t*)dst;
    switch (ksize) {
        case 3: rc = launch<3>(ctx, s, d, n, h, w, taps, border); break;
        case 5: rc = launch<5>(ctx, s, d, n, h, w, taps, border); break;
        case 7: rc = launch<7>(ctx, s, d, n, h, w, taps, border); break;
        case 9: rc = launch<9>(ctx, s, d, n, h, w, taps, border); break;
        case 11: rc = launch<11>(ctx, s, d, n, h, w, taps, border); break;
        case 13: rc = launch<13>(ctx, s, d, n, h, w, taps, border); break;
        case 15: rc = launch<15>(ctx, s, d, n, h, w, taps, border); break;
    }
    if (rc == YAM_OK) *handled = 1;
    return rc;
}
