// K6 fast path: rectangular, symmetric (odd k) erode / dilate / open / close / open+close with
// compile-time radii.  Same semantics as the generic chain kernel in yam_morph.cu (constant border
// = identity element, out-of-image intermediates = identity of the next stage), restructured for
// throughput:
//   * pixels live in shared memory as packed u16x2 words; all min/max are native VIMNMX3.U16x2
//     (three operands, two pixels per instruction); odd pixel shifts are one PRMT.
//   * each pass is register tiled: the horizontal pass produces 4 words (8 px) per item from one
//     128-bit + two margin loads, the vertical pass produces 8 rows x 2 words per item from a
//     sliding window held in registers.
//   * every pass only computes the region the later stages still need (compile-time extents).
// One launch per chain: 1 read + 1 write of the image in HBM.
#include "yam_common.cuh"

namespace yam_morph_fast {

constexpr int kThreads = 256;
constexpr int TW = 128;  // tile width in pixels (64 packed words)
constexpr int TH = 64;
constexpr int PADW = 4;  // spare words on both sides of every buffer row (group-aligned margin loads)

template <int DIL>
__device__ __forceinline__ uint32_t mm3(uint32_t a, uint32_t b, uint32_t c) {
    return DIL ? __vimax3_u16x2(a, b, c) : __vimin3_u16x2(a, b, c);
}

// ---- horizontal pass: rows [row0, row0+NROWS), word groups [G0, G0+NG) of 4 words ------------------
// Margin words come from the neighbouring lanes by shuffle (adjacent lanes hold adjacent groups of
// the same row); only lanes at a row start / warp edge load them.  Scalar margin loads at a 16-byte
// lane stride were 4-way bank conflicts and dominated the shared-memory wavefronts.
template <int R, int DIL, int NROWS, int G0, int NG, int STRIDE>
__device__ __forceinline__ void hpass(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int row0) {
    constexpr int C = (R + 1) / 2;  // margin words on each side (1 or 2)
    static_assert(C <= 2, "fast path covers radii up to 4");
    constexpr int NW = 4 + 2 * C;
    constexpr int TOTAL = NROWS * NG;
    const int lane = threadIdx.x & 31;
    for (int item0 = 0; item0 < TOTAL; item0 += kThreads) {  // warp-uniform trip count (shuffles inside)
        const int item = item0 + threadIdx.x;
        const bool valid = item < TOTAL;
        const int it = valid ? item : TOTAL - 1;
        const int r = it / NG, g = it - r * NG;
        const int base = (row0 + r) * STRIDE + PADW + 4 * (G0 + g);
        uint32_t w[NW];
        const uint4 mid = *reinterpret_cast<const uint4*>(src + base);
        w[C] = mid.x; w[C + 1] = mid.y; w[C + 2] = mid.z; w[C + 3] = mid.w;
        // left margin = last C words of the previous group, right margin = first C words of the next
        uint32_t lft[2], rgt[2];
        lft[0] = __shfl_up_sync(0xffffffffu, mid.w, 1);
        rgt[0] = __shfl_down_sync(0xffffffffu, mid.x, 1);
        if (C == 2) {
            lft[1] = lft[0];
            lft[0] = __shfl_up_sync(0xffffffffu, mid.z, 1);
            rgt[1] = __shfl_down_sync(0xffffffffu, mid.y, 1);
        }
        if (lane == 0 || g == 0) {
#pragma unroll
            for (int i = 0; i < C; i++) lft[i] = src[base - C + i];
        }
        if (lane == 31 || g == NG - 1 || item + 1 >= TOTAL) {
#pragma unroll
            for (int i = 0; i < C; i++) rgt[i] = src[base + 4 + i];
        }
#pragma unroll
        for (int i = 0; i < C; i++) {
            w[i] = lft[i];
            w[C + 4 + i] = rgt[i];
        }
        uint32_t s[NW - 1];  // s[i] = pixels (hi of w[i], lo of w[i+1]) : shift by one pixel
#pragma unroll
        for (int i = 0; i < NW - 1; i++) s[i] = __byte_perm(w[i], w[i + 1], 0x5432);
        uint32_t out[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int c = C + k;
            // operands for pixel offsets -R..R: even offset 2d -> w[c+d]; odd offset 2m+1 -> s[c+m]
            uint32_t ops[2 * R + 1];
#pragma unroll
            for (int o = -R; o <= R; o++) ops[o + R] = (o & 1) ? s[c + ((o - 1) >> 1)] : w[c + (o >> 1)];
            uint32_t acc = ops[0];
#pragma unroll
            for (int i = 0; i < R; i++) acc = mm3<DIL>(acc, ops[2 * i + 1], ops[2 * i + 2]);
            out[k] = acc;
        }
        if (valid) *reinterpret_cast<uint4*>(dst + base) = make_uint4(out[0], out[1], out[2], out[3]);
    }
}

// ---- vertical pass: output rows [row0, row0+NROWS), word pairs [P0, P0+NP) -------------------------
// FIX = 1: positions outside the image are overwritten with `next_id` (identity of the next stage)
template <int R, int DIL, int NROWS, int P0, int NP, int STRIDE, int FIX>
__device__ __forceinline__ void vpass(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int row0,
                                      int gy_org, int gx_org, int h, int w, uint32_t next_id) {
    constexpr int RG = 8;
    constexpr int NGRP = (NROWS + RG - 1) / RG;
    for (int item = threadIdx.x; item < NGRP * NP; item += kThreads) {
        const int rg = item / NP, p = item - rg * NP;
        const int r_first = row0 + rg * RG;
        const int col = PADW + 2 * (P0 + p);
        uint2 win[RG + 2 * R];
#pragma unroll
        for (int i = 0; i < RG + 2 * R; i++) {
            // rows beyond the computed range of the last group are clamped (their outputs are dropped)
            const int rr = min(r_first - R + i, row0 + NROWS - 1 + R);
            win[i] = *reinterpret_cast<const uint2*>(src + rr * STRIDE + col);
        }
#pragma unroll
        for (int j = 0; j < RG; j++) {
            if (rg * RG + j >= NROWS) break;
            uint32_t ax = win[j].x, ay = win[j].y;
#pragma unroll
            for (int i = 0; i < R; i++) {
                ax = mm3<DIL>(ax, win[j + 2 * i + 1].x, win[j + 2 * i + 2].x);
                ay = mm3<DIL>(ay, win[j + 2 * i + 1].y, win[j + 2 * i + 2].y);
            }
            const int by = r_first + j;
            if (FIX) {
                const int gy = gy_org + by;
                const int gx = gx_org + 4 * (P0 + p);  // first pixel of the word pair
                if ((unsigned)gy >= (unsigned)h) {
                    ax = ay = next_id | (next_id << 16);
                } else if (gx < 0 || gx + 3 >= w) {
                    uint32_t px[4] = {ax & 0xffffu, ax >> 16, ay & 0xffffu, ay >> 16};
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if ((unsigned)(gx + q) >= (unsigned)w) px[q] = next_id;
                    ax = px[0] | (px[1] << 16);
                    ay = px[2] | (px[3] << 16);
                }
            }
            *reinterpret_cast<uint2*>(dst + by * STRIDE + col) = make_uint2(ax, ay);
        }
    }
}

template <int R, int DIL, int REM, int H, int HA, int STRIDE, int LAST>
__device__ __forceinline__ void stage(uint32_t* bufA, uint32_t* bufB, int gy_org, int gx_org, int h, int w,
                                      bool border_tile, uint32_t next_id) {
    // columns still needed after this stage: pixels [HA-REM, HA+TW+REM)
    constexpr int C0W = (HA - REM) / 2, C1W = (HA + TW + REM + 1) / 2;  // word range
    constexpr int G0 = C0W / 4, G1 = (C1W + 3) / 4;
    constexpr int P0 = C0W / 2, P1 = (C1W + 1) / 2;
    constexpr int HR0 = H - REM - R, HNR = TH + 2 * (REM + R);  // rows the vertical pass will read
    constexpr int VR0 = H - REM, VNR = TH + 2 * REM;
    hpass<R, DIL, HNR, G0, G1 - G0, STRIDE>(bufA, bufB, HR0);
    __syncthreads();
    if (!LAST && border_tile)
        vpass<R, DIL, VNR, P0, P1 - P0, STRIDE, 1>(bufB, bufA, VR0, gy_org, gx_org, h, w, next_id);
    else
        vpass<R, DIL, VNR, P0, P1 - P0, STRIDE, 0>(bufB, bufA, VR0, gy_org, gx_org, h, w, next_id);
    __syncthreads();
}

// OPS bit i = 1 -> stage i dilates.  R2 / R3 = 0 -> stage absent.
template <typename T, int R1, int R2, int R3, int OPS>
__global__ void __launch_bounds__(kThreads) morph_fast_kernel(const T* __restrict__ src, T* __restrict__ dst,
                                                              int h, int w) {
    constexpr int VEC = 16 / sizeof(T);
    constexpr int H = R1 + R2 + R3;
    constexpr int HA = (H + VEC - 1) / VEC * VEC;
    constexpr int BWW = (TW + 2 * HA) / 2;
    constexpr int STRIDE = BWW + 2 * PADW;
    constexpr int BH = TH + 2 * H;
    constexpr int D1 = OPS & 1, D2 = (OPS >> 1) & 1, D3 = (OPS >> 2) & 1;
    extern __shared__ __align__(16) uint32_t smem_words[];
    uint32_t* bufA = smem_words;
    uint32_t* bufB = smem_words + BH * STRIDE;

    src += (int64_t)blockIdx.z * h * w;
    dst += (int64_t)blockIdx.z * h * w;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int gx_org = x0 - HA, gy_org = y0 - H;
    const bool border_tile = (gx_org < 0) || (gy_org < 0) || (x0 + TW + HA > w) || (y0 + TH + H > h);

    // ---- load (outside the image = identity of stage 1); 8 pixels = 4 packed words per thread-item
    {
        constexpr int UPR = (TW + 2 * HA) / 8;   // 8-pixel units per row
        const uint32_t id1 = D1 ? 0u : 0xffffu;
        const uint32_t idw = id1 | (id1 << 16);
        const bool row_aligned = ((w % 8) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
        for (int v = threadIdx.x; v < BH * UPR; v += kThreads) {
            const int by = v / UPR, ux = v - by * UPR;
            const int gy = gy_org + by, gx = gx_org + ux * 8;
            uint32_t wd[4];
            if ((unsigned)gy >= (unsigned)h) {
                wd[0] = wd[1] = wd[2] = wd[3] = idw;
            } else if (row_aligned && gx >= 0 && gx + 8 <= w) {
                const T* p = src + (int64_t)gy * w + gx;
                if (sizeof(T) == 2) {
                    const uint4 q = *reinterpret_cast<const uint4*>(p);
                    wd[0] = q.x; wd[1] = q.y; wd[2] = q.z; wd[3] = q.w;
                } else {
                    const uint2 q = *reinterpret_cast<const uint2*>(p);
                    wd[0] = __byte_perm(q.x, 0, 0x4140); wd[1] = __byte_perm(q.x, 0, 0x4342);
                    wd[2] = __byte_perm(q.y, 0, 0x4140); wd[3] = __byte_perm(q.y, 0, 0x4342);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int xa = gx + 2 * i, xb = xa + 1;
                    const uint32_t a = (unsigned)xa < (unsigned)w ? (uint32_t)src[(int64_t)gy * w + xa] : id1;
                    const uint32_t b = (unsigned)xb < (unsigned)w ? (uint32_t)src[(int64_t)gy * w + xb] : id1;
                    wd[i] = a | (b << 16);
                }
            }
            *reinterpret_cast<uint4*>(bufA + by * STRIDE + PADW + ux * 4) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
        }
    }
    __syncthreads();

    stage<R1, D1, R2 + R3, H, HA, STRIDE, (R2 == 0)>(bufA, bufB, gy_org, gx_org, h, w, border_tile, D2 ? 0u : 0xffffu);
    if (R2 > 0)
        stage<(R2 > 0 ? R2 : 1), D2, R3, H, HA, STRIDE, (R3 == 0)>(bufA, bufB, gy_org, gx_org, h, w, border_tile,
                                                                  D3 ? 0u : 0xffffu);
    if (R3 > 0)
        stage<(R3 > 0 ? R3 : 1), D3, 0, H, HA, STRIDE, 1>(bufA, bufB, gy_org, gx_org, h, w, border_tile, 0u);

    // ---- store the core, 8 pixels per thread-item
    {
        constexpr int UPR = TW / 8;
        const bool row_aligned = ((w % 8) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
        for (int v = threadIdx.x; v < TH * UPR; v += kThreads) {
            const int ty = v / UPR, ux = v - ty * UPR;
            const int gy = y0 + ty, gx = x0 + ux * 8;
            if (gy >= h || gx >= w) continue;
            const uint4 q = *reinterpret_cast<const uint4*>(bufA + (H + ty) * STRIDE + PADW + HA / 2 + ux * 4);
            T* d = dst + (int64_t)gy * w + gx;
            if (row_aligned && gx + 8 <= w) {
                if (sizeof(T) == 2) {
                    *reinterpret_cast<uint4*>(d) = q;
                } else {
                    // low bytes of the four packed pixels pairs -> 8 bytes
                    const uint32_t lo = __byte_perm(q.x, q.y, 0x6420), hi = __byte_perm(q.z, q.w, 0x6420);
                    *reinterpret_cast<uint2*>(d) = make_uint2(lo, hi);
                }
            } else {
                const uint32_t wd[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (gx + 2 * i < w) d[2 * i] = (T)(wd[i] & 0xffffu);
                    if (gx + 2 * i + 1 < w) d[2 * i + 1] = (T)(wd[i] >> 16);
                }
            }
        }
    }
}

template <typename T, int R1, int R2, int R3>
constexpr size_t smem_bytes() {
    constexpr int VEC = 16 / sizeof(T);
    constexpr int H = R1 + R2 + R3;
    constexpr int HA = (H + VEC - 1) / VEC * VEC;
    constexpr int STRIDE = (TW + 2 * HA) / 2 + 2 * PADW;
    return (size_t)2 * (TH + 2 * H) * STRIDE * 4;
}

template <typename T, int R1, int R2, int R3, int OPS>
int launch(yam_ctx* ctx, const T* src, T* dst, int64_t n, int64_t h, int64_t w) {
    constexpr size_t smem = smem_bytes<T, R1, R2, R3>();
    auto kern = morph_fast_kernel<T, R1, R2, R3, OPS>;
    if (smem > 48 * 1024) YAM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((w + TW - 1) / TW), (unsigned)((h + TH - 1) / TH), (unsigned)n);
    kern<<<grid, kThreads, smem, ctx->stream>>>(src, dst, (int)h, (int)w);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

}  // namespace yam_morph_fast

// Try the fast path for a chain of up to 3 symmetric stages; returns 1 if it launched, 0 if the
// chain is not covered (caller falls back to the generic kernel), negative on error.
template <typename T>
int yam_morph_fast_try(yam_ctx* ctx, const T* src, T* dst, int64_t n, int64_t h, int64_t w, int nstages,
                       const int* dilate, const int* radius) {
    using namespace yam_morph_fast;
    int ops = 0;
    for (int i = 0; i < nstages; i++) ops |= (dilate[i] ? 1 : 0) << i;
    int rc = 0;
#define YAM_TRY(R1, R2, R3, OPS)                                                      \
    if (!rc && nstages == ((R1) > 0) + ((R2) > 0) + ((R3) > 0) && radius[0] == (R1) && \
        (nstages < 2 || radius[1] == (R2)) && (nstages < 3 || radius[2] == (R3)) && ops == (OPS)) { \
        rc = launch<T, R1, R2, R3, OPS>(ctx, src, dst, n, h, w);                       \
        return rc ? rc : 1;                                                            \
    }
    // single erode / dilate, k = 3, 5, 7
    YAM_TRY(1, 0, 0, 0) YAM_TRY(1, 0, 0, 1) YAM_TRY(2, 0, 0, 0) YAM_TRY(2, 0, 0, 1)
    YAM_TRY(3, 0, 0, 0) YAM_TRY(3, 0, 0, 1)
    // open (E,D) / close (D,E), k = 3, 5
    YAM_TRY(1, 1, 0, 2) YAM_TRY(1, 1, 0, 1) YAM_TRY(2, 2, 0, 2) YAM_TRY(2, 2, 0, 1)
    // open followed by close: E, D(2r), E
    YAM_TRY(1, 2, 1, 2) YAM_TRY(2, 4, 2, 2)
#undef YAM_TRY
    return 0;
}

template int yam_morph_fast_try<uint8_t>(yam_ctx*, const uint8_t*, uint8_t*, int64_t, int64_t, int64_t, int,
                                         const int*, const int*);
template int yam_morph_fast_try<uint16_t>(yam_ctx*, const uint16_t*, uint16_t*, int64_t, int64_t, int64_t, int,
                                          const int*, const int*);
