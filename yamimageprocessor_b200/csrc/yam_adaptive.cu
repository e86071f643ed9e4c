// K9 fast path: Gaussian adaptive threshold -> packed 1-bit mask (cv2.adaptiveThreshold GAUSSIAN_C,
// core/segmentation.py:91-94; float32 blur in cv2 4.13's summation order, oracle adaptive_threshold).
//
// Design: PERSISTENT kernel, one CTA of 16 warps per SM, tiles of 240 (16-bit) or 224 (8-bit) x TH
// outputs handed out round-robin; the raw tile of the NEXT-BUT-ONE tile is in flight while a tile is
// computed (two TMA buffers, one mbarrier each), so the load latency never shows.
//   stage   the RAW integer tile + halo (256 px x (TH + 2r) rows) lands in shared memory by ONE TMA
//           tensor copy (cp.async.bulk.tensor.3d, zero issue slots); only image-border tiles touch
//           it again (replicate fix-up of the zero-filled out-of-range pixels).
//   H pass  a warp owns a pair of rows; lane l holds 8 pixels of both rows as packed float2
//           {row a, row b} (PRMT + FADD2 conversion), gets its r-pixel halos from the neighbour lanes
//           by SHFL, and runs the row recurrence on the packed fp32 pipe (FFMA2: both halves are IEEE
//           fp32 operations, so the result is bit-identical to scalar code).  Lanes 0 and 31 only
//           feed halos; lanes 1..30 write 8 x 2 results to the intermediate tile.
//   V pass  a warp owns 64 columns x RB rows; lane l holds columns (l, l + 32) of RB + 2r
//           intermediate rows in registers (packed), runs the column recurrence (centre, then
//           symmetric pairs) with FMUL2 / FADD2 / FFMA2, rounds with the 1.5 * 2^23 magic add and
//           compares IN THE INTEGER DOMAIN against the raw pixel (src - rint(mean) > -C  <=>
//           bits(mean + 1.5 * 2^23) < src + 0x4B400000 + C); two ballots give two finished mask
//           words per row, stored as 16-bit halves (tiles are 240 px wide, i.e. 16-bit aligned).
// Compared with sep_f32_tiled (yam_filter.cu) this moves ~3x fewer shared-memory bytes per pixel
// (raw pixels instead of floats, no staged byte tile, no pack pass) and halves the H-pass issue
// slots.  HBM traffic is unchanged: 1 read of the pixels + 1/8 byte per pixel written.
#include <cuda.h>  // CUtensorMap types; the encoder itself is resolved at run time (no libcuda link)

#include "yam_common.cuh"
#include "yam_host.h"

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kBoxW = 256;  // staged pixels per row: 32 lanes x 8 px
// The innermost TMA box coordinate must fall on a 16-byte boundary, so the left margin is 8 px for
// 16-bit pixels (lane 0 = left halo lane, outputs from lanes 1..30 = 240 px) and 16 px for 8-bit pixels
// (lanes 0-1 / 30-31 are halo lanes, outputs from lanes 2..29 = 224 px).
template <typename Tin>
struct LaneGeom {
    static constexpr int LM = sizeof(Tin) == 2 ? 1 : 2;         // left-margin lanes
    static constexpr int OUTW = sizeof(Tin) == 2 ? 240 : 224;   // outputs per tile row
    static constexpr int MARGIN = 8 * LM;
    static constexpr int UNITS = OUTW / 16;                     // finished 16-bit mask units per tile row
};
constexpr int kTP = 240;    // pitch of the intermediate tile in floats

struct Taps16 {
    float v[16];
};

template <int KS>
struct TileGeom {
    static constexpr int R = KS / 2;
    // at most 94 staged rows = 47 row pairs = 3 H-pass rounds of 16 warps; TH a multiple of 4 so the
    // V pass is 4 row blocks x 4 column groups = one item per warp
    static constexpr int TH = ((94 - 2 * R) / 4) * 4;
    static constexpr int ROWS = TH + 2 * R;
    static constexpr int RP = ROWS / 2;
    static constexpr int RB = TH / 4;
    static_assert(ROWS % 2 == 0 && TH % 4 == 0 && R <= 8 && RP <= 3 * kWarps, "tile geometry");
};

template <typename Tin, int KS>
constexpr size_t tile_smem() {
    return 2 * (size_t)TileGeom<KS>::ROWS * kBoxW * sizeof(Tin) + (size_t)TileGeom<KS>::ROWS * kTP * sizeof(float) + 16 +
           (size_t)TileGeom<KS>::TH * 16 * sizeof(uint16_t);
}

// ---- mbarrier / TMA (PTX ISA: mbarrier.*, cp.async.bulk.tensor) -------------------------------------
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

// ---- integer pixel -> float, two rows at a time ----------------------------------------------------
// (0x4B000000 | v) is the float 2^23 + v; subtracting 2^23 (one packed FADD2 for both rows) leaves v.
__device__ __forceinline__ float2 cvt_pair(uint32_t a, uint32_t b) {
    return __fadd2_rn(make_float2(__uint_as_float(a), __uint_as_float(b)), make_float2(-8388608.0f, -8388608.0f));
}

template <typename Tin>
__device__ __forceinline__ void load_convert8(const Tin* ra, const Tin* rb, float2* x);
template <>
__device__ __forceinline__ void load_convert8<uint16_t>(const uint16_t* ra, const uint16_t* rb, float2* x) {
    const uint4 A = *reinterpret_cast<const uint4*>(ra), B = *reinterpret_cast<const uint4*>(rb);
    const uint32_t aw[4] = {A.x, A.y, A.z, A.w}, bw[4] = {B.x, B.y, B.z, B.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        x[2 * k] = cvt_pair(__byte_perm(aw[k], 0x4B000000u, 0x7410), __byte_perm(bw[k], 0x4B000000u, 0x7410));
        x[2 * k + 1] = cvt_pair(__byte_perm(aw[k], 0x4B000000u, 0x7432), __byte_perm(bw[k], 0x4B000000u, 0x7432));
    }
}
template <>
__device__ __forceinline__ void load_convert8<uint8_t>(const uint8_t* ra, const uint8_t* rb, float2* x) {
    const uint2 A = *reinterpret_cast<const uint2*>(ra), B = *reinterpret_cast<const uint2*>(rb);
    const uint32_t aw[2] = {A.x, A.y}, bw[2] = {B.x, B.y};
#pragma unroll
    for (int k = 0; k < 2; k++) {
        x[4 * k] = cvt_pair(__byte_perm(aw[k], 0x4B000000u, 0x7440), __byte_perm(bw[k], 0x4B000000u, 0x7440));
        x[4 * k + 1] = cvt_pair(__byte_perm(aw[k], 0x4B000000u, 0x7441), __byte_perm(bw[k], 0x4B000000u, 0x7441));
        x[4 * k + 2] = cvt_pair(__byte_perm(aw[k], 0x4B000000u, 0x7442), __byte_perm(bw[k], 0x4B000000u, 0x7442));
        x[4 * k + 3] = cvt_pair(__byte_perm(aw[k], 0x4B000000u, 0x7443), __byte_perm(bw[k], 0x4B000000u, 0x7443));
    }
}

__device__ __forceinline__ float2 bcast(float v) { return make_float2(v, v); }

// cv2 4.13 row pass on two rows at once (same operation order as row_dot in yam_filter.cu):
//   K>=7: s = k0*x0; s = fma(x_i, k_i, s)      K==5: (x1+x3)*k3, fma(x2,k2,.), fma(x0+x4,k4,.)
//   K==3: fma(x0+x2, k2, x1*k1)
template <int KS>
__device__ __forceinline__ float2 row_dot2(const float2* x, const Taps16& t) {
    if (KS == 3) return __ffma2_rn(__fadd2_rn(x[0], x[2]), bcast(t.v[2]), __fmul2_rn(x[1], bcast(t.v[1])));
    if (KS == 5) {
        float2 s = __fmul2_rn(__fadd2_rn(x[1], x[3]), bcast(t.v[3]));
        s = __ffma2_rn(x[2], bcast(t.v[2]), s);
        return __ffma2_rn(__fadd2_rn(x[0], x[4]), bcast(t.v[4]), s);
    }
    float2 s = __fmul2_rn(bcast(t.v[0]), x[0]);
#pragma unroll
    for (int i = 1; i < KS; i++) s = __ffma2_rn(x[i], bcast(t.v[i]), s);
    return s;
}

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// MASK: the kernel also writes the GLOBAL threshold mask of the same pixels (dst = src > t[frame] ? maxval : 0,
// cv2.threshold(..., THRESH_BINARY) with the Otsu value that is already on the device): the raw tile is in shared
// memory anyway, so the separate threshold pass (one more read of the image, DRAM-bound) disappears.
template <typename Tin, int KS, bool MASK>
__global__ void __launch_bounds__(kThreads, 1)
adaptive_bits_tma_kernel(const __grid_constant__ CUtensorMap tmap, uint16_t* __restrict__ bits16, int h, int w,
                         int wpr, Taps16 taps, int ci, int tiles_x, int tiles_y, int total_tiles,
                         const int32_t* __restrict__ t_dev, uint16_t* __restrict__ mask_out, uint32_t mv2) {
    typedef TileGeom<KS> G;
    typedef LaneGeom<Tin> L;
    constexpr int R = G::R, TH = G::TH, ROWS = G::ROWS, RP = G::RP, RB = G::RB;
    constexpr int OUTW = L::OUTW, MARGIN = L::MARGIN;
    constexpr size_t RAW_BYTES = (size_t)ROWS * kBoxW * sizeof(Tin);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s_t = reinterpret_cast<float*>(smem_raw + 2 * RAW_BYTES);                        // [ROWS][240] row-pass results
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + 2 * RAW_BYTES + (size_t)ROWS * kTP * 4);
    uint16_t* s_bits = reinterpret_cast<uint16_t*>(bar + 2);                                // [TH][16] finished 16-bit mask units
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // tile k of this CTA = global tile blockIdx.x + k * gridDim.x; x fastest so neighbouring CTAs share halos in L2
    auto issue = [&](int k) {
        const int tile = blockIdx.x + k * gridDim.x;
        if (tile >= total_tiles) return;
        const int tx = tile % tiles_x, rest = tile / tiles_x;
        const int ty = rest % tiles_y, fr = rest / tiles_y;
        mbar_expect_tx(bar + (k & 1), (uint32_t)RAW_BYTES);
        tma_load_3d(smem_raw + (k & 1) * RAW_BYTES, &tmap, bar + (k & 1), tx * OUTW - MARGIN, ty * TH - R, fr);
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

  // this CTA's tile coordinates advance incrementally (no divisions in the loop)
  int tx = (int)(blockIdx.x % tiles_x), ty = (int)((blockIdx.x / tiles_x) % tiles_y), frame = (int)(blockIdx.x / (tiles_x * tiles_y));
  for (int k = 0;; k++) {
    const int tile = blockIdx.x + k * gridDim.x;
    if (tile >= total_tiles) break;
    const int x0 = tx * OUTW, y0 = ty * TH;
    const int bx0 = x0 - MARGIN, by0 = y0 - R;  // box origin: a multiple of 16 bytes in x
    Tin* s_raw = reinterpret_cast<Tin*>(smem_raw + (k & 1) * RAW_BYTES);                    // [ROWS][256] raw pixels
    mbar_wait(bar + (k & 1), (k >> 1) & 1);

    // ---- BORDER_REPLICATE for tiles that hang over the image: TMA zero-fills out-of-range pixels
    if (bx0 < 0 || by0 < 0 || bx0 + kBoxW > w || by0 + ROWS > h) {
        const int cl = max(0, -bx0), cr = min(kBoxW, w - bx0);    // in-range columns [cl, cr)
        const int rt = max(0, -by0), rbm = min(ROWS, h - by0);     // in-range rows    [rt, rbm)
        const int ncol = cl + (kBoxW - cr);
        for (int i = tid; i < (rbm - rt) * ncol; i += kThreads) {
            const int ry = rt + i / ncol, kk = i - (i / ncol) * ncol;
            const int cx = kk < cl ? kk : cr + (kk - cl);
            s_raw[ry * kBoxW + cx] = s_raw[ry * kBoxW + (kk < cl ? cl : cr - 1)];
        }
        __syncthreads();
        const int nrow = rt + (ROWS - rbm);
        for (int i = tid; i < nrow * kBoxW; i += kThreads) {
            const int kk = i / kBoxW, cx = i - kk * kBoxW;
            const int ry = kk < rt ? kk : rbm + (kk - rt);
            s_raw[ry * kBoxW + cx] = s_raw[(kk < rt ? rt : rbm - 1) * kBoxW + cx];
        }
        __syncthreads();
    }

    // ---- horizontal pass: one warp per row pair.  Trip counts are compile-time constants (the last
    // round recomputes row pair RP - 1 in the warps that ran out of work: same values, benign) so the
    // compiler keeps the warp converged around the shuffles.
    {
        // intermediate tile layout: 16-byte chunks of every second 128-byte group are pair-swapped
        // (chunk ^ ((column >> 5) & 1)) so that the 8-float rows of lanes m and m + 4 do not collide
        const int m = lane - L::LM;
        const int sw = (m >> 2) & 1;
        const int off0 = ((2 * m) ^ sw) * 4, off1 = ((2 * m + 1) ^ sw) * 4;
        const bool stores = m >= 0 && m < OUTW / 8;
        constexpr int ROUNDS = (RP + kThreads / 32 - 1) / (kThreads / 32);
        // software pipeline: the window (load, convert, halo shuffles) of round it + 1 is prepared
        // before the 88 packed FMAs of round it are issued, so LDS / SHFL latency hides behind them
        // MASK: the lane has 8 raw pixels of two rows in hand here: the global threshold mask of those 16 pixels
        // goes out from the H pass (two 16-byte streaming stores per lane and row pair, spread over the FMA-bound
        // rounds; a separate store loop after the V pass ran at DRAM write speed with nothing to overlap it)
        const int tv = MASK ? t_dev[frame] : 0;
        const bool m_all = tv < 0;                                 // src > t holds for every pixel
        const uint32_t t2 = (uint32_t)min(max(tv, 0), 65535) * 0x10001u;
        auto mask_row = [&](const uint4& q, int row) {             // row = tile row (0 .. ROWS), this lane's 8 pixels
            const int gy = y0 - R + row, gx = x0 + 8 * m;
            if (row < R || row >= R + TH || gy >= h || m < 0 || m >= OUTW / 8 || gx >= w) return;
            uint4 o;
            o.x = m_all ? mv2 : (__vcmpgtu2(q.x, t2) & mv2);
            o.y = m_all ? mv2 : (__vcmpgtu2(q.y, t2) & mv2);
            o.z = m_all ? mv2 : (__vcmpgtu2(q.z, t2) & mv2);
            o.w = m_all ? mv2 : (__vcmpgtu2(q.w, t2) & mv2);
            yam_st_stream(reinterpret_cast<uint4*>(mask_out + ((int64_t)frame * h + gy) * w + gx), o);
        };
        auto prepare = [&](int it, float2* e) {
            const int rp = min(warp + it * kWarps, RP - 1);
            const Tin* ra = s_raw + (2 * rp) * kBoxW + 8 * lane;
            if (MASK && sizeof(Tin) == 2 && warp + it * kWarps < RP) {
                mask_row(*reinterpret_cast<const uint4*>(ra), 2 * rp);
                mask_row(*reinterpret_cast<const uint4*>(ra + kBoxW), 2 * rp + 1);
            }
            load_convert8<Tin>(ra, ra + kBoxW, e + R);
#pragma unroll
            for (int i = 0; i < R; i++) {
                // left halo = the neighbour's last R pixels, right halo = the other neighbour's first R
                const float2 lo = e[R + 8 - R + i], hi = e[R + i];
                e[i] = make_float2(__shfl_up_sync(0xffffffffu, lo.x, 1), __shfl_up_sync(0xffffffffu, lo.y, 1));
                e[R + 8 + i] = make_float2(__shfl_down_sync(0xffffffffu, hi.x, 1), __shfl_down_sync(0xffffffffu, hi.y, 1));
            }
        };
        float2 win[2][8 + 2 * R];
        prepare(0, win[0]);
#pragma unroll
        for (int it = 0; it < ROUNDS; it++) {
            if (it + 1 < ROUNDS) prepare(it + 1, win[(it + 1) & 1]);
            const float2* e = win[it & 1];
            const int rp = min(warp + it * kWarps, RP - 1);
            float2 o[8];
#pragma unroll
            for (int j = 0; j < 8; j++) o[j] = row_dot2<KS>(e + j, taps);
            if (stores) {
                float* ta = s_t + (2 * rp) * kTP;
                float* tb = ta + kTP;
                *reinterpret_cast<float4*>(ta + off0) = make_float4(o[0].x, o[1].x, o[2].x, o[3].x);
                *reinterpret_cast<float4*>(ta + off1) = make_float4(o[4].x, o[5].x, o[6].x, o[7].x);
                *reinterpret_cast<float4*>(tb + off0) = make_float4(o[0].y, o[1].y, o[2].y, o[3].y);
                *reinterpret_cast<float4*>(tb + off1) = make_float4(o[4].y, o[5].y, o[6].y, o[7].y);
            }
        }
    }
    __syncthreads();

    // ---- vertical pass + compare + ballot: 4 row blocks x 4 column groups of 64 = 16 items, one per warp
    float2 kv[R + 1];
#pragma unroll
    for (int q = 0; q <= R; q++) kv[q] = bcast(taps.v[R + q]);
    static_assert(kWarps == 16, "one V-pass item per warp");
    {
        const int item = warp;
        const int rbk = item >> 2, cg = item & 3;
        const int r0 = rbk * RB;
        const bool has1 = 64 * cg + 32 + lane < OUTW;  // the last group is only partly populated
        const int i0 = 64 * cg + lane;
        const int i1 = has1 ? 64 * cg + 32 + (lane ^ 4) : i0;  // pair-swapped chunk (see the H pass)
        const float* t0 = s_t + r0 * kTP + i0;
        const float* t1 = s_t + r0 * kTP + i1;
        float2 c[RB + 2 * R];
#pragma unroll
        for (int i = 0; i < RB + 2 * R; i++) c[i] = make_float2(t0[i * kTP], t1[i * kTP]);
        const Tin* sp0 = s_raw + (R + r0) * kBoxW + MARGIN + 64 * cg + lane;
        const Tin* sp1 = sp0 + (has1 ? 32 : 0);
        const int gx0 = x0 + 64 * cg + lane;
        // src - rint(mean) > -C  <=>  bits(mean + 1.5 * 2^23) - (0x4B400000 + C) < src.  Out-of-image
        // columns use an offset that makes the left side huge, so the loop body needs no extra predicate.
        const int ci0 = gx0 < w ? ci : -0x30000000, ci1 = (has1 && gx0 + 32 < w) ? ci : -0x30000000;
        uint2* sb = reinterpret_cast<uint2*>(s_bits + r0 * 16 + 4 * cg);   // two finished words per row and group
#pragma unroll
        for (int j = 0; j < RB; j++) {
            float2 a = __fmul2_rn(kv[0], c[j + R]);
#pragma unroll
            for (int q = 1; q <= R; q++) a = __ffma2_rn(__fadd2_rn(c[j + R + q], c[j + R - q]), kv[q], a);
            // bits(a + 1.5 * 2^23) = 0x4B400000 + rint(a) (round-half-even; 0 <= a < 2^22)
            const float2 t = __fadd2_rn(a, make_float2(12582912.0f, 12582912.0f));
            const int s0 = (int)sp0[j * kBoxW];
            const int s1 = (int)sp1[j * kBoxW];
            const uint32_t w0 = __ballot_sync(0xffffffffu, __float_as_int(t.x) - ci0 < s0);
            const uint32_t w1 = __ballot_sync(0xffffffffu, __float_as_int(t.y) - ci1 < s1);
            if (lane == 0) sb[j * 4] = make_uint2(w0, w1);
        }
    }
    __syncthreads();
    // every read of this tile's raw buffer and of s_t is done: refill the buffer with tile k + 2
    if (tid == 0) {
        fence_proxy_async();   // the buffer was touched through the generic proxy (reads, border fix-up)
        issue(k + 2);
    }

    // ---- finished mask units -> global: OUTW / 16 16-bit units per tile row (the rest belongs to the next tile)
    {
        const int half = (x0 >> 4) + lane;
        if (lane < L::UNITS && half < 2 * wpr) {
            uint16_t* op = bits16 + ((int64_t)frame * h + y0) * (2 * wpr) + half;
            const int rows = min(TH, h - y0);
            for (int row = warp; row < rows; row += kWarps) op[(int64_t)row * (2 * wpr)] = s_bits[row * 16 + lane];
        }
    }
    // next tile of this CTA
    tx += (int)gridDim.x;
    while (tx >= tiles_x) {
        tx -= tiles_x;
        if (++ty == tiles_y) {
            ty = 0;
            frame++;
        }
    }
    // (s_bits is rewritten only after the next tile's post-H-pass barrier, which every thread reaches
    // after finishing this loop)
  }
}

// ---- host side -----------------------------------------------------------------------------------------
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

template <typename Tin, int KS>
int launch(yam_ctx* ctx, const Tin* src, uint32_t* bits, int64_t n, int64_t h, int64_t w, const Taps16& taps, int idelta,
           const int32_t* t_dev, uint16_t* mask_out, uint32_t mv2) {
    typedef TileGeom<KS> G;
    CUtensorMap map;
    const cuuint64_t gdim[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    const cuuint64_t gstride[2] = {(cuuint64_t)w * sizeof(Tin), (cuuint64_t)w * h * sizeof(Tin)};
    const cuuint32_t box[3] = {kBoxW, (cuuint32_t)G::ROWS, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult rc = encode_tiled()(&map, sizeof(Tin) == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 3,
                                       const_cast<Tin*>(src), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        yam_set_error("adaptive_threshold_bits: cuTensorMapEncodeTiled failed (%d)", (int)rc);
        return YAM_ECUDA;
    }
    constexpr size_t smem = tile_smem<Tin, KS>();
    YAM_CUDA(cudaFuncSetAttribute(adaptive_bits_tma_kernel<Tin, KS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    YAM_CUDA(cudaFuncSetAttribute(adaptive_bits_tma_kernel<Tin, KS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int wpr = (int)((w + 31) / 32);
    // 16-bit units past the last tile column (w % 32 != 0 only) are never written by a tile: clear them
    constexpr int kOutW = LaneGeom<Tin>::OUTW;
    if ((w & 31) && ((w + kOutW - 1) / kOutW) * kOutW < (int64_t)wpr * 32)
        YAM_CUDA(cudaMemsetAsync(bits, 0, (size_t)n * h * wpr * 4, ctx->stream));
    const int tiles_x = (int)((w + kOutW - 1) / kOutW), tiles_y = (int)((h + G::TH - 1) / G::TH);
    const int64_t total = (int64_t)tiles_x * tiles_y * n;
    if (total >= (1ll << 31)) {
        yam_set_error("adaptive_threshold_bits: too many tiles");
        return YAM_EINVAL;
    }
    const unsigned grid = (unsigned)(total < ctx->num_sms ? total : ctx->num_sms);   // persistent: one CTA per SM
    if (mask_out)
        adaptive_bits_tma_kernel<Tin, KS, true><<<grid, kThreads, smem, ctx->stream>>>(
            map, reinterpret_cast<uint16_t*>(bits), (int)h, (int)w, wpr, taps, 0x4B400000 + idelta, tiles_x, tiles_y, (int)total,
            t_dev, mask_out, mv2);
    else
        adaptive_bits_tma_kernel<Tin, KS, false><<<grid, kThreads, smem, ctx->stream>>>(
            map, reinterpret_cast<uint16_t*>(bits), (int)h, (int)w, wpr, taps, 0x4B400000 + idelta, tiles_x, tiles_y, (int)total,
            nullptr, nullptr, 0u);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

template <typename Tin>
int dispatch(yam_ctx* ctx, const Tin* src, uint32_t* bits, int64_t n, int64_t h, int64_t w, int ks, const Taps16& taps,
             int idelta, const int32_t* t_dev, uint16_t* mask_out, uint32_t mv2) {
    switch (ks) {
        case 3: return launch<Tin, 3>(ctx, src, bits, n, h, w, taps, idelta, t_dev, mask_out, mv2);
        case 5: return launch<Tin, 5>(ctx, src, bits, n, h, w, taps, idelta, t_dev, mask_out, mv2);
        case 7: return launch<Tin, 7>(ctx, src, bits, n, h, w, taps, idelta, t_dev, mask_out, mv2);
        case 11: return launch<Tin, 11>(ctx, src, bits, n, h, w, taps, idelta, t_dev, mask_out, mv2);
        case 15: return launch<Tin, 15>(ctx, src, bits, n, h, w, taps, idelta, t_dev, mask_out, mv2);
    }
    return YAM_EINVAL;
}

}  // namespace

// Returns YAM_OK with *handled = 1 when the TMA kernel took the call; *handled = 0 (and YAM_OK) when the
// shape does not qualify (the caller then uses sep_f32_tiled): rows must be 16-byte multiples of a
// 16-byte aligned base (tensor-map rule), the frame at least one box wide and high, block size one
// of 3, 5, 7, 11, 15, and the integer compare needs |C| < 2^20.
// t_dev / mask_out (both or neither; 16-bit input only): also write the global threshold mask, see the kernel.
int yam_adaptive_bits_tma(yam_ctx* ctx, const void* src, uint32_t* bits, int64_t n, int64_t h, int64_t w, int dtype,
                          int block_size, const float* taps_f, int idelta, int* handled, const int32_t* t_dev, void* mask_out,
                          double maxval) {
    *handled = 0;
    if (mask_out && (dtype != YAM_U16 || !t_dev)) return YAM_OK;
    const int es = dtype == YAM_U8 ? 1 : 2;
    const bool ks_ok = block_size == 3 || block_size == 5 || block_size == 7 || block_size == 11 || block_size == 15;
    if (!ks_ok || (dtype != YAM_U8 && dtype != YAM_U16)) return YAM_OK;
    if ((w * es) % 16 || (reinterpret_cast<uintptr_t>(src) & 15) || w < kBoxW || h < 96 || n > 65535) return YAM_OK;
    if (idelta <= -(1 << 20) || idelta >= (1 << 20)) return YAM_OK;
    const char* legacy = getenv("YAM_ADAPTIVE_LEGACY");
    if (legacy && legacy[0] == '1') return YAM_OK;
    if (!encode_tiled()) return YAM_OK;
    Taps16 taps;
    for (int i = 0; i < 16; i++) taps.v[i] = i < block_size ? taps_f[i] : 0.0f;
    const double mr = rint(maxval);
    const uint32_t mv2 = (uint32_t)(mr < 0 ? 0 : mr > 65535 ? 65535 : mr) * 0x10001u;
    const int rc = dtype == YAM_U8 ? dispatch<uint8_t>(ctx, (const uint8_t*)src, bits, n, h, w, block_size, taps, idelta, nullptr, nullptr, 0u)
                                   : dispatch<uint16_t>(ctx, (const uint16_t*)src, bits, n, h, w, block_size, taps, idelta, t_dev,
                                                        (uint16_t*)mask_out, mv2);
    if (rc == YAM_OK) *handled = 1;
    return rc;
}
