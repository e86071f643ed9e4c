// Bit-parallel morphology on 1-bit-per-pixel masks (rectangular structuring elements).
//
// A thresholded mask is binary by construction, so the segmentation chain
//   adaptive threshold -> open -> close -> connected components
// never needs the 8-bit mask in HBM: the threshold kernel emits packed bits
// (yam_adaptive_threshold_bits), this file erodes / dilates 32 pixels per instruction with
// funnel shifts + LOP3, and yam_ccl_label_bits labels the packed mask directly.
// Semantics are those of cv2.erode / dilate / morphologyEx on a {0,255} mask: constant border with
// the identity element (1 for erode, 0 for dilate), out-of-image intermediates = identity of the
// next stage, n iterations of a k x k rectangle = one (n(k-1)+1)^2 rectangle.
// Layout: row-major words, wpr = ceil(w/32) words per row, bit i of word j = pixel 32*j+i,
// bits beyond the image width are 0.
#include "yam_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int TR = 32;    // output rows per block
constexpr int TWW = 32;   // output words per block row (1024 pixels)
constexpr int kMaxStages = 4;
constexpr int kMaxHalo = 31;  // summed radius per side; one halo word per side covers 32 pixels

struct BitStage {
    int dilate;
    int L, R;  // window = pixel offsets [-L, +R] in both axes
};
struct BitChain {
    int n;
    BitStage s[kMaxStages];
    int HL, HR;
};

// pixel (x + s) for every bit x of `cur`, with neighbours `prev` (pixels -32..-1) and `next` (+32..+63)
__device__ __forceinline__ uint32_t shift_px(uint32_t prev, uint32_t cur, uint32_t next, int s) {
    if (s == 0) return cur;
    if (s > 0) return __funnelshift_r(cur, next, s);   // (next:cur) >> s
    return __funnelshift_l(prev, cur, -s);             // (cur:prev) << |s|, upper word
}

__device__ __forceinline__ uint32_t valid_mask(int j, int w) {
    const int rem = w - j * 32;
    if (rem >= 32) return 0xffffffffu;
    if (rem <= 0) return 0u;
    return (1u << rem) - 1u;
}

// One block = TR x (32*TWW) output pixels.  The buffer carries one extra word per side and HL/HR
// extra rows; every pass runs over the WHOLE buffer: bits whose window leaves the buffer come out
// wrong, but wrongness travels at most (window radius) pixels per stage and the summed radius is
// <= 31 < 32, so it never reaches the tile.
__global__ void __launch_bounds__(kThreads) bit_morph_chain_kernel(const uint32_t* __restrict__ in,
                                                                   uint32_t* __restrict__ out, int h, int w, int wpr,
                                                                   BitChain ch) {
    const int BH = TR + ch.HL + ch.HR;
    constexpr int BW = TWW + 2;
    extern __shared__ uint32_t smem_bits[];
    uint32_t* A = smem_bits;
    uint32_t* B = smem_bits + BH * BW;
    in += (int64_t)blockIdx.z * h * wpr;
    out += (int64_t)blockIdx.z * h * wpr;
    const int j0 = blockIdx.x * TWW - 1;  // word index of buffer column 0
    const int y0 = blockIdx.y * TR - ch.HL;
    const uint32_t id0 = ch.s[0].dilate ? 0u : 0xffffffffu;
    for (int i = threadIdx.x; i < BH * BW; i += kThreads) {
        const int by = i / BW, bj = i - by * BW;
        const int gy = y0 + by, gj = j0 + bj;
        uint32_t v = id0;
        if ((unsigned)gy < (unsigned)h && (unsigned)gj < (unsigned)wpr) {
            const uint32_t vm = valid_mask(gj, w);
            v = (in[(int64_t)gy * wpr + gj] & vm) | (id0 & ~vm);
        }
        A[i] = v;
    }
    __syncthreads();
    for (int si = 0; si < ch.n; si++) {
        const int dil = ch.s[si].dilate, L = ch.s[si].L, R = ch.s[si].R;
        // horizontal pass A -> B
        for (int i = threadIdx.x; i < BH * BW; i += kThreads) {
            const int by = i / BW, bj = i - by * BW;
            const uint32_t* row = A + by * BW;
            const uint32_t cur = row[bj];
            const uint32_t prev = bj > 0 ? row[bj - 1] : cur;
            const uint32_t next = bj + 1 < BW ? row[bj + 1] : cur;
            uint32_t acc = cur;
            if (dil) {
                for (int s = 1; s <= L; s++) acc |= shift_px(prev, cur, next, -s);
                for (int s = 1; s <= R; s++) acc |= shift_px(prev, cur, next, s);
            } else {
                for (int s = 1; s <= L; s++) acc &= shift_px(prev, cur, next, -s);
                for (int s = 1; s <= R; s++) acc &= shift_px(prev, cur, next, s);
            }
            B[i] = acc;
        }
        __syncthreads();
        // vertical pass B -> A; outside the image the next stage must see its identity element
        const bool last = (si + 1 == ch.n);
        const uint32_t next_id = (!last && !ch.s[si + 1].dilate) ? 0xffffffffu : 0u;
        for (int i = threadIdx.x; i < BH * BW; i += kThreads) {
            const int by = i / BW, bj = i - by * BW;
            if (by < L || by > BH - 1 - R) continue;
            uint32_t acc = B[i];
            if (dil) {
                for (int o = -L; o <= R; o++) acc |= B[(by + o) * BW + bj];
            } else {
                for (int o = -L; o <= R; o++) acc &= B[(by + o) * BW + bj];
            }
            const int gy = y0 + by, gj = j0 + bj;
            uint32_t vm = 0u;
            if ((unsigned)gy < (unsigned)h && (unsigned)gj < (unsigned)wpr) vm = valid_mask(gj, w);
            A[i] = (acc & vm) | (next_id & ~vm);
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < TR * TWW; i += kThreads) {
        const int ty = i / TWW, tj = i - ty * TWW;
        const int gy = blockIdx.y * TR + ty, gj = blockIdx.x * TWW + tj;
        if (gy < h && gj < wpr) out[(int64_t)gy * wpr + gj] = A[(ch.HL + ty) * BW + 1 + tj] & valid_mask(gj, w);
    }
}

// ------------------------------------------------------------------------------------------------
// Register-resident fast path (odd k x k rectangles, summed radius <= 12 per side).
// One WARP owns a column of 30 output words x TRR output rows: lane l holds word column
// (30*warp_col - 1 + l) for every buffer row in registers (lanes 0 / 31 are the halo words).  The
// horizontal pass takes its neighbour words by warp shuffle, the vertical pass is a compile-time
// sliding window over the register column (doubling: windows of 2, 4, 8 rows combined with one
// overlapping AND/OR), every stage of the chain stays in registers: no shared memory, no barriers.
// The "wrongness travels <= summed radius" argument of the kernel above holds unchanged.
template <bool DIL>
__device__ __forceinline__ uint32_t bcomb(uint32_t a, uint32_t b) { return DIL ? (a | b) : (a & b); }

template <bool DIL, int RAD, int BH>
__device__ __forceinline__ void reg_stage(uint32_t (&v)[BH], uint32_t next_id, uint32_t vm_lane, int y_first, int h) {
    uint32_t t[BH];
#pragma unroll
    for (int i = 0; i < BH; i++) {
        const uint32_t cur = v[i];
        const uint32_t prev = __shfl_up_sync(0xffffffffu, cur, 1);
        const uint32_t next = __shfl_down_sync(0xffffffffu, cur, 1);
        uint32_t acc = cur;
#pragma unroll
        for (int s = 1; s <= RAD; s++) {
            acc = bcomb<DIL>(acc, __funnelshift_r(cur, next, s));  // pixel x + s
            acc = bcomb<DIL>(acc, __funnelshift_l(prev, cur, s));  // pixel x - s
        }
        t[i] = acc;
    }
    constexpr int K = 2 * RAD + 1;
    constexpr int P = K >= 8 ? 8 : (K >= 4 ? 4 : 2);
    uint32_t wp[BH];
#pragma unroll
    for (int i = 0; i < BH; i++) wp[i] = t[i];
#pragma unroll
    for (int step = 1; step < P; step <<= 1) {
#pragma unroll
        for (int i = 0; i + step < BH; i++) wp[i] = bcomb<DIL>(wp[i], wp[i + step]);  // window 2*step starting at i
    }
#pragma unroll
    for (int i = 0; i < BH; i++) {
        uint32_t acc = t[i];
        if (i >= RAD && i + RAD < BH) acc = bcomb<DIL>(wp[i - RAD], wp[i + RAD - P + 1]);
        const uint32_t m = ((unsigned)(y_first + i) < (unsigned)h) ? vm_lane : 0u;
        v[i] = (acc & m) | (next_id & ~m);
    }
}

template <int OP, int A>
struct RegChain {
    static constexpr int HALO = OP == YAM_MORPH_OPEN_CLOSE ? 4 * A : (OP == YAM_MORPH_OPEN || OP == YAM_MORPH_CLOSE) ? 2 * A : A;
};

constexpr int kRegWarps = 4;  // warps per block (independent: no block-level cooperation)

template <int OP, int A, int TRR>
__global__ void __launch_bounds__(kRegWarps * 32) bit_morph_reg_kernel(const uint32_t* __restrict__ in,
                                                                       uint32_t* __restrict__ out, int h, int w,
                                                                       int wpr, int col_groups) {
    constexpr int HALO = RegChain<OP, A>::HALO;
    constexpr int BH = TRR + 2 * HALO;
    constexpr uint32_t ONES = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int cgp = blockIdx.x * kRegWarps + (threadIdx.x >> 5);
    if (cgp >= col_groups) return;  // surplus warp of the last block (warp-uniform exit)
    const int band = blockIdx.y;
    in += (int64_t)blockIdx.z * h * wpr;
    out += (int64_t)blockIdx.z * h * wpr;
    const int gj = cgp * 30 - 1 + lane;
    const int y_first = band * TRR - HALO;
    const uint32_t vm_lane = ((unsigned)gj < (unsigned)wpr) ? valid_mask(gj, w) : 0u;
    constexpr bool first_dil = (OP == YAM_MORPH_DILATE || OP == YAM_MORPH_CLOSE);
    const uint32_t id0 = first_dil ? 0u : ONES;
    uint32_t v[BH];
#pragma unroll
    for (int i = 0; i < BH; i++) {
        const int gy = y_first + i;
        uint32_t x = 0u, m = 0u;
        if ((unsigned)gy < (unsigned)h && vm_lane) {
            x = __ldg(in + (int64_t)gy * wpr + gj);
            m = vm_lane;
        }
        v[i] = (x & m) | (id0 & ~m);
    }
    if (OP == YAM_MORPH_ERODE) {
        reg_stage<false, A, BH>(v, 0u, vm_lane, y_first, h);
    } else if (OP == YAM_MORPH_DILATE) {
        reg_stage<true, A, BH>(v, 0u, vm_lane, y_first, h);
    } else if (OP == YAM_MORPH_OPEN) {
        reg_stage<false, A, BH>(v, 0u, vm_lane, y_first, h);
        reg_stage<true, A, BH>(v, 0u, vm_lane, y_first, h);
    } else if (OP == YAM_MORPH_CLOSE) {
        reg_stage<true, A, BH>(v, ONES, vm_lane, y_first, h);
        reg_stage<false, A, BH>(v, 0u, vm_lane, y_first, h);
    } else {  // open then close: erode A, dilate 2A, erode A
        reg_stage<false, A, BH>(v, 0u, vm_lane, y_first, h);
        reg_stage<true, 2 * A, BH>(v, ONES, vm_lane, y_first, h);
        reg_stage<false, A, BH>(v, 0u, vm_lane, y_first, h);
    }
    if (lane >= 1 && lane <= 30 && gj < wpr) {
#pragma unroll
        for (int i = 0; i < TRR; i++) {
            const int gy = band * TRR + i;
            if (gy < h) out[(int64_t)gy * wpr + gj] = v[HALO + i] & vm_lane;
        }
    }
}

template <int OP, int A, int TRR>
int launch_bit_reg(yam_ctx* ctx, const uint32_t* in, uint32_t* out, int64_t n, int64_t h, int64_t w) {
    const int wpr = (int)((w + 31) / 32);
    const int col_groups = (wpr + 29) / 30;
    // grid.x blocks of kRegWarps warps cover the col_groups of one band; surplus warps exit
    dim3 grid((unsigned)((col_groups + kRegWarps - 1) / kRegWarps), (unsigned)((h + TRR - 1) / TRR), (unsigned)n);
    bit_morph_reg_kernel<OP, A, TRR><<<grid, kRegWarps * 32, 0, ctx->stream>>>(in, out, (int)h, (int)w, wpr, col_groups);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

template <int OP>
int dispatch_bit_reg(yam_ctx* ctx, const uint32_t* in, uint32_t* out, int64_t n, int64_t h, int64_t w, int a) {
    switch (a) {
        case 1: return launch_bit_reg<OP, 1, 16>(ctx, in, out, n, h, w);
        case 2: return launch_bit_reg<OP, 2, 16>(ctx, in, out, n, h, w);
        case 3: return launch_bit_reg<OP, 3, 16>(ctx, in, out, n, h, w);
    }
    return -1;
}

// bits -> u8 mask {0, 255}: 32 pixels per thread
__global__ void __launch_bounds__(kThreads) bits_unpack_kernel(const uint32_t* __restrict__ bits, uint8_t* __restrict__ mask,
                                                               int64_t rows, int w, int wpr) {
    const int64_t gw = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gw >= rows * wpr) return;
    const int64_t row = gw / wpr;
    const int j = (int)(gw - row * wpr);
    const uint32_t b = bits[gw];
    uint8_t* d = mask + row * (int64_t)w + (int64_t)j * 32;
    const int valid = min(32, w - j * 32);
    if (valid == 32 && ((reinterpret_cast<uintptr_t>(d) & 15) == 0)) {
        uint32_t wd[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t nib = (b >> (4 * i)) & 0xfu;
            // spread 4 bits to 4 bytes of 0x00 / 0xff
            const uint32_t spread = (nib * 0x00204081u) & 0x01010101u;
            wd[i] = spread * 0xffu;
        }
        reinterpret_cast<uint4*>(d)[0] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
        reinterpret_cast<uint4*>(d)[1] = make_uint4(wd[4], wd[5], wd[6], wd[7]);
    } else {
        for (int i = 0; i < valid; i++) d[i] = ((b >> i) & 1u) ? 255 : 0;
    }
}

int launch_bit_chain(yam_ctx* ctx, const uint32_t* in, uint32_t* out, int64_t n, int64_t h, int64_t w,
                     const BitChain& ch) {
    const int wpr = (int)((w + 31) / 32);
    const size_t smem = (size_t)2 * (TR + ch.HL + ch.HR) * (TWW + 2) * sizeof(uint32_t);
    dim3 grid((unsigned)((wpr + TWW - 1) / TWW), (unsigned)((h + TR - 1) / TR), (unsigned)n);
    bit_morph_chain_kernel<<<grid, kThreads, smem, ctx->stream>>>(in, out, (int)h, (int)w, wpr, ch);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

// stages -> launches whose summed radius stays <= kMaxHalo per side (ping-pong through scratch)
int run_bit_stages(yam_ctx* ctx, const uint32_t* in, uint32_t* out, int64_t n, int64_t h, int64_t w,
                   const BitStage* stages, int count) {
    BitStage pieces[64];
    int np = 0;
    for (int i = 0; i < count; i++) {
        int L = stages[i].L, R = stages[i].R;
        while (L > 0 || R > 0) {
            const int l = L < kMaxHalo ? L : kMaxHalo, r = R < kMaxHalo ? R : kMaxHalo;
            YAM_REQUIRE(np < 64, "bits_morph: window too large");
            pieces[np++] = BitStage{stages[i].dilate, l, r};
            L -= l;
            R -= r;
        }
    }
    const int64_t words = n * h * ((w + 31) / 32);
    if (np == 0) {
        if (in != out) YAM_CUDA(cudaMemcpyAsync(out, in, (size_t)words * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        return YAM_OK;
    }
    BitChain chains[64];
    int nc = 0;
    BitChain cur = {};
    for (int i = 0; i < np; i++) {
        const bool fits = cur.n < kMaxStages && cur.HL + pieces[i].L <= kMaxHalo && cur.HR + pieces[i].R <= kMaxHalo;
        if (cur.n > 0 && !fits) {
            chains[nc++] = cur;
            cur = BitChain{};
        }
        if (cur.n > 0 && cur.s[cur.n - 1].dilate == pieces[i].dilate) {
            cur.s[cur.n - 1].L += pieces[i].L;
            cur.s[cur.n - 1].R += pieces[i].R;
        } else {
            cur.s[cur.n++] = pieces[i];
        }
        cur.HL += pieces[i].L;
        cur.HR += pieces[i].R;
    }
    chains[nc++] = cur;
    if (nc == 1) return launch_bit_chain(ctx, in, out, n, h, w, chains[0]);
    void* scratch = nullptr;
    if (int rc = yam_scratch(ctx, (size_t)words * 4, &scratch)) return rc;
    uint32_t* tmp = (uint32_t*)scratch;
    const uint32_t* src = in;
    for (int i = 0; i < nc; i++) {
        uint32_t* dst = ((nc - 1 - i) % 2 == 0) ? out : tmp;
        if (int rc = launch_bit_chain(ctx, src, dst, n, h, w, chains[i])) return rc;
        src = dst;
    }
    return YAM_OK;
}

}  // namespace

extern "C" {

int yam_bits_morph(yam_ctx* ctx, const uint32_t* bits_in, uint32_t* bits_out, int64_t n, int64_t h, int64_t w, int op,
                   int ksize, int iterations) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(bits_in && bits_out && bits_in != bits_out && n > 0 && h > 0 && w > 0 && n <= 65535, "bits_morph: bad arguments");
    YAM_REQUIRE(ksize >= 1 && ksize <= 31 && iterations >= 1 && iterations <= 64, "bits_morph: bad kernel size / iterations");
    const int a = ksize / 2;
    const int L = a * iterations, R = (ksize - 1 - a) * iterations;
    if ((ksize & 1) && L >= 1 && L <= 3 && h < (1 << 30)) {
        // register-resident kernel (n iterations of a k x k rectangle = one window of radius n*(k/2))
        switch (op) {
            case YAM_MORPH_ERODE: return dispatch_bit_reg<YAM_MORPH_ERODE>(ctx, bits_in, bits_out, n, h, w, L);
            case YAM_MORPH_DILATE: return dispatch_bit_reg<YAM_MORPH_DILATE>(ctx, bits_in, bits_out, n, h, w, L);
            case YAM_MORPH_OPEN: return dispatch_bit_reg<YAM_MORPH_OPEN>(ctx, bits_in, bits_out, n, h, w, L);
            case YAM_MORPH_CLOSE: return dispatch_bit_reg<YAM_MORPH_CLOSE>(ctx, bits_in, bits_out, n, h, w, L);
            case YAM_MORPH_OPEN_CLOSE: return dispatch_bit_reg<YAM_MORPH_OPEN_CLOSE>(ctx, bits_in, bits_out, n, h, w, L);
            default: break;
        }
    }
    BitStage st[4];
    int ns = 0;
    switch (op) {
        case YAM_MORPH_ERODE: st[ns++] = BitStage{0, L, R}; break;
        case YAM_MORPH_DILATE: st[ns++] = BitStage{1, L, R}; break;
        case YAM_MORPH_OPEN: st[ns++] = BitStage{0, L, R}; st[ns++] = BitStage{1, L, R}; break;
        case YAM_MORPH_CLOSE: st[ns++] = BitStage{1, L, R}; st[ns++] = BitStage{0, L, R}; break;
        case YAM_MORPH_OPEN_CLOSE:
            st[ns++] = BitStage{0, L, R}; st[ns++] = BitStage{1, L, R};
            st[ns++] = BitStage{1, L, R}; st[ns++] = BitStage{0, L, R};
            break;
        default: YAM_REQUIRE(false, "bits_morph: unknown op %d", op);
    }
    return run_bit_stages(ctx, bits_in, bits_out, n, h, w, st, ns);
}

int yam_bits_unpack(yam_ctx* ctx, const uint32_t* bits, void* mask_u8, int64_t n, int64_t h, int64_t w) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(bits && mask_u8 && n > 0 && h > 0 && w > 0, "bits_unpack: bad arguments");
    const int wpr = (int)((w + 31) / 32);
    const int64_t words = n * h * wpr;
    bits_unpack_kernel<<<(unsigned)((words + kThreads - 1) / kThreads), kThreads, 0, ctx->stream>>>(
        bits, (uint8_t*)mask_u8, n * h, (int)w, wpr);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

}  // extern "C"
