// K6 morphology: erode / dilate / open / close (cv2 semantics: anchor k/2, constant border with
// the identity element, `iterations`).
//
// Rectangular structuring elements are separable and n iterations of a k x k rectangle equal one
// pass with the n-fold Minkowski sum, so a whole open / close / open+close is planned on the host
// as a short chain of (op, left, right) stages and executed by ONE kernel per chain: the halo
// tile is staged once in shared memory (pixels widened to u16 so the native VIMNMX.U16x2 /
// VIMNMX3.U16x2 packed min/max apply to u8 and u16 alike), each stage runs a horizontal and a
// vertical pass between two shared-memory buffers, and only the final tile is written back.
// HBM traffic: 1 read + 1 write per pixel for the whole chain.
//
// Elliptical and cross elements are not separable: one launch per iteration, footprint rows
// applied as horizontal runs from a shared-memory tile.
#include "yam_common.cuh"
#include "yam_host.h"

// fast path for short symmetric chains (yam_morph_fast.cu): 1 = launched, 0 = not covered, < 0 = error
template <typename T>
int yam_morph_fast_try(yam_ctx* ctx, const T* src, T* dst, int64_t n, int64_t h, int64_t w, int nstages,
                       const int* dilate, const int* radius);

namespace {

constexpr int kThreads = 256;
constexpr int TW = 64;
constexpr int TH = 64;
constexpr int kMaxStages = 4;
constexpr int kMaxHalo = 24;  // per side, per launch

struct Stage {
    int dilate;  // 0 erode (min), 1 dilate (max)
    int L, R;    // window = offsets [-L, +R] in both axes
};
struct Chain {
    int n;
    Stage s[kMaxStages];
    int HL, HR;  // summed extents
};

__device__ __forceinline__ uint32_t op2(uint32_t a, uint32_t b, int dilate) {
    return dilate ? __vmaxu2(a, b) : __vminu2(a, b);
}
__device__ __forceinline__ uint32_t op1(uint32_t a, uint32_t b, int dilate) {
    return dilate ? max(a, b) : min(a, b);
}

template <typename T>
__global__ void __launch_bounds__(kThreads) morph_rect_chain_kernel(const T* __restrict__ src,
                                                                    T* __restrict__ dst, int h, int w,
                                                                    Chain ch) {
    constexpr int VEC = 16 / sizeof(T);
    // aligned horizontal halo, one spare pixel so packed words at the valid-region edge stay in bounds
    const int HA = (max(ch.HL, ch.HR) + 1 + VEC - 1) / VEC * VEC;
    const int BW = TW + 2 * HA;                                 // buffer width in pixels (even)
    const int BH = TH + ch.HL + ch.HR;
    const int BWW = BW / 2;                                     // packed words per row
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint16_t* bufA = reinterpret_cast<uint16_t*>(smem_raw);
    uint16_t* bufB = bufA + BH * BW;

    const int64_t frame = blockIdx.z;
    src += frame * (int64_t)h * w;
    dst += frame * (int64_t)h * w;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int gx_org = x0 - HA, gy_org = y0 - ch.HL;  // global coords of buffer (0,0)

    // ---- load: outside the image = identity of the first stage
    {
        const uint16_t ident = ch.s[0].dilate ? 0 : 0xffff;
        const int vec_per_row = BW / VEC;
        const bool row_aligned = ((w % VEC) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
        for (int v = threadIdx.x; v < BH * vec_per_row; v += kThreads) {
            const int by = v / vec_per_row, vx = v - by * vec_per_row;
            const int gy = gy_org + by, gx = gx_org + vx * VEC;
            uint16_t* d = bufA + by * BW + vx * VEC;
            if ((unsigned)gy >= (unsigned)h) {
#pragma unroll
                for (int i = 0; i < VEC; i++) d[i] = ident;
            } else if (row_aligned && gx >= 0 && gx + VEC <= w) {
                const uint4 q = *reinterpret_cast<const uint4*>(src + (int64_t)gy * w + gx);
                const T* e = reinterpret_cast<const T*>(&q);
#pragma unroll
                for (int i = 0; i < VEC; i++) d[i] = e[i];
            } else {
#pragma unroll 4
                for (int i = 0; i < VEC; i++)
                    d[i] = ((unsigned)(gx + i) < (unsigned)w) ? (uint16_t)src[(int64_t)gy * w + gx + i] : ident;
            }
        }
    }
    __syncthreads();

    for (int si = 0; si < ch.n; si++) {
        const int dil = ch.s[si].dilate, L = ch.s[si].L, R = ch.s[si].R;
        const uint32_t id1 = dil ? 0u : 0xffffu;
        // ---- horizontal pass A -> B : words j with 2j-L >= 0 and 2j+1+R <= BW-1
        {
            const int j_lo = (L + 1) / 2, j_hi = (BW - 2 - R) / 2;  // inclusive
            const int nj = j_hi - j_lo + 1;
            if (nj > 0) {
                for (int item = threadIdx.x; item < BH * nj; item += kThreads) {
                    const int by = item / nj, j = j_lo + (item - by * nj);
                    const uint16_t* row = bufA + by * BW;
                    uint32_t common = id1;
                    for (int p = 2 * j + 1 - L; p <= 2 * j + R; p++) common = op1(common, row[p], dil);
                    const uint32_t lo = op1(common, row[2 * j - L], dil);
                    const uint32_t hi = op1(common, row[2 * j + 1 + R], dil);
                    reinterpret_cast<uint32_t*>(bufB + by * BW)[j] = lo | (hi << 16);
                }
            }
        }
        __syncthreads();
        // ---- vertical pass B -> A : rows y with y-L >= 0 and y+R <= BH-1; packed u16x2
        {
            const int y_lo = L, y_hi = BH - 1 - R;
            const int ny = y_hi - y_lo + 1;
            const uint32_t id2 = dil ? 0u : 0xffffffffu;
            const int next_dil = (si + 1 < ch.n) ? ch.s[si + 1].dilate : dil;
            const uint32_t next_id = next_dil ? 0u : 0xffffu;
            const bool last = (si + 1 == ch.n);
            if (ny > 0) {
                for (int item = threadIdx.x; item < ny * BWW; item += kThreads) {
                    const int yy = item / BWW, j = item - yy * BWW;
                    const int by = y_lo + yy;
                    const uint32_t* col = reinterpret_cast<const uint32_t*>(bufB) + j;
                    uint32_t acc = id2;
                    for (int o = -L; o <= R; o++) acc = op2(acc, col[(by + o) * BWW], dil);
                    if (!last) {
                        // positions outside the image act as the next stage's identity
                        const int gy = gy_org + by, gx = gx_org + 2 * j;
                        const bool rowin = (unsigned)gy < (unsigned)h;
                        uint32_t lo = acc & 0xffffu, hi = acc >> 16;
                        if (!rowin || (unsigned)gx >= (unsigned)w) lo = next_id;
                        if (!rowin || (unsigned)(gx + 1) >= (unsigned)w) hi = next_id;
                        acc = lo | (hi << 16);
                    }
                    reinterpret_cast<uint32_t*>(bufA)[by * BWW + j] = acc;
                }
            }
        }
        __syncthreads();
    }

    // ---- store the TW x TH core: 8 pixels (u16) / 16 pixels (u8) per thread-item
    {
        const int vec_per_row = TW / VEC;
        const bool row_aligned = ((w % VEC) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
        for (int v = threadIdx.x; v < TH * vec_per_row; v += kThreads) {
            const int ty = v / vec_per_row, vx = v - ty * vec_per_row;
            const int gy = y0 + ty, gx = x0 + vx * VEC;
            if (gy >= h || gx >= w) continue;
            const uint16_t* s = bufA + (ch.HL + ty) * BW + HA + vx * VEC;
            T* d = dst + (int64_t)gy * w + gx;
            if (row_aligned && gx + VEC <= w) {
                uint4 q;
                T* e = reinterpret_cast<T*>(&q);
#pragma unroll
                for (int i = 0; i < VEC; i++) e[i] = (T)s[i];
                *reinterpret_cast<uint4*>(d) = q;
            } else {
                for (int i = 0; i < VEC && gx + i < w; i++) d[i] = (T)s[i];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// arbitrary structuring element (ellipse, cross): per footprint row a horizontal run [c0,c1]
struct SeRows {
    int k, anchor;
    int8_t c0[YAM_MAX_SE], c1[YAM_MAX_SE];  // inclusive column range per row, c0 > c1 = empty
};

template <typename T>
__global__ void __launch_bounds__(kThreads) morph_se_kernel(const T* __restrict__ src, T* __restrict__ dst,
                                                            int h, int w, SeRows se, int dilate) {
    const int k = se.k, a = se.anchor;
    const int BW = TW + k - 1, BH = TH + k - 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint16_t* buf = reinterpret_cast<uint16_t*>(smem_raw);
    const int64_t frame = blockIdx.z;
    src += frame * (int64_t)h * w;
    dst += frame * (int64_t)h * w;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const uint16_t ident = dilate ? 0 : 0xffff;
    for (int i = threadIdx.x; i < BW * BH; i += kThreads) {
        const int by = i / BW, bx = i - by * BW;
        const int gy = y0 - a + by, gx = x0 - a + bx;
        buf[i] = ((unsigned)gy < (unsigned)h && (unsigned)gx < (unsigned)w) ? (uint16_t)src[(int64_t)gy * w + gx] : ident;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TW * TH; i += kThreads) {
        const int ty = i / TW, tx = i - ty * TW;
        const int gy = y0 + ty, gx = x0 + tx;
        if (gy >= h || gx >= w) continue;
        uint32_t acc = ident;
        for (int r = 0; r < k; r++) {
            const uint16_t* row = buf + (ty + r) * BW + tx;
            for (int c = se.c0[r]; c <= se.c1[r]; c++) acc = op1(acc, row[c], dilate);
        }
        dst[(int64_t)gy * w + gx] = (T)acc;
    }
}

// ------------------------------------------------------------------------------------------------
// host-side planning

size_t chain_smem(const Chain& ch, int elem_size) {
    const int VEC = 16 / elem_size;
    const int halo = ch.HL > ch.HR ? ch.HL : ch.HR;
    const int HA = (halo + 1 + VEC - 1) / VEC * VEC;
    const int BW = TW + 2 * HA, BH = TH + ch.HL + ch.HR;
    return (size_t)2 * BH * BW * sizeof(uint16_t);
}

template <typename T>
int launch_chain(yam_ctx* ctx, const T* src, T* dst, int64_t n, int64_t h, int64_t w, const Chain& ch) {
    const size_t smem = chain_smem(ch, sizeof(T));
    if (smem > 48 * 1024)
        YAM_CUDA(cudaFuncSetAttribute(morph_rect_chain_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((w + TW - 1) / TW), (unsigned)((h + TH - 1) / TH), (unsigned)n);
    morph_rect_chain_kernel<T><<<grid, kThreads, smem, ctx->stream>>>(src, dst, (int)h, (int)w, ch);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

// Split a list of stages into launches whose summed halo stays <= kMaxHalo per side.
// Intermediate results ping-pong between dst and a scratch image.
template <typename T>
int run_rect_stages(yam_ctx* ctx, const T* src, T* dst, int64_t n, int64_t h, int64_t w,
                    const Stage* stages, int count) {
    // expand stages so that each piece fits the halo budget
    Stage pieces[64];
    int np = 0;
    for (int i = 0; i < count; i++) {
        int L = stages[i].L, R = stages[i].R;
        if (L == 0 && R == 0) continue;
        while (L > 0 || R > 0) {
            const int l = L < kMaxHalo ? L : kMaxHalo, r = R < kMaxHalo ? R : kMaxHalo;
            YAM_REQUIRE(np < 64, "morph: window too large");
            pieces[np++] = Stage{stages[i].dilate, l, r};
            L -= l;
            R -= r;
        }
    }
    if (np == 0) {
        if ((const void*)src != (void*)dst)
            YAM_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * h * w * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
        return YAM_OK;
    }
    // group pieces into launches
    Chain chains[64];
    int nc = 0;
    Chain cur = {};
    for (int i = 0; i < np; i++) {
        const bool fits = cur.n < kMaxStages && cur.HL + pieces[i].L <= kMaxHalo && cur.HR + pieces[i].R <= kMaxHalo;
        if (cur.n > 0 && !fits) {
            chains[nc++] = cur;
            cur = Chain{};
        }
        // merge with previous piece of the same op inside a chain
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
    if (nc == 1) {
        const Chain& c = chains[0];
        bool symmetric = c.n <= 3;
        int dil[3] = {0, 0, 0}, rad[3] = {0, 0, 0};
        for (int i = 0; i < c.n && symmetric; i++) {
            symmetric = c.s[i].L == c.s[i].R;
            dil[i] = c.s[i].dilate;
            rad[i] = c.s[i].L;
        }
        if (symmetric) {
            const int rc = yam_morph_fast_try<T>(ctx, src, dst, n, h, w, c.n, dil, rad);
            if (rc != 0) return rc < 0 ? rc : YAM_OK;
        }
        return launch_chain<T>(ctx, src, dst, n, h, w, c);
    }
    // multi-launch: need a scratch image; the final launch must land in dst
    void* scratch = nullptr;
    if (int rc = yam_scratch(ctx, (size_t)n * h * w * sizeof(T), &scratch)) return rc;
    T* tmp = (T*)scratch;
    const T* in = src;
    for (int i = 0; i < nc; i++) {
        // choose outputs so that launch nc-1 writes dst: alternate backwards
        T* out = ((nc - 1 - i) % 2 == 0) ? dst : tmp;
        if (int rc = launch_chain<T>(ctx, in, out, n, h, w, chains[i])) return rc;
        in = out;
    }
    return YAM_OK;
}

template <typename T>
int run_se(yam_ctx* ctx, const T* src, T* dst, int64_t n, int64_t h, int64_t w, const SeRows& se, int dilate) {
    const size_t smem = (size_t)(TW + se.k - 1) * (TH + se.k - 1) * sizeof(uint16_t);
    dim3 grid((unsigned)((w + TW - 1) / TW), (unsigned)((h + TH - 1) / TH), (unsigned)n);
    morph_se_kernel<T><<<grid, kThreads, smem, ctx->stream>>>(src, dst, (int)h, (int)w, se, dilate);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

template <typename T>
int morph_typed(yam_ctx* ctx, const T* src, T* dst, int64_t n, int64_t h, int64_t w, int op, int shape,
                int k, int iterations) {
    const int a = k / 2;
    bool rect = (shape == YAM_SHAPE_RECT);
    uint8_t se_bytes[YAM_MAX_SE * YAM_MAX_SE];
    SeRows se;
    if (!rect) {
        yam_host_structuring_element(shape, k, se_bytes);
        bool all = true;
        for (int i = 0; i < k * k; i++) all = all && se_bytes[i];
        rect = all;
        se.k = k;
        se.anchor = a;
        for (int r = 0; r < k; r++) {
            int c0 = k, c1 = -1;
            for (int c = 0; c < k; c++)
                if (se_bytes[r * k + c]) {
                    if (c < c0) c0 = c;
                    c1 = c;
                }
            // ellipse and cross rows are contiguous runs except the cross's off-centre rows,
            // which are single pixels: both are runs.
            se.c0[r] = (int8_t)c0;
            se.c1[r] = (int8_t)c1;
        }
    }
    // op sequence: erode^n, dilate^n, open = E^n D^n, close = D^n E^n
    int seq[2];
    int ns = 0;
    switch (op) {
        case YAM_MORPH_ERODE: seq[ns++] = 0; break;
        case YAM_MORPH_DILATE: seq[ns++] = 1; break;
        case YAM_MORPH_OPEN: seq[ns++] = 0; seq[ns++] = 1; break;
        case YAM_MORPH_CLOSE: seq[ns++] = 1; seq[ns++] = 0; break;
        default: YAM_REQUIRE(false, "morph: unknown op %d", op);
    }
    if (rect) {
        Stage st[2];
        for (int i = 0; i < ns; i++) st[i] = Stage{seq[i], a * iterations, (k - 1 - a) * iterations};
        return run_rect_stages<T>(ctx, src, dst, n, h, w, st, ns);
    }
    // non-separable: one launch per iteration, ping-pong through scratch
    const int total = ns * iterations;
    void* scratch = nullptr;
    if (total > 1)
        if (int rc = yam_scratch(ctx, (size_t)n * h * w * sizeof(T), &scratch)) return rc;
    T* tmp = (T*)scratch;
    const T* in = src;
    int idx = 0;
    for (int s = 0; s < ns; s++)
        for (int it = 0; it < iterations; it++, idx++) {
            T* out = ((total - 1 - idx) % 2 == 0) ? dst : tmp;
            if (int rc = run_se<T>(ctx, in, out, n, h, w, se, seq[s])) return rc;
            in = out;
        }
    return YAM_OK;
}

int check_args(const void* src, const void* dst, int64_t n, int64_t h, int64_t w, int dtype, int k, int it) {
    YAM_REQUIRE(src && dst && src != dst, "morph: NULL or aliased image pointers");
    YAM_REQUIRE(n > 0 && h > 0 && w > 0 && n <= 65535, "morph: bad shape");
    YAM_REQUIRE(dtype == YAM_U8 || dtype == YAM_U16, "morph: unsupported dtype %d", dtype);
    YAM_REQUIRE(k >= 1 && k <= YAM_MAX_SE, "morph: kernel_size must be in [1,%d], got %d", YAM_MAX_SE, k);
    YAM_REQUIRE(it >= 1 && it <= 64, "morph: iterations must be in [1,64], got %d", it);
    return YAM_OK;
}

}  // namespace

extern "C" {

int yam_morph(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w, int dtype, int op,
              int shape, int ksize, int iterations) {
    if (int rc = yam_enter(ctx)) return rc;
    if (int rc = check_args(src, dst, n, h, w, dtype, ksize, iterations)) return rc;
    YAM_REQUIRE(shape >= 0 && shape <= 2, "morph: unknown shape %d", shape);
    if (dtype == YAM_U8)
        return morph_typed<uint8_t>(ctx, (const uint8_t*)src, (uint8_t*)dst, n, h, w, op, shape, ksize, iterations);
    return morph_typed<uint16_t>(ctx, (const uint16_t*)src, (uint16_t*)dst, n, h, w, op, shape, ksize, iterations);
}

int yam_morph_open_close(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w, int dtype,
                         int ksize, int iterations) {
    if (int rc = yam_enter(ctx)) return rc;
    if (int rc = check_args(src, dst, n, h, w, dtype, ksize, iterations)) return rc;
    const int a = ksize / 2;
    const int L = a * iterations, R = (ksize - 1 - a) * iterations;
    // open = E D ; close = D E  ->  E, D, D, E (adjacent dilations merge inside a launch)
    Stage st[4] = {{0, L, R}, {1, L, R}, {1, L, R}, {0, L, R}};
    if (dtype == YAM_U8) return run_rect_stages<uint8_t>(ctx, (const uint8_t*)src, (uint8_t*)dst, n, h, w, st, 4);
    return run_rect_stages<uint16_t>(ctx, (const uint16_t*)src, (uint16_t*)dst, n, h, w, st, 4);
}

}  // extern "C"
