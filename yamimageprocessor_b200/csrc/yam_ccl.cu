// K10 connected-component labelling and K11 region properties.
//
// CCL (8-connectivity, raster-first canonical numbering) as a union-find over RUN SEGMENTS:
//   1. pack     the u8 mask is packed to 1 bit/pixel; every maximal run of set bits inside one
//               32-pixel word is a node whose id is the (padded) linear index of its first pixel;
//               parent[id] = id.  Nodes are ~10-30x fewer than pixels, and only they are ever
//               touched by the union-find, so its traffic stays in L2.
//   2. union    one thread per word links each of its segments to the segment that continues it
//               in the previous word and to every 8-connected segment in the row above
//               (lock-free union by minimum index with atomicMin).
//   3. flatten  parent[id] = root(id); segments that are their own root set a bit in a root mask.
//               Linking by minimum index makes the root the component's first pixel in raster
//               order, so rank(root) among roots IS the canonical label.
//   4. scan     popcount prefix over the root mask (per frame) -> rank of every root.
//   5. final    one thread per word writes the 32 int32 labels (1 + rank of the segment's root).
// HBM traffic: 1 B/px mask read + 4 B/px label write + O(segments).
//
// Region properties: one pass over labels (+ intensity); every thread folds the runs of equal
// label inside its 8-pixel chunk and issues one set of 64-bit atomics per run, bbox atomics are
// skipped when a (monotonic) pre-read shows they cannot change the value.
#include "yam_common.cuh"

namespace {

constexpr int kThreads = 256;

struct CclGeom {
    int h, w;
    int wpr;        // words per row
    int wp;         // padded row pitch in pixels = 32 * wpr
    int64_t words_per_frame;
    int64_t total_words;  // frames * words_per_frame
};

__device__ __forceinline__ int ld_parent(const int* p) { return __ldcg(p); }

__device__ __forceinline__ int find_root(const int* __restrict__ P, int x) {
    while (true) {
        const int p = ld_parent(P + x);
        if (p == x) return x;
        x = p;
    }
}

__device__ __forceinline__ void unite(int* __restrict__ P, int a, int b) {
    while (true) {
        a = find_root(P, a);
        b = find_root(P, b);
        if (a == b) return;
        if (a > b) {
            const int t = a;
            a = b;
            b = t;
        }
        const int old = atomicMin(P + b, a);
        if (old == b) return;
        b = old;
    }
}

// start (bit index) of the run of ones in `wv` that contains bit `b` (bit b must be set)
__device__ __forceinline__ int run_start(uint32_t wv, int b) {
    const uint32_t zeros_below = ~wv & ((1u << b) - 1u);
    return zeros_below ? 32 - __clz(zeros_below) : 0;
}

// ---- 1. pack + init ----------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) ccl_pack_kernel(const uint8_t* __restrict__ mask, CclGeom g,
                                                            uint32_t* __restrict__ bits, int* __restrict__ P) {
    const int64_t gw = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gw >= g.total_words) return;
    const int64_t frame = gw / g.words_per_frame;
    const int64_t wf = gw - frame * g.words_per_frame;
    const int y = (int)(wf / g.wpr), j = (int)(wf - (int64_t)y * g.wpr);
    const uint8_t* row = mask + (frame * g.h + y) * (int64_t)g.w;
    const int x0 = j * 32;
    uint32_t b = 0;
    if (x0 + 32 <= g.w && ((reinterpret_cast<uintptr_t>(row + x0) & 15) == 0)) {
        const uint4 q0 = yam_ld_stream(reinterpret_cast<const uint4*>(row + x0));
        const uint4 q1 = yam_ld_stream(reinterpret_cast<const uint4*>(row + x0) + 1);
        const uint32_t wd[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t v = wd[i];
            b |= ((v & 0xffu) ? 1u : 0u) << (4 * i);
            b |= ((v & 0xff00u) ? 1u : 0u) << (4 * i + 1);
            b |= ((v & 0xff0000u) ? 1u : 0u) << (4 * i + 2);
            b |= ((v & 0xff000000u) ? 1u : 0u) << (4 * i + 3);
        }
    } else {
        for (int i = 0; i < 32 && x0 + i < g.w; i++) b |= (row[x0 + i] ? 1u : 0u) << i;
    }
    bits[gw] = b;
    // segment starts: bit set and the bit below clear (bit 0 starts a segment of this word)
    uint32_t starts = b & ~(b << 1);
    int* Pf = P + frame * (int64_t)g.h * g.wp;
    const int base = (y * g.wpr + j) * 32;
    while (starts) {
        const int s = __ffs(starts) - 1;
        starts &= starts - 1;
        Pf[base + s] = base + s;
    }
}

// ---- 2. union ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) ccl_union_kernel(const uint32_t* __restrict__ bits, CclGeom g,
                                                             int* __restrict__ P) {
    const int64_t gw = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gw >= g.total_words) return;
    const uint32_t b = bits[gw];
    if (!b) return;
    const int64_t frame = gw / g.words_per_frame;
    const int64_t wf = gw - frame * g.words_per_frame;
    const int y = (int)(wf / g.wpr), j = (int)(wf - (int64_t)y * g.wpr);
    int* Pf = P + frame * (int64_t)g.h * g.wp;
    const int base = (y * g.wpr + j) * 32;
    // horizontal: the segment at bit 0 continues the segment that ends at bit 31 of the previous word
    if ((b & 1u) && j > 0) {
        const uint32_t pv = bits[gw - 1];
        if (pv >> 31) unite(Pf, base, base - 32 + run_start(pv, 31));
    }
    if (y == 0) return;
    // vertical: 34-column window of the row above; bit k of `above` <-> column 32*j + k - 1
    const uint32_t u = bits[gw - g.wpr];
    const uint32_t up = j > 0 ? bits[gw - g.wpr - 1] : 0u;
    const uint32_t un = j + 1 < g.wpr ? bits[gw - g.wpr + 1] : 0u;
    const unsigned long long above =
        (unsigned long long)(up >> 31) | ((unsigned long long)u << 1) | ((unsigned long long)(un & 1u) << 33);
    if (!above) return;
    const int base_up = base - g.wp;  // node id of bit 0 of word (y-1, j)
    uint32_t starts = b & ~(b << 1);
    while (starts) {
        const int s = __ffs(starts) - 1;
        starts &= starts - 1;
        // segment [s, e]
        const uint32_t from_s = b >> s;
        const int len = (~from_s) ? __ffs(~from_s) - 1 : 32 - s;
        const int e = s + len - 1;
        // window bits s .. e+2
        const unsigned long long wmask = ((e + 3 >= 64) ? ~0ull : ((1ull << (e + 3)) - 1ull)) & ~((1ull << s) - 1ull);
        unsigned long long m = above & wmask;
        while (m) {
            const int k0 = __ffsll((long long)m) - 1;
            // clear the lowest run of ones
            m &= m + (1ull << k0);
            int node;
            if (k0 == 0) {
                node = base_up - 32 + run_start(up, 31);
            } else if (k0 == 33) {
                node = base_up + 32;  // bit 0 of the next word starts its segment
            } else {
                node = base_up + run_start(u, k0 - 1);
            }
            unite(Pf, base + s, node);
        }
    }
}

// ---- 3. flatten + root mask + per-block root counts ----------------------------------------------
__global__ void __launch_bounds__(kThreads) ccl_flatten_kernel(const uint32_t* __restrict__ bits, CclGeom g,
                                                               int* __restrict__ P, uint32_t* __restrict__ rootmask,
                                                               uint32_t* __restrict__ blocksums, int blocks_per_frame) {
    // grid: (blocks_per_frame, frames)
    const int64_t wf = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t frame = blockIdx.y;
    uint32_t roots = 0;
    if (wf < g.words_per_frame) {
        const int64_t gw = frame * g.words_per_frame + wf;
        const uint32_t b = bits[gw];
        if (b) {
            int* Pf = P + frame * (int64_t)g.h * g.wp;
            const int base = (int)wf * 32;
            uint32_t starts = b & ~(b << 1);
            while (starts) {
                const int s = __ffs(starts) - 1;
                starts &= starts - 1;
                const int r = find_root(Pf, base + s);
                if (r == base + s) roots |= 1u << s;
                else Pf[base + s] = r;
            }
        }
        rootmask[gw] = roots;
    }
    // block sum of popcounts
    uint32_t c = __popc(roots);
    c = yam_warp_sum(c);
    __shared__ uint32_t s_c[kThreads / 32];
    if ((threadIdx.x & 31) == 0) s_c[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < kThreads / 32; i++) t += s_c[i];
        blocksums[frame * blocks_per_frame + blockIdx.x] = t;
    }
}

// ---- 4a. exclusive scan of block sums, one block per frame --------------------------------------
__global__ void __launch_bounds__(1024) ccl_scan_blocks_kernel(uint32_t* __restrict__ blocksums, int blocks_per_frame,
                                                               int32_t* __restrict__ counts) {
    uint32_t* bs = blocksums + (int64_t)blockIdx.x * blocks_per_frame;
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < blocks_per_frame; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < blocks_per_frame ? bs[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t wv = s_warp[lane];
            uint32_t wi = wv;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += up;
            }
            s_warp[lane] = wi - wv;  // exclusive
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        const uint32_t excl = carry + s_warp[warp] + incl - v;
        if (i < blocks_per_frame) bs[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0 && counts) counts[blockIdx.x] = (int32_t)s_carry;
}

// ---- 4b. per-word exclusive prefix -----------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) ccl_word_prefix_kernel(const uint32_t* __restrict__ rootmask, CclGeom g,
                                                                   const uint32_t* __restrict__ blocksums,
                                                                   int blocks_per_frame,
                                                                   uint32_t* __restrict__ wordprefix) {
    const int64_t wf = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t frame = blockIdx.y;
    const int64_t gw = frame * g.words_per_frame + wf;
    const uint32_t c = wf < g.words_per_frame ? __popc(rootmask[gw]) : 0u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    __shared__ uint32_t s_w[kThreads / 32];
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t off = blocksums[frame * blocks_per_frame + blockIdx.x];
    for (int i = 0; i < warp; i++) off += s_w[i];
    if (wf < g.words_per_frame) wordprefix[gw] = off + incl - c;
}

// ---- 5. final labels ----------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) ccl_final_kernel(const uint32_t* __restrict__ bits, CclGeom g,
                                                             const int* P, const uint32_t* __restrict__ rootmask,
                                                             const uint32_t* __restrict__ wordprefix,
                                                             int32_t* labels) {
    const int64_t gw = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gw >= g.total_words) return;
    const int64_t frame = gw / g.words_per_frame;
    const int64_t wf = gw - frame * g.words_per_frame;
    const int y = (int)(wf / g.wpr), j = (int)(wf - (int64_t)y * g.wpr);
    const uint32_t b = bits[gw];
    const int* Pf = P + frame * (int64_t)g.h * g.wp;
    const uint32_t* rm = rootmask + frame * g.words_per_frame;
    const uint32_t* wpf = wordprefix + frame * g.words_per_frame;
    const int base = (int)wf * 32;
    int32_t out[32];
    int cur = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) {
        const bool set = (b >> i) & 1u;
        const bool start = set && (i == 0 || !((b >> (i - 1)) & 1u));
        if (start) {
            const int r = ld_parent(Pf + base + i);  // flattened: parent is the root (or itself)
            const uint32_t rw = (uint32_t)r >> 5, rb = (uint32_t)r & 31u;
            cur = 1 + (int)(wpf[rw] + __popc(rm[rw] & ((1u << rb) - 1u)));
        }
        out[i] = set ? cur : 0;
    }
    int32_t* drow = labels + (frame * g.h + y) * (int64_t)g.w + (int64_t)j * 32;
    const int valid = min(32, g.w - j * 32);
    if (valid == 32 && ((reinterpret_cast<uintptr_t>(drow) & 15) == 0)) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            *reinterpret_cast<int4*>(drow + 4 * i) = make_int4(out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < 32; i++)
            if (i < valid) drow[i] = out[i];
    }
}

// ------------------------------------------------------------------------------------------------
// region properties
constexpr int kArea = 0, kSumR = 1, kSumC = 2, kSumI = 3, kMinR = 4, kMinC = 5, kMaxR = 6, kMaxC = 7;

__global__ void props_init_kernel(long long* __restrict__ props, int64_t n_labels) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_labels * YAM_PROPS_STRIDE) return;
    const int f = (int)(i % YAM_PROPS_STRIDE);
    props[i] = (f == kMinR || f == kMinC) ? 0x7fffffffffffffffLL : 0LL;
}

// Per-thread open run: the label last seen by this thread with its partial sums.  A thread owns an
// 8-pixel-wide column strip over a band of rows, so consecutive rows of the same region merge in
// registers and each region costs one flush per strip instead of one per row.
struct OpenRun {
    int lab;
    uint32_t area, sumr, sumc, sumi;  // band <= 32 rows x 8 px: 32-bit partials cannot overflow
    int minr, maxr, minc, maxc;
};

__device__ __forceinline__ void flush_open(long long* __restrict__ props, int64_t n_labels, const OpenRun& r) {
    if (r.lab <= 0 || r.lab > n_labels) return;
    long long* p = props + (int64_t)(r.lab - 1) * YAM_PROPS_STRIDE;
    atomicAdd((unsigned long long*)&p[kArea], (unsigned long long)r.area);
    atomicAdd((unsigned long long*)&p[kSumR], (unsigned long long)r.sumr);
    atomicAdd((unsigned long long*)&p[kSumC], (unsigned long long)r.sumc);
    if (r.sumi) atomicAdd((unsigned long long*)&p[kSumI], (unsigned long long)r.sumi);
    atomicMin(&p[kMinR], (long long)r.minr);
    atomicMax(&p[kMaxR], (long long)r.maxr);
    atomicMin(&p[kMinC], (long long)r.minc);
    atomicMax(&p[kMaxC], (long long)r.maxc);
}

constexpr int kBand = 32;   // rows per block band
constexpr int kRowsPerIter = 4;

template <typename TI>
__global__ void __launch_bounds__(kThreads) props_kernel(const int32_t* __restrict__ labels,
                                                         const TI* __restrict__ intensity, int h, int w,
                                                         int64_t n_labels, long long* __restrict__ props) {
    const int chunks = (w + 7) / 8;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= chunks) return;
    const int x0 = c * 8;
    const bool aligned = (w % 8) == 0 && ((reinterpret_cast<uintptr_t>(labels) & 15) == 0);
    const bool ialigned = intensity && (w % 8) == 0 && ((reinterpret_cast<uintptr_t>(intensity) & (8 * sizeof(TI) - 1)) == 0);
    OpenRun run;
    run.lab = 0;
    run.area = run.sumr = run.sumc = run.sumi = 0;
    run.minr = run.maxr = run.minc = run.maxc = 0;
    for (int64_t band = blockIdx.y; band * kBand < h; band += gridDim.y) {
        const int y_begin = (int)(band * kBand);
        const int y_end = min(h, y_begin + kBand);
        for (int yb = y_begin; yb < y_end; yb += kRowsPerIter) {
            int32_t l[kRowsPerIter][8];
            // issue all label loads of this iteration first (memory-level parallelism)
#pragma unroll
            for (int r = 0; r < kRowsPerIter; r++) {
                const int y = yb + r;
                if (y < y_end) {
                    const int32_t* lrow = labels + (int64_t)y * w;
                    if (aligned) {
                        const int4 a = __ldcs(reinterpret_cast<const int4*>(lrow + x0));
                        const int4 b = __ldcs(reinterpret_cast<const int4*>(lrow + x0 + 4));
                        l[r][0] = a.x; l[r][1] = a.y; l[r][2] = a.z; l[r][3] = a.w;
                        l[r][4] = b.x; l[r][5] = b.y; l[r][6] = b.z; l[r][7] = b.w;
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; i++) l[r][i] = (x0 + i < w) ? lrow[x0 + i] : 0;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; i++) l[r][i] = 0;
                }
            }
#pragma unroll
            for (int r = 0; r < kRowsPerIter; r++) {
                const int y = yb + r;
                uint32_t any = 0;
#pragma unroll
                for (int i = 0; i < 8; i++) any |= (uint32_t)l[r][i];
                if (!any) continue;
                uint32_t iv[8];
                if (ialigned) {
                    const TI* irow = intensity + (int64_t)y * w + x0;
                    if (sizeof(TI) == 2) {
                        const uint4 q = __ldcs(reinterpret_cast<const uint4*>(irow));
                        iv[0] = q.x & 0xffffu; iv[1] = q.x >> 16; iv[2] = q.y & 0xffffu; iv[3] = q.y >> 16;
                        iv[4] = q.z & 0xffffu; iv[5] = q.z >> 16; iv[6] = q.w & 0xffffu; iv[7] = q.w >> 16;
                    } else {
                        const uint2 q = __ldcs(reinterpret_cast<const uint2*>(irow));
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            iv[i] = (q.x >> (8 * i)) & 0xffu;
                            iv[4 + i] = (q.y >> (8 * i)) & 0xffu;
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; i++)
                        iv[i] = (intensity && x0 + i < w) ? (uint32_t)intensity[(int64_t)y * w + x0 + i] : 0u;
                }
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int lab = l[r][i];
                    if (lab == 0) continue;
                    const int x = x0 + i;
                    if (lab != run.lab) {
                        flush_open(props, n_labels, run);
                        run.lab = lab;
                        run.area = run.sumr = run.sumc = run.sumi = 0;
                        run.minr = y;
                        run.minc = x;
                        run.maxc = x + 1;
                    }
                    run.area += 1;
                    run.sumr += (uint32_t)y;
                    run.sumc += (uint32_t)x;
                    run.sumi += iv[i];
                    run.minc = min(run.minc, x);
                    run.maxc = max(run.maxc, x + 1);
                    run.maxr = y + 1;
                }
            }
        }
        flush_open(props, n_labels, run);
        run.lab = 0;
    }
}

}  // namespace

extern "C" {

int yam_ccl_label(yam_ctx* ctx, const void* mask, int32_t* labels, int64_t n, int64_t h, int64_t w,
                  int32_t* counts_dev, int32_t* counts_host) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(mask && labels && n > 0 && h > 0 && w > 0, "ccl: bad arguments");
    YAM_REQUIRE(n <= 65535, "ccl: at most 65535 frames per call");
    CclGeom g;
    g.h = (int)h;
    g.w = (int)w;
    g.wpr = (int)((w + 31) / 32);
    g.wp = g.wpr * 32;
    g.words_per_frame = (int64_t)h * g.wpr;
    g.total_words = g.words_per_frame * n;
    YAM_REQUIRE(g.words_per_frame * 32 < (1ll << 31), "ccl: frame too large for int32 labels (%lld x %lld)",
                (long long)h, (long long)w);
    const int blocks_per_frame = (int)((g.words_per_frame + kThreads - 1) / kThreads);
    const bool alias = (w % 32) == 0;  // labels buffer doubles as the parent array
    // scratch layout
    const size_t words_bytes = yam_align_up((size_t)g.total_words * 4, 256);
    const size_t bs_bytes = yam_align_up((size_t)n * blocks_per_frame * 4, 256);
    const size_t cnt_bytes = yam_align_up((size_t)n * 4, 256);
    const size_t p_bytes = alias ? 0 : yam_align_up((size_t)n * h * g.wp * 4, 256);
    void* scratch = nullptr;
    if (int rc = yam_scratch(ctx, 3 * words_bytes + bs_bytes + cnt_bytes + p_bytes, &scratch)) return rc;
    char* sp = (char*)scratch;
    uint32_t* bits = (uint32_t*)sp;
    uint32_t* rootmask = (uint32_t*)(sp + words_bytes);
    uint32_t* wordprefix = (uint32_t*)(sp + 2 * words_bytes);
    uint32_t* blocksums = (uint32_t*)(sp + 3 * words_bytes);
    int32_t* counts = counts_dev ? counts_dev : (int32_t*)(sp + 3 * words_bytes + bs_bytes);
    int* P = alias ? (int*)labels : (int*)(sp + 3 * words_bytes + bs_bytes + cnt_bytes);

    const unsigned gblocks = (unsigned)((g.total_words + kThreads - 1) / kThreads);
    ccl_pack_kernel<<<gblocks, kThreads, 0, ctx->stream>>>((const uint8_t*)mask, g, bits, P);
    YAM_LAUNCHED(ctx);
    ccl_union_kernel<<<gblocks, kThreads, 0, ctx->stream>>>(bits, g, P);
    YAM_LAUNCHED(ctx);
    ccl_flatten_kernel<<<dim3((unsigned)blocks_per_frame, (unsigned)n), kThreads, 0, ctx->stream>>>(
        bits, g, P, rootmask, blocksums, blocks_per_frame);
    YAM_LAUNCHED(ctx);
    ccl_scan_blocks_kernel<<<(unsigned)n, 1024, 0, ctx->stream>>>(blocksums, blocks_per_frame, counts);
    YAM_LAUNCHED(ctx);
    ccl_word_prefix_kernel<<<dim3((unsigned)blocks_per_frame, (unsigned)n), kThreads, 0, ctx->stream>>>(
        rootmask, g, blocksums, blocks_per_frame, wordprefix);
    YAM_LAUNCHED(ctx);
    ccl_final_kernel<<<gblocks, kThreads, 0, ctx->stream>>>(bits, g, P, rootmask, wordprefix, labels);
    YAM_LAUNCHED(ctx);
    if (counts_host) {
        YAM_CUDA(cudaMemcpyAsync(counts_host, counts, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
        YAM_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return YAM_OK;
}

int yam_region_props(yam_ctx* ctx, const int32_t* labels, const void* intensity, int intensity_dtype, int64_t h,
                     int64_t w, int64_t n_labels, int64_t* props_dev) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(labels && h > 0 && w > 0 && n_labels >= 0, "region_props: bad arguments");
    YAM_REQUIRE(h < (1 << 30) && w < (1 << 30), "region_props: image side too large");
    if (n_labels == 0) return YAM_OK;
    YAM_REQUIRE(props_dev, "region_props: props_dev is NULL");
    YAM_REQUIRE(!intensity || intensity_dtype == YAM_U8 || intensity_dtype == YAM_U16,
                "region_props: unsupported intensity dtype %d", intensity_dtype);
    long long* props = (long long*)props_dev;
    const int64_t cells = n_labels * YAM_PROPS_STRIDE;
    props_init_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, ctx->stream>>>(props, n_labels);
    YAM_LAUNCHED(ctx);
    const int chunks = (int)((w + 7) / 8);
    unsigned gx = (unsigned)((chunks + kThreads - 1) / kThreads);
    const int64_t bands = (h + kBand - 1) / kBand;
    unsigned gy = (unsigned)(bands < 65535 ? bands : 65535);
    dim3 grid(gx, gy, 1);
    if (!intensity || intensity_dtype == YAM_U16)
        props_kernel<uint16_t><<<grid, kThreads, 0, ctx->stream>>>(labels, (const uint16_t*)intensity, (int)h, (int)w, n_labels, props);
    else
        props_kernel<uint8_t><<<grid, kThreads, 0, ctx->stream>>>(labels, (const uint8_t*)intensity, (int)h, (int)w, n_labels, props);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

}  // extern "C"
