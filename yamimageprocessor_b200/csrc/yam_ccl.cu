// K10 connected-component labelling and K11 region properties.
//
// CCL (8-connectivity, raster-first canonical numbering) as a union-find over RUN SEGMENTS:
//   pack      (u8 masks only) the mask is packed to 1 bit/pixel; the fused path hands bits in directly.
//   scan      every maximal run of set bits inside one 32-pixel word is a node.  Nodes are numbered
//             COMPACTLY in raster order by a single-pass chained scan (decoupled look-back) of the
//             per-word segment counts, so the parent array has one int per segment (~1-2 % of the
//             pixels) and every union-find access stays in L2.
//   tile      one CTA per 32-row x 32-word tile (32 x 1024 px): the tile's nodes get local ids, all
//             links between words of the tile are resolved by a union-find that lives entirely in
//             SHARED memory (atomicMin linking to the minimum index, path halving), and the flattened
//             result is written as global parent pointers.  Tiles with more than kTileCap nodes
//             (dense noise) run the same links on the global parent array instead.
//   border    only the links that cross a tile edge (top rows and the two edge word-columns of every
//             tile, ~3 % of the words) go through global atomics.
//   rank      one pass over the nodes: walk to the root (read-only), store it, count roots and rank
//             them with a chained scan.  Linking by minimum index makes the root the segment holding
//             the component's first pixel in raster order, so rank(root) + 1 IS the canonical label;
//             it is stored negated in the root's slot.
//   final     one thread per word resolves its segments (at most two dependent loads) and the block
//             writes the int32 labels through a swizzled shared-memory tile with 128-bit coalesced
//             stores.
// HBM traffic: 1 B/px mask read (1/8 from bits) + 4 B/px label write + O(words) bookkeeping.
//
// Region properties: threads own an 8-pixel-wide column strip over a band of rows and keep the
// current label's partial sums in registers; one set of 64-bit atomics per label change.
#include "yam_common.cuh"

namespace {

constexpr int kThreads = 256;

struct CclGeom {
    int h, w;
    int wpr;                  // words per row
    int64_t words_per_frame;
    int64_t total_words;      // frames * words_per_frame
    int frames;
};

__device__ __forceinline__ int find_root(int* __restrict__ P, int x) {
    // path halving: every visited node is re-pointed to its grandparent (plain stores are safe:
    // parents only ever move to smaller ancestors of the same set)
    int p = __ldcg(P + x);
    while (p != x) {
        const int gp = __ldcg(P + p);
        if (gp != p) P[x] = gp;
        x = p;
        p = gp;
    }
    return x;
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

// the same union-find on a shared-memory parent array (tile-local node ids)
__device__ __forceinline__ int s_find_root(volatile int* lp, int x) {
    int p = lp[x];
    while (p != x) {
        const int gp = lp[p];
        if (gp != p) lp[x] = gp;
        x = p;
        p = gp;
    }
    return x;
}
__device__ __forceinline__ void s_unite(int* lp, int a, int b) {
    while (true) {
        a = s_find_root(lp, a);
        b = s_find_root(lp, b);
        if (a == b) return;
        if (a > b) {
            const int t = a;
            a = b;
            b = t;
        }
        const int old = atomicMin(lp + b, a);
        if (old == b) return;
        b = old;
    }
}

__device__ __forceinline__ uint32_t seg_starts(uint32_t b) { return b & ~(b << 1); }
__device__ __forceinline__ int seg_count(uint32_t b) { return __popc(seg_starts(b)); }

// start (bit index) of the run of ones in `wv` that contains bit `b` (bit b must be set)
__device__ __forceinline__ int run_start(uint32_t wv, int b) {
    const uint32_t zeros_below = ~wv & ((1u << b) - 1u);
    return zeros_below ? 32 - __clz(zeros_below) : 0;
}

// index (within its word) of the segment of `wv` that contains bit `b`
__device__ __forceinline__ int seg_index(uint32_t wv, int b) {
    const int s = run_start(wv, b);
    return __popc(seg_starts(wv) & ((1u << s) - 1u));
}

// Links of one segment (start bit s, index k_self inside its word b) to the word on its left
// (`left`, same row) and to the three words above it (`up` = column j-1, `u` = column j, `un` =
// column j+1 of the previous row).  A neighbour that does not belong to the current pass is handed in
// as 0.  link(kind, k_other):
//   kind 0: last segment of `left`   kind 1: last segment of `up`
//   kind 2: segment k_other of `u`   kind 3: first segment of `un`
// A run of the row above that spans two of the words is linked through its first word only; the
// horizontal link inside that row connects the rest.
template <typename F>
__device__ __forceinline__ void seg_links(uint32_t b, int s, uint32_t left, unsigned long long above, uint32_t u, F&& link) {
    if (s == 0 && (left >> 31)) link(0, 0);
    if (!above) return;
    const uint32_t from_s = b >> s;
    const int len = (~from_s) ? __ffs(~from_s) - 1 : 32 - s;
    const int e = s + len - 1;
    // 8-connectivity: columns s-1 .. e+1 of the row above = bits s .. e+2 of `above`
    unsigned long long m = above & (((1ull << (e + 3)) - 1ull) & ~((1ull << s) - 1ull));
    while (m) {
        const int k0 = __ffsll((long long)m) - 1;
        m &= m + (1ull << k0);  // clear the lowest run of ones
        if (k0 == 0) link(1, 0);
        else if (k0 == 33) link(3, 0);
        else link(2, seg_index(u, k0 - 1));
    }
}
__device__ __forceinline__ unsigned long long above_window(uint32_t up, uint32_t u, uint32_t un) {
    return (unsigned long long)(up >> 31) | ((unsigned long long)u << 1) | ((unsigned long long)(un & 1u) << 33);
}
// all segments of a word: link(kind, k_self, k_other)
template <typename F>
__device__ __forceinline__ void word_links(uint32_t b, uint32_t left, uint32_t up, uint32_t u, uint32_t un, F&& link) {
    const unsigned long long above = above_window(up, u, un);
    uint32_t starts = seg_starts(b);
    int k = 0;
    while (starts) {
        const int s = __ffs(starts) - 1;
        starts &= starts - 1;
        seg_links(b, s, left, above, u, [&](int kind, int ko) { link(kind, k, ko); });
        k++;
    }
}

// exclusive prefix of v within a block of NT threads; *total (optional, shared) = block sum.
// Ends with the data in s_tmp still live: callers separate two scans with __syncthreads().
template <int NT = kThreads>
__device__ __forceinline__ uint32_t block_excl_scan_u32(uint32_t v, uint32_t* s_tmp, uint32_t* total = nullptr) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) s_tmp[warp] = incl;
    __syncthreads();
    uint32_t off = 0, all = 0;
#pragma unroll
    for (int i = 0; i < NT / 32; i++) {
        const uint32_t t = s_tmp[i];
        if (i < warp) off += t;
        all += t;
    }
    if (total && threadIdx.x == 0) *total = all;
    return off + incl - v;
}

// ---- ordered chunk prefixes ---------------------------------------------------------------------------
// The scans below run at most one chunk per SM, claimed in ticket order.  A chunk publishes its
// aggregate in a 64-bit status word (bit 63 = valid, low 32 = value) and obtains its exclusive prefix
// by summing the aggregates of ALL earlier chunks (<= num_sms values, one warp-parallel wait).  Every
// earlier ticket is held by a block that is already running, so the wait cannot deadlock.
constexpr unsigned long long kFlagValid = 1ull << 63;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Called by every thread of the block (contains __syncthreads); `aggregate` is read from thread 0.
__device__ uint32_t ordered_exclusive(unsigned long long* status, int chunk, uint32_t aggregate, uint32_t* s_prefix) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        if (lane == 0) st_status(status + chunk, kFlagValid | aggregate);
        uint32_t excl = 0;
        // 8 status words per lane are requested together (one L2 round trip instead of eight), then
        // the few that were not published yet are polled
        for (int base = 0; base < chunk; base += 256) {
            unsigned long long st[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int idx = base + 32 * j + lane;
                st[j] = idx < chunk ? ld_status(status + idx) : kFlagValid;
            }
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int idx = base + 32 * j + lane;
                while (!(st[j] & kFlagValid)) st[j] = ld_status(status + idx);
                excl += (uint32_t)st[j];
            }
        }
        excl = yam_warp_sum(excl);
        if (lane == 0) *s_prefix = excl;
    }
    __syncthreads();
    return *s_prefix;
}

// ---- 0. pack a u8 mask to bits (the fused path skips this) --------------------------------------
__global__ void __launch_bounds__(kThreads) ccl_pack_kernel(const uint8_t* __restrict__ mask, CclGeom g,
                                                            uint32_t* __restrict__ bits) {
    const int64_t gw = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gw >= g.total_words) return;
    const int64_t row_id = gw / g.wpr;  // frame * h + y
    const int j = (int)(gw - row_id * g.wpr);
    const uint8_t* row = mask + row_id * (int64_t)g.w;
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
}

// ---- 1. scan: nbase[word] = number of segments in all earlier words -------------------------------
constexpr int kBigThreads = 1024;
constexpr int kScanWords = 16;                           // words per thread and sub-chunk
constexpr int kScanSub = kBigThreads * kScanWords;       // words per sub-chunk

// segment counts of kScanWords consecutive words, packed 8 bits each (a word has at most 16 segments)
__device__ __forceinline__ void load_seg_counts(const uint32_t* __restrict__ bits, int64_t w0, int64_t end, bool aligned,
                                                uint32_t (&c)[kScanWords / 4]) {
    if (aligned && w0 + kScanWords <= end) {
        uint4 q[kScanWords / 4];
#pragma unroll
        for (int v = 0; v < kScanWords / 4; v++) q[v] = __ldcg(reinterpret_cast<const uint4*>(bits + w0) + v);
#pragma unroll
        for (int v = 0; v < kScanWords / 4; v++)
            c[v] = seg_count(q[v].x) | (seg_count(q[v].y) << 8) | (seg_count(q[v].z) << 16) | (seg_count(q[v].w) << 24);
    } else {
#pragma unroll
        for (int v = 0; v < kScanWords / 4; v++) {
            c[v] = 0;
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (w0 + 4 * v + i < end) c[v] |= (uint32_t)seg_count(__ldcg(bits + w0 + 4 * v + i)) << (8 * i);
        }
    }
}
__device__ __forceinline__ uint32_t packed_sum(const uint32_t (&c)[kScanWords / 4]) {
    uint32_t t = 0;
#pragma unroll
    for (int v = 0; v < kScanWords / 4; v++) t += c[v];   // byte lanes: 4 x 16 <= 64 each, no carry
    return (t & 0xffu) + ((t >> 8) & 0xffu) + ((t >> 16) & 0xffu) + (t >> 24);
}

// grid = chunks (<= num_sms); chunk c owns words [c * per, (c + 1) * per), per a multiple of kScanSub
__global__ void __launch_bounds__(kBigThreads) ccl_scan_kernel(const uint32_t* __restrict__ bits, int64_t total_words,
                                                               int64_t per, uint32_t* __restrict__ nbase,
                                                               unsigned long long* __restrict__ status,
                                                               unsigned int* __restrict__ ticket,
                                                               uint32_t* __restrict__ total_nodes) {
    __shared__ uint32_t s_tmp[kBigThreads / 32];
    __shared__ uint32_t s_prefix, s_total;
    __shared__ int s_chunk;
    if (threadIdx.x == 0) s_chunk = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int chunk = s_chunk;
    const int64_t begin = (int64_t)chunk * per;
    const int64_t end = begin + per < total_words ? begin + per : total_words;
    const bool aligned = (reinterpret_cast<uintptr_t>(bits) & 15) == 0;
    // pass 1: segments in this chunk (the first sub-chunk's counts stay in registers for pass 2)
    uint32_t c0[kScanWords / 4], c[kScanWords / 4];
    load_seg_counts(bits, begin + (int64_t)threadIdx.x * kScanWords, end, aligned, c0);
    uint32_t mine = packed_sum(c0);
    for (int64_t sub = begin + kScanSub; sub < end; sub += kScanSub) {
        load_seg_counts(bits, sub + (int64_t)threadIdx.x * kScanWords, end, aligned, c);
        mine += packed_sum(c);
    }
    block_excl_scan_u32<kBigThreads>(mine, s_tmp, &s_total);
    __syncthreads();
    const uint32_t agg = s_total;
    uint32_t carry = ordered_exclusive(status, chunk, agg, &s_prefix);
    if (threadIdx.x == 0 && end >= total_words) *total_nodes = carry + agg;
    // pass 2: prefixes (later sub-chunks re-read their bits from L2)
    for (int64_t sub = begin; sub < end; sub += kScanSub) {
        const int64_t w0 = sub + (int64_t)threadIdx.x * kScanWords;
        if (sub == begin) {
#pragma unroll
            for (int v = 0; v < kScanWords / 4; v++) c[v] = c0[v];
        } else {
            load_seg_counts(bits, w0, end, aligned, c);
        }
        __syncthreads();
        uint32_t run = carry + block_excl_scan_u32<kBigThreads>(packed_sum(c), s_tmp, &s_total);
        uint32_t o[kScanWords];
#pragma unroll
        for (int i = 0; i < kScanWords; i++) {
            o[i] = run;
            run += (c[i >> 2] >> (8 * (i & 3))) & 0xffu;
        }
        if (w0 + kScanWords <= end) {
#pragma unroll
            for (int v = 0; v < kScanWords / 4; v++)
                reinterpret_cast<uint4*>(nbase + w0)[v] = make_uint4(o[4 * v], o[4 * v + 1], o[4 * v + 2], o[4 * v + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < kScanWords; i++)
                if (w0 + i < end) nbase[w0 + i] = o[i];
        }
        __syncthreads();
        carry += s_total;
    }
}

// ---- 2. tile-local union-find ------------------------------------------------------------------------
constexpr int kTileR = 32;     // rows per tile
constexpr int kTileC = 32;     // words per tile row (1024 px)
constexpr int kTileWords = kTileR * kTileC;
constexpr int kTileCap = 2048; // nodes a tile can hold in shared memory (2 per word; denser tiles use global parents)
constexpr int kTilePer = kTileWords / kThreads;  // consecutive words per thread (same tile row)
static_assert(kTileC % kTilePer == 0, "a thread's words must share a tile row");

struct TileSmem {
    uint32_t bits[kTileWords];
    uint32_t lbase[kTileWords];   // tile-local id of the word's first segment
    uint32_t gbase[kTileWords];   // global id of the word's first segment
    int lp[kTileCap];             // local parents
    uint32_t info[kTileCap];      // local node -> (word index in tile << 5) | start bit
};

__global__ void __launch_bounds__(kThreads) ccl_tile_kernel(const uint32_t* __restrict__ bits,
                                                            const uint32_t* __restrict__ nbase, CclGeom g,
                                                            int tiles_x, int tiles_y, int* __restrict__ P) {
    extern __shared__ __align__(16) unsigned char tile_smem_raw[];
    TileSmem& S = *reinterpret_cast<TileSmem*>(tile_smem_raw);
    __shared__ uint32_t s_tmp[kThreads / 32];
    __shared__ uint32_t s_total;
    const int tx = blockIdx.x % tiles_x;
    const int ty = (blockIdx.x / tiles_x) % tiles_y;
    const int frame = blockIdx.x / (tiles_x * tiles_y);
    const int j0 = tx * kTileC, y0 = ty * kTileR;
    const int64_t fbase = (int64_t)frame * g.words_per_frame;
    for (int i = threadIdx.x; i < kTileWords; i += kThreads) {
        const int r = i / kTileC, c = i - r * kTileC;
        uint32_t b = 0, nb = 0;
        if (y0 + r < g.h && j0 + c < g.wpr) {
            const int64_t gw = fbase + (int64_t)(y0 + r) * g.wpr + j0 + c;
            b = __ldg(bits + gw);
            nb = __ldg(nbase + gw);
        }
        S.bits[i] = b;
        S.gbase[i] = nb;
    }
    __syncthreads();
    const int i0 = threadIdx.x * kTilePer;
    uint32_t wb[kTilePer];
    uint32_t mine = 0;
#pragma unroll
    for (int q = 0; q < kTilePer; q++) {
        wb[q] = S.bits[i0 + q];
        mine += seg_count(wb[q]);
    }
    uint32_t lb = block_excl_scan_u32(mine, s_tmp, &s_total);
#pragma unroll
    for (int q = 0; q < kTilePer; q++) {
        S.lbase[i0 + q] = lb;
        lb += seg_count(wb[q]);
    }
    __syncthreads();
    const int nl = (int)s_total;
    if (nl == 0) return;
    if (nl > kTileCap) {
        // dense tile: the same links, word-parallel, on the global parent array
#pragma unroll
        for (int q = 0; q < kTilePer; q++) {
            const int n = seg_count(wb[q]);
            const uint32_t gb = S.gbase[i0 + q];
            for (int k = 0; k < n; k++) P[gb + k] = (int)(gb + k);
        }
        __syncthreads();
        const int r = i0 / kTileC;
#pragma unroll
        for (int q = 0; q < kTilePer; q++) {
            const uint32_t b = wb[q];
            if (!b) continue;
            const int i = i0 + q, c = i - r * kTileC;
            const uint32_t left = c > 0 ? S.bits[i - 1] : 0u;
            uint32_t up = 0, u = 0, un = 0;
            if (r > 0) {
                up = c > 0 ? S.bits[i - kTileC - 1] : 0u;
                u = S.bits[i - kTileC];
                un = c + 1 < kTileC ? S.bits[i - kTileC + 1] : 0u;
            }
            const int self = (int)S.gbase[i];
            word_links(b, left, up, u, un, [&](int kind, int ks, int ko) {
                int other;
                if (kind == 0) other = (int)S.gbase[i - 1] + seg_count(left) - 1;
                else if (kind == 1) other = (int)S.gbase[i - kTileC - 1] + seg_count(up) - 1;
                else if (kind == 2) other = (int)S.gbase[i - kTileC] + ko;
                else other = (int)S.gbase[i - kTileC + 1];
                unite(P, self + ks, other);
            });
        }
        return;
    }
    // parents = identity, node -> (word, start bit)
#pragma unroll
    for (int q = 0; q < kTilePer; q++) {
        uint32_t starts = seg_starts(wb[q]);
        uint32_t l = S.lbase[i0 + q];
        while (starts) {
            const int sbit = __ffs(starts) - 1;
            starts &= starts - 1;
            S.lp[l] = (int)l;
            S.info[l] = ((uint32_t)(i0 + q) << 5) | (uint32_t)sbit;
            l++;
        }
    }
    __syncthreads();
    // links between words of this tile: one thread per NODE (all lanes busy)
    for (int l = threadIdx.x; l < nl; l += kThreads) {
        const uint32_t info = S.info[l];
        const int i = (int)(info >> 5), sbit = (int)(info & 31u);
        const int r = i / kTileC, c = i - r * kTileC;
        const uint32_t b = S.bits[i];
        const uint32_t left = (sbit == 0 && c > 0) ? S.bits[i - 1] : 0u;
        uint32_t up = 0, u = 0, un = 0;
        if (r > 0) {
            up = c > 0 ? S.bits[i - kTileC - 1] : 0u;
            u = S.bits[i - kTileC];
            un = c + 1 < kTileC ? S.bits[i - kTileC + 1] : 0u;
        }
        seg_links(b, sbit, left, above_window(up, u, un), u, [&](int kind, int ko) {
            int other;
            if (kind == 0) other = (int)S.lbase[i - 1] + seg_count(left) - 1;
            else if (kind == 1) other = (int)S.lbase[i - kTileC - 1] + seg_count(up) - 1;
            else if (kind == 2) other = (int)S.lbase[i - kTileC] + ko;
            else other = (int)S.lbase[i - kTileC + 1];
            s_unite(S.lp, l, other);
        });
    }
    __syncthreads();
    // flattened local forest -> global parents
    for (int l = threadIdx.x; l < nl; l += kThreads) {
        int x = l, p = S.lp[x];
        while (p != x) {
            x = p;
            p = S.lp[x];
        }
        const int wi = (int)(S.info[l] >> 5), wr = (int)(S.info[x] >> 5);
        const uint32_t gid = S.gbase[wi] + ((uint32_t)l - S.lbase[wi]);
        const uint32_t gr = S.gbase[wr] + ((uint32_t)x - S.lbase[wr]);
        P[gid] = (int)gr;
    }
}

// ---- 3. links that cross a tile edge --------------------------------------------------------------------
// thread space: [0, na) top rows of the tile bands (all links to the row above + the left link at
// tile column 0); [na, na + nb) the two edge word-columns of every tile in the other rows.
__global__ void __launch_bounds__(kThreads) ccl_border_kernel(const uint32_t* __restrict__ bits,
                                                              const uint32_t* __restrict__ nbase, CclGeom g,
                                                              int tiles_x, int tiles_y, int64_t na, int64_t nb,
                                                              int* __restrict__ P) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= na + nb) return;
    int y, j, frame;
    bool top, left_edge;
    if (t < na) {
        const int64_t band = t / g.wpr;  // frame * tiles_y + ty
        j = (int)(t - band * g.wpr);
        frame = (int)(band / tiles_y);
        y = (int)(band - (int64_t)frame * tiles_y) * kTileR;
        top = true;
        left_edge = (j % kTileC) == 0;
    } else {
        const int64_t u = t - na;
        const int side = (int)(u & 1);
        const int64_t v = u >> 1;
        const int txi = (int)(v % tiles_x);
        const int64_t row_id = v / tiles_x;  // frame * h + y
        frame = (int)(row_id / g.h);
        y = (int)(row_id - (int64_t)frame * g.h);
        if ((y % kTileR) == 0) return;  // top rows are handled above
        j = txi * kTileC + (side ? kTileC - 1 : 0);
        if (j >= g.wpr) return;
        top = false;
        left_edge = side == 0;
    }
    const int64_t gw = (int64_t)frame * g.words_per_frame + (int64_t)y * g.wpr + j;
    const uint32_t b = __ldg(bits + gw);
    if (!b) return;
    const bool right_edge = (j % kTileC) == kTileC - 1;
    uint32_t left = 0, up = 0, u = 0, un = 0;
    if (left_edge && j > 0) left = __ldg(bits + gw - 1);
    if (y > 0) {
        if ((top || left_edge) && j > 0) up = __ldg(bits + gw - g.wpr - 1);
        if (top) u = __ldg(bits + gw - g.wpr);
        if ((top || right_edge) && j + 1 < g.wpr) un = __ldg(bits + gw - g.wpr + 1);
    }
    if (!((b & 1u) && (left >> 31)) && !(up >> 31) && !u && !(un & 1u)) return;
    const int self = (int)__ldg(nbase + gw);
    word_links(b, left, up, u, un, [&](int kind, int ks, int ko) {
        int other;
        if (kind == 0) other = (int)__ldg(nbase + gw - 1) + seg_count(left) - 1;
        else if (kind == 1) other = (int)__ldg(nbase + gw - g.wpr - 1) + seg_count(up) - 1;
        else if (kind == 2) other = (int)__ldg(nbase + gw - g.wpr) + ko;
        else other = (int)__ldg(nbase + gw - g.wpr + 1);
        unite(P, self + ks, other);
    });
}

// ---- 4. flatten + rank: P[node] = root, P[root] = -(rank + 1) ---------------------------------------------
constexpr int kRankPer = 8;
constexpr int kRankSub = kBigThreads * kRankPer;   // nodes per sub-chunk
constexpr int kRankCtasPerSm = 2;                  // 2 x 1024 threads per SM stay co-resident (32 registers)

// walk kRankPer consecutive nodes to their roots (read-only), store the roots; returns the root bitmask
__device__ __forceinline__ uint32_t flatten_nodes(int* __restrict__ P, int n0, int total) {
    int par[kRankPer];
    if (n0 + kRankPer <= total) {
#pragma unroll
        for (int v = 0; v < kRankPer / 4; v++) {
            const int4 q = __ldcg(reinterpret_cast<const int4*>(P + n0) + v);
            par[4 * v] = q.x; par[4 * v + 1] = q.y; par[4 * v + 2] = q.z; par[4 * v + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < kRankPer; i++) par[i] = n0 + i < total ? __ldcg(P + n0 + i) : -1;
    }
    // first hop of all nodes at once (after the tile pass most nodes point at a root already);
    // a negative parent marks a root that already holds its rank
    int gp[kRankPer];
#pragma unroll
    for (int i = 0; i < kRankPer; i++) {
        const int node = n0 + i;
        gp[i] = (node < total && par[i] != node) ? __ldcg(P + par[i]) : -1;
    }
    uint32_t roots = 0;
#pragma unroll
    for (int i = 0; i < kRankPer; i++) {
        const int node = n0 + i;
        if (node < total) {
            if (par[i] == node) {
                roots |= 1u << i;
            } else {
                int x = par[i], p = gp[i];
                while (p >= 0 && p != x) {
                    x = p;
                    p = __ldcg(P + x);
                }
                if (par[i] != x) P[node] = x;
            }
        }
    }
    return roots;
}

// grid = kRankCtasPerSm x num_sms; chunks = min(grid, ceil(total / kRankSub)), each a whole number of sub-chunks
__global__ void __launch_bounds__(kBigThreads, kRankCtasPerSm) ccl_rank_kernel(const uint32_t* __restrict__ total_nodes,
                                                               int* __restrict__ P,
                                                               unsigned long long* __restrict__ status,
                                                               unsigned int* __restrict__ ticket,
                                                               uint32_t* __restrict__ sub_excl,
                                                               uint32_t* __restrict__ total_roots,
                                                               int32_t* __restrict__ counts_single) {
    __shared__ uint32_t s_tmp[kBigThreads / 32];
    __shared__ uint32_t s_prefix, s_total;
    __shared__ int s_chunk;
    const int total = (int)*total_nodes;
    const int subs = (total + kRankSub - 1) / kRankSub;
    if (subs == 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            *total_roots = 0;
            if (counts_single) *counts_single = 0;
        }
        return;
    }
    const int chunks = subs < (int)gridDim.x ? subs : (int)gridDim.x;
    const int subs_per = (subs + chunks - 1) / chunks;
    if (threadIdx.x == 0) s_chunk = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int chunk = s_chunk;
    const int sub_begin = chunk * subs_per;
    if (chunk >= chunks || sub_begin >= subs) {
        // nothing to rank here, but later chunks still sum over this ticket's status word
        if (chunk < (int)gridDim.x && threadIdx.x == 0) st_status(status + chunk, kFlagValid);
        return;
    }
    const int sub_end = sub_begin + subs_per < subs ? sub_begin + subs_per : subs;
    // pass 1: flatten, count roots (the first sub-chunk's root mask stays in a register for pass 2)
    const uint32_t roots0 = flatten_nodes(P, sub_begin * kRankSub + threadIdx.x * kRankPer, total);
    uint32_t mine = __popc(roots0);
    for (int sub = sub_begin + 1; sub < sub_end; sub++)
        mine += __popc(flatten_nodes(P, sub * kRankSub + threadIdx.x * kRankPer, total));
    block_excl_scan_u32<kBigThreads>(mine, s_tmp, &s_total);
    __syncthreads();
    const uint32_t agg = s_total;
    uint32_t carry = ordered_exclusive(status, chunk, agg, &s_prefix);
    if (threadIdx.x == 0 && sub_end >= subs) {
        *total_roots = carry + agg;
        if (counts_single) *counts_single = (int32_t)(carry + agg);
    }
    // pass 2: rank the roots (a root's slot still holds its own index)
    for (int sub = sub_begin; sub < sub_end; sub++) {
        const int n0 = sub * kRankSub + threadIdx.x * kRankPer;
        uint32_t roots = roots0;
        if (sub != sub_begin) {
            roots = 0;
#pragma unroll
            for (int i = 0; i < kRankPer; i++)
                if (n0 + i < total && __ldcg(P + n0 + i) == n0 + i) roots |= 1u << i;
        }
        __syncthreads();
        uint32_t rank = carry + block_excl_scan_u32<kBigThreads>(__popc(roots), s_tmp, &s_total);
        if (threadIdx.x == 0) sub_excl[sub] = carry;
#pragma unroll
        for (int i = 0; i < kRankPer; i++) {
            if ((roots >> i) & 1u) {
                P[n0 + i] = -(int)(rank + 1);
                rank++;
            }
        }
        __syncthreads();
        carry += s_total;
    }
}

// ---- 5. per-frame root offsets and component counts (one warp per frame; stacks only) ---------------------
__global__ void __launch_bounds__(kThreads) ccl_frame_offsets_kernel(const uint32_t* __restrict__ nbase, CclGeom g,
                                                                     const uint32_t* __restrict__ total_nodes,
                                                                     const uint32_t* __restrict__ total_roots,
                                                                     const uint32_t* __restrict__ chunk_excl,  // per rank sub-chunk
                                                                     const int* __restrict__ P,
                                                                     uint32_t* __restrict__ frame_off,
                                                                     int32_t* __restrict__ counts) {
    const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (f >= g.frames) return;
    auto roots_before_frame = [&](int frame) -> uint32_t {
        if (frame >= g.frames) return *total_roots;
        const uint32_t n0 = nbase[(int64_t)frame * g.words_per_frame];
        if (n0 >= *total_nodes) return *total_roots;
        const uint32_t c0 = n0 / kRankSub;
        uint32_t r = 0;
        for (uint32_t i = c0 * kRankSub + lane; i < n0; i += 32) r += (P[i] < 0) ? 1u : 0u;
        return chunk_excl[c0] + yam_warp_sum(r);
    };
    const uint32_t here = roots_before_frame(f), next = roots_before_frame(f + 1);
    if (lane == 0) {
        frame_off[f] = here;
        if (counts) counts[f] = (int32_t)(next - here);
    }
}

// ---- 6. final labels ----------------------------------------------------------------------------
// Generic widths: one thread per NIBBLE (4 pixels = one 16-byte store where the row pitch allows).
// A set nibble holds at most two run segments, each resolved with at most two dependent loads
// (node -> root -> -(label)).
constexpr int kFinalIter = 4;  // nibbles per thread

// Whole-word rows (w % 32 == 0): lane = word.  Each lane resolves the labels of its word's first two
// segments, then the warp writes the 32 words cooperatively: 8 rounds x (lane = nibble of one of 4
// words), the owner's bits and labels arrive by shuffle.  Words with more than two segments (noise)
// are written by their owner lane afterwards.
__global__ void __launch_bounds__(kThreads) ccl_final_warp_kernel(const uint32_t* __restrict__ bits,
                                                                  const uint32_t* __restrict__ nbase,
                                                                  const uint32_t* __restrict__ frame_off, CclGeom g,
                                                                  const int* __restrict__ P,
                                                                  int32_t* __restrict__ labels,
                                                                  const int32_t* __restrict__ remap, int remap_size,
                                                                  int64_t word_begin, int64_t word_end) {
    // words [word_begin, word_end) are written; labels points at the pixel of word_begin
    // (remap_size > 0: labels beyond the table map to 0 instead of reading past it -- bounded merge tables)
    const int lane = threadIdx.x & 31;
    const int64_t warp_base = word_begin + ((int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * 32;
    if (warp_base >= word_end) return;
    const int64_t gw = warp_base + lane;
    const uint32_t b = gw < word_end ? __ldg(bits + gw) : 0u;
    int4* dst = reinterpret_cast<int4*>(labels + (warp_base - word_begin) * 32);
    const int valid_words = word_end - warp_base < 32 ? (int)(word_end - warp_base) : 32;
    if (!__any_sync(0xffffffffu, b != 0u)) {
#pragma unroll
        for (int it = 0; it < 8; it++)
            if (4 * it + (lane >> 3) < valid_words) __stcs(dst + it * 32 + lane, make_int4(0, 0, 0, 0));
        return;
    }
    int l0 = 0, l1 = 0, base = 0, off = 0;
    const uint32_t starts = seg_starts(b);
    const int nseg = __popc(starts);
    if (b) {
        base = (int)__ldg(nbase + gw);
        if (g.frames > 1) off = (int)__ldg(frame_off + gw / g.words_per_frame);
        int v0 = __ldg(P + base);
        int v1 = nseg > 1 ? __ldg(P + base + 1) : -1;
        if (v0 >= 0) v0 = __ldg(P + v0);  // non-root: its root holds -(label)
        if (v1 >= 0) v1 = __ldg(P + v1);
        l0 = -v0 - off;
        l1 = -v1 - off;
        if (remap) {  // strip-local label -> global label (cross-strip merge)
            l0 = (remap_size <= 0 || l0 < remap_size) ? __ldg(remap + l0) : 0;
            if (nseg > 1) l1 = (remap_size <= 0 || l1 < remap_size) ? __ldg(remap + l1) : 0;
        }
    }
    // bits that belong to the second (or a later) segment
    const uint32_t later = starts & (starts - 1u);
    const uint32_t hi = later ? ~((later & (0u - later)) - 1u) : 0u;
    const bool many = nseg > 2;
    const bool any_many = __any_sync(0xffffffffu, many);
#pragma unroll
    for (int it = 0; it < 8; it++) {
        const int src = 4 * it + (lane >> 3), q = lane & 7;
        const uint32_t bb = __shfl_sync(0xffffffffu, b, src);
        const uint32_t hh = __shfl_sync(0xffffffffu, hi, src);
        const int a0 = __shfl_sync(0xffffffffu, l0, src);
        const int a1 = __shfl_sync(0xffffffffu, l1, src);
        const uint32_t nib = (bb >> (4 * q)) & 0xfu;
        const uint32_t sec = (hh >> (4 * q)) & 0xfu;
        int4 out;
        out.x = (nib & 1u) ? ((sec & 1u) ? a1 : a0) : 0;
        out.y = (nib & 2u) ? ((sec & 2u) ? a1 : a0) : 0;
        out.z = (nib & 4u) ? ((sec & 4u) ? a1 : a0) : 0;
        out.w = (nib & 8u) ? ((sec & 8u) ? a1 : a0) : 0;
        bool skip = src >= valid_words;
        if (any_many) skip = skip || __shfl_sync(0xffffffffu, many ? 1 : 0, src);
        if (!skip) __stcs(dst + it * 32 + lane, out);
    }
    if (many) {
        int32_t* d = labels + (gw - word_begin) * 32;
        uint32_t st = starts;
        int cur = 0, k = 0;
#pragma unroll 1
        for (int i = 0; i < 32; i++) {
            if ((st >> i) & 1u) {
                int v = __ldg(P + base + k);
                if (v >= 0) v = __ldg(P + v);
                cur = -v - off;
                if (remap) cur = (remap_size <= 0 || cur < remap_size) ? __ldg(remap + cur) : 0;
                k++;
            }
            d[i] = ((b >> i) & 1u) ? cur : 0;
        }
    }
}

__global__ void __launch_bounds__(kThreads) ccl_final_kernel(const uint32_t* __restrict__ bits,
                                                             const uint32_t* __restrict__ nbase,
                                                             const uint32_t* __restrict__ frame_off, CclGeom g,
                                                             const int* __restrict__ P, int32_t* __restrict__ labels,
                                                             bool vec_ok, const int32_t* __restrict__ remap, int remap_size,
                                                             int64_t word_begin, int64_t word_end) {
    // words [word_begin, word_end) (whole rows) are written; labels points at the first pixel of that range
    const int64_t t_base = word_begin * 8 + (int64_t)blockIdx.x * (kThreads * kFinalIter) + threadIdx.x;
    const int64_t total_nibbles = word_end * 8;
    uint32_t bw[kFinalIter];
#pragma unroll
    for (int it = 0; it < kFinalIter; it++) {
        const int64_t t = t_base + (int64_t)it * kThreads;
        bw[it] = t < total_nibbles ? __ldg(bits + (t >> 3)) : 0u;
    }
#pragma unroll
    for (int it = 0; it < kFinalIter; it++) {
        const int64_t t = t_base + (int64_t)it * kThreads;
        if (t >= total_nibbles) continue;
        const int64_t gw = t >> 3;
        const int q = (int)(t & 7);
        const uint32_t b = bw[it];
        const uint32_t nib = (b >> (4 * q)) & 0xfu;
        int4 out = make_int4(0, 0, 0, 0);
        if (nib) {
            const int base = (int)__ldg(nbase + gw);
            const int off = g.frames > 1 ? (int)__ldg(frame_off + gw / g.words_per_frame) : 0;
            const int first = __ffs(nib) - 1;
            const uint32_t second = nib & (nib + (1u << first)) & 0xfu;  // bits of a second run, if any
            const int k0 = seg_index(b, 4 * q + first);
            int v0 = __ldg(P + base + k0);
            int v1 = second ? __ldg(P + base + k0 + 1) : -1;
            if (v0 >= 0) v0 = __ldg(P + v0);   // non-root: its root holds -(label)
            if (v1 >= 0) v1 = __ldg(P + v1);
            int l0 = -v0 - off, l1 = -v1 - off;
            if (remap) {
                l0 = (remap_size <= 0 || l0 < remap_size) ? __ldg(remap + l0) : 0;
                if (second) l1 = (remap_size <= 0 || l1 < remap_size) ? __ldg(remap + l1) : 0;
            }
            out.x = (nib & 1u) ? ((second & 1u) ? l1 : l0) : 0;
            out.y = (nib & 2u) ? ((second & 2u) ? l1 : l0) : 0;
            out.z = (nib & 4u) ? ((second & 4u) ? l1 : l0) : 0;
            out.w = (nib & 8u) ? ((second & 8u) ? l1 : l0) : 0;
        }
        {
            const int64_t row_id = gw / g.wpr;
            const int x = (int)(gw - row_id * g.wpr) * 32 + 4 * q;
            if (x >= g.w) continue;
            int32_t* d = labels + (row_id - word_begin / g.wpr) * (int64_t)g.w + x;
            if (vec_ok && x + 4 <= g.w) {
                __stcs(reinterpret_cast<int4*>(d), out);
            } else {
                d[0] = out.x;
                if (x + 1 < g.w) d[1] = out.y;
                if (x + 2 < g.w) d[2] = out.z;
                if (x + 3 < g.w) d[3] = out.w;
            }
        }
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
__global__ void __launch_bounds__(kThreads, 4) props_kernel(const int32_t* __restrict__ labels,
                                                         const TI* __restrict__ intensity, int h, int w,
                                                         int64_t n_labels, long long* __restrict__ props,
                                                         const int64_t* __restrict__ frame_offsets) {
    if (frame_offsets) {
        // stack: blockIdx.z = frame; its rows of the table start at frame_offsets[frame]
        const int64_t f = blockIdx.z, o0 = frame_offsets[f];
        n_labels = frame_offsets[f + 1] - o0;
        if (n_labels <= 0) return;
        labels += f * (int64_t)h * w;
        if (intensity) intensity += f * (int64_t)h * w;
        props += o0 * YAM_PROPS_STRIDE;
    }
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

__global__ void __launch_bounds__(kThreads) relabel_kernel(int32_t* __restrict__ labels, int64_t count,
                                                           const int32_t* __restrict__ remap, int64_t remap_size) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool aligned = (reinterpret_cast<uintptr_t>(labels) & 15) == 0;
    int64_t done = 0;
    if (aligned) {
        const int64_t groups = count / 4;
        for (int64_t g = tid; g < groups; g += stride) {
            int4 v = reinterpret_cast<int4*>(labels)[g];
            if (v.x | v.y | v.z | v.w) {
                v.x = (v.x > 0 && v.x < remap_size) ? __ldg(remap + v.x) : 0;
                v.y = (v.y > 0 && v.y < remap_size) ? __ldg(remap + v.y) : 0;
                v.z = (v.z > 0 && v.z < remap_size) ? __ldg(remap + v.z) : 0;
                v.w = (v.w > 0 && v.w < remap_size) ? __ldg(remap + v.w) : 0;
                reinterpret_cast<int4*>(labels)[g] = v;
            }
        }
        done = groups * 4;
    }
    for (int64_t i = done + tid; i < count; i += stride) {
        const int v = labels[i];
        if (v) labels[i] = (v > 0 && v < remap_size) ? remap[v] : 0;
    }
}

// ---- cross-strip merge: union-find over the global label ids that meet at strip boundaries --------
__global__ void __launch_bounds__(kThreads) iota_kernel(int32_t* __restrict__ p, int64_t count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) p[i] = (int32_t)i;
}

// thread = (boundary r | r+1, column x): bottom row of strip r against the three neighbours in the
// top row of strip r + 1
__global__ void __launch_bounds__(kThreads) strip_union_kernel(const int32_t* __restrict__ edges,
                                                               const int64_t* __restrict__ offs, int world, int64_t w,
                                                               int* __restrict__ P) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)(world - 1) * w) return;
    const int r = (int)(t / w);
    const int64_t x = t - (int64_t)r * w;
    const int32_t a = edges[((int64_t)r * 2 + 1) * w + x];
    if (a <= 0) return;
    const int ga = (int)(offs[r] + a);
    const int32_t* top = edges + ((int64_t)(r + 1) * 2) * w;
    const int64_t ob = offs[r + 1];
    int last = 0;
#pragma unroll
    for (int dx = -1; dx <= 1; dx++) {
        const int64_t xx = x + dx;
        if (xx < 0 || xx >= w) continue;
        const int32_t b = top[xx];
        if (b > 0 && b != last) {
            unite(P, ga, (int)(ob + b));
            last = b;
        }
    }
}

__global__ void __launch_bounds__(kThreads) flatten_all_kernel(int* __restrict__ P, int64_t count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    int x = (int)i, p = __ldcg(P + x);
    if (p == x) return;
    while (p != x) {
        x = p;
        p = __ldcg(P + x);
    }
    P[i] = x;  // owner-only store: readers see either an ancestor or the root
}

// ---- one-call merge + renumbering (yam_merge_strips_remap) -----------------------------------------
struct StripOffsets {
    long long v[65];  // exclusive prefix of the per-strip component counts, by value (world <= 64)
};

// P = identity over [0, total]; non-root bitmap cleared
__global__ void __launch_bounds__(kThreads) merge_init_kernel(int32_t* __restrict__ P, int64_t count,
                                                              uint32_t* __restrict__ nonroot, int64_t nwords) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) P[i] = (int32_t)i;
    if (i < nwords) nonroot[i] = 0u;
}

// The same when the offsets are UPPER BOUNDS (offs.v[r + 1] - offs.v[r] >= the strip's real count, which only
// the device knows: counts[r * counts_stride]): ids between a strip's count and its bound do not exist and are
// marked non-root, so they take no rank.  *overflow = 1 if a count exceeds its bound (the caller then redoes
// the merge with exact offsets).
__global__ void __launch_bounds__(kThreads) merge_init_bounded_kernel(int32_t* __restrict__ P, int64_t count,
                                                                      uint32_t* __restrict__ nonroot, int64_t nwords,
                                                                      StripOffsets offs, int world,
                                                                      const int32_t* __restrict__ counts, int64_t counts_stride,
                                                                      int32_t* __restrict__ overflow) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) P[i] = (int32_t)i;
    if (i < nwords) {
        uint32_t m = 0;
        const int64_t id0 = i * 32;
        int r = 0;
        while (r < world && offs.v[r + 1] < id0) r++;          // strip r holds ids (offs[r], offs[r + 1]]
        for (int b = 0; b < 32; b++) {
            const int64_t id = id0 + b;
            if (id == 0 || id >= count) continue;
            while (r < world && offs.v[r + 1] < id) r++;
            if (r < world && id - offs.v[r] > (int64_t)counts[(int64_t)r * counts_stride]) m |= 1u << b;
        }
        nonroot[i] = m;
    }
    if (i == 0) {
        int bad = 0;
        for (int r = 0; r < world; r++)
            if ((int64_t)counts[(int64_t)r * counts_stride] > offs.v[r + 1] - offs.v[r]) bad = 1;
        *overflow = bad;
    }
}

// thread = (boundary r | r+1, column x) over the packed rows [world][stride]: row 0 = first label row,
// row 1 = last label row of a strip
__global__ void __launch_bounds__(kThreads) merge_union_kernel(const int32_t* __restrict__ packed, int64_t stride,
                                                               StripOffsets offs, int world, int64_t w,
                                                               int* __restrict__ P) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)(world - 1) * w) return;
    const int r = (int)(t / w);
    const int64_t x = t - (int64_t)r * w;
    const int32_t a = packed[(int64_t)r * stride + w + x];
    if (a <= 0 || a > offs.v[r + 1] - offs.v[r]) return;      // (beyond the strip's range: only with overflowed bounds)
    const int ga = (int)(offs.v[r] + a);
    const int32_t* top = packed + (int64_t)(r + 1) * stride;
    const long long ob = offs.v[r + 1];
    int last = 0;
#pragma unroll
    for (int dx = -1; dx <= 1; dx++) {
        const int64_t xx = x + dx;
        if (xx < 0 || xx >= w) continue;
        const int32_t b = top[xx];
        if (b > 0 && b != last && b <= offs.v[r + 2] - ob) {
            unite(P, ga, (int)(ob + b));
            last = b;
        }
    }
}

// Only ids that sit on a strip boundary row can have lost their root status: flatten those and mark
// the non-roots in the bitmap (thread = one pixel of one packed row; duplicates are idempotent).
__global__ void __launch_bounds__(kThreads) merge_flatten_kernel(const int32_t* __restrict__ packed, int64_t stride,
                                                                 StripOffsets offs, int world, int64_t w,
                                                                 int* __restrict__ P, uint32_t* __restrict__ nonroot) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)world * 2 * w) return;
    const int r = (int)(t / (2 * w));
    const int64_t x = t - (int64_t)r * 2 * w;
    const int32_t a = packed[(int64_t)r * stride + x];
    if (a <= 0 || a > offs.v[r + 1] - offs.v[r]) return;
    if (x > 0 && x != w && packed[(int64_t)r * stride + x - 1] == a) return;  // one thread per run of a label
    const int g = (int)(offs.v[r] + a);
    int cur = g, p = __ldcg(P + cur);
    while (p != cur) {
        cur = p;
        p = __ldcg(P + cur);
    }
    if (cur != g) {
        P[g] = cur;  // owner-only store of the final root (readers see an ancestor or the root)
        atomicOr(nonroot + (g >> 5), 1u << (g & 31));
    }
}

// exclusive prefix of the per-word non-root counts.  One block of 32 warps; warp w owns a contiguous
// segment, lanes stride through it (coalesced): pass 1 sums the segments, pass 2 writes the prefix with
// a shuffle scan per 32-word group and a running carry.
__global__ void __launch_bounds__(1024) merge_scan_kernel(const uint32_t* __restrict__ nonroot, int64_t nwords,
                                                          uint32_t* __restrict__ prefix, int64_t total,
                                                          int32_t* __restrict__ total_out) {
    __shared__ uint32_t seg_sum[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t seg = ((nwords + 31) / 32 + 31) / 32 * 32;          // words per warp, a multiple of 32
    const int64_t b0 = min(nwords, (int64_t)warp * seg), b1 = min(nwords, b0 + seg);
    uint32_t s = 0;
    for (int64_t i = b0 + lane; i < b1; i += 32) s += __popc(nonroot[i]);
    s = yam_warp_sum(s);
    if (lane == 0) seg_sum[warp] = s;
    __syncthreads();
    uint32_t carry = 0, all = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) {
        const uint32_t v = seg_sum[i];
        if (i < warp) carry += v;
        all += v;
    }
    for (int64_t g = b0; g < b1; g += 32) {
        const int64_t i = g + lane;
        const uint32_t c = i < b1 ? __popc(nonroot[i]) : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (i < b1) prefix[i] = carry + incl - c;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (threadIdx.x == 0) total_out[0] = (int32_t)(total - (int64_t)all);
}

// remap[l] = raster-first global label of local label l of strip `rank`: the rank of its root among
// the roots = root id minus the number of non-roots below it
__global__ void __launch_bounds__(kThreads) merge_remap_kernel(const int* __restrict__ P,
                                                               const uint32_t* __restrict__ nonroot,
                                                               const uint32_t* __restrict__ prefix, long long base,
                                                               int64_t count, int32_t* __restrict__ remap) {
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l > count) return;
    if (l == 0) {
        remap[0] = 0;
        return;
    }
    const int rt = P[base + l];
    const uint32_t below = prefix[rt >> 5] + __popc(nonroot[rt >> 5] & ((1u << (rt & 31)) - 1u));
    remap[l] = rt - (int)below;
}

}  // namespace

extern "C" {

int64_t yam_merge_strips_workspace_bytes(int64_t total) {
    if (total < 0 || total >= (1ll << 31) - 1) return -1;
    const int64_t nwords = (total + 32) / 32;
    return (int64_t)yam_align_up((size_t)(total + 1) * 4, 256) + 2 * (int64_t)yam_align_up((size_t)nwords * 4, 256);
}

static int merge_strips_remap_impl(yam_ctx* ctx, const int32_t* packed_dev, int64_t stride, int world, int64_t w,
                                   const int64_t* offsets_host, int rank, int rank_count, void* workspace, int32_t* remap_dev,
                                   int32_t* total_dev, const int32_t* counts_dev, int64_t counts_stride, int32_t* overflow_dev);

int yam_merge_strips_remap(yam_ctx* ctx, const int32_t* packed_dev, int64_t stride, int world, int64_t w,
                           const int64_t* offsets_host, int rank, int rank_count, void* workspace, int32_t* remap_dev,
                           int32_t* total_dev) {
    if (int rc = yam_enter(ctx)) return rc;
    return merge_strips_remap_impl(ctx, packed_dev, stride, world, w, offsets_host, rank, rank_count, workspace, remap_dev, total_dev,
                                   nullptr, 0, nullptr);
}

int yam_merge_strips_remap_bounded(yam_ctx* ctx, const int32_t* packed_dev, int64_t stride, int world, int64_t w,
                                   const int64_t* offsets_bound_host, const int32_t* counts_dev, int64_t counts_stride, int rank,
                                   int rank_count, void* workspace, int32_t* remap_dev, int32_t* total_dev, int32_t* overflow_dev) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(counts_dev && counts_stride >= 1 && overflow_dev, "merge_strips_remap_bounded: NULL argument");
    return merge_strips_remap_impl(ctx, packed_dev, stride, world, w, offsets_bound_host, rank, rank_count, workspace, remap_dev,
                                   total_dev, counts_dev, counts_stride, overflow_dev);
}

static int merge_strips_remap_impl(yam_ctx* ctx, const int32_t* packed_dev, int64_t stride, int world, int64_t w,
                                   const int64_t* offsets_host, int rank, int rank_count, void* workspace, int32_t* remap_dev,
                                   int32_t* total_dev, const int32_t* counts_dev, int64_t counts_stride, int32_t* overflow_dev) {
    YAM_REQUIRE(packed_dev && offsets_host && workspace && remap_dev && total_dev, "merge_strips_remap: NULL argument");
    YAM_REQUIRE(world >= 1 && world <= 64 && w > 0 && stride >= 2 * w && rank >= 0 && rank_count >= 1 && rank + rank_count <= world,
                "merge_strips_remap: bad geometry (world %d, w %lld, stride %lld, strips %d..+%d)", world, (long long)w,
                (long long)stride, rank, rank_count);
    StripOffsets offs;
    for (int i = 0; i <= world; i++) {
        offs.v[i] = offsets_host[i];
        YAM_REQUIRE(offs.v[i] >= 0 && (i == 0 || offs.v[i] >= offs.v[i - 1]), "merge_strips_remap: offsets must ascend");
    }
    YAM_REQUIRE(offs.v[0] == 0, "merge_strips_remap: offsets[0] must be 0");
    const int64_t total = offs.v[world];
    YAM_REQUIRE(total < (1ll << 31) - 1, "merge_strips_remap: more than 2^31 components");
    const int64_t count = total + 1, nwords = (total + 32) / 32;
    int* P = (int*)workspace;
    uint32_t* nonroot = (uint32_t*)((char*)workspace + yam_align_up((size_t)count * 4, 256));
    uint32_t* prefix = (uint32_t*)((char*)nonroot + yam_align_up((size_t)nwords * 4, 256));
    const auto blocks = [](int64_t items) { return (unsigned)((items + kThreads - 1) / kThreads); };
    if (counts_dev)
        merge_init_bounded_kernel<<<blocks(count), kThreads, 0, ctx->stream>>>(P, count, nonroot, nwords, offs, world, counts_dev,
                                                                               counts_stride, overflow_dev);
    else
        merge_init_kernel<<<blocks(count), kThreads, 0, ctx->stream>>>(P, count, nonroot, nwords);
    YAM_LAUNCHED(ctx);
    if (world > 1) {
        merge_union_kernel<<<blocks((int64_t)(world - 1) * w), kThreads, 0, ctx->stream>>>(packed_dev, stride, offs,
                                                                                          world, w, P);
        YAM_LAUNCHED(ctx);
        merge_flatten_kernel<<<blocks((int64_t)world * 2 * w), kThreads, 0, ctx->stream>>>(packed_dev, stride, offs,
                                                                                          world, w, P, nonroot);
        YAM_LAUNCHED(ctx);
    }
    merge_scan_kernel<<<1, 1024, 0, ctx->stream>>>(nonroot, nwords, prefix, total, total_dev);
    YAM_LAUNCHED(ctx);
    // tables of strips rank .. rank + rank_count - 1, back to back (count_i + 1 entries each)
    int64_t at = 0;
    for (int r = rank; r < rank + rank_count; r++) {
        const int64_t mine = offs.v[r + 1] - offs.v[r];
        merge_remap_kernel<<<blocks(mine + 1), kThreads, 0, ctx->stream>>>(P, nonroot, prefix, offs.v[r], mine, remap_dev + at);
        YAM_LAUNCHED(ctx);
        at += mine + 1;
    }
    return YAM_OK;
}

int yam_merge_strip_labels(yam_ctx* ctx, const int32_t* edges_dev, const int64_t* offsets_dev, int world, int64_t w,
                           int64_t total, int32_t* root_dev) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(edges_dev && offsets_dev && root_dev && world >= 1 && w > 0 && total >= 0 && total < (1ll << 31) - 1,
                "merge_strip_labels: bad arguments");
    const int64_t count = total + 1;
    const unsigned nb = (unsigned)((count + kThreads - 1) / kThreads);
    iota_kernel<<<nb, kThreads, 0, ctx->stream>>>(root_dev, count);
    YAM_LAUNCHED(ctx);
    if (world > 1) {
        const int64_t pairs = (int64_t)(world - 1) * w;
        strip_union_kernel<<<(unsigned)((pairs + kThreads - 1) / kThreads), kThreads, 0, ctx->stream>>>(
            edges_dev, offsets_dev, world, w, root_dev);
        YAM_LAUNCHED(ctx);
        flatten_all_kernel<<<nb, kThreads, 0, ctx->stream>>>(root_dev, count);
        YAM_LAUNCHED(ctx);
    }
    return YAM_OK;
}

int yam_relabel(yam_ctx* ctx, int32_t* labels, int64_t count, const int32_t* remap_dev, int64_t remap_size) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(labels && remap_dev && count > 0 && remap_size > 0, "relabel: bad arguments");
    int64_t bx = (count / 4 + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)ctx->num_sms * 8;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    relabel_kernel<<<(unsigned)bx, kThreads, 0, ctx->stream>>>(labels, count, remap_dev, remap_size);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

// Workspace of one labelling job (caller-owned device memory, or the context scratch for the
// one-shot entry points): everything `emit` needs survives between `resolve` and `emit`.
struct CclWorkspace {
    uint32_t* bits_scratch;  // packed mask when the input was a byte mask
    uint32_t* nbase;
    char* sync_base;
    size_t sync_bytes;
    unsigned int* tickets;
    unsigned long long *statusA, *statusB;
    uint32_t* chunk_excl;
    uint32_t* frame_off;
    uint32_t* totals;
    int32_t* counts;
    int* P;
    size_t bytes;
};

static int ccl_geometry(int64_t n, int64_t h, int64_t w, CclGeom* g) {
    YAM_REQUIRE(n > 0 && h > 0 && w > 0, "ccl: bad arguments");
    YAM_REQUIRE(n <= 65535, "ccl: at most 65535 frames per call");
    g->h = (int)h;
    g->w = (int)w;
    g->wpr = (int)((w + 31) / 32);
    g->words_per_frame = (int64_t)h * g->wpr;
    g->total_words = g->words_per_frame * n;
    g->frames = (int)n;
    YAM_REQUIRE(g->words_per_frame * 32 < (1ll << 31), "ccl: frame too large for int32 labels (%lld x %lld)",
                (long long)h, (long long)w);
    YAM_REQUIRE(g->total_words * 16 < (1ll << 31) && g->total_words < (1ll << 27),
                "ccl: stack too large for one call (%lld words); split the stack", (long long)g->total_words);
    return YAM_OK;
}

// layout: bits | nbase | [tickets | statusA | statusB] | chunk_excl | frame_off | totals | counts | P
static void ccl_layout(const CclGeom& g, int num_sms, char* base, CclWorkspace* ws) {
    const int64_t max_nodes = g.total_words * 16;  // at most 16 run segments per 32-pixel word
    const int64_t scan_subs = (g.total_words + kScanSub - 1) / kScanSub;
    const int64_t scan_chunks = scan_subs < num_sms ? scan_subs : num_sms;
    const int64_t rank_subs = (max_nodes + kRankSub - 1) / kRankSub;
    const size_t words_bytes = yam_align_up((size_t)g.total_words * 4, 256);
    const size_t sync_bytes = yam_align_up(256 + (size_t)(scan_chunks + kRankCtasPerSm * num_sms) * 8, 256);
    const size_t chunk_bytes = yam_align_up((size_t)rank_subs * 4, 256);
    const size_t frame_bytes = yam_align_up((size_t)g.frames * 4, 256);
    const size_t node_bytes = yam_align_up((size_t)max_nodes * 4, 256);
    char* sp = base;
    ws->bits_scratch = (uint32_t*)sp; sp += words_bytes;
    ws->nbase = (uint32_t*)sp; sp += words_bytes;
    ws->sync_base = sp;
    ws->sync_bytes = sync_bytes;
    ws->tickets = (unsigned int*)sp;                       // [0] scan, [1] rank
    ws->statusA = (unsigned long long*)(sp + 256);
    ws->statusB = ws->statusA + scan_chunks;
    sp += sync_bytes;
    ws->chunk_excl = (uint32_t*)sp; sp += chunk_bytes;
    ws->frame_off = (uint32_t*)sp; sp += frame_bytes;
    ws->totals = (uint32_t*)sp; sp += 256;                 // [0] nodes, [1] roots
    ws->counts = (int32_t*)sp; sp += frame_bytes;
    ws->P = (int*)sp; sp += node_bytes;
    ws->bytes = (size_t)(sp - base);
}

// scan .. rank: after this the workspace holds, for every run segment, its root and the root's label
static int ccl_resolve(yam_ctx* ctx, const void* mask, const uint32_t* bits_in, const CclGeom& g, const CclWorkspace& ws,
                       int32_t* counts) {
    const int64_t n = g.frames;
    const int64_t scan_subs = (g.total_words + kScanSub - 1) / kScanSub;
    const int64_t scan_chunks = scan_subs < ctx->num_sms ? scan_subs : ctx->num_sms;
    const int64_t scan_per = (scan_subs + scan_chunks - 1) / scan_chunks * kScanSub;  // words per chunk
    const uint32_t* bits = bits_in ? bits_in : ws.bits_scratch;
    YAM_CUDA(cudaMemsetAsync(ws.sync_base, 0, ws.sync_bytes, ctx->stream));
    if (n == 1) YAM_CUDA(cudaMemsetAsync(ws.frame_off, 0, sizeof(uint32_t), ctx->stream));
    const unsigned wblocks = (unsigned)((g.total_words + kThreads - 1) / kThreads);
    if (!bits_in) {
        ccl_pack_kernel<<<wblocks, kThreads, 0, ctx->stream>>>((const uint8_t*)mask, g, ws.bits_scratch);
        YAM_LAUNCHED(ctx);
    }
    ccl_scan_kernel<<<(unsigned)scan_chunks, kBigThreads, 0, ctx->stream>>>(bits, g.total_words, scan_per, ws.nbase, ws.statusA,
                                                                          ws.tickets, ws.totals);
    YAM_LAUNCHED(ctx);
    const int tiles_x = (g.wpr + kTileC - 1) / kTileC, tiles_y = (g.h + kTileR - 1) / kTileR;
    const int64_t tiles = (int64_t)tiles_x * tiles_y * n;
    YAM_REQUIRE(tiles < (1ll << 31), "ccl: too many tiles");
    ccl_tile_kernel<<<(unsigned)tiles, kThreads, sizeof(TileSmem), ctx->stream>>>(bits, ws.nbase, g, tiles_x, tiles_y, ws.P);
    YAM_LAUNCHED(ctx);
    const int64_t na = (int64_t)tiles_y * n * g.wpr, nb = (int64_t)g.h * n * tiles_x * 2;
    ccl_border_kernel<<<(unsigned)((na + nb + kThreads - 1) / kThreads), kThreads, 0, ctx->stream>>>(bits, ws.nbase, g, tiles_x,
                                                                                                  tiles_y, na, nb, ws.P);
    YAM_LAUNCHED(ctx);
    ccl_rank_kernel<<<(unsigned)(kRankCtasPerSm * ctx->num_sms), kBigThreads, 0, ctx->stream>>>(ws.totals, ws.P, ws.statusB, ws.tickets + 1,
                                                                            ws.chunk_excl, ws.totals + 1,
                                                                            n == 1 ? counts : nullptr);
    YAM_LAUNCHED(ctx);
    if (n > 1) {
        ccl_frame_offsets_kernel<<<(unsigned)((n * 32 + kThreads - 1) / kThreads), kThreads, 0, ctx->stream>>>(
            ws.nbase, g, ws.totals, ws.totals + 1, ws.chunk_excl, ws.P, ws.frame_off, counts);
        YAM_LAUNCHED(ctx);
    }
    return YAM_OK;
}

// labels of the words [word_begin, word_end) (whole rows), optionally mapped through remap
static int ccl_emit(yam_ctx* ctx, const uint32_t* bits, const CclGeom& g, const CclWorkspace& ws, int32_t* labels,
                    const int32_t* remap, int64_t word_begin, int64_t word_end, int remap_size = 0) {
    const int64_t words = word_end - word_begin;
    if (words <= 0) return YAM_OK;
    const bool lab_aligned = (reinterpret_cast<uintptr_t>(labels) & 15) == 0;
    if ((g.w & 31) == 0 && lab_aligned) {
        const unsigned wblocks32 = (unsigned)((words + kThreads - 1) / kThreads);  // 8 warps x 32 words
        ccl_final_warp_kernel<<<wblocks32, kThreads, 0, ctx->stream>>>(bits, ws.nbase, ws.frame_off, g, ws.P, labels, remap,
                                                                      remap_size, word_begin, word_end);
    } else {
        const unsigned fblocks = (unsigned)((words * 8 + kThreads * kFinalIter - 1) / (kThreads * kFinalIter));
        ccl_final_kernel<<<fblocks, kThreads, 0, ctx->stream>>>(bits, ws.nbase, ws.frame_off, g, ws.P, labels,
                                                                lab_aligned && (g.w & 3) == 0, remap, remap_size, word_begin, word_end);
    }
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

static int ccl_label_impl(yam_ctx* ctx, const void* mask, const uint32_t* bits_in, int32_t* labels, int64_t n,
                          int64_t h, int64_t w, int32_t* counts_dev, int32_t* counts_host) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE((mask || bits_in) && labels, "ccl: bad arguments");
    CclGeom g;
    if (int rc = ccl_geometry(n, h, w, &g)) return rc;
    CclWorkspace ws;
    ccl_layout(g, ctx->num_sms, nullptr, &ws);
    void* scratch = nullptr;
    if (int rc = yam_scratch(ctx, ws.bytes, &scratch)) return rc;
    ccl_layout(g, ctx->num_sms, (char*)scratch, &ws);
    int32_t* counts = counts_dev ? counts_dev : ws.counts;
    if (int rc = ccl_resolve(ctx, mask, bits_in, g, ws, counts)) return rc;
    if (int rc = ccl_emit(ctx, bits_in ? bits_in : ws.bits_scratch, g, ws, labels, nullptr, 0, g.total_words)) return rc;
    if (counts_host) {
        YAM_CUDA(cudaMemcpyAsync(counts_host, counts, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
        YAM_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return YAM_OK;
}

int yam_ccl_label(yam_ctx* ctx, const void* mask, int32_t* labels, int64_t n, int64_t h, int64_t w,
                  int32_t* counts_dev, int32_t* counts_host) {
    YAM_REQUIRE(mask != nullptr, "ccl: mask is NULL");
    return ccl_label_impl(ctx, mask, nullptr, labels, n, h, w, counts_dev, counts_host);
}

int yam_ccl_label_bits(yam_ctx* ctx, const uint32_t* bits, int32_t* labels, int64_t n, int64_t h, int64_t w,
                       int32_t* counts_dev, int32_t* counts_host) {
    YAM_REQUIRE(bits != nullptr, "ccl: bits is NULL");
    return ccl_label_impl(ctx, nullptr, bits, labels, n, h, w, counts_dev, counts_host);
}

static int region_props_impl(yam_ctx* ctx, const int32_t* labels, const void* intensity, int intensity_dtype, int64_t n,
                             int64_t h, int64_t w, const int64_t* offsets_dev, int64_t n_labels, int64_t* props_dev) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(labels && n > 0 && n <= 65535 && h > 0 && w > 0 && n_labels >= 0, "region_props: bad arguments");
    YAM_REQUIRE(h < (1 << 24) && w < (1 << 24), "region_props: image side must be below 2^24 (32-bit band partial sums)");
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
    dim3 grid(gx, gy, (unsigned)n);
    if (!intensity || intensity_dtype == YAM_U16)
        props_kernel<uint16_t><<<grid, kThreads, 0, ctx->stream>>>(labels, (const uint16_t*)intensity, (int)h, (int)w, n_labels,
                                                                   props, offsets_dev);
    else
        props_kernel<uint8_t><<<grid, kThreads, 0, ctx->stream>>>(labels, (const uint8_t*)intensity, (int)h, (int)w, n_labels,
                                                                  props, offsets_dev);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

int64_t yam_ccl_workspace_bytes(yam_ctx* ctx, int64_t n, int64_t h, int64_t w) {
    if (!ctx) return -1;
    CclGeom g;
    if (ccl_geometry(n, h, w, &g)) return -1;
    CclWorkspace ws;
    ccl_layout(g, ctx->num_sms, nullptr, &ws);
    return (int64_t)ws.bytes;
}

int yam_ccl_resolve_bits(yam_ctx* ctx, const uint32_t* bits, int64_t n, int64_t h, int64_t w, void* workspace,
                         int32_t* counts_dev) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(bits && workspace && counts_dev, "ccl_resolve_bits: NULL argument");
    YAM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "ccl_resolve_bits: workspace must be 256-byte aligned");
    CclGeom g;
    if (int rc = ccl_geometry(n, h, w, &g)) return rc;
    CclWorkspace ws;
    ccl_layout(g, ctx->num_sms, (char*)workspace, &ws);
    return ccl_resolve(ctx, nullptr, bits, g, ws, counts_dev);
}

int yam_ccl_emit_rows_bounded(yam_ctx* ctx, const uint32_t* bits, int64_t n, int64_t h, int64_t w, const void* workspace,
                              const int32_t* remap_dev, int64_t remap_size, int64_t row_begin, int64_t row_end, int32_t* labels);

int yam_ccl_emit_rows(yam_ctx* ctx, const uint32_t* bits, int64_t n, int64_t h, int64_t w, const void* workspace,
                      const int32_t* remap_dev, int64_t row_begin, int64_t row_end, int32_t* labels) {
    return yam_ccl_emit_rows_bounded(ctx, bits, n, h, w, workspace, remap_dev, 0, row_begin, row_end, labels);
}

int yam_ccl_emit_rows_bounded(yam_ctx* ctx, const uint32_t* bits, int64_t n, int64_t h, int64_t w, const void* workspace,
                              const int32_t* remap_dev, int64_t remap_size, int64_t row_begin, int64_t row_end, int32_t* labels) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(remap_size >= 0 && remap_size < (1ll << 31), "ccl_emit_rows: bad remap_size");
    YAM_REQUIRE(bits && workspace && labels, "ccl_emit_rows: NULL argument");
    CclGeom g;
    if (int rc = ccl_geometry(n, h, w, &g)) return rc;
    const int64_t total_rows = n * h;
    YAM_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= total_rows, "ccl_emit_rows: bad row range [%lld, %lld)",
                (long long)row_begin, (long long)row_end);
    YAM_REQUIRE(!remap_dev || n == 1, "ccl_emit_rows: a remap table applies to a single frame");
    CclWorkspace ws;
    ccl_layout(g, ctx->num_sms, (char*)const_cast<void*>(workspace), &ws);
    return ccl_emit(ctx, bits, g, ws, labels, remap_dev, row_begin * g.wpr, row_end * g.wpr, (int)remap_size);
}

int yam_region_props(yam_ctx* ctx, const int32_t* labels, const void* intensity, int intensity_dtype, int64_t h,
                     int64_t w, int64_t n_labels, int64_t* props_dev) {
    return region_props_impl(ctx, labels, intensity, intensity_dtype, 1, h, w, nullptr, n_labels, props_dev);
}

int yam_region_props_stack(yam_ctx* ctx, const int32_t* labels, const void* intensity, int intensity_dtype, int64_t n,
                           int64_t h, int64_t w, const int64_t* offsets_dev, int64_t total, int64_t* props_dev) {
    YAM_REQUIRE(offsets_dev != nullptr, "region_props_stack: offsets_dev is NULL");
    return region_props_impl(ctx, labels, intensity, intensity_dtype, n, h, w, offsets_dev, total, props_dev);
}

}  // extern "C"
