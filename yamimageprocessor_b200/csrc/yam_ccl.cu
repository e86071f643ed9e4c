// K10 connected-component labelling and K11 region properties.
//
// CCL (8-connectivity, raster-first canonical numbering) as a union-find over RUN SEGMENTS:
//   pack      the u8 mask is packed to 1 bit/pixel; every maximal run of set bits inside one 32-pixel
//             word is a node.  Nodes are numbered COMPACTLY in raster order (exclusive prefix sum of
//             the per-word segment counts), so the parent array has one int per segment (~1-2 % of
//             the pixels) and every union-find access stays in L2.
//   union     one thread per word links each of its segments to the segment that continues it in
//             the previous word and to every 8-connected segment in the row above (lock-free union
//             by minimum index with atomicMin, path halving in find).
//   flatten   parent[node] = root(node); roots are counted per word.
//   rootlabel exclusive prefix of the root counts = rank of every root.  Linking by minimum index
//             makes the root the segment holding the component's first pixel in raster order, so
//             rank(root) + 1 IS the canonical label; it is stored negated in the root's slot.
//   final     one thread per word resolves its segments (at most two dependent loads) and the block
//             writes the int32 labels through a swizzled shared-memory tile with 128-bit coalesced
//             stores.
// HBM traffic: 1 B/px mask read + 4 B/px label write + O(words) bookkeeping (12 B per 32 px).
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

// read-only find for the flatten pass: there every slot is written by its owner only, so that
// "parent == root" holds for all nodes afterwards (halving stores from other threads would race)
__device__ __forceinline__ int find_root_ro(const int* __restrict__ P, int x) {
    int p = __ldcg(P + x);
    while (p != x) {
        x = p;
        p = __ldcg(P + x);
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

__device__ __forceinline__ uint32_t seg_starts(uint32_t b) { return b & ~(b << 1); }

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

__device__ __forceinline__ uint32_t block_sum_u32(uint32_t v, uint32_t* s_tmp) {
    v = yam_warp_sum(v);
    if ((threadIdx.x & 31) == 0) s_tmp[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t t = 0;
    if (threadIdx.x < 32) {
        t = threadIdx.x < kThreads / 32 ? s_tmp[threadIdx.x] : 0u;
        t = yam_warp_sum(t);
    }
    return t;  // valid in warp 0
}

// exclusive prefix of v within the block (kThreads threads); returns the prefix for this thread
__device__ __forceinline__ uint32_t block_excl_scan_u32(uint32_t v, uint32_t* s_tmp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) s_tmp[warp] = incl;
    __syncthreads();
    uint32_t off = 0;
#pragma unroll
    for (int i = 0; i < kThreads / 32; i++)
        if (i < warp) off += s_tmp[i];
    return off + incl - v;
}

// In-place exclusive scan of data[0..n) by ONE block (all kThreads threads); total -> *total_out.
__device__ void block_scan_array(uint32_t* __restrict__ data, int64_t n, uint32_t* __restrict__ total_out,
                                 uint32_t* s_tmp, uint32_t* s_carry) {
    if (threadIdx.x == 0) *s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += kThreads * 4) {
        const int64_t i0 = base + (int64_t)threadIdx.x * 4;
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = (i0 + k < n) ? __ldcg(data + i0 + k) : 0u;
        const uint32_t mine = v[0] + v[1] + v[2] + v[3];
        uint32_t run = *s_carry + block_excl_scan_u32(mine, s_tmp);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (i0 + k < n) data[i0 + k] = run;
            run += v[k];
        }
        __syncthreads();
        if (threadIdx.x == kThreads - 1) *s_carry = run;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = *s_carry;
}

// "last block done" election: returns true in every thread of the block that finishes last.
// `counter` must be 0 on entry and is reset to 0 by the elected block.
__device__ bool last_block_done(unsigned int* counter, unsigned int nblocks, int* s_flag) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicAdd(counter, 1u);
        *s_flag = (ticket == nblocks - 1);
        if (*s_flag) *counter = 0;
    }
    __syncthreads();
    if (*s_flag) __threadfence();
    return *s_flag != 0;
}

// ---- 1. pack: bits + per-block segment counts; the last block scans the counts ------------------
__global__ void __launch_bounds__(kThreads) ccl_pack_kernel(const uint8_t* __restrict__ mask, CclGeom g,
                                                            uint32_t* __restrict__ bits,
                                                            uint32_t* __restrict__ block_counts,
                                                            uint32_t* __restrict__ total_nodes,
                                                            unsigned int* __restrict__ counter) {
    __shared__ uint32_t s_tmp[kThreads / 32];
    __shared__ uint32_t s_carry;
    __shared__ int s_flag;
    const int64_t gw = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t b = 0;
    if (gw < g.total_words) {
        const int64_t row_id = gw / g.wpr;  // frame * h + y
        const int j = (int)(gw - row_id * g.wpr);
        const uint8_t* row = mask + row_id * (int64_t)g.w;
        const int x0 = j * 32;
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
    const uint32_t total = block_sum_u32(__popc(seg_starts(b)), s_tmp);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
    if (last_block_done(counter, gridDim.x, &s_flag))
        block_scan_array(block_counts, gridDim.x, total_nodes, s_tmp, &s_carry);
}

// ---- 1b. the same bookkeeping when the mask already is bit-packed (fused segmentation path) --------
__global__ void __launch_bounds__(kThreads) ccl_count_kernel(const uint32_t* __restrict__ bits, CclGeom g,
                                                             uint32_t* __restrict__ block_counts,
                                                             uint32_t* __restrict__ total_nodes,
                                                             unsigned int* __restrict__ counter) {
    __shared__ uint32_t s_tmp[kThreads / 32];
    __shared__ uint32_t s_carry;
    __shared__ int s_flag;
    const int64_t gw = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t b = gw < g.total_words ? bits[gw] : 0u;
    const uint32_t total = block_sum_u32(__popc(seg_starts(b)), s_tmp);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
    if (last_block_done(counter, gridDim.x, &s_flag))
        block_scan_array(block_counts, gridDim.x, total_nodes, s_tmp, &s_carry);
}

// ---- 2. node base per word; parent init; node -> (word, start bit) ----------------------------------
__global__ void __launch_bounds__(kThreads) ccl_nodebase_kernel(const uint32_t* __restrict__ bits, CclGeom g,
                                                                const uint32_t* __restrict__ block_offsets,
                                                                uint32_t* __restrict__ nbase, int* __restrict__ P,
                                                                uint32_t* __restrict__ node_info) {
    __shared__ uint32_t s_tmp[kThreads / 32];
    const int64_t gw = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t b = gw < g.total_words ? bits[gw] : 0u;
    uint32_t starts = seg_starts(b);
    const uint32_t base = block_offsets[blockIdx.x] + block_excl_scan_u32(__popc(starts), s_tmp);
    if (gw < g.total_words) {
        nbase[gw] = base;
        uint32_t k = 0;
        while (starts) {
            const int sbit = __ffs(starts) - 1;
            starts &= starts - 1;
            P[base + k] = (int)(base + k);
            node_info[base + k] = ((uint32_t)gw << 5) | (uint32_t)sbit;
            k++;
        }
    }
}

// ---- 3. union: one thread per segment ----------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) ccl_union_kernel(const uint32_t* __restrict__ bits,
                                                             const uint32_t* __restrict__ nbase,
                                                             const uint32_t* __restrict__ node_info, CclGeom g,
                                                             const uint32_t* __restrict__ total_nodes,
                                                             int* __restrict__ P) {
    const int total = (int)*total_nodes;
    for (int node = blockIdx.x * blockDim.x + threadIdx.x; node < total; node += gridDim.x * blockDim.x) {
        const uint32_t info = node_info[node];
        const int64_t gw = info >> 5;
        const int s = (int)(info & 31u);
        const uint32_t b = bits[gw];
        const int64_t row_id = gw / g.wpr;
        const int j = (int)(gw - row_id * g.wpr);
        const int y = (int)(row_id % g.h);
        // horizontal: a segment at bit 0 continues the segment that ends at bit 31 of the previous word
        if (s == 0 && j > 0) {
            const uint32_t pv = bits[gw - 1];
            if (pv >> 31) unite(P, node, (int)nbase[gw - 1] + seg_index(pv, 31));
        }
        if (y == 0) continue;
        // vertical: 34-column window of the row above; bit k of `above` <-> column 32*j + k - 1
        const uint32_t u = bits[gw - g.wpr];
        const uint32_t up = j > 0 ? bits[gw - g.wpr - 1] : 0u;
        const uint32_t un = j + 1 < g.wpr ? bits[gw - g.wpr + 1] : 0u;
        const unsigned long long above =
            (unsigned long long)(up >> 31) | ((unsigned long long)u << 1) | ((unsigned long long)(un & 1u) << 33);
        const uint32_t from_s = b >> s;
        const int len = (~from_s) ? __ffs(~from_s) - 1 : 32 - s;
        const int e = s + len - 1;
        const unsigned long long wmask = ((1ull << (e + 3)) - 1ull) & ~((1ull << s) - 1ull);  // bits s .. e+2
        unsigned long long m = above & wmask;
        if (!m) continue;
        const int base_u = u ? (int)nbase[gw - g.wpr] : 0;
        while (m) {
            const int k0 = __ffsll((long long)m) - 1;
            m &= m + (1ull << k0);  // clear the lowest run of ones
            int other;
            if (k0 == 0) {
                other = (int)nbase[gw - g.wpr - 1] + seg_index(up, 31);
            } else if (k0 == 33) {
                other = (int)nbase[gw - g.wpr + 1];  // bit 0 of the next word starts its first segment
            } else {
                other = base_u + seg_index(u, k0 - 1);
            }
            unite(P, node, other);
        }
    }
}

// ---- 4. flatten (one thread per node) + root counts per 256-node chunk; last block scans them ---------
__global__ void __launch_bounds__(kThreads) ccl_flatten_kernel(const uint32_t* __restrict__ total_nodes,
                                                               int* __restrict__ P,
                                                               uint32_t* __restrict__ chunk_roots,
                                                               uint32_t* __restrict__ total_roots,
                                                               unsigned int* __restrict__ counter) {
    __shared__ uint32_t s_tmp[kThreads / 32];
    __shared__ uint32_t s_carry;
    __shared__ int s_flag;
    const int total = (int)*total_nodes;
    const int chunks = (total + kThreads - 1) / kThreads;
    for (int c = blockIdx.x; c < chunks; c += gridDim.x) {
        const int node = c * kThreads + threadIdx.x;
        uint32_t is_root = 0;
        if (node < total) {
            const int r = find_root_ro(P, node);
            if (r == node) is_root = 1;
            else P[node] = r;
        }
        const uint32_t t = block_sum_u32(is_root, s_tmp);
        if (threadIdx.x == 0) chunk_roots[c] = t;
        __syncthreads();
    }
    if (last_block_done(counter, gridDim.x, &s_flag))
        block_scan_array(chunk_roots, chunks, total_roots, s_tmp, &s_carry);
}

// ---- 5. root labels: P[root] = -(rank + 1) ---------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) ccl_rootlabel_kernel(const uint32_t* __restrict__ total_nodes,
                                                                 const uint32_t* __restrict__ chunk_offsets,
                                                                 int* __restrict__ P) {
    __shared__ uint32_t s_tmp[kThreads / 32];
    const int total = (int)*total_nodes;
    const int chunks = (total + kThreads - 1) / kThreads;
    for (int c = blockIdx.x; c < chunks; c += gridDim.x) {
        const int node = c * kThreads + threadIdx.x;
        const uint32_t is_root = (node < total && __ldcg(P + node) == node) ? 1u : 0u;
        const uint32_t rank = chunk_offsets[c] + block_excl_scan_u32(is_root, s_tmp);
        if (is_root) P[node] = -(int)(rank + 1);
        __syncthreads();
    }
}

// ---- 6. per-frame root offsets and component counts (one thread per frame) ---------------------------------
__global__ void ccl_frame_offsets_kernel(const uint32_t* __restrict__ nbase, CclGeom g,
                                         const uint32_t* __restrict__ total_nodes,
                                         const uint32_t* __restrict__ total_roots,
                                         const uint32_t* __restrict__ chunk_offsets, const int* __restrict__ P,
                                         uint32_t* __restrict__ frame_off, int32_t* __restrict__ counts) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= g.frames) return;
    auto roots_before_frame = [&](int frame) -> uint32_t {
        if (frame >= g.frames) return *total_roots;
        const uint32_t n0 = nbase[(int64_t)frame * g.words_per_frame];
        if (n0 >= *total_nodes) return *total_roots;
        uint32_t r = chunk_offsets[n0 / kThreads];
        for (uint32_t i = (n0 / kThreads) * kThreads; i < n0; i++) r += (P[i] < 0) ? 1u : 0u;
        return r;
    };
    const uint32_t here = roots_before_frame(f), next = roots_before_frame(f + 1);
    frame_off[f] = here;
    if (counts) counts[f] = (int32_t)(next - here);
}

// ---- 7. final labels ----------------------------------------------------------------------------
// block = kFinalThreads words; labels are staged in shared memory (swizzled at int4 granularity:
// both the per-thread row writes and the coalesced read-out are conflict free)
constexpr int kFinalThreads = 128;

__global__ void __launch_bounds__(kFinalThreads) ccl_final_kernel(const uint32_t* __restrict__ bits,
                                                                  const uint32_t* __restrict__ nbase,
                                                                  const uint32_t* __restrict__ frame_off, CclGeom g,
                                                                  const int* __restrict__ P,
                                                                  int32_t* __restrict__ labels) {
    __shared__ int4 s_out[kFinalThreads * 8];
    const int64_t gw0 = (int64_t)blockIdx.x * kFinalThreads;
    const int64_t gw = gw0 + threadIdx.x;
    const int t = threadIdx.x;
    uint32_t b = 0;
    int off = 0, base = 0;
    if (gw < g.total_words) {
        b = bits[gw];
        if (b) {
            base = (int)nbase[gw];
            if (g.frames > 1) off = (int)frame_off[gw / g.words_per_frame];
        }
    }
    int32_t out[32];
    if (b) {
        // first-level loads for up to 4 segments are issued together; more segments are rare
        uint32_t starts = seg_starts(b);
        const int nseg = __popc(starts);
        int lab[4];
#pragma unroll
        for (int k = 0; k < 4; k++) lab[k] = k < nseg ? __ldg(P + base + k) : -1;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (lab[k] >= 0) lab[k] = __ldg(P + lab[k]);  // non-root: its root holds -(label)
        int cur = 0, k = 0;
#pragma unroll
        for (int i = 0; i < 32; i++) {
            const bool set = (b >> i) & 1u;
            const bool start = set && (i == 0 || !((b >> (i - 1)) & 1u));
            if (start) {
                int v;
                if (k < 4) {
                    v = k == 0 ? lab[0] : k == 1 ? lab[1] : k == 2 ? lab[2] : lab[3];
                } else {
                    v = __ldg(P + base + k);
                    if (v >= 0) v = __ldg(P + v);
                }
                cur = -v - off;
                k++;
            }
            out[i] = set ? cur : 0;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 32; i++) out[i] = 0;
    }
#pragma unroll
    for (int q = 0; q < 8; q++)
        s_out[t * 8 + ((q + t) & 7)] = make_int4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
    __syncthreads();
    if ((g.w & 31) == 0 && ((reinterpret_cast<uintptr_t>(labels) & 15) == 0)) {
        // rows are whole words: the block's words are consecutive labels
        int4* dst = reinterpret_cast<int4*>(labels + gw0 * 32);
        const int64_t valid_words = g.total_words - gw0 < kFinalThreads ? g.total_words - gw0 : kFinalThreads;
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int idx = it * kFinalThreads + t;  // int4 index inside the block tile
            const int wd = idx >> 3, q = idx & 7;
            if (wd < valid_words) __stcs(dst + idx, s_out[wd * 8 + ((q + wd) & 7)]);
        }
    } else if (gw < g.total_words) {
        const int64_t row_id = gw / g.wpr;
        const int j = (int)(gw - row_id * g.wpr);
        int32_t* drow = labels + row_id * (int64_t)g.w + (int64_t)j * 32;
        const int valid = min(32, g.w - j * 32);
        for (int i = 0; i < valid; i++) {
            const int4 v = s_out[t * 8 + (((i >> 2) + t) & 7)];
            drow[i] = (i & 3) == 0 ? v.x : (i & 3) == 1 ? v.y : (i & 3) == 2 ? v.z : v.w;
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

}  // namespace

extern "C" {

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

static int ccl_label_impl(yam_ctx* ctx, const void* mask, const uint32_t* bits_in, int32_t* labels, int64_t n,
                          int64_t h, int64_t w, int32_t* counts_dev, int32_t* counts_host) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE((mask || bits_in) && labels && n > 0 && h > 0 && w > 0, "ccl: bad arguments");
    YAM_REQUIRE(n <= 65535, "ccl: at most 65535 frames per call");
    CclGeom g;
    g.h = (int)h;
    g.w = (int)w;
    g.wpr = (int)((w + 31) / 32);
    g.words_per_frame = (int64_t)h * g.wpr;
    g.total_words = g.words_per_frame * n;
    g.frames = (int)n;
    YAM_REQUIRE(g.words_per_frame * 32 < (1ll << 31), "ccl: frame too large for int32 labels (%lld x %lld)",
                (long long)h, (long long)w);
    YAM_REQUIRE(g.total_words * 16 < (1ll << 31) && g.total_words < (1ll << 27),
                "ccl: stack too large for one call (%lld words); split the stack", (long long)g.total_words);
    const int64_t nblocks = (g.total_words + kThreads - 1) / kThreads;
    const int64_t max_nodes = g.total_words * 16;  // at most 16 run segments per 32-pixel word
    const int64_t max_chunks = (max_nodes + kThreads - 1) / kThreads;
    // scratch: bits | nbase | blockA | chunkB | frame_off | misc(totals[2], counter) | counts | node_info | P
    const size_t words_bytes = yam_align_up((size_t)g.total_words * 4, 256);
    const size_t blk_bytes = yam_align_up((size_t)nblocks * 4, 256);
    const size_t chunk_bytes = yam_align_up((size_t)max_chunks * 4, 256);
    const size_t frame_bytes = yam_align_up((size_t)n * 4, 256);
    const size_t node_bytes = yam_align_up((size_t)max_nodes * 4, 256);
    void* scratch = nullptr;
    if (int rc = yam_scratch(ctx, 2 * words_bytes + blk_bytes + chunk_bytes + 2 * frame_bytes + 256 + 2 * node_bytes, &scratch)) return rc;
    char* sp = (char*)scratch;
    uint32_t* bits_scratch = (uint32_t*)sp; sp += words_bytes;
    const uint32_t* bits = bits_in ? bits_in : bits_scratch;
    uint32_t* nbase = (uint32_t*)sp; sp += words_bytes;
    uint32_t* blockA = (uint32_t*)sp; sp += blk_bytes;
    uint32_t* chunkB = (uint32_t*)sp; sp += chunk_bytes;
    uint32_t* frame_off = (uint32_t*)sp; sp += frame_bytes;
    uint32_t* totals = (uint32_t*)sp;            // [0] nodes, [1] roots
    unsigned int* counter = (unsigned int*)(sp + 16); sp += 256;
    int32_t* counts = counts_dev ? counts_dev : (int32_t*)sp; sp += frame_bytes;
    uint32_t* node_info = (uint32_t*)sp; sp += node_bytes;
    int* P = (int*)sp;

    // the election counter is self-resetting, but scratch is shared with other operators
    YAM_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), ctx->stream));
    const unsigned gblocks = (unsigned)nblocks;
    const unsigned pgrid = (unsigned)(ctx->num_sms * 8);  // persistent grids for the node-parallel kernels
    if (bits_in)
        ccl_count_kernel<<<gblocks, kThreads, 0, ctx->stream>>>(bits, g, blockA, totals, counter);
    else
        ccl_pack_kernel<<<gblocks, kThreads, 0, ctx->stream>>>((const uint8_t*)mask, g, bits_scratch, blockA, totals, counter);
    YAM_LAUNCHED(ctx);
    ccl_nodebase_kernel<<<gblocks, kThreads, 0, ctx->stream>>>(bits, g, blockA, nbase, P, node_info);
    YAM_LAUNCHED(ctx);
    ccl_union_kernel<<<pgrid, kThreads, 0, ctx->stream>>>(bits, nbase, node_info, g, totals, P);
    YAM_LAUNCHED(ctx);
    ccl_flatten_kernel<<<pgrid, kThreads, 0, ctx->stream>>>(totals, P, chunkB, totals + 1, counter);
    YAM_LAUNCHED(ctx);
    ccl_rootlabel_kernel<<<pgrid, kThreads, 0, ctx->stream>>>(totals, chunkB, P);
    YAM_LAUNCHED(ctx);
    ccl_frame_offsets_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(nbase, g, totals, totals + 1, chunkB, P,
                                                                                 frame_off, counts);
    YAM_LAUNCHED(ctx);
    const unsigned fblocks = (unsigned)((g.total_words + kFinalThreads - 1) / kFinalThreads);
    ccl_final_kernel<<<fblocks, kFinalThreads, 0, ctx->stream>>>(bits, nbase, frame_off, g, P, labels);
    YAM_LAUNCHED(ctx);
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
