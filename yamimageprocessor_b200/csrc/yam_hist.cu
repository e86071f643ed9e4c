// K7 / K8 histogram family: histogram, Otsu threshold, equalizeHist, CLAHE.
//
// 16-bit histograms: 65536 bins x 4 B = 256 KiB does not fit the 227 KiB of shared memory an SM
// offers, so each CTA keeps a privatised histogram of PACKED 16-bit counters (65536 x 2 B =
// 128 KiB, two bins per 32-bit word, shared-memory atomics).  A counter that reaches 0x8000 is
// "spilled": 0x8000 is subtracted from it and added to a global overflow histogram, so a half
// never carries into its neighbour and counts stay exact for any region size.
//   * Otsu / plain histogram: P CTAs per frame (row slabs); every CTA stores its packed histogram to a
//     scratch slab and a second small kernel adds the slabs per bin (no flush atomics).
//   * CLAHE: ONE CTA per CLAHE tile does histogram -> clip -> redistribute -> prefix sum -> LUT
//     entirely on chip and writes only the 128 KiB LUT (no histogram ever reaches HBM).
// 8-bit histograms use per-warp privatised 256-bin shared histograms.
// Otsu's fp64 recurrence over the 65536 bins is sequential per frame, but its result is an arg-max:
// otsu_certify_kernel proves it in parallel from exact integer prefix sums and an error bound of the
// recurrence; frames it cannot certify go through the exact staged chain kernels.  Nothing is read
// back and no host thread takes part (the host scan functions below serve the host-only C-ABI helpers).
#include <cooperative_groups.h>
#include <math.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "yam_common.cuh"
#include "yam_host.h"

int yam_threshold_dev(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t frame_px, int dtype,
                      const int32_t* t_dev, double maxval);

namespace {

namespace cg = cooperative_groups;
constexpr int kHistThreads = 1024;
constexpr int kBins16 = 65536;
constexpr int kWords16 = kBins16 / 2;
constexpr size_t kSmem16 = (size_t)kWords16 * 4;

// ---------------------------------------------------------------------------------------------
// packed 16-bit shared-memory counters with spill

template <typename CntT>
__device__ __forceinline__ void bump16(uint32_t* __restrict__ sh, CntT* __restrict__ overflow,
                                       int* __restrict__ spill_flag, uint32_t v, uint32_t count) {
    const uint32_t shift = (v & 1u) * 16u;
    const uint32_t old = (atomicAdd(&sh[v >> 1], count << shift) >> shift) & 0xffffu;
    if (old < 0x8000u && old + count >= 0x8000u) {
        atomicSub(&sh[v >> 1], 0x8000u << shift);
        atomicAdd(&overflow[v], (CntT)0x8000u);
        if (spill_flag) *spill_flag = 1;
    }
}

// Accumulate rows [r0, r1) x cols [c0, c1) of a (possibly reflect-padded) region into `sh`.
template <typename CntT, bool NARROW = false>
__device__ __forceinline__ void accumulate16(uint32_t* __restrict__ sh, CntT* __restrict__ overflow,
                                             int* __restrict__ spill_flag, const uint16_t* __restrict__ src, int h, int w, int c0, int c1,
                                             int r0, int r1) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const bool aligned = ((w & 7) == 0) && ((c0 & 7) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    const int cw = c1 - c0;
    const int cin = min(c1, w) - c0;               // columns inside the image
    const int nvec = aligned ? (cin > 0 ? cin / 8 : 0) : 0;
    // Eight pixels per call: all eight shared-memory atomics are issued back to back (independent), the
    // returned words are checked together afterwards (one predicate per call instead of a dependent
    // atomic -> compare -> branch chain per pixel).  A counter spills when THIS increment took its half
    // to 0x8000 (bit 15 of the half goes 0 -> 1); between that atomic and the subtraction in the rare
    // path at most threads x 8 further increments can land, far below the 0x8000 of head room, so a
    // half never carries into its neighbour.
    auto consume = [&](const uint4& q) {
        const uint32_t wd[4] = {q.x, q.y, q.z, q.w};
        uint32_t old[8];
        uint32_t flips = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t v = (i & 1) ? (wd[i >> 1] >> 16) : (wd[i >> 1] & 0xffffu);
            const uint32_t inc = 1u + (v & 1u) * 0xffffu;             // 1 or 0x10000
            old[i] = atomicAdd(&sh[v >> 1], inc);
            flips |= ~old[i] & (old[i] + inc) & (inc << 15);
        }
        if (flips) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint32_t v = (i & 1) ? (wd[i >> 1] >> 16) : (wd[i >> 1] & 0xffffu);
                const uint32_t inc = 1u + (v & 1u) * 0xffffu;
                if (~old[i] & (old[i] + inc) & (inc << 15)) {
                    atomicSub(&sh[v >> 1], inc << 15);
                    atomicAdd(&overflow[v], (CntT)0x8000u);
                    if (spill_flag) *spill_flag = 1;
                }
            }
        }
    };
    // Narrow regions (CLAHE tiles of single frames: 256 or 512 px wide = one or two vectors per lane and row):
    // a warp issues the loads of FOUR vectors (four rows x 1 or two rows x 2) before the first atomic; row by
    // row the kernel paid one exposed memory round trip per vector (8-32 sequential round trips per warp and tile).
    if (NARROW && nvec > 0 && nvec <= 64) {
        const int vpr = nvec <= 32 ? 1 : 2;          // vectors per lane and row
        const int rows_per = 4 / vpr;                // rows per batch
        for (int r = r0 + warp; r < r1; r += rows_per * nwarps) {
            uint4 q[4];
#pragma unroll
            for (int s4 = 0; s4 < 4; s4++) {
                const int rr = r + (s4 / vpr) * nwarps, vi = lane + 32 * (s4 % vpr);
                if (rr < r1 && vi < nvec)
                    q[s4] = yam_ld_stream(reinterpret_cast<const uint4*>(src + (int64_t)yam_border(rr, h, YAM_BORDER_REFLECT101) * w + c0) + vi);
            }
#pragma unroll
            for (int s4 = 0; s4 < 4; s4++) {
                const int rr = r + (s4 / vpr) * nwarps, vi = lane + 32 * (s4 % vpr);
                if (rr < r1 && vi < nvec) consume(q[s4]);
            }
            if (nvec * 8 < cw) {                     // columns past the last whole vector / the image edge
                for (int j = 0; j < rows_per; j++) {
                    const int rr = r + j * nwarps;
                    if (rr >= r1) break;
                    const uint16_t* row = src + (int64_t)yam_border(rr, h, YAM_BORDER_REFLECT101) * w;
                    for (int c = nvec * 8 + lane; c < cw; c += 32) bump16(sh, overflow, spill_flag, row[yam_border(c0 + c, w, YAM_BORDER_REFLECT101)], 1u);
                }
            }
        }
        return;
    }
    for (int r = r0 + warp; r < r1; r += nwarps) {
        const int gy = yam_border(r, h, YAM_BORDER_REFLECT101);
        const uint16_t* row = src + (int64_t)gy * w;
        const uint4* vrow = reinterpret_cast<const uint4*>(row + c0);
        int v = lane;
        // four 16-byte loads in flight per thread AT ALL TIMES: every consumed vector is replaced by the
        // load of the vector four steps ahead before the next one is consumed (with one CTA of 1024 threads
        // per SM a single load per thread, 16 KiB in flight, left the kernel waiting on HBM latency)
        if (v + 96 < nvec) {
            uint4 q0 = yam_ld_stream(vrow + v), q1 = yam_ld_stream(vrow + v + 32);
            uint4 q2 = yam_ld_stream(vrow + v + 64), q3 = yam_ld_stream(vrow + v + 96);
            v += 128;
            for (; v + 96 < nvec; v += 128) {
                consume(q0);
                q0 = yam_ld_stream(vrow + v);
                consume(q1);
                q1 = yam_ld_stream(vrow + v + 32);
                consume(q2);
                q2 = yam_ld_stream(vrow + v + 64);
                consume(q3);
                q3 = yam_ld_stream(vrow + v + 96);
            }
            consume(q0);
            consume(q1);
            consume(q2);
            consume(q3);
        }
        for (; v < nvec; v += 32) consume(yam_ld_stream(vrow + v));
        for (int c = nvec * 8 + lane; c < cw; c += 32) {
            const int gx = yam_border(c0 + c, w, YAM_BORDER_REFLECT101);
            bump16(sh, overflow, spill_flag, row[gx], 1u);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// plain histogram (Otsu): grid (parts, 1, frames)
// Flush without atomics: every CTA stores its packed private histogram (128 KiB, coalesced 16-byte
// stores, zeros included) to a scratch slab and hist16_reduce_kernel adds the slabs per bin.  The
// flush by RED (one per non-empty bin and CTA: ~30 000 x 148 for a 4096^2 frame) was 61 % of the
// plain kernel's 43 us there (ncu source view: the warps sit on the scoreboard behind REDG.64).
__device__ __forceinline__ void store_packed16(const uint32_t* __restrict__ sh, uint32_t* __restrict__ slab) {
    const uint4* s4 = reinterpret_cast<const uint4*>(sh);
    uint4* d4 = reinterpret_cast<uint4*>(slab);
    for (int i = threadIdx.x; i < kWords16 / 4; i += kHistThreads) d4[i] = s4[i];
}

// grid (kWords16 / 256, slots): thread = one packed word (two bins) of one slot, summed over `nparts` slabs
// and ADDED to the counts already in `out` (the spill path of the private histograms adds there directly)
template <typename CntT>
__global__ void __launch_bounds__(256) hist16_reduce_kernel(const uint32_t* __restrict__ slabs, int nparts,
                                                            CntT* __restrict__ out) {
    const int wi = blockIdx.x * 256 + threadIdx.x;
    const uint32_t* p = slabs + (int64_t)blockIdx.y * nparts * kWords16 + wi;
    uint32_t c0 = 0, c1 = 0;     // nparts x 0x7fff < 2^32
    int q = 0;
    for (; q + 4 <= nparts; q += 4) {
        const uint32_t a = __ldcg(p + (int64_t)q * kWords16), b = __ldcg(p + (int64_t)(q + 1) * kWords16);
        const uint32_t c = __ldcg(p + (int64_t)(q + 2) * kWords16), d = __ldcg(p + (int64_t)(q + 3) * kWords16);
        c0 += (a & 0xffffu) + (b & 0xffffu) + (c & 0xffffu) + (d & 0xffffu);
        c1 += (a >> 16) + (b >> 16) + (c >> 16) + (d >> 16);
    }
    for (; q < nparts; q++) {
        const uint32_t a = __ldcg(p + (int64_t)q * kWords16);
        c0 += a & 0xffffu;
        c1 += a >> 16;
    }
    CntT* o = out + (int64_t)blockIdx.y * kBins16 + 2 * wi;
    o[0] += (CntT)c0;
    o[1] += (CntT)c1;
}

__global__ void __launch_bounds__(kHistThreads, 1) hist16_slab_kernel(const uint16_t* __restrict__ src, int h, int w,
                                                                      unsigned long long* __restrict__ hist,
                                                                      uint32_t* __restrict__ slabs) {
    extern __shared__ __align__(16) uint32_t sh[];
    src += (int64_t)blockIdx.z * h * w;
    unsigned long long* out = hist + (int64_t)blockIdx.z * kBins16;
    for (int i = threadIdx.x; i < kWords16; i += kHistThreads) sh[i] = 0;
    __syncthreads();
    const int parts = gridDim.x;
    const int rows_per = (h + parts - 1) / parts;
    const int r0 = min(h, (int)blockIdx.x * rows_per), r1 = min(h, r0 + rows_per);
    accumulate16<unsigned long long>(sh, out, nullptr, src, h, w, 0, w, r0, r1);
    __syncthreads();
    store_packed16(sh, slabs + ((int64_t)blockIdx.z * parts + blockIdx.x) * kWords16);
}

// 8-bit histogram: grid (blocks, 1, frames); per-warp private histograms
__global__ void __launch_bounds__(256) hist8_kernel(const uint8_t* __restrict__ src, int64_t frame_px,
                                                    unsigned long long* __restrict__ hist) {
    __shared__ uint32_t sh[8][256];
    src += (int64_t)blockIdx.z * frame_px;
    unsigned long long* out = hist + (int64_t)blockIdx.z * 256;
    for (int i = threadIdx.x; i < 8 * 256; i += 256) (&sh[0][0])[i] = 0;
    __syncthreads();
    uint32_t* mine = sh[threadIdx.x >> 5];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool aligned = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    int64_t done = 0;
    if (aligned) {
        const int64_t groups = frame_px / 16;
        for (int64_t g = tid; g < groups; g += stride) {
            const uint4 q = yam_ld_stream(reinterpret_cast<const uint4*>(src) + g);
            const uint32_t wd[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                atomicAdd(&mine[wd[i] & 0xff], 1u);
                atomicAdd(&mine[(wd[i] >> 8) & 0xff], 1u);
                atomicAdd(&mine[(wd[i] >> 16) & 0xff], 1u);
                atomicAdd(&mine[wd[i] >> 24], 1u);
            }
        }
        done = groups * 16;
    }
    for (int64_t i = done + tid; i < frame_px; i += stride) atomicAdd(&mine[src[i]], 1u);
    __syncthreads();
    uint32_t total = 0;
#pragma unroll
    for (int wv = 0; wv < 8; wv++) total += sh[wv][threadIdx.x];
    if (total) atomicAdd(&out[threadIdx.x], (unsigned long long)total);
}

// ---------------------------------------------------------------------------------------------
// Otsu scan on device, one thread per frame (used for stacks; a single frame is scanned on the
// host because 65536 dependent fp64 divisions are ~10x faster on a CPU core than on one GPU thread).
__global__ void otsu_scan_kernel(const unsigned long long* __restrict__ hist, int bins, int64_t n,
                                 int32_t* __restrict__ out) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    const unsigned long long* h = hist + f * bins;
    double total = 0.0, mu = 0.0;
    for (int i = 0; i < bins; i++) {
        const double c = (double)h[i];
        total = __dadd_rn(total, c);
        mu = __dadd_rn(mu, __dmul_rn((double)i, c));
    }
    if (!(total > 0.0)) {
        out[f] = 0;
        return;
    }
    const double scale = __ddiv_rn(1.0, total);
    mu = __dmul_rn(mu, scale);
    double mu1 = 0.0, q1 = 0.0, best = 0.0;
    int best_i = 0;
    const double eps = 1.1920928955078125e-07;
    for (int i = 0; i < bins; i++) {
        const double p = __dmul_rn((double)h[i], scale);
        mu1 = __dmul_rn(mu1, q1);
        q1 = __dadd_rn(q1, p);
        const double q2 = __dsub_rn(1.0, q1);
        if (fmin(q1, q2) < eps || fmax(q1, q2) > 1.0 - eps) continue;
        mu1 = __ddiv_rn(__dadd_rn(mu1, __dmul_rn((double)i, p)), q1);
        const double mu2 = __ddiv_rn(__dsub_rn(mu, __dmul_rn(q1, mu1)), q2);
        const double d = __dsub_rn(mu1, mu2);
        const double sigma = __dmul_rn(__dmul_rn(__dmul_rn(q1, q2), d), d);
        if (sigma > best) {
            best = sigma;
            best_i = i;
        }
    }
    out[f] = best_i;
}

// ---------------------------------------------------------------------------------------------
// Exact staged Otsu scan (the fall-back behind otsu_certify_kernel, which fills OtsuMeta).  The recurrence
//     mu1 <- (mu1 * q1_prev + i * p_i) / q1_i,   q1_i = q1_prev + p_i
// has two sequential chains.  q1 does not depend on mu1, so it runs one tile ahead (one DADD per
// bin), the reciprocals 1/q1_i of a finished tile are computed by all lanes in parallel, and the mu1
// chain replaces the division by a multiplication with two FMA corrections (q = n*r;
// q += fma(-q, q1, n) * r): five dependent fp64 operations per bin instead of a ~100-cycle division.  The last stage evaluates sigma in
// parallel AND re-derives every mu1_i from its stored predecessor with a true IEEE division; a
// single differing bit makes the block redo its frame with the plain sequential loop, so the
// result is the reference's by construction, whatever the corrected quotient did.
struct OtsuMeta {
    double scale, mu;
    int first, last;  // occupied bin range the chain has to cover; first < 0: empty frame
    int certified;    // 1: otsu_certify_kernel proved the threshold, the chain / sigma kernels skip this frame
    int pad;
};
constexpr double kOtsuEps = 1.1920928955078125e-07;  // FLT_EPSILON

// Both chains in one warp per frame.  Per tile of kOtsuTile bins: all lanes stage p_i = h_i * scale
// of the NEXT tile, the reciprocals 1/q1_i and i * p_i of the CURRENT tile in shared memory (coalesced
// loads: nothing on a chain waits for DRAM); lane 0 then advances the q1 chain over the next tile
// and the mu1 chain over the current tile in ONE loop, so the single DADD of the q1 chain issues in
// the shadow of the five dependent operations of the mu1 chain; all lanes write the tiles back.
constexpr int kOtsuTile = 512;

__global__ void __launch_bounds__(32) otsu_chain_kernel(const unsigned long long* __restrict__ hist, int bins,
                                                        const OtsuMeta* __restrict__ meta,
                                                        double* __restrict__ q1arr, double* __restrict__ mu1arr) {
    __shared__ double s_pn[kOtsuTile], s_qn[kOtsuTile];                     // next tile: p in, q1 out
    __shared__ double s_q[kOtsuTile], s_r[kOtsuTile], s_ip[kOtsuTile], s_m[kOtsuTile];  // current tile
    const OtsuMeta m = meta[blockIdx.x];
    if (m.first < 0 || m.certified) return;
    const int64_t base = (int64_t)blockIdx.x * bins;
    const unsigned long long* h = hist + base;
    const int lane = threadIdx.x;
    const int span = m.last + 1 - m.first;
    const int tiles = (span + kOtsuTile - 1) / kOtsuTile;
    double q1 = 0.0, mu1 = 0.0, qprev = 0.0;
    bool plateau = false;
    // prologue: q1 chain over tile 0
    {
        const int cnt = min(kOtsuTile, span);
        for (int k = lane; k < cnt; k += 32) s_pn[k] = __dmul_rn((double)h[m.first + k], m.scale);
        __syncwarp();
        if (lane == 0)
            for (int k = 0; k < cnt; k++) {
                q1 = __dadd_rn(q1, s_pn[k]);
                s_qn[k] = q1;
            }
        __syncwarp();
    }
    for (int b = 0; b < tiles; b++) {
        const int t0 = m.first + b * kOtsuTile;
        const int cnt = min(kOtsuTile, m.last + 1 - t0);
        const int cnt_next = b + 1 < tiles ? min(kOtsuTile, m.last + 1 - (t0 + kOtsuTile)) : 0;
        for (int k = lane; k < cnt; k += 32) {
            const int i = t0 + k;
            const double q = s_qn[k];
            const double q2 = __dsub_rn(1.0, q);
            const bool skip = fmin(q, q2) < kOtsuEps || fmax(q, q2) > 1.0 - kOtsuEps;
            s_q[k] = q;
            s_r[k] = skip ? 0.0 : __drcp_rn(q);   // 0 marks a bin the reference skips
            s_ip[k] = __dmul_rn((double)i, __dmul_rn((double)h[i], m.scale));
            q1arr[base + i] = q;
        }
        __syncwarp();  // s_qn has been consumed: the next tile may overwrite s_pn / s_qn
        for (int k = lane; k < cnt_next; k += 32) s_pn[k] = __dmul_rn((double)h[t0 + kOtsuTile + k], m.scale);
        __syncwarp();
        if (lane == 0) {
            // blocks of 4 bins with the operands preloaded: shared-memory latency stays off the chains
            auto mu_step = [&](int k, double q, double r, double ip) {
                // Runs of EMPTY bins (sparse histograms: 12-bit data, small frames): q1 does not move and mu1
                // goes through x -> fl(fl(x q1) / q1); once an empty bin leaves mu1 unchanged, so does every
                // following empty bin: copy instead of five dependent fp64 operations per bin.
                // (bit-pattern compares: integer ALU latency instead of the fp64 pipe's; all values are >= +0)
                const bool empty_bin = __double_as_longlong(ip) == 0ll && __double_as_longlong(q) == __double_as_longlong(qprev);
                if (plateau && empty_bin) {
                    s_m[k] = mu1;
                    return;
                }
                // mu1 = (mu1 * q1_prev + i p_i) / q1_i with the corrected quotient
                const double t = __dmul_rn(mu1, qprev);
                const double nsum = __dadd_rn(t, ip);
                const double qq = __dmul_rn(nsum, r);
                const double e = __fma_rn(-qq, q, nsum);
                const double quo = __fma_rn(e, r, qq);
                const double nm = r == 0.0 ? t : quo;  // skipped bin: the reference multiplied by q1 before `continue`
                plateau = empty_bin && __double_as_longlong(r) != 0ll && __double_as_longlong(nm) == __double_as_longlong(mu1);
                mu1 = nm;
                s_m[k] = mu1;
                qprev = q;
            };
            const int both = (cnt < cnt_next ? cnt : cnt_next) & ~3;
            int k = 0;
            for (; k < both; k += 4) {
                double q[4], r[4], ip[4], pn[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    q[j] = s_q[k + j];
                    r[j] = s_r[k + j];
                    ip[j] = s_ip[k + j];
                    pn[j] = s_pn[k + j];
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (__double_as_longlong(pn[j]) != 0ll) q1 = __dadd_rn(q1, pn[j]);   // q1 chain, one tile ahead (x + 0 = x: no dependent add)
                    s_qn[k + j] = q1;
                    mu_step(k + j, q[j], r[j], ip[j]);
                }
            }
            for (int kk = k; kk < cnt_next; kk++) {
                q1 = __dadd_rn(q1, s_pn[kk]);
                s_qn[kk] = q1;
            }
            for (; k + 4 <= cnt; k += 4) {
                double q[4], r[4], ip[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    q[j] = s_q[k + j];
                    r[j] = s_r[k + j];
                    ip[j] = s_ip[k + j];
                }
#pragma unroll
                for (int j = 0; j < 4; j++) mu_step(k + j, q[j], r[j], ip[j]);
            }
            for (; k < cnt; k++) mu_step(k, s_q[k], s_r[k], s_ip[k]);
        }
        __syncwarp();
        for (int k = lane; k < cnt; k += 32) mu1arr[base + t0 + k] = s_m[k];
        __syncwarp();
    }
}

__device__ int otsu_scan_exact(const unsigned long long* __restrict__ h, int bins) {
    double total = 0.0, mu = 0.0;
    for (int i = 0; i < bins; i++) {
        const double c = (double)h[i];
        total = __dadd_rn(total, c);
        mu = __dadd_rn(mu, __dmul_rn((double)i, c));
    }
    if (!(total > 0.0)) return 0;
    const double scale = __ddiv_rn(1.0, total);
    mu = __dmul_rn(mu, scale);
    double mu1 = 0.0, q1 = 0.0, best = 0.0;
    int best_i = 0;
    for (int i = 0; i < bins; i++) {
        const double p = __dmul_rn((double)h[i], scale);
        mu1 = __dmul_rn(mu1, q1);
        q1 = __dadd_rn(q1, p);
        const double q2 = __dsub_rn(1.0, q1);
        if (fmin(q1, q2) < kOtsuEps || fmax(q1, q2) > 1.0 - kOtsuEps) continue;
        mu1 = __ddiv_rn(__dadd_rn(mu1, __dmul_rn((double)i, p)), q1);
        const double mu2 = __ddiv_rn(__dsub_rn(mu, __dmul_rn(q1, mu1)), q2);
        const double d = __dsub_rn(mu1, mu2);
        const double sigma = __dmul_rn(__dmul_rn(__dmul_rn(q1, q2), d), d);
        if (sigma > best) {
            best = sigma;
            best_i = i;
        }
    }
    return best_i;
}

// parallel: verify the chain bit for bit, evaluate sigma, first maximum
__global__ void __launch_bounds__(1024) otsu_sigma_kernel(const unsigned long long* __restrict__ hist, int bins,
                                                          const OtsuMeta* __restrict__ meta,
                                                          const double* __restrict__ q1arr,
                                                          const double* __restrict__ mu1arr, int32_t* __restrict__ out,
                                                          int* __restrict__ redo_count) {
    __shared__ double s_best[32];
    __shared__ int s_idx[32];
    __shared__ int s_fail;
    const OtsuMeta m = meta[blockIdx.x];
    if (m.certified) return;  // out[] already holds the proven threshold
    if (m.first < 0) {
        if (threadIdx.x == 0) out[blockIdx.x] = 0;
        return;
    }
    if (threadIdx.x == 0) s_fail = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * bins;
    const unsigned long long* h = hist + base;
    double best = 0.0;
    int best_i = 0;
    bool fail = false;
    for (int i = m.first + threadIdx.x; i <= m.last; i += blockDim.x) {
        const double q1 = q1arr[base + i], mu1 = mu1arr[base + i];
        const double qprev = i > m.first ? q1arr[base + i - 1] : 0.0;
        const double mprev = i > m.first ? mu1arr[base + i - 1] : 0.0;
        const double q2 = __dsub_rn(1.0, q1);
        const bool skip = fmin(q1, q2) < kOtsuEps || fmax(q1, q2) > 1.0 - kOtsuEps;
        const double t = __dmul_rn(mprev, qprev);
        const double p = __dmul_rn((double)h[i], m.scale);
        const double expect = skip ? t : __ddiv_rn(__dadd_rn(t, __dmul_rn((double)i, p)), q1);
        if (__double_as_longlong(expect) != __double_as_longlong(mu1)) fail = true;
        if (skip) continue;
        const double mu2 = __ddiv_rn(__dsub_rn(m.mu, __dmul_rn(q1, mu1)), q2);
        const double d = __dsub_rn(mu1, mu2);
        const double sigma = __dmul_rn(__dmul_rn(__dmul_rn(q1, q2), d), d);
        if (sigma > best) {  // ascending i per thread: strict '>' keeps the first maximum
            best = sigma;
            best_i = i;
        }
    }
    if (fail) s_fail = 1;
    // (sigma, index) max-reduce; equal sigma -> smaller index
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ob > best || (ob == best && oi < best_i)) {
            best = ob;
            best_i = oi;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        s_best[warp] = best;
        s_idx[warp] = best_i;
    }
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        best = lane < nw ? s_best[lane] : 0.0;
        best_i = lane < nw ? s_idx[lane] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (ob > best || (ob == best && oi < best_i)) {
                best = ob;
                best_i = oi;
            }
        }
        if (lane == 0) {
            if (s_fail) {
                atomicAdd(redo_count, 1);
                best_i = otsu_scan_exact(h, bins);
            } else if (!(best > 0.0)) {
                best_i = 0;
            }
            out[blockIdx.x] = best_i;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Certified parallel Otsu.  The reference's recurrence (above) is sequential, but its RESULT is an
// argmax, and the value it maximises is known in closed form from exact integer prefix sums:
//     C_i = sum_{j<=i} h_j,  S'_i = sum_{f<=j<=i} j h_j      (f = first bin the recurrence evaluates; the
//     first moments of the bins it skips in front are dropped by its `mu1 *= q1; continue`)
//     sigma_i = q1 q2 (mu1 - mu2)^2,  q1 = C_i/N, q2 = 1 - q1, mu1 = S'_i/C_i, mu2 = (S_T - S'_i)/(N - C_i).
// A forward error analysis of the recurrence bounds how far its fp64 value can be from that:
//     q1^ = q1 (1 + th),  |th| <= (i - first + 3) u        (one rounding per addition; exact when N = 2^k)
//     (mu1 q1)^ = (S'_i/N)(1 + th), |th| <= (3 (i - f) + 5) u   (three roundings per bin, all terms >= 0)
// and the remaining operations (mu2's cancellation, q2, the difference, the products) propagate as
// interval bounds e_num, e_q2, e_mu2, e_d below (u = 2^-53, every bound inflated by 1 % and a few u to
// cover second-order terms and the rounding of the bound's own evaluation).  Each bin gets
// sigma_lo <= sigma^ <= sigma_up.  If one bin's sigma_lo exceeds every other bin's sigma_up, that bin IS
// the reference's threshold, whatever the roundings were: the frame is "certified" and the sequential
// scan is skipped.  Otherwise (ties between empty bins, flat maxima, > 2^53 moments, a skip decision
// inside its error band) the exact chain kernels run, and only up to the last bin that could still
// win (kmax).  Either way the threshold is the reference's; tests/test_gpu_ops.py sweeps both outcomes
// and tests/test_otsu_certify.py holds the NumPy model of this kernel against the sequential scan.
// One 8-CTA cluster per frame: 8 bins per thread, prefix sums and reductions exchanged through
// distributed shared memory.
constexpr int kCertCluster = 8, kCertThreads = 1024, kCertPer = 8;
static_assert(kCertCluster * kCertThreads * kCertPer == kBins16, "one thread per 8 bins");
constexpr double kU = 1.1102230246251565e-16;  // 2^-53
constexpr double kInfl = 1.01;

struct CertShared {
    unsigned long long cnt, mom;   // totals of this CTA's bins
    int first, last;               // occupied bins of this CTA (INT_MAX / -1 when none)
    int f, amb;                    // first bin certainly evaluated, first ambiguous bin (INT_MAX when none)
    unsigned long long s_excl_f;   // first moment of all bins before f (valid in the CTA that owns f)
    double best_lo;
    int best_i;
    int ncand, kmax;
    double star[7];                // q1, q2, mu1, mu2, T, rq, rT of the winning bin (valid in the CTA that owns it)
};

// closed-form ("star") quantities of one bin from the exact prefix sums, and the accumulated-error radii
struct BinStar {
    double q1, q2, mu1, mu2, T, rq, rT;
};
// (x / N is computed as x * (1 / N): exact when N is a power of two, else 1.5 u instead of 0.5 u, inside the slack)
__device__ __forceinline__ BinStar bin_star(unsigned long long C, unsigned long long Sp, unsigned long long N,
                                            unsigned long long ST, double invN, int i, int first, int f, bool pow2) {
    BinStar b;
    const unsigned long long rest = ST - Sp;
    b.q1 = (double)C * invN;
    b.q2 = (double)(N - C) * invN;
    b.rq = pow2 ? 0.0 : ((double)max(i - first + 1, 0) + 2.0) * kU * kInfl;
    b.rT = (3.0 * (double)max(i - f, 0) + 5.0) * kU * kInfl;
    b.T = (double)Sp * invN;
    b.mu1 = (double)Sp / (double)C;
    b.mu2 = (double)rest / (double)(N - C);
    return b;
}

// Second chance for a bin k whose sigma interval overlaps the winner's (bin i): the accumulated errors of
// the two chains at k and at i are the SAME numbers up to the |k - i| steps between them, and
// F = A^2 / (q (1 - q)), A = T - M q, reacts to them almost identically at neighbouring bins.  With
// q^_k = q_k + Eq + lq, T^_k = T_k + Et + lt (Eq, Et: the errors at i; lq, lt: the few roundings between
// the bins), ln F_i - ln F_k moves with (Eq, Et, the rounding of mu) only through the DIFFERENCE of the
// log-derivatives at the two bins.  True: k provably loses.  (tests/otsu_certify_model.py
// differential_excludes is the same arithmetic.)
__device__ __forceinline__ bool cert_differential_excludes(const BinStar& bi, const BinStar& bk, int dist, double M, bool pow2) {
    const double D = (double)dist;
    const double di = bi.mu1 - bi.mu2, dk = bk.mu1 - bk.mu2;
    const double Ai = bi.q1 * bi.q2 * di, Ak = bk.q1 * bk.q2 * dk;
    const double si = bi.q1 * bi.q2 * di * di, sk = bk.q1 * bk.q2 * dk * dk;
    const double dT = 2.0 * fabs(1.0 / Ai - 1.0 / Ak);
    const double dq = fabs(-2.0 * M * (1.0 / Ai - 1.0 / Ak) - (1.0 / bi.q1 - 1.0 / bk.q1) + (1.0 / bi.q2 - 1.0 / bk.q2));
    const double dM = 2.0 * fabs(bi.q1 / Ai - bk.q1 / Ak);
    const double common = 1.5 * (dT * bi.T * bi.rT + dq * bi.q1 * bi.rq + dM * M * 3.0 * kU);
    const double lq = pow2 ? 0.0 : bk.q1 * (D + 3.0) * kU * kInfl;
    const double lt = bk.T * (3.0 * D + 5.0) * kU * kInfl;
    const double local = 1.5 * ((2.0 / fabs(Ak)) * lt + (fabs(2.0 * M / Ak) + 1.0 / bk.q1 + 1.0 / bk.q2) * lq);
    auto eps = [&](const BinStar& b, double d) {
        const double kappa = (fabs(b.T / (M - b.T)) + 4.0) * kU;
        return 7.0 * kU + 2.0 * fabs(b.mu2 / d) * kappa;
    };
    auto star_err = [&](const BinStar& b, double d) { return 16.0 * kU * (1.0 + (fabs(b.mu1) + fabs(b.mu2)) / fabs(d)); };
    const double gap = (si - sk) / si;
    const double need = (common + local + eps(bi, di) + eps(bk, dk) + star_err(bi, di) + star_err(bk, dk)) * kInfl + 8.0 * kU;
    return gap > need && need < INFINITY;   // NaN compares false
}

template <typename T, typename Op>
__device__ __forceinline__ T cert_block_reduce(T v, T* s_tmp, Op op, T identity) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();               // s_tmp may still be read from the previous reduction
    if (lane == 0) s_tmp[warp] = v;
    __syncthreads();
    v = lane < (kCertThreads >> 5) ? s_tmp[lane] : identity;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;                      // every thread holds the block result
}

__global__ void __cluster_dims__(kCertCluster, 1, 1) __launch_bounds__(kCertThreads, 1)
otsu_certify_kernel(const unsigned long long* __restrict__ hist, OtsuMeta* __restrict__ meta, int32_t* __restrict__ out,
                    int32_t* __restrict__ certified_out, int force_chain) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ CertShared sh;
    __shared__ unsigned long long s_w0[32], s_w1[32];
    __shared__ double s_d[32];
    __shared__ int s_i[32];
    const int frame = blockIdx.y, rank = (int)cluster.block_rank(), tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const unsigned long long* h = hist + (int64_t)frame * kBins16;
    const int bin0 = (rank * kCertThreads + tid) * kCertPer;

    // ---- exact prefix sums of counts and first moments
    unsigned long long c[kCertPer];
    {
        const ulonglong2* hv = reinterpret_cast<const ulonglong2*>(h + bin0);
#pragma unroll
        for (int k = 0; k < kCertPer / 2; k++) {
            const ulonglong2 v = hv[k];
            c[2 * k] = v.x;
            c[2 * k + 1] = v.y;
        }
    }
    unsigned long long tc = 0, tm = 0;
    int nzf = 0x7fffffff, nzl = -1;
#pragma unroll
    for (int k = 0; k < kCertPer; k++) {
        tc += c[k];
        tm += c[k] * (unsigned long long)(bin0 + k);
        if (c[k]) {
            nzf = min(nzf, bin0 + k);
            nzl = bin0 + k;
        }
    }
    unsigned long long ic = tc, im = tm;  // inclusive scan over the CTA's threads
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long a = __shfl_up_sync(0xffffffffu, ic, o), b = __shfl_up_sync(0xffffffffu, im, o);
        if (lane >= o) {
            ic += a;
            im += b;
        }
    }
    if (lane == 31) {
        s_w0[warp] = ic;
        s_w1[warp] = im;
    }
    __syncthreads();
    if (warp == 0) {
        unsigned long long a = s_w0[lane], b = s_w1[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long x = __shfl_up_sync(0xffffffffu, a, o), y = __shfl_up_sync(0xffffffffu, b, o);
            if (lane >= o) {
                a += x;
                b += y;
            }
        }
        s_w0[lane] = a;
        s_w1[lane] = b;
    }
    __syncthreads();
    unsigned long long ec = ic - tc + (warp ? s_w0[warp - 1] : 0ull);  // exclusive, within the CTA
    unsigned long long em = im - tm + (warp ? s_w1[warp - 1] : 0ull);
    const int cta_first = cert_block_reduce<int>(nzf, s_i, [](int a, int b) { return min(a, b); }, 0x7fffffff);
    const int cta_last = cert_block_reduce<int>(nzl, s_i, [](int a, int b) { return max(a, b); }, -1);
    if (tid == 0) {
        sh.cnt = s_w0[31];
        sh.mom = s_w1[31];
        sh.first = cta_first;
        sh.last = cta_last;
    }
    cluster.sync();                                                     // (1) CTA totals visible
    unsigned long long N = 0, ST = 0;
    int first = 0x7fffffff, last = -1;
#pragma unroll
    for (int r = 0; r < kCertCluster; r++) {
        const CertShared* p = cluster.map_shared_rank(&sh, r);
        const unsigned long long rc = p->cnt, rm = p->mom;
        if (r < rank) {
            ec += rc;
            em += rm;
        }
        N += rc;
        ST += rm;
        first = min(first, p->first);
        last = max(last, p->last);
    }
    OtsuMeta m;
    m.first = last < 0 ? -1 : first;
    m.last = last;
    m.scale = last < 0 ? 0.0 : __ddiv_rn(1.0, (double)N);
    m.mu = __dmul_rn((double)ST, m.scale);
    m.certified = 0;
    m.pad = 0;
    const bool leader = rank == 0 && tid == 0;
    if (last < 0 || ST >= (1ull << 53) || force_chain) {   // empty frame: t = 0;  moments beyond 2^53: exact chain
        if (leader) {
            m.certified = last < 0 ? 1 : 0;
            meta[frame] = m;
            if (last < 0) out[frame] = 0;
            if (certified_out) certified_out[frame] = m.certified;
        }
        cluster.sync();
        return;
    }

    // ---- which bins does the recurrence evaluate?  skip <=> q1^ < eps or q1^ > 1 - eps
    const bool pow2 = (N & (N - 1)) == 0;
    const double Nf = (double)N, invN = 1.0 / Nf, slop = 4.0 * kU;
    unsigned nonc_mask = 0, amb_mask = 0;
    {
        unsigned long long C = ec;
#pragma unroll
        for (int k = 0; k < kCertPer; k++) {
            C += c[k];
            const int i = bin0 + k;
            const double q1 = (double)C * invN;
            const double rq = pow2 ? 0.0 : ((double)max(i - first + 1, 0) + 2.0) * kU * kInfl;
            bool sf, st, nonc;
            if (pow2) {                          // the q1 chain is exact (multiples of 1/N): so are the decisions
                sf = q1 < kOtsuEps;
                st = q1 > 1.0 - kOtsuEps;
                nonc = !(sf || st);
            } else {
                sf = q1 * (1.0 + rq + slop) < kOtsuEps;
                st = q1 * (1.0 - rq - slop) > 1.0 - kOtsuEps;
                nonc = q1 * (1.0 - rq - slop) >= kOtsuEps * (1.0 + slop) && q1 * (1.0 + rq + slop) <= (1.0 - kOtsuEps) * (1.0 - slop);
            }
            const bool live = i >= first;
            if (nonc && live) nonc_mask |= 1u << k;
            if (!(sf || st || nonc) && live) amb_mask |= 1u << k;
        }
    }
    const int my_f = nonc_mask ? bin0 + __ffs(nonc_mask) - 1 : 0x7fffffff;
    const int my_amb = amb_mask ? bin0 + __ffs(amb_mask) - 1 : 0x7fffffff;
    const int cta_f = cert_block_reduce<int>(my_f, s_i, [](int a, int b) { return min(a, b); }, 0x7fffffff);
    const int cta_amb = cert_block_reduce<int>(my_amb, s_i, [](int a, int b) { return min(a, b); }, 0x7fffffff);
    if (my_f == cta_f && my_f != 0x7fffffff) {   // the thread that owns the CTA's first evaluated bin
        unsigned long long sx = em;
        for (int k = 0; k < my_f - bin0; k++) sx += c[k] * (unsigned long long)(bin0 + k);
        sh.s_excl_f = sx;
    }
    if (tid == 0) {
        sh.f = cta_f;
        sh.amb = cta_amb;
    }
    cluster.sync();                                                     // (2) f, ambiguity, S before f
    int f = 0x7fffffff, amb0 = 0x7fffffff;
    unsigned long long S0 = 0;
#pragma unroll
    for (int r = 0; r < kCertCluster; r++) {
        const CertShared* p = cluster.map_shared_rank(&sh, r);
        const int rf = p->f;
        if (rf < f) {
            f = rf;
            S0 = p->s_excl_f;
        }
        amb0 = min(amb0, p->amb);
    }
    if (f == 0x7fffffff || amb0 < f) {
        // nothing certainly evaluated: the reference returns 0 unless an ambiguous bin might count;
        // an ambiguous bin in front of f leaves f itself unknown -> exact chain
        if (leader) {
            m.certified = (f == 0x7fffffff && amb0 == 0x7fffffff) ? 1 : 0;
            meta[frame] = m;
            if (m.certified) out[frame] = 0;
            if (certified_out) certified_out[frame] = m.certified;
        }
        cluster.sync();
        return;
    }

    // ---- interval for the recurrence's sigma at every bin it may evaluate
    double up[kCertPer];
    double my_lo = 0.0;
    int my_i = 0x7fffffff;
    {
        const double mu_star = (double)ST * invN;
        unsigned long long C = ec, S = em;
#pragma unroll
        for (int k = 0; k < kCertPer; k++) {
            C += c[k];
            S += c[k] * (unsigned long long)(bin0 + k);
            const int i = bin0 + k;
            const bool nonc = (nonc_mask >> k) & 1u, cand = (nonc || ((amb_mask >> k) & 1u)) && i >= f && i <= last;
            up[k] = -1.0;
            if (!cand) continue;
            const BinStar b = bin_star(C, S - S0, N, ST, invN, i, first, f, pow2);
            const double q1 = b.q1, q2 = b.q2, rq = b.rq, rT = b.rT, B = b.T, mu1 = b.mu1, mu2 = b.mu2;
            const double num = (double)(ST - (S - S0)) * invN;
            const double e_m1 = mu1 * ((rT + rq) * kInfl + 2.0 * kU);
            const double e_num = (mu_star * 3.0 * kU + B * (rT + kU) + kU * (num + mu_star)) * kInfl + 4.0 * kU * mu_star;
            const double e_q2 = (q1 * rq + kU * q2) * kInfl + 2.0 * kU * q2;
            const double den = q2 - e_q2;
            const double e_mu2 = den > 0.0 ? (num + e_num) / den * (1.0 + 4.0 * kU) - mu2 * (1.0 - 4.0 * kU) : INFINITY;
            const double d = fabs(mu1 - mu2);
            const double e_d = (e_m1 + e_mu2) * (1.0 + 4.0 * kU) + 4.0 * kU * (fabs(mu1) + fabs(mu2));
            double u_ = q1 * (1.0 + rq) * (q2 + e_q2) * (d + e_d) * (d + e_d) * (1.0 + 16.0 * kU);
            double l_ = q1 * (1.0 - rq) * fmax(q2 - e_q2, 0.0) * fmax(d - e_d, 0.0) * fmax(d - e_d, 0.0) * (1.0 - 16.0 * kU);
            if (u_ != u_) u_ = INFINITY;
            if (l_ != l_ || !nonc) l_ = 0.0;
            up[k] = u_;
            if (l_ > my_lo) {            // ascending i: strict '>' keeps the first maximum
                my_lo = l_;
                my_i = i;
            }
        }
    }
    // (lo, index) arg-max: larger lo wins, equal lo -> smaller index
    {
        double b = my_lo;
        int bi = my_i;
        auto better = [](double ob, int oi, double b, int bi) { return ob > b || (ob == b && oi < bi); };
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, b, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (better(ob, oi, b, bi)) {
                b = ob;
                bi = oi;
            }
        }
        __syncthreads();
        if (lane == 0) {
            s_d[warp] = b;
            s_i[warp] = bi;
        }
        __syncthreads();
        if (warp == 0) {
            b = s_d[lane];
            bi = s_i[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, b, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (better(ob, oi, b, bi)) {
                    b = ob;
                    bi = oi;
                }
            }
            if (lane == 0) {
                sh.best_lo = b;
                sh.best_i = bi;
            }
        }
    }
    cluster.sync();                                                     // (3) per-CTA best lower bounds
    double Lmax = 0.0;
    int istar = 0x7fffffff;
#pragma unroll
    for (int r = 0; r < kCertCluster; r++) {
        const CertShared* p = cluster.map_shared_rank(&sh, r);
        const double ob = p->best_lo;
        const int oi = p->best_i;
        if (ob > Lmax || (ob == Lmax && oi < istar)) {
            Lmax = ob;
            istar = oi;
        }
    }
    // the owner of the winning bin publishes its closed-form values for the differential test
    if (Lmax > 0.0 && istar >= bin0 && istar < bin0 + kCertPer) {
        unsigned long long C = ec, S = em;
        for (int k = 0; k <= istar - bin0; k++) {
            C += c[k];
            S += c[k] * (unsigned long long)(bin0 + k);
        }
        const BinStar b = bin_star(C, S - S0, N, ST, invN, istar, first, f, pow2);
        sh.star[0] = b.q1; sh.star[1] = b.q2; sh.star[2] = b.mu1; sh.star[3] = b.mu2;
        sh.star[4] = b.T; sh.star[5] = b.rq; sh.star[6] = b.rT;
    }
    cluster.sync();                                                     // (3b) winner's values visible
    int ncand = 0, kmax = -1;
    if (Lmax > 0.0) {
        bool any = false;
#pragma unroll
        for (int k = 0; k < kCertPer; k++) any |= up[k] >= Lmax && bin0 + k != istar;
        if (any) {   // rare: a sigma interval of this thread overlaps the winner's
            const double* ps = cluster.map_shared_rank(&sh, istar / (kCertThreads * kCertPer))->star;
            BinStar bi;
            bi.q1 = ps[0]; bi.q2 = ps[1]; bi.mu1 = ps[2]; bi.mu2 = ps[3]; bi.T = ps[4]; bi.rq = ps[5]; bi.rT = ps[6];
            const double M = (double)ST * invN;
            unsigned long long C = ec, S = em;
            for (int k = 0; k < kCertPer; k++) {
                C += c[k];
                S += c[k] * (unsigned long long)(bin0 + k);
                const int i = bin0 + k;
                if (!(up[k] >= Lmax) || i == istar) continue;
                bool excluded = false;
                if ((nonc_mask >> k) & 1u) {
                    const BinStar bk = bin_star(C, S - S0, N, ST, invN, i, first, f, pow2);
                    excluded = cert_differential_excludes(bi, bk, abs(i - istar), M, pow2);
                }
                if (!excluded) {
                    ncand++;
                    kmax = i;
                }
            }
        }
    }
    const int cta_n = cert_block_reduce<int>(ncand, s_i, [](int a, int b) { return a + b; }, 0);
    const int cta_k = cert_block_reduce<int>(kmax, s_i, [](int a, int b) { return max(a, b); }, -1);
    if (tid == 0) {
        sh.ncand = cta_n;
        sh.kmax = cta_k;
    }
    cluster.sync();                                                     // (4) competitor counts
    if (leader) {
        int n = 0, km = -1;
        for (int r = 0; r < kCertCluster; r++) {
            const CertShared* p = cluster.map_shared_rank(&sh, r);
            n += p->ncand;
            km = max(km, p->kmax);
        }
        if (Lmax > 0.0 && n == 0) {
            m.certified = 1;
            out[frame] = istar;
        } else if (Lmax > 0.0) {
            m.last = max(km, istar);     // no bin beyond the last competitor can win: the chain stops there
        }
        meta[frame] = m;
        if (certified_out) certified_out[frame] = m.certified;
    }
    cluster.sync();                                                     // (5) peers' shared memory read
}

// ---------------------------------------------------------------------------------------------
// equalizeHist (u8): LUT from the global histogram, one block per frame, then gather
__global__ void __launch_bounds__(256) equalize_lut_kernel(const unsigned long long* __restrict__ hist,
                                                           int64_t total, uint8_t* __restrict__ lut) {
    __shared__ unsigned long long s_h[256];
    __shared__ unsigned long long s_cum[256];
    const unsigned long long* h = hist + (int64_t)blockIdx.x * 256;
    uint8_t* out = lut + (int64_t)blockIdx.x * 256;
    s_h[threadIdx.x] = h[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long c = 0;
        for (int i = 0; i < 256; i++) {
            c += s_h[i];
            s_cum[i] = c;
        }
    }
    __syncthreads();
    // first non-empty bin
    __shared__ int s_i0;
    if (threadIdx.x == 0) {
        int i0 = 0;
        while (i0 < 255 && s_h[i0] == 0) i0++;
        s_i0 = i0;
    }
    __syncthreads();
    const int i0 = s_i0;
    const int b = threadIdx.x;
    if (s_h[i0] == (unsigned long long)total) {
        out[b] = (uint8_t)i0;  // constant image: every pixel maps to itself (only bin i0 occurs)
        return;
    }
    const float scale = __fdiv_rn(255.0f, (float)(total - (int64_t)s_h[i0]));
    if (b <= i0) {
        out[b] = 0;
    } else {
        const float s = (float)(long long)(s_cum[b] - s_h[i0]);
        out[b] = (uint8_t)yam_rint_sat(__fmul_rn(s, scale), 255);
    }
}

__global__ void __launch_bounds__(256) gather_lut8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                          int64_t frame_px, const uint8_t* __restrict__ luts) {
    __shared__ uint8_t s_lut[256];
    s_lut[threadIdx.x] = luts[(int64_t)blockIdx.z * 256 + threadIdx.x];
    __syncthreads();
    src += (int64_t)blockIdx.z * frame_px;
    dst += (int64_t)blockIdx.z * frame_px;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
    int64_t done = 0;
    if (aligned) {
        const int64_t groups = frame_px / 16;
        for (int64_t g = tid; g < groups; g += stride) {
            uint4 q = yam_ld_stream(reinterpret_cast<const uint4*>(src) + g);
            uint8_t* e = reinterpret_cast<uint8_t*>(&q);
#pragma unroll
            for (int i = 0; i < 16; i++) e[i] = s_lut[e[i]];
            yam_st_stream(reinterpret_cast<uint4*>(dst) + g, q);
        }
        done = groups * 16;
    }
    for (int64_t i = done + tid; i < frame_px; i += stride) dst[i] = s_lut[src[i]];
}

// ---------------------------------------------------------------------------------------------
// CLAHE
struct ClaheGeom {
    int h, w;            // image
    int tiles_x, tiles_y;
    int tw, th;          // tile size (of the padded image)
    int clip;            // 0 = no clipping
    float lut_scale;     // (histSize-1)/area
    float inv_tw, inv_th;
    int y_off;           // global row index of local row 0 (row-strip sharding; 0 for whole images)
};

// counts of bin pair `wi` (bins 2*wi, 2*wi+1) from the packed shared histogram (+ spilled part)
struct SmemCounts {
    const uint32_t* sh;
    const uint32_t* ovf;
    bool use_ovf;
    __device__ __forceinline__ void get32(int wi, uint32_t& c0, uint32_t& c1) const {
        const uint32_t wv = sh[wi];
        c0 = wv & 0xffffu;
        c1 = wv >> 16;
        if (use_ovf) {
            c0 += __ldcg(&ovf[2 * wi]);
            c1 += __ldcg(&ovf[2 * wi + 1]);
        }
    }
};
// counts from a global 32-bit histogram (multi-CTA tiles)
struct GlobalCounts {
    const uint32_t* gh;
    __device__ __forceinline__ void get32(int wi, uint32_t& c0, uint32_t& c1) const {
        const uint2 v = __ldcg(reinterpret_cast<const uint2*>(gh) + wi);
        c0 = v.x;
        c1 = v.y;
    }
};

// clip -> redistribute -> ordered prefix sum -> LUT for one tile; all kHistThreads threads take part.
// 32-bit arithmetic throughout (a tile has < 2^31 pixels, host-checked), two passes over the bins:
//   pass 0  per-warp sums of the clipped counts -> excess, batch, residual, step (cv2's redistribution)
//   pass 1  ordered prefix sum of the adjusted counts -> LUT.  The warp bases come from pass 0 in
//           closed form: adjusted total = clipped total + batch * bins + (residual increments that
//           fall into the warp's bin range), so no third pass is needed.
// "bin is one of the first `residual` multiples of `step`" is tested without a division:
// q = umulhi(bin, ceil(2^32 / step)) == bin / step exactly for bin, step <= 65536.
template <class Counts>
__device__ __forceinline__ void clahe_lut_passes(const Counts& cnt, const ClaheGeom& g, uint16_t* __restrict__ lut,
                                                 unsigned long long* s_red, unsigned long long* s_base) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t area = (uint32_t)((long long)g.tw * g.th);
    const int wbase = warp * 1024;  // each warp owns words [1024*warp, +1024); lane reads base+32*i+lane
    const uint32_t clipv = g.clip > 0 ? (uint32_t)g.clip : 0xffffffffu;
    uint32_t* s_red32 = reinterpret_cast<uint32_t*>(s_red);
    uint32_t* s_base32 = reinterpret_cast<uint32_t*>(s_base);
    // pass 0: sum of clipped counts -> excess
    uint32_t part = 0;
#pragma unroll 4
    for (int i = 0; i < 32; i++) {
        uint32_t c0, c1;
        cnt.get32(wbase + 32 * i + lane, c0, c1);
        part += min(c0, clipv) + min(c1, clipv);
    }
    part = yam_warp_sum(part);
    if (lane == 0) s_red32[warp] = part;
    __syncthreads();
    uint32_t clipped_total = 0, clipped_before = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) {
        const uint32_t v = s_red32[i];
        clipped_total += v;
        if (i < warp) clipped_before += v;
    }
    __syncthreads();
    const uint32_t excess = area - clipped_total;
    const uint32_t batch = g.clip > 0 ? excess / kBins16 : 0;
    const uint32_t residual = g.clip > 0 ? excess - batch * kBins16 : 0;
    const uint32_t step = residual ? max((uint32_t)kBins16 / residual, 1u) : 1u;
    // step == 1: magic would be 2^32; q == bin then (handled explicitly)
    const uint32_t magic = step > 1 ? (uint32_t)((0x100000000ull + step - 1) / step) : 0u;
    auto bumped = [&](uint32_t bin) -> uint32_t {   // 1 if cv2 increments this bin with the residual
        if (!residual) return 0u;
        const uint32_t q = step > 1 ? __umulhi(bin, magic) : bin;
        return (q * step == bin && q < residual) ? 1u : 0u;
    };
    // residual increments in bins [0, first_bin): multiples of step below first_bin, at most `residual`
    const uint32_t first_bin = (uint32_t)warp * 2048u;
    uint32_t bumps_before = 0;
    if (residual && first_bin) bumps_before = min((first_bin - 1u) / step + 1u, residual);
    uint32_t run = clipped_before + batch * first_bin + bumps_before;

    // pass 1: ordered prefix sum -> LUT
#pragma unroll 2
    for (int i = 0; i < 32; i++) {
        const int wi = wbase + 32 * i + lane;
        uint32_t c0, c1;
        cnt.get32(wi, c0, c1);
        const uint32_t a0 = min(c0, clipv) + batch + bumped(2u * wi);
        const uint32_t a1 = min(c1, clipv) + batch + bumped(2u * wi + 1u);
        uint32_t incl = a0 + a1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        const uint32_t cum0 = run + incl - a1, cum1 = run + incl;
        // cv2: saturate_cast<ushort>(sum * lutScale) with sum an int converted to float
        const uint32_t l0 = (uint32_t)yam_rint_sat(__fmul_rn((float)(int)cum0, g.lut_scale), 65535);
        const uint32_t l1 = (uint32_t)yam_rint_sat(__fmul_rn((float)(int)cum1, g.lut_scale), 65535);
        reinterpret_cast<uint32_t*>(lut)[wi] = l0 | (l1 << 16);
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    (void)s_base32;
}

// u16: one CTA per tile at a time (persistent CTAs loop over the tiles of a frame chunk):
// histogram -> clip -> redistribute -> scan -> LUT.  grid (min(tiles_total, SMs)); each CTA owns
// one 256 KiB overflow slot that is all-zero between tiles.
__global__ void __launch_bounds__(kHistThreads, 1) clahe_lut16_kernel(const uint16_t* __restrict__ src_all,
                                                                      ClaheGeom g, int frames,
                                                                      uint32_t* __restrict__ overflow,
                                                                      uint16_t* __restrict__ luts) {
    extern __shared__ __align__(16) uint32_t sh[];
    __shared__ unsigned long long s_red[32];
    __shared__ unsigned long long s_base[32];
    __shared__ int s_spilled;
    const int tiles_per_frame = g.tiles_x * g.tiles_y;
    const int total_tiles = tiles_per_frame * frames;
    uint32_t* ovf = overflow + (int64_t)blockIdx.x * kBins16;

    for (int slot = blockIdx.x; slot < total_tiles; slot += gridDim.x) {
        const int frame = slot / tiles_per_frame, tile = slot - frame * tiles_per_frame;
        const int tyi = tile / g.tiles_x, txi = tile - tyi * g.tiles_x;
        const uint16_t* src = src_all + (int64_t)frame * g.h * g.w;
        uint16_t* lut = luts + (int64_t)slot * kBins16;

        for (int i = threadIdx.x; i < kWords16; i += kHistThreads) sh[i] = 0;
        if (threadIdx.x == 0) s_spilled = 0;
        __syncthreads();
        accumulate16<uint32_t, true>(sh, ovf, &s_spilled, src, g.h, g.w, txi * g.tw, (txi + 1) * g.tw, tyi * g.th,
                               (tyi + 1) * g.th);
        __threadfence();
        __syncthreads();
        const bool use_ovf = s_spilled != 0;

        {
            SmemCounts cnt{sh, ovf, use_ovf};
            clahe_lut_passes(cnt, g, lut, s_red, s_base);
        }
        __syncthreads();
        // leave the overflow slot zero for the next tile / call
        if (use_ovf) {
            for (int i = threadIdx.x; i < kBins16; i += kHistThreads) ovf[i] = 0;
            __threadfence();
            __syncthreads();
        }
    }
}

// Big tiles (mosaic strips: 8 tiles of 67 Mpx) would leave most SMs idle with one CTA per tile:
// `parts` CTAs per tile accumulate row slabs into packed shared counters and flush the non-zero
// bins into the tile's 32-bit global histogram (zeroed by the caller); a second kernel builds the LUT.
__global__ void __launch_bounds__(kHistThreads, 1) clahe_hist16_parts_kernel(const uint16_t* __restrict__ src_all,
                                                                             ClaheGeom g,
                                                                             uint32_t* __restrict__ ghist,
                                                                             uint32_t* __restrict__ slabs) {
    extern __shared__ __align__(16) uint32_t sh[];
    const int tiles_per_frame = g.tiles_x * g.tiles_y;
    const int slot = blockIdx.y;
    const int frame = slot / tiles_per_frame, tile = slot - frame * tiles_per_frame;
    const int tyi = tile / g.tiles_x, txi = tile - tyi * g.tiles_x;
    const uint16_t* src = src_all + (int64_t)frame * g.h * g.w;
    uint32_t* gh = ghist + (int64_t)slot * kBins16;
    for (int i = threadIdx.x; i < kWords16; i += kHistThreads) sh[i] = 0;
    __syncthreads();
    const int parts = gridDim.x;
    const int rows_per = (g.th + parts - 1) / parts;
    const int r0 = tyi * g.th + blockIdx.x * rows_per;
    const int r1 = min((tyi + 1) * g.th, r0 + rows_per);
    accumulate16<uint32_t>(sh, gh, nullptr, src, g.h, g.w, txi * g.tw, (txi + 1) * g.tw, r0, r1);
    __syncthreads();
    store_packed16(sh, slabs + ((int64_t)slot * parts + blockIdx.x) * kWords16);   // hist16_reduce_kernel adds the slabs
}

__global__ void __launch_bounds__(kHistThreads, 1) clahe_lut_from_hist_kernel(const uint32_t* __restrict__ ghist,
                                                                              ClaheGeom g,
                                                                              uint16_t* __restrict__ luts) {
    __shared__ unsigned long long s_red[32];
    __shared__ unsigned long long s_base[32];
    const int slot = blockIdx.x;
    GlobalCounts cnt{ghist + (int64_t)slot * kBins16};
    clahe_lut_passes(cnt, g, luts + (int64_t)slot * kBins16, s_red, s_base);
}

// u8: one CTA (256 threads) per tile
__global__ void __launch_bounds__(256) clahe_lut8_kernel(const uint8_t* __restrict__ src, ClaheGeom g,
                                                         uint8_t* __restrict__ luts) {
    __shared__ uint32_t sh[8][256];
    __shared__ unsigned long long s_cnt[256];
    __shared__ unsigned long long s_cum[256];
    __shared__ unsigned long long s_excess;
    const int tile = blockIdx.x;
    const int tyi = tile / g.tiles_x, txi = tile - tyi * g.tiles_x;
    src += (int64_t)blockIdx.z * g.h * g.w;
    uint8_t* lut = luts + ((int64_t)blockIdx.z * g.tiles_x * g.tiles_y + tile) * 256;
    for (int i = threadIdx.x; i < 8 * 256; i += 256) (&sh[0][0])[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* mine = sh[warp];
    const int c0 = txi * g.tw, r0 = tyi * g.th;
    for (int r = r0 + warp; r < r0 + g.th; r += 8) {
        const int gy = yam_border(r, g.h, YAM_BORDER_REFLECT101);
        const uint8_t* row = src + (int64_t)gy * g.w;
        for (int c = lane; c < g.tw; c += 32)
            atomicAdd(&mine[row[yam_border(c0 + c, g.w, YAM_BORDER_REFLECT101)]], 1u);
    }
    __syncthreads();
    const int b = threadIdx.x;
    unsigned long long cnt = 0;
#pragma unroll
    for (int wv = 0; wv < 8; wv++) cnt += sh[wv][b];
    const unsigned long long clipv = g.clip > 0 ? (unsigned long long)g.clip : ~0ull;
    s_cnt[b] = cnt < clipv ? cnt : clipv;
    __syncthreads();
    if (b == 0) {
        unsigned long long s = 0;
        for (int i = 0; i < 256; i++) s += s_cnt[i];
        s_excess = (unsigned long long)g.tw * g.th - s;
    }
    __syncthreads();
    if (g.clip > 0) {
        const unsigned long long excess = s_excess;
        const unsigned long long batch = excess / 256;
        const uint32_t residual = (uint32_t)(excess - batch * 256);
        const uint32_t step = residual ? max(256u / residual, 1u) : 1u;
        unsigned long long a = s_cnt[b] + batch;
        if (residual && (b % step) == 0 && (b / step) < residual) a += 1;
        s_cnt[b] = a;
    }
    __syncthreads();
    if (b == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < 256; i++) {
            run += s_cnt[i];
            s_cum[i] = run;
        }
    }
    __syncthreads();
    lut[b] = (uint8_t)yam_rint_sat(__fmul_rn((float)(long long)s_cum[b], g.lut_scale), 255);
}

// bilinear blend of the four neighbouring tile LUTs, cv2 order, no FMA contraction
template <typename T>
__global__ void __launch_bounds__(256) clahe_apply_kernel(const T* __restrict__ src, T* __restrict__ dst,
                                                          ClaheGeom g, const T* __restrict__ luts) {
    constexpr int VEC = 16 / sizeof(T);
    constexpr int BINS = sizeof(T) == 1 ? 256 : 65536;
    constexpr int HI = BINS - 1;
    const int64_t frame_px = (int64_t)g.h * g.w;
    src += (int64_t)blockIdx.z * frame_px;
    dst += (int64_t)blockIdx.z * frame_px;
    luts += (int64_t)blockIdx.z * g.tiles_x * g.tiles_y * BINS;
    const int y = blockIdx.x;
    const float tyf = __fsub_rn(__fmul_rn((float)(y + g.y_off), g.inv_th), 0.5f);
    int ty1 = (int)floorf(tyf);
    const float ya = __fsub_rn(tyf, (float)ty1), ya1 = __fsub_rn(1.0f, ya);
    const int ty2 = min(ty1 + 1, g.tiles_y - 1);
    ty1 = max(ty1, 0);
    const T* lrow1 = luts + (int64_t)ty1 * g.tiles_x * BINS;
    const T* lrow2 = luts + (int64_t)ty2 * g.tiles_x * BINS;
    const T* srow = src + (int64_t)y * g.w;
    T* drow = dst + (int64_t)y * g.w;
    const bool aligned = ((g.w % VEC) == 0) &&
                         (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0);
    const int groups = (g.w + VEC - 1) / VEC;
    for (int gi = threadIdx.x; gi < groups; gi += blockDim.x) {
        const int x0 = gi * VEC;
        T in[VEC], out[VEC];
        if (aligned) {
            const uint4 q = yam_ld_stream(reinterpret_cast<const uint4*>(srow + x0));
            *reinterpret_cast<uint4*>(in) = q;
        } else {
#pragma unroll
            for (int i = 0; i < VEC; i++) in[i] = (x0 + i < g.w) ? srow[x0 + i] : (T)0;
        }
#pragma unroll
        for (int i = 0; i < VEC; i++) {
            const int x = x0 + i;
            const float txf = __fsub_rn(__fmul_rn((float)x, g.inv_tw), 0.5f);
            int tx1 = (int)floorf(txf);
            const float xa = __fsub_rn(txf, (float)tx1), xa1 = __fsub_rn(1.0f, xa);
            const int tx2 = min(tx1 + 1, g.tiles_x - 1);
            tx1 = max(tx1, 0);
            const int v = in[i];
            const float l11 = (float)__ldg(lrow1 + (int64_t)tx1 * BINS + v);
            const float l12 = (float)__ldg(lrow1 + (int64_t)tx2 * BINS + v);
            const float l21 = (float)__ldg(lrow2 + (int64_t)tx1 * BINS + v);
            const float l22 = (float)__ldg(lrow2 + (int64_t)tx2 * BINS + v);
            const float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa));
            const float bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa));
            const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
            out[i] = (T)yam_rint_sat(res, HI);
        }
        if (aligned) {
            yam_st_stream(reinterpret_cast<uint4*>(drow + x0), *reinterpret_cast<const uint4*>(out));
        } else {
#pragma unroll
            for (int i = 0; i < VEC; i++)
                if (x0 + i < g.w) drow[x0 + i] = out[i];
        }
    }
}

// Otsu recurrence on host threads (the fp64 recurrence is sequential per frame).  One process-wide
// pool of workers, created on first use and never torn down (its threads sleep on a condition
// variable); callers hand in frames as their histograms arrive and wait for the batch at the end.
static std::atomic<int> g_host_threads{0};  // 0 = not configured: hardware_concurrency()

static int host_threads() {
    int t = g_host_threads.load();
    if (t <= 0) {
        const unsigned hw = std::thread::hardware_concurrency();
        t = hw == 0 ? 4 : (int)hw;
    }
    return t < 32 ? t : 32;
}

class ScanPool {
public:
    static ScanPool& instance() {
        static ScanPool* pool = new ScanPool();  // leaked on purpose: no static-destruction order issues
        return *pool;
    }
    void submit(std::function<void()> fn) {
        {
            std::lock_guard<std::mutex> lk(m_);
            queue_.push_back(std::move(fn));
            pending_++;
        }
        cv_.notify_one();
    }
    // the calling thread helps until every submitted task has finished
    void wait_all() {
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
            if (!queue_.empty()) {
                auto fn = std::move(queue_.front());
                queue_.pop_front();
                lk.unlock();
                fn();
                lk.lock();
                if (--pending_ == 0) done_.notify_all();
                continue;
            }
            if (pending_ == 0) return;
            done_.wait(lk);
        }
    }

private:
    ScanPool() {
        const int nt = host_threads() - 1;  // the caller is the last worker
        for (int i = 0; i < nt; i++) {
            try {
                std::thread([this] { run(); }).detach();
            } catch (...) {
                break;  // fewer helpers: wait_all() still drains the queue on the calling thread
            }
        }
    }
    void run() {
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
            cv_.wait(lk, [this] { return !queue_.empty(); });
            auto fn = std::move(queue_.front());
            queue_.pop_front();
            lk.unlock();
            fn();
            lk.lock();
            if (--pending_ == 0) done_.notify_all();
        }
    }
    std::mutex m_;
    std::condition_variable cv_, done_;
    std::deque<std::function<void()>> queue_;
    int64_t pending_ = 0;
};

// 64-bit device histograms -> 32-bit counts for the host scan (halves the read-back)
// thread = 8 consecutive values: four 16-byte LUT loads, four 16-byte stores (two interleaved entries each)
__global__ void __launch_bounds__(256) clahe_quad_build_kernel(const uint16_t* __restrict__ luts, int tiles_x, int tiles_y,
                                                              ushort4* __restrict__ quad) {
    const int cell = blockIdx.y;
    const int cy = cell / (tiles_x + 1), cx = cell - cy * (tiles_x + 1);
    const int ty1 = max(cy - 1, 0), ty2 = min(cy, tiles_y - 1);
    const int tx1 = max(cx - 1, 0), tx2 = min(cx, tiles_x - 1);
    const int v = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(luts + ((int64_t)ty1 * tiles_x + tx1) * kBins16 + v));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(luts + ((int64_t)ty1 * tiles_x + tx2) * kBins16 + v));
    const uint4 c = __ldg(reinterpret_cast<const uint4*>(luts + ((int64_t)ty2 * tiles_x + tx1) * kBins16 + v));
    const uint4 d = __ldg(reinterpret_cast<const uint4*>(luts + ((int64_t)ty2 * tiles_x + tx2) * kBins16 + v));
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
    const uint32_t cw[4] = {c.x, c.y, c.z, c.w}, dw[4] = {d.x, d.y, d.z, d.w};
    uint4* out = reinterpret_cast<uint4*>(quad + (int64_t)cell * kBins16 + v);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        // entries v + 2k and v + 2k + 1: (l11 | l12 << 16, l21 | l22 << 16) each
        uint4 o;
        o.x = __byte_perm(aw[k], bw[k], 0x5410);
        o.y = __byte_perm(cw[k], dw[k], 0x5410);
        o.z = __byte_perm(aw[k], bw[k], 0x7632);
        o.w = __byte_perm(cw[k], dw[k], 0x7632);
        out[k] = o;
    }
}

__global__ void __launch_bounds__(256) clahe_apply_quad_kernel(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst,
                                                              ClaheGeom g, const ushort4* __restrict__ quad) {
    const int y = blockIdx.x;
    const float tyf = __fsub_rn(__fmul_rn((float)(y + g.y_off), g.inv_th), 0.5f);
    const int fy = (int)floorf(tyf);
    const float ya = __fsub_rn(tyf, (float)fy), ya1 = __fsub_rn(1.0f, ya);
    const int cy = min(max(fy + 1, 0), g.tiles_y);
    const uint2* qrow = reinterpret_cast<const uint2*>(quad) + (int64_t)cy * (g.tiles_x + 1) * kBins16;
    const uint16_t* srow = src + (int64_t)y * g.w;
    uint16_t* drow = dst + (int64_t)y * g.w;
    const bool aligned = ((g.w % 8) == 0) &&
                         (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0);
    const int groups = (g.w + 7) / 8;
    for (int gi = threadIdx.x; gi < groups; gi += blockDim.x) {
        const int x0 = gi * 8;
        uint16_t in[8], out[8];
        if (aligned) {
            *reinterpret_cast<uint4*>(in) = yam_ld_stream(reinterpret_cast<const uint4*>(srow + x0));
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) in[i] = (x0 + i < g.w) ? srow[x0 + i] : (uint16_t)0;
        }
        uint2 q[8];
        float xa[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float txf = __fsub_rn(__fmul_rn((float)(x0 + i), g.inv_tw), 0.5f);
            const int fx = (int)floorf(txf);
            xa[i] = __fsub_rn(txf, (float)fx);
            const int cx = min(max(fx + 1, 0), g.tiles_x);
            q[i] = __ldg(qrow + (int64_t)cx * kBins16 + in[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float xa1 = __fsub_rn(1.0f, xa[i]);
            const float l11 = (float)(q[i].x & 0xffffu), l12 = (float)(q[i].x >> 16);
            const float l21 = (float)(q[i].y & 0xffffu), l22 = (float)(q[i].y >> 16);
            const float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa[i]));
            const float bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa[i]));
            const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
            out[i] = (uint16_t)yam_rint_sat(res, 65535);
        }
        if (aligned) {
            yam_st_stream(reinterpret_cast<uint4*>(drow + x0), *reinterpret_cast<const uint4*>(out));
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (x0 + i < g.w) drow[x0 + i] = out[i];
        }
    }
}

// ---- table-driven apply kernels for 16-bit frames ---------------------------------------------------
// The x interpolation weight and cell of a column are the same for every row: a one-off table
// (clahe_xtab_kernel, float xa[w] + uint16 cell[w]) replaces ~12 per-pixel instructions (I2F, FMUL, FADD,
// floor, F2I, min / max) by coalesced L1-resident loads; LUT values become floats with PRMT + FADD2
// (no I2F on the quarter-rate conversion pipe) and the blend runs on the packed fp32 pipe.  Each
// packed half is an IEEE fp32 operation in cv2's order (no FMA), so results are bit-identical.
__global__ void __launch_bounds__(256) clahe_xtab_kernel(ClaheGeom g, float* __restrict__ xa, uint16_t* __restrict__ cell) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= g.w) return;
    const float txf = __fsub_rn(__fmul_rn((float)x, g.inv_tw), 0.5f);
    const int fx = (int)floorf(txf);
    xa[x] = __fsub_rn(txf, (float)fx);
    cell[x] = (uint16_t)min(max(fx + 1, 0), g.tiles_x);   // tiles (max(cell-1,0), min(cell,tiles_x-1))
}

// q01 = l11 | l12 << 16, q23 = l21 | l22 << 16 -> blended, rounded 16-bit value
__device__ __forceinline__ uint32_t clahe_blend16(uint32_t q01, uint32_t q23, float xa, float2 yw) {
    const float2 nb = make_float2(-8388608.0f, -8388608.0f);
    const float2 a = __fadd2_rn(make_float2(__uint_as_float(__byte_perm(q01, 0x4B000000u, 0x7410)),
                                            __uint_as_float(__byte_perm(q01, 0x4B000000u, 0x7432))), nb);
    const float2 b = __fadd2_rn(make_float2(__uint_as_float(__byte_perm(q23, 0x4B000000u, 0x7410)),
                                            __uint_as_float(__byte_perm(q23, 0x4B000000u, 0x7432))), nb);
    const float2 xw = make_float2(__fsub_rn(1.0f, xa), xa);
    const float2 pa = __fmul2_rn(a, xw), pb = __fmul2_rn(b, xw);
    const float2 pr = __fmul2_rn(make_float2(__fadd_rn(pa.x, pa.y), __fadd_rn(pb.x, pb.y)), yw);
    const float res = __fadd_rn(pr.x, pr.y);
    // a convex combination of 16-bit values: 0 <= res <= 65535, so saturation cannot trigger and
    // rint(res) sits in the low mantissa bits of res + 1.5 * 2^23 (round-half-even)
    return __float_as_uint(__fadd_rn(res, 12582912.0f)) & 0xffffu;
}

// grid (rows, ceil(w / 2048), frames); thread = 8 pixels.  QUAD: one 8-byte gather per pixel from the
// interleaved cell table; otherwise four 2-byte gathers from the per-tile LUTs (stacks of small frames).
template <bool QUAD>
__global__ void __launch_bounds__(256) clahe_apply16_tab_kernel(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst,
                                                               ClaheGeom g, const void* __restrict__ table,
                                                               const float* __restrict__ xa_tab,
                                                               const uint16_t* __restrict__ cell_tab) {
    const int y = blockIdx.x;
    const int x0 = (blockIdx.y * 256 + threadIdx.x) * 8;
    if (x0 >= g.w) return;
    const int64_t frame_px = (int64_t)g.h * g.w;
    src += (int64_t)blockIdx.z * frame_px;
    dst += (int64_t)blockIdx.z * frame_px;
    const float tyf = __fsub_rn(__fmul_rn((float)(y + g.y_off), g.inv_th), 0.5f);
    const int fy = (int)floorf(tyf);
    const float ya = __fsub_rn(tyf, (float)fy);
    const float2 yw = make_float2(__fsub_rn(1.0f, ya), ya);
    const int cy = min(max(fy + 1, 0), g.tiles_y);
    const uint4 pix = yam_ld_stream(reinterpret_cast<const uint4*>(src + (int64_t)y * g.w + x0));
    const float4 xa_lo = __ldg(reinterpret_cast<const float4*>(xa_tab + x0));
    const float4 xa_hi = __ldg(reinterpret_cast<const float4*>(xa_tab + x0 + 4));
    const uint4 cells = __ldg(reinterpret_cast<const uint4*>(cell_tab + x0));
    const float xa[8] = {xa_lo.x, xa_lo.y, xa_lo.z, xa_lo.w, xa_hi.x, xa_hi.y, xa_hi.z, xa_hi.w};
    const uint32_t pw[4] = {pix.x, pix.y, pix.z, pix.w}, cw[4] = {cells.x, cells.y, cells.z, cells.w};
    uint32_t q01[8], q23[8];
    if (QUAD) {
        const uint2* qrow = reinterpret_cast<const uint2*>(table) + (int64_t)cy * (g.tiles_x + 1) * kBins16;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t v = (pw[i >> 1] >> (16 * (i & 1))) & 0xffffu, cx = (cw[i >> 1] >> (16 * (i & 1))) & 0xffffu;
            const uint2 q = __ldg(qrow + (int64_t)cx * kBins16 + v);
            q01[i] = q.x;
            q23[i] = q.y;
        }
    } else {
        const uint16_t* luts = reinterpret_cast<const uint16_t*>(table) + (int64_t)blockIdx.z * g.tiles_x * g.tiles_y * kBins16;
        const uint16_t* lrow1 = luts + (int64_t)max(cy - 1, 0) * g.tiles_x * kBins16;
        const uint16_t* lrow2 = luts + (int64_t)min(cy, g.tiles_y - 1) * g.tiles_x * kBins16;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t v = (pw[i >> 1] >> (16 * (i & 1))) & 0xffffu, cx = (cw[i >> 1] >> (16 * (i & 1))) & 0xffffu;
            const int64_t o1 = (int64_t)max((int)cx - 1, 0) * kBins16 + v, o2 = (int64_t)min((int)cx, g.tiles_x - 1) * kBins16 + v;
            q01[i] = (uint32_t)__ldg(lrow1 + o1) | ((uint32_t)__ldg(lrow1 + o2) << 16);
            q23[i] = (uint32_t)__ldg(lrow2 + o1) | ((uint32_t)__ldg(lrow2 + o2) << 16);
        }
    }
    uint32_t out[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
        out[i] = clahe_blend16(q01[2 * i], q23[2 * i], xa[2 * i], yw) |
                 (clahe_blend16(q01[2 * i + 1], q23[2 * i + 1], xa[2 * i + 1], yw) << 16);
    yam_st_stream(reinterpret_cast<uint4*>(dst + (int64_t)y * g.w + x0), make_uint4(out[0], out[1], out[2], out[3]));
}

inline bool clahe_tab_ok(const void* src, const void* dst, int64_t w) {
    return (w % 8) == 0 && (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0);
}
inline size_t clahe_tab_bytes(int64_t w) { return yam_align_up((size_t)w * 4, 256) + yam_align_up((size_t)w * 2, 256); }

// builds the column tables at `tab` (unless the caller already did: build_tab = false) and applies
// rows x frames; `table` = quad cells (QUAD) or the LUT set
template <bool QUAD>
int clahe_apply16_tab(yam_ctx* ctx, const uint16_t* s_ptr, uint16_t* d_ptr, int64_t rows, int64_t frames, const ClaheGeom& g,
                      const void* table, void* tab, bool build_tab = true) {
    float* xa = (float*)tab;
    uint16_t* cell = (uint16_t*)((char*)tab + yam_align_up((size_t)g.w * 4, 256));
    if (build_tab) {
        clahe_xtab_kernel<<<(unsigned)((g.w + 255) / 256), 256, 0, ctx->stream>>>(g, xa, cell);
        YAM_LAUNCHED(ctx);
    }
    dim3 grid((unsigned)rows, (unsigned)((g.w + 2047) / 2048), (unsigned)frames);
    clahe_apply16_tab_kernel<QUAD><<<grid, 256, 0, ctx->stream>>>(s_ptr, d_ptr, g, table, xa, cell);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

constexpr int64_t kQuadMinPixels = 8ll << 20;  // below this the 40 MB cell-table build costs more than it saves (measured again in round 2: 32 x 2048^2 frames 1.45 ms with four 2-byte gathers, 1.63 ms through per-frame cell tables)
constexpr int kQuadMaxCells = 128;

inline bool clahe_use_quad(int64_t pixels, int tiles_x, int tiles_y) {
    return pixels >= kQuadMinPixels && (tiles_x + 1) * (tiles_y + 1) <= kQuadMaxCells;
}
inline size_t clahe_quad_bytes(int tiles_x, int tiles_y) {
    return (size_t)(tiles_x + 1) * (tiles_y + 1) * kBins16 * sizeof(ushort4);
}
inline size_t clahe_tab_bytes(int64_t w);

// apply `rows` rows of one frame with the quad path: build quad from luts, then gather
int clahe_apply16_quad(yam_ctx* ctx, const uint16_t* s_ptr, uint16_t* d_ptr, int64_t rows, const ClaheGeom& g,
                       const uint16_t* luts, ushort4* quad, bool build_tab = true) {
    // scratch layout: [quad cells][column tables]
    const int cells = (g.tiles_x + 1) * (g.tiles_y + 1);
    clahe_quad_build_kernel<<<dim3(kBins16 / (8 * 256), (unsigned)cells), 256, 0, ctx->stream>>>(luts, g.tiles_x, g.tiles_y, quad);
    YAM_LAUNCHED(ctx);
    if (clahe_tab_ok(s_ptr, d_ptr, g.w)) {
        ClaheGeom gr = g;
        gr.h = (int)rows;
        return clahe_apply16_tab<true>(ctx, s_ptr, d_ptr, rows, 1, gr, quad, (char*)quad + clahe_quad_bytes(g.tiles_x, g.tiles_y), build_tab);
    }
    clahe_apply_quad_kernel<<<(unsigned)rows, 256, 0, ctx->stream>>>(s_ptr, d_ptr, g, quad);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

// cv2 CLAHE geometry: pad right/bottom (REFLECT_101) when either side is not divisible by the grid
// (both sides get padded then), tile area, clip limit, LUT scale
int clahe_geometry(int64_t h, int64_t w, int dtype, double clip_limit, int tiles_x, int tiles_y, ClaheGeom* out) {
    const int bins = dtype == YAM_U8 ? 256 : kBins16;
    int64_t pw = w, ph = h;
    if (w % tiles_x || h % tiles_y) {
        pw = w + (tiles_x - w % tiles_x);
        ph = h + (tiles_y - h % tiles_y);
    }
    ClaheGeom g;
    g.h = (int)h;
    g.w = (int)w;
    g.tiles_x = tiles_x;
    g.tiles_y = tiles_y;
    g.tw = (int)(pw / tiles_x);
    g.th = (int)(ph / tiles_y);
    const long long area = (long long)g.tw * g.th;
    YAM_REQUIRE(area < (1ll << 31), "clahe: tile area too large");
    g.lut_scale = (float)(bins - 1) / (float)area;
    g.clip = 0;
    if (clip_limit > 0) {
        int c = (int)(clip_limit * (double)area / bins);
        g.clip = c < 1 ? 1 : c;
    }
    g.inv_tw = 1.0f / (float)g.tw;
    g.inv_th = 1.0f / (float)g.th;
    g.y_off = 0;
    *out = g;
    return YAM_OK;
}

// LUTs of `nf` frames starting at s_ptr into `luts` (tiles x 65536 u16 per frame).  `scratch_tail`
// provides max(slots x 256 KiB overflow, tiles_total x 256 KiB histograms) of scratch.
int clahe_luts16(yam_ctx* ctx, const uint16_t* s_ptr, const ClaheGeom& g, int64_t nf, uint16_t* luts, void* scratch_tail) {
    const int64_t tiles = (int64_t)g.tiles_x * g.tiles_y;
    const int64_t total_tiles = tiles * nf;
    const int slots = ctx->num_sms;
    const long long area = (long long)g.tw * g.th;
    // function attributes are per DEVICE: one flag per device id (a process may own several contexts)
    static bool attr_set[64] = {};
    if (ctx->device < 0 || ctx->device >= 64 || !attr_set[ctx->device]) {
        YAM_CUDA(cudaFuncSetAttribute(clahe_lut16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem16));
        YAM_CUDA(cudaFuncSetAttribute(clahe_hist16_parts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem16));
        if (ctx->device >= 0 && ctx->device < 64) attr_set[ctx->device] = true;
    }
    // (YAM_CLAHE_PARTS_MIN_AREA: experiment knob, pixels per tile from which tiles are split over several CTAs)
    static const long long parts_min_area = [] {
        const char* e = getenv("YAM_CLAHE_PARTS_MIN_AREA");
        return e ? atoll(e) : (4ll << 20);
    }();
    if (total_tiles * 2 <= slots && area >= parts_min_area) {
        // few huge tiles: split every tile over several CTAs
        int64_t parts = (2 * (int64_t)slots + total_tiles - 1) / total_tiles;
        if (parts > g.th) parts = g.th;
        YAM_CUDA(cudaMemsetAsync(scratch_tail, 0, (size_t)total_tiles * kBins16 * sizeof(uint32_t), ctx->stream));
        void* slabs = nullptr;
        if (int rc = yam_scratch2(ctx, (size_t)total_tiles * parts * kSmem16, &slabs)) return rc;
        clahe_hist16_parts_kernel<<<dim3((unsigned)parts, (unsigned)total_tiles), kHistThreads, kSmem16, ctx->stream>>>(
            s_ptr, g, (uint32_t*)scratch_tail, (uint32_t*)slabs);
        YAM_LAUNCHED(ctx);
        hist16_reduce_kernel<uint32_t><<<dim3(kWords16 / 256, (unsigned)total_tiles), 256, 0, ctx->stream>>>(
            (const uint32_t*)slabs, (int)parts, (uint32_t*)scratch_tail);
        YAM_LAUNCHED(ctx);
        clahe_lut_from_hist_kernel<<<(unsigned)total_tiles, kHistThreads, 0, ctx->stream>>>((const uint32_t*)scratch_tail, g, luts);
        YAM_LAUNCHED(ctx);
        return YAM_OK;
    }
    // the kernel restores zeros after use, but scratch is shared with other ops: clear it
    YAM_CUDA(cudaMemsetAsync(scratch_tail, 0, (size_t)slots * kBins16 * sizeof(uint32_t), ctx->stream));
    const unsigned gridx = (unsigned)(total_tiles < slots ? total_tiles : slots);
    clahe_lut16_kernel<<<gridx, kHistThreads, kSmem16, ctx->stream>>>(s_ptr, g, (int)nf, (uint32_t*)scratch_tail, luts);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

int hist_into(yam_ctx* ctx, const void* src, int64_t n, int64_t h, int64_t w, int dtype,
              unsigned long long* hist) {
    const int bins = dtype == YAM_U8 ? 256 : kBins16;
    YAM_CUDA(cudaMemsetAsync(hist, 0, sizeof(unsigned long long) * bins * n, ctx->stream));
    if (dtype == YAM_U8) {
        const int64_t frame_px = h * w;
        int64_t bx = (frame_px / 16 + 255) / 256;
        int64_t cap = (int64_t)ctx->num_sms * 8 / n;
        if (cap < 1) cap = 1;
        if (bx > cap) bx = cap;
        if (bx < 1) bx = 1;
        hist8_kernel<<<dim3((unsigned)bx, 1, (unsigned)n), 256, 0, ctx->stream>>>((const uint8_t*)src, frame_px, hist);
    } else {
        static bool attr_set[64] = {};  // per device, see clahe_luts16
        if (ctx->device < 0 || ctx->device >= 64 || !attr_set[ctx->device]) {
            YAM_CUDA(cudaFuncSetAttribute(hist16_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem16));
            if (ctx->device >= 0 && ctx->device < 64) attr_set[ctx->device] = true;
        }
        int64_t parts = ctx->num_sms / n;
        if (parts < 1) parts = 1;
        if (parts > h) parts = h;
        // frames go through in groups so the slab scratch stays bounded (<= 296 slabs of 128 KiB)
        int64_t group = (2 * (int64_t)ctx->num_sms) / parts;
        if (group < 1) group = 1;
        if (group > n) group = n;
        void* slabs = nullptr;
        if (int rc = yam_scratch2(ctx, (size_t)group * parts * kSmem16, &slabs)) return rc;
        const int64_t frame_px = h * w;
        for (int64_t f0 = 0; f0 < n; f0 += group) {
            const int64_t nf = (n - f0) < group ? (n - f0) : group;
            hist16_slab_kernel<<<dim3((unsigned)parts, 1, (unsigned)nf), kHistThreads, kSmem16, ctx->stream>>>(
                (const uint16_t*)src + f0 * frame_px, (int)h, (int)w, hist + f0 * kBins16, (uint32_t*)slabs);
            YAM_LAUNCHED(ctx);
            hist16_reduce_kernel<unsigned long long><<<dim3(kWords16 / 256, (unsigned)nf), 256, 0, ctx->stream>>>(
                (const uint32_t*)slabs, (int)parts, hist + f0 * kBins16);
            if (f0 + group < n) YAM_LAUNCHED(ctx);
        }
    }
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

}  // namespace

extern "C" {

int yam_set_host_threads(int threads) {
    if (threads > 0) g_host_threads.store(threads);
    return host_threads();
}

// frames one pool task scans in lock step: 1 while there are idle threads, up to 4 beyond that
static int64_t otsu_group_size(int64_t frames) {
    const int64_t ht = host_threads();
    int64_t g = (frames + ht - 1) / ht;
    return g < 1 ? 1 : (g > 4 ? 4 : g);
}

int yam_otsu_from_hists(const uint64_t* hists, int bins, int64_t n, int32_t* out_thresholds) {
    YAM_REQUIRE(hists && out_thresholds && bins > 0 && n > 0, "yam_otsu_from_hists: bad arguments");
    if (n == 1) {
        out_thresholds[0] = yam_host_otsu(hists, bins);
        return YAM_OK;
    }
    ScanPool& pool = ScanPool::instance();
    const int64_t group = otsu_group_size(n);
    for (int64_t f = 0; f < n; f += group) {
        const int cnt = (int)((n - f) < group ? (n - f) : group);
        const uint64_t* hp = hists + f * bins;
        int32_t* outp = out_thresholds + f;
        pool.submit([hp, outp, bins, cnt] { yam_host_otsu_group(hp, sizeof(uint64_t) * (size_t)bins, cnt, bins, 0, outp); });
    }
    pool.wait_all();
    return YAM_OK;
}

int yam_threshold_frames(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w, int dtype,
                         const int32_t* thresh_dev, double maxval) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(src && dst && thresh_dev && n > 0 && h > 0 && w > 0 && n <= 65535, "threshold_frames: bad arguments");
    YAM_REQUIRE(dtype == YAM_U8 || dtype == YAM_U16, "threshold_frames: unsupported dtype %d", dtype);
    return yam_threshold_dev(ctx, src, dst, n, h * w, dtype, thresh_dev, maxval);
}

int yam_histogram(yam_ctx* ctx, const void* src, int64_t n, int64_t h, int64_t w, int dtype, uint64_t* hist_dev) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(src && hist_dev && n > 0 && h > 0 && w > 0 && n <= 65535, "histogram: bad arguments");
    YAM_REQUIRE(dtype == YAM_U8 || dtype == YAM_U16, "histogram: unsupported dtype %d", dtype);
    YAM_REQUIRE(h < (1 << 30) && w < (1 << 30), "histogram: image side too large");
    return hist_into(ctx, src, n, h, w, dtype, (unsigned long long*)hist_dev);
}

// Thresholds of n histograms that already sit in device memory, enqueued on ctx->stream with no host
// synchronisation: 65536 bins -> certified parallel scan, exact chain kernels for the frames it could
// not certify (they exit at once for the others); 256 bins -> one device thread per frame.
// `stage` = otsu_stage_bytes(n) bytes of scratch.
static constexpr int64_t kStageChunk = 256;
static size_t otsu_stage_bytes(int64_t n, int bins) {
    if (bins != kBins16) return 0;
    const int64_t nf_max = n < kStageChunk ? n : kStageChunk;
    return 2 * yam_align_up(sizeof(double) * bins * nf_max, 256) + yam_align_up(sizeof(OtsuMeta) * nf_max, 256) + 256;
}
static std::atomic<int> g_otsu_force_chain{0};
static int otsu_scan_device(yam_ctx* ctx, const unsigned long long* hist, int bins, int64_t n, int32_t* t_dev, void* stage,
                            int32_t* certified_dev = nullptr) {
    if (bins != kBins16) {
        if (certified_dev) YAM_CUDA(cudaMemsetAsync(certified_dev, 0, sizeof(int32_t) * n, ctx->stream));
        otsu_scan_kernel<<<(unsigned)((n + 31) / 32), 32, 0, ctx->stream>>>(hist, bins, n, t_dev);
        YAM_LAUNCHED(ctx);
        return YAM_OK;
    }
    const int64_t nf_max = n < kStageChunk ? n : kStageChunk;
    const size_t arr_bytes = yam_align_up(sizeof(double) * bins * nf_max, 256);
    char* sp = (char*)stage;
    double* q1arr = (double*)sp; sp += arr_bytes;
    double* mu1arr = (double*)sp; sp += arr_bytes;
    OtsuMeta* meta = (OtsuMeta*)sp; sp += yam_align_up(sizeof(OtsuMeta) * nf_max, 256);
    int* redo = (int*)sp;
    YAM_CUDA(cudaMemsetAsync(redo, 0, sizeof(int), ctx->stream));
    for (int64_t f0 = 0; f0 < n; f0 += kStageChunk) {
        const unsigned nf = (unsigned)((n - f0) < kStageChunk ? (n - f0) : kStageChunk);
        const unsigned long long* hc = hist + f0 * bins;
        otsu_certify_kernel<<<dim3(kCertCluster, nf), kCertThreads, 0, ctx->stream>>>(hc, meta, t_dev + f0, certified_dev ? certified_dev + f0 : nullptr,
                                                                                      g_otsu_force_chain.load());
        YAM_LAUNCHED(ctx);
        otsu_chain_kernel<<<nf, 32, 0, ctx->stream>>>(hc, bins, meta, q1arr, mu1arr);
        YAM_LAUNCHED(ctx);
        otsu_sigma_kernel<<<nf, 1024, 0, ctx->stream>>>(hc, bins, meta, q1arr, mu1arr, t_dev + f0, redo);
        YAM_LAUNCHED(ctx);
    }
    return YAM_OK;
}

int yam_otsu_set_force_chain(int on) { return g_otsu_force_chain.exchange(on ? 1 : 0); }

int yam_otsu_from_hist_dev(yam_ctx* ctx, const uint64_t* hist_dev, int bins, int64_t n, int32_t* thresh_dev,
                           int32_t* certified_dev) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(hist_dev && thresh_dev && n > 0 && n <= 65535 && (bins == 256 || bins == kBins16), "otsu_from_hist_dev: bad arguments");
    void* stage = nullptr;
    if (int rc = yam_scratch(ctx, otsu_stage_bytes(n, bins) + 256, &stage)) return rc;
    return otsu_scan_device(ctx, (const unsigned long long*)hist_dev, bins, n, thresh_dev, stage, certified_dev);
}

int yam_otsu_threshold(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w, int dtype,
                       double maxval, int32_t* thresh_dev, int32_t* thresh_host) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(src && n > 0 && h > 0 && w > 0 && n <= 65535, "otsu: bad arguments");
    YAM_REQUIRE(dtype == YAM_U8 || dtype == YAM_U16, "otsu: unsupported dtype %d", dtype);
    YAM_REQUIRE(h < (1 << 30) && w < (1 << 30), "otsu: image side too large");
    const int bins = dtype == YAM_U8 ? 256 : kBins16;
    const size_t hist_bytes = sizeof(unsigned long long) * bins * n;
    // histogram -> scan -> threshold, all on the stream: nothing is read back and the host does not wait
    // (the round-1/2 schedule read the histograms back and ran the sequential fp64 recurrence on host threads)
    const size_t hist_block = yam_align_up(yam_align_up(hist_bytes, 256) + yam_align_up(sizeof(int32_t) * n, 256), 256);
    void* scratch = nullptr;
    if (int rc = yam_scratch(ctx, hist_block + otsu_stage_bytes(n, bins) + 256, &scratch)) return rc;
    unsigned long long* hist = (unsigned long long*)scratch;
    int32_t* t_dev = thresh_dev ? thresh_dev : (int32_t*)((char*)scratch + yam_align_up(hist_bytes, 256));
    if (int rc = hist_into(ctx, src, n, h, w, dtype, hist)) return rc;
    if (int rc = otsu_scan_device(ctx, hist, bins, n, t_dev, (char*)scratch + hist_block)) return rc;
    if (dst) {
        if (int rc = yam_threshold_dev(ctx, src, dst, n, h * w, dtype, t_dev, maxval)) return rc;
    }
    if (thresh_host) {
        YAM_CUDA(cudaMemcpyAsync(thresh_host, t_dev, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
        YAM_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return YAM_OK;
}

int yam_equalize_hist(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(src && dst && n > 0 && h > 0 && w > 0 && n <= 65535, "equalize_hist: bad arguments");
    const size_t hist_bytes = sizeof(unsigned long long) * 256 * n;
    void* scratch = nullptr;
    if (int rc = yam_scratch(ctx, hist_bytes + 256 * n, &scratch)) return rc;
    unsigned long long* hist = (unsigned long long*)scratch;
    uint8_t* luts = (uint8_t*)scratch + hist_bytes;
    if (int rc = hist_into(ctx, src, n, h, w, YAM_U8, hist)) return rc;
    equalize_lut_kernel<<<(unsigned)n, 256, 0, ctx->stream>>>(hist, h * w, luts);
    YAM_LAUNCHED(ctx);
    const int64_t frame_px = h * w;
    int64_t bx = (frame_px / 16 + 255) / 256;
    int64_t cap = (int64_t)ctx->num_sms * 8 / n;
    if (cap < 1) cap = 1;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    gather_lut8_kernel<<<dim3((unsigned)bx, 1, (unsigned)n), 256, 0, ctx->stream>>>((const uint8_t*)src, (uint8_t*)dst,
                                                                                  frame_px, luts);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

int yam_clahe(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w, int dtype,
              double clip_limit, int tiles_x, int tiles_y) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(src && dst && src != dst && n > 0 && h > 0 && w > 0 && n <= 65535, "clahe: bad arguments");
    YAM_REQUIRE(dtype == YAM_U8 || dtype == YAM_U16, "clahe: unsupported dtype %d", dtype);
    YAM_REQUIRE(tiles_x >= 1 && tiles_y >= 1 && tiles_x <= 256 && tiles_y <= 256, "clahe: bad tile grid %dx%d", tiles_x, tiles_y);
    YAM_REQUIRE(h < (1 << 30) && w < (1 << 30), "clahe: image side too large");
    const int bins = dtype == YAM_U8 ? 256 : kBins16;
    ClaheGeom g;
    if (int rc = clahe_geometry(h, w, dtype, clip_limit, tiles_x, tiles_y, &g)) return rc;
    const int64_t tiles = (int64_t)tiles_x * tiles_y;
    const size_t lut_bytes = (size_t)n * tiles * bins * yam_dtype_size(dtype);
    if (dtype == YAM_U16) {
        // frames are processed in chunks so the LUT set of a chunk (tiles x 128 KiB per frame) stays
        // L2-resident between the LUT kernel and the apply kernel, and scratch stays bounded.
        const size_t lut_frame = (size_t)tiles * kBins16 * sizeof(uint16_t);
        int64_t chunk = (int64_t)((64u << 20) / lut_frame);
        if (chunk < 1) chunk = 1;
        if (chunk > n) chunk = n;
        // quad path: the LUTs of a chunk are built together (enough tiles to fill the SMs), then every frame
        // interleaves its cell table and gathers from it while it is still in L2
        const bool quad_path = clahe_use_quad(h * w, tiles_x, tiles_y);
        const int64_t tail_slots = (tiles * chunk > ctx->num_sms) ? tiles * chunk : ctx->num_sms;
        const size_t ovf_bytes = yam_align_up((size_t)tail_slots * kBins16 * sizeof(uint32_t), 256);
        const size_t luts_bytes = yam_align_up(lut_frame * chunk, 256);
        const size_t quad_bytes = quad_path ? clahe_quad_bytes(tiles_x, tiles_y) : 0;
        const size_t tab_bytes = clahe_tab_bytes(w);
        void* scratch = nullptr;
        if (int rc = yam_scratch(ctx, luts_bytes + ovf_bytes + quad_bytes + tab_bytes, &scratch)) return rc;
        uint16_t* luts = (uint16_t*)scratch;
        void* tail = (char*)scratch + luts_bytes;
        ushort4* quad = (ushort4*)((char*)scratch + luts_bytes + ovf_bytes);
        const int64_t frame_px = h * w;
        for (int64_t f0 = 0; f0 < n; f0 += chunk) {
            const int64_t nf = (n - f0) < chunk ? (n - f0) : chunk;
            const uint16_t* s_ptr = (const uint16_t*)src + f0 * frame_px;
            uint16_t* d_ptr = (uint16_t*)dst + f0 * frame_px;
            if (int rc = clahe_luts16(ctx, s_ptr, g, nf, luts, tail)) return rc;
            if (quad_path) {
                for (int64_t f = 0; f < nf; f++)
                    if (int rc = clahe_apply16_quad(ctx, s_ptr + f * frame_px, d_ptr + f * frame_px, h, g, luts + f * tiles * kBins16,
                                                    quad, f0 + f == 0))
                        return rc;
            } else if (clahe_tab_ok(s_ptr, d_ptr, w)) {
                if (int rc = clahe_apply16_tab<false>(ctx, s_ptr, d_ptr, h, nf, g, luts, (char*)scratch + luts_bytes + ovf_bytes)) return rc;
            } else {
                dim3 grid((unsigned)h, 1, (unsigned)nf);
                clahe_apply_kernel<uint16_t><<<grid, 256, 0, ctx->stream>>>(s_ptr, d_ptr, g, luts);
                YAM_LAUNCHED(ctx);
            }
        }
    } else {
        void* scratch = nullptr;
        if (int rc = yam_scratch(ctx, lut_bytes, &scratch)) return rc;
        uint8_t* luts = (uint8_t*)scratch;
        clahe_lut8_kernel<<<dim3((unsigned)tiles, 1, (unsigned)n), 256, 0, ctx->stream>>>((const uint8_t*)src, g, luts);
        YAM_LAUNCHED(ctx);
        dim3 grid((unsigned)h, 1, (unsigned)n);
        clahe_apply_kernel<uint8_t><<<grid, 256, 0, ctx->stream>>>((const uint8_t*)src, (uint8_t*)dst, g, luts);
        YAM_LAUNCHED(ctx);
    }
    return YAM_OK;
}

int yam_clahe_luts(yam_ctx* ctx, const void* src, int64_t h, int64_t w, int dtype, double clip_limit, int tiles_x,
                   int tiles_y, void* luts_dev) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(src && luts_dev && h > 0 && w > 0, "clahe_luts: bad arguments");
    YAM_REQUIRE(dtype == YAM_U8 || dtype == YAM_U16, "clahe_luts: unsupported dtype %d", dtype);
    YAM_REQUIRE(tiles_x >= 1 && tiles_y >= 1 && tiles_x <= 256 && tiles_y <= 256, "clahe_luts: bad tile grid");
    YAM_REQUIRE(h < (1 << 30) && w < (1 << 30), "clahe_luts: image side too large");
    ClaheGeom g;
    if (int rc = clahe_geometry(h, w, dtype, clip_limit, tiles_x, tiles_y, &g)) return rc;
    const int64_t tiles = (int64_t)tiles_x * tiles_y;
    if (dtype == YAM_U16) {
        const int64_t tail_slots = tiles > ctx->num_sms ? tiles : ctx->num_sms;
        void* scratch = nullptr;
        if (int rc = yam_scratch(ctx, (size_t)tail_slots * kBins16 * sizeof(uint32_t), &scratch)) return rc;
        return clahe_luts16(ctx, (const uint16_t*)src, g, 1, (uint16_t*)luts_dev, scratch);
    } else {
        clahe_lut8_kernel<<<dim3((unsigned)tiles, 1, 1), 256, 0, ctx->stream>>>((const uint8_t*)src, g, (uint8_t*)luts_dev);
    }
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

int yam_clahe_apply(yam_ctx* ctx, const void* src, void* dst, int64_t rows, int64_t w, int dtype, const void* luts_dev,
                    int tiles_x, int tiles_y, int tile_w, int tile_h, int64_t y_offset) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(src && dst && src != dst && luts_dev && rows > 0 && w > 0, "clahe_apply: bad arguments");
    YAM_REQUIRE(dtype == YAM_U8 || dtype == YAM_U16, "clahe_apply: unsupported dtype %d", dtype);
    YAM_REQUIRE(tiles_x >= 1 && tiles_y >= 1 && tile_w >= 1 && tile_h >= 1 && y_offset >= 0, "clahe_apply: bad geometry");
    ClaheGeom g;
    g.h = (int)rows;
    g.w = (int)w;
    g.tiles_x = tiles_x;
    g.tiles_y = tiles_y;
    g.tw = tile_w;
    g.th = tile_h;
    g.clip = 0;
    g.lut_scale = 0.f;
    g.inv_tw = 1.0f / (float)tile_w;
    g.inv_th = 1.0f / (float)tile_h;
    g.y_off = (int)y_offset;
    dim3 grid((unsigned)rows, 1, 1);
    if (dtype == YAM_U16 && clahe_use_quad(rows * w, tiles_x, tiles_y)) {
        void* scratch = nullptr;
        if (int rc = yam_scratch(ctx, clahe_quad_bytes(tiles_x, tiles_y) + clahe_tab_bytes(w), &scratch)) return rc;
        return clahe_apply16_quad(ctx, (const uint16_t*)src, (uint16_t*)dst, rows, g, (const uint16_t*)luts_dev, (ushort4*)scratch);
    }
    if (dtype == YAM_U16 && clahe_tab_ok(src, dst, w)) {
        void* scratch = nullptr;
        if (int rc = yam_scratch(ctx, clahe_tab_bytes(w), &scratch)) return rc;
        return clahe_apply16_tab<false>(ctx, (const uint16_t*)src, (uint16_t*)dst, rows, 1, g, luts_dev, scratch);
    }
    if (dtype == YAM_U16)
        clahe_apply_kernel<uint16_t><<<grid, 256, 0, ctx->stream>>>((const uint16_t*)src, (uint16_t*)dst, g, (const uint16_t*)luts_dev);
    else
        clahe_apply_kernel<uint8_t><<<grid, 256, 0, ctx->stream>>>((const uint8_t*)src, (uint8_t*)dst, g, (const uint8_t*)luts_dev);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

}  // extern "C"
