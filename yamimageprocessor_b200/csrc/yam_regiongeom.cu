// Contour / hull geometry of labelled regions: the `perimeter` and `solidity` columns of
// region_properties_data (core/extraction.py:80,83 -> skimage.measure.regionprops).
//
// perimeter  = skimage.measure.perimeter(region image, neighborhood=4): a region pixel is a BORDER pixel
//   when one of its four neighbours is not in the region (outside the image counts as background); a
//   border pixel is classed by v = 1 + 2 * #(4-neighbours that are border pixels of the same region)
//   + 10 * #(diagonal neighbours that are); classes {5, 7, 15, 17, 25, 27} weigh 1, {21, 33} sqrt 2,
//   {13, 23} (1 + sqrt 2) / 2.  Everything is local to a 5 x 5 window of the LABEL image, so all regions
//   are done in one pass; the kernel counts the three classes per label (integers, exact) and the host
//   forms n1 + n2 * sqrt 2 + n3 * (1 + sqrt 2) / 2 in float64.
// solidity   = area / area_convex, area_convex = pixels whose centre lies in the closed convex hull of the
//   edge midpoints of the region's pixels (skimage.morphology.convex_hull_image, offset_coordinates=True,
//   include_borders=True).  In doubled coordinates the midpoints are integers; per doubled row only the
//   smallest / largest column can be a hull vertex, so each region has two point lists of 2 * height + 1
//   entries (filled with atomicMin by one pass over the label image), each reduced to a lower convex hull
//   by a monotone chain (one thread per region and side), and the hull's pixel count is a sum over the
//   region's rows of exact integer ceil / floor divisions.  tests/region_geometry_model.py is the same
//   arithmetic in Python, checked on the CPU against a literal restatement of the skimage calls.
#include "yam_common.cuh"

namespace {

constexpr int kPT = 32;        // perimeter tile side
constexpr int kPS = kPT + 4;   // + two rings: the class of a pixel needs the border flags of its 3 x 3
                               //   neighbourhood, each of which needs its own four neighbours
constexpr int kMissing = 0x7f7f7f7f;   // cudaMemset(0x7f): no pixel contributed to this doubled row

// class of a border pixel by v = 1 + 2 a + 10 d (a <= 4, d <= 4 -> v <= 49); 0 = weighs nothing
__constant__ uint8_t kPerimeterClass[50] = {
    /* 0*/ 0, 0, 0, 0, 0, /* 5*/ 1, 0, /* 7*/ 1, 0, 0,
    /*10*/ 0, 0, 0, /*13*/ 3, 0, /*15*/ 1, 0, /*17*/ 1, 0, 0,
    /*20*/ 0, /*21*/ 2, 0, /*23*/ 3, 0, /*25*/ 1, 0, /*27*/ 1, 0, 0,
    /*30*/ 0, 0, 0, /*33*/ 2, 0, 0, 0, 0, 0, 0,
    /*40*/ 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

__global__ void __launch_bounds__(256) region_perimeter_kernel(const int32_t* __restrict__ labels, int h, int w, int64_t n_labels,
                                                               int tiles_x, int64_t tiles, unsigned long long* __restrict__ out) {
    __shared__ int s_lab[kPS][kPS + 1];
    __shared__ uint8_t s_border[kPS][kPS + 4];
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int ty0 = (int)(t / tiles_x) * kPT, tx0 = (int)(t % tiles_x) * kPT;
        __syncthreads();   // the previous tile's readers are done
        for (int i = threadIdx.x; i < kPS * kPS; i += 256) {
            const int ry = i / kPS, rx = i - ry * kPS;
            const int gy = ty0 - 2 + ry, gx = tx0 - 2 + rx;
            s_lab[ry][rx] = (gy >= 0 && gy < h && gx >= 0 && gx < w) ? labels[(int64_t)gy * w + gx] : 0;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < kPS * kPS; i += 256) {
            const int ry = i / kPS, rx = i - ry * kPS;
            bool b = false;
            if (ry >= 1 && ry < kPS - 1 && rx >= 1 && rx < kPS - 1) {
                const int l = s_lab[ry][rx];
                b = l > 0 && (s_lab[ry - 1][rx] != l || s_lab[ry + 1][rx] != l || s_lab[ry][rx - 1] != l || s_lab[ry][rx + 1] != l);
            }
            s_border[ry][rx] = b ? 1 : 0;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < kPT * kPT; i += 256) {
            const int py = i / kPT + 2, px = i % kPT + 2;
            if (!s_border[py][px]) continue;   // background, interior and out-of-image pixels
            const int l = s_lab[py][px];
            if (l > n_labels) continue;
            int a = 0, d = 0;
#pragma unroll
            for (int dy = -1; dy <= 1; dy++)
#pragma unroll
                for (int dx = -1; dx <= 1; dx++) {
                    if (dy == 0 && dx == 0) continue;
                    const int hit = (s_lab[py + dy][px + dx] == l && s_border[py + dy][px + dx]) ? 1 : 0;
                    if (dy == 0 || dx == 0) a += hit;
                    else d += hit;
                }
            const int cls = kPerimeterClass[1 + 2 * a + 10 * d];
            if (cls) atomicAdd(out + (int64_t)(l - 1) * 3 + (cls - 1), 1ULL);
        }
    }
}

// points of region i's two chains: one per doubled row 2 r_min - 1 .. 2 r_max + 1
__device__ __forceinline__ int64_t hull_points(const int64_t* __restrict__ props, int64_t i) {
    const int64_t* p = props + i * YAM_PROPS_STRIDE;
    return p[0] > 0 ? 2 * (p[6] - p[4]) + 1 : 0;
}

// exclusive prefix of the per-region point counts; off[n] = total.  One CTA: the table has one entry per
// region (10^5 .. 10^7), the pass over the label image dominates
__global__ void __launch_bounds__(1024) hull_offsets_kernel(const int64_t* __restrict__ props, int64_t n, int64_t* __restrict__ off) {
    __shared__ int64_t s_part[1024];
    const int t = threadIdx.x;
    const int64_t chunk = (n + 1023) / 1024;
    int64_t lo = (int64_t)t * chunk;
    if (lo > n) lo = n;
    int64_t hi = lo + chunk;
    if (hi > n) hi = n;
    int64_t sum = 0;
    for (int64_t i = lo; i < hi; i++) sum += hull_points(props, i);
    s_part[t] = sum;
    __syncthreads();
    if (t == 0) {
        int64_t run = 0;
        for (int j = 0; j < 1024; j++) {
            const int64_t v = s_part[j];
            s_part[j] = run;
            run += v;
        }
        off[n] = run;
    }
    __syncthreads();
    int64_t run = s_part[t];
    for (int64_t i = lo; i < hi; i++) {
        off[i] = run;
        run += hull_points(props, i);
    }
}

// thread = 8 consecutive pixels of a row; a run of label L in row y touches the doubled rows 2y - 1, 2y,
// 2y + 1 (k = base, base + 1, base + 2) with its first pixel (left chain) and its last (right, negated)
__global__ void __launch_bounds__(256) hull_fill_kernel(const int32_t* __restrict__ labels, int h, int w, int64_t n_labels,
                                                        const int64_t* __restrict__ props, const int64_t* __restrict__ off,
                                                        int* __restrict__ vl, int* __restrict__ vr) {
    const int groups = (w + 7) / 8;
    const int64_t total = (int64_t)h * groups;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
        const int y = (int)(g / groups), x0 = (int)(g - (int64_t)y * groups) * 8;
        const int32_t* row = labels + (int64_t)y * w;
        int cur = 0, xf = 0, xl = 0;
        auto flush = [&]() {
            if (cur > 0 && cur <= n_labels) {
                const int64_t* p = props + (int64_t)(cur - 1) * YAM_PROPS_STRIDE;
                const int64_t dy = (int64_t)y - p[4];
                if (dy < 0 || dy >= p[6] - p[4]) return;   // props not of this label image: leave the row out
                const int64_t base = off[cur - 1] + 2 * dy;
                atomicMin(vl + base, 2 * xf);
                atomicMin(vl + base + 1, 2 * xf - 1);
                atomicMin(vl + base + 2, 2 * xf);
                atomicMin(vr + base, -2 * xl);
                atomicMin(vr + base + 1, -(2 * xl + 1));
                atomicMin(vr + base + 2, -2 * xl);
            }
        };
        for (int i = 0; i < 8 && x0 + i < w; i++) {
            const int l = row[x0 + i];
            if (l != cur) {
                flush();
                cur = l;
                xf = x0 + i;
            }
            xl = x0 + i;
        }
        flush();
    }
}

// thread = (region, side): lower convex hull of (k, v[k]) by a monotone chain (stack of k in scratch),
// then for every pixel row (odd k) under hull edge (ka, va) - (kb, vb) the bound va + (vb - va)(k - ka) /
// (kb - ka) in doubled columns; the side adds  side - ceil(bound / 2)  (left: -first column, right: last
// column + 1), so the two sides of a region sum to its hull's pixel count
__global__ void __launch_bounds__(128) hull_chain_kernel(const int64_t* __restrict__ off, int64_t n_labels, const int* __restrict__ vl,
                                                         const int* __restrict__ vr, int* __restrict__ stack_l, int* __restrict__ stack_r,
                                                         unsigned long long* __restrict__ convex_area) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n_labels) return;
    const int64_t region = t >> 1;
    const int side = (int)(t & 1);
    const int64_t first = off[region];
    const int np = (int)(off[region + 1] - first);
    if (np <= 0) return;
    const int* v = (side ? vr : vl) + first;
    int* st = (side ? stack_r : stack_l) + first;
    int m = 0;
    for (int k = 0; k < np; k++) {
        const int vk = v[k];
        if (vk == kMissing) continue;
        while (m >= 2) {
            const int ka = st[m - 2], kb = st[m - 1];
            const long long va = v[ka], vb = v[kb];
            const long long cross = (long long)(kb - ka) * ((long long)vk - va) - (vb - va) * (long long)(k - ka);
            if (cross <= 0) m--;   // kb is on or above the chord ka -> k
            else break;
        }
        st[m++] = k;
    }
    long long total = 0;
    for (int j = 0; j + 1 < m; j++) {
        const int ka = st[j], kb = st[j + 1];
        const long long va = v[ka], vb = v[kb];
        const long long den = kb - ka, den2 = 2 * den;
        for (int k = (ka & 1) ? ka : ka + 1; k < kb; k += 2) {   // pixel rows with ka <= k < kb
            const long long num = va * den + (vb - va) * (long long)(k - ka);
            const long long c = num >= 0 ? (num + den2 - 1) / den2 : -((-num) / den2);   // ceil(num / den2)
            total += side - c;
        }
    }
    atomicAdd(convex_area + region, (unsigned long long)total);   // the left side is negative: wraps back
}

}  // namespace

extern "C" {

int yam_region_perimeter(yam_ctx* ctx, const int32_t* labels, int64_t h, int64_t w, int64_t n_labels, int64_t* counts_dev) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(labels && counts_dev && h > 0 && w > 0 && n_labels >= 0, "region_perimeter: bad arguments");
    YAM_REQUIRE(h < (1 << 24) && w < (1 << 24), "region_perimeter: image side must be below 2^24");
    if (n_labels == 0) return YAM_OK;
    YAM_CUDA(cudaMemsetAsync(counts_dev, 0, (size_t)n_labels * 3 * sizeof(int64_t), ctx->stream));
    const int64_t tiles_x = (w + kPT - 1) / kPT, tiles_y = (h + kPT - 1) / kPT;
    const int64_t tiles = tiles_x * tiles_y;
    int64_t bx = tiles;
    const int64_t cap = (int64_t)ctx->num_sms * 8;
    if (bx > cap) bx = cap;
    region_perimeter_kernel<<<(unsigned)bx, 256, 0, ctx->stream>>>(labels, (int)h, (int)w, n_labels, (int)tiles_x, tiles,
                                                                  (unsigned long long*)counts_dev);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

int yam_region_convex_area(yam_ctx* ctx, const int32_t* labels, int64_t h, int64_t w, int64_t n_labels, const int64_t* props_dev,
                           int64_t* convex_area_dev) {
    if (int rc = yam_enter(ctx)) return rc;
    YAM_REQUIRE(labels && props_dev && convex_area_dev && h > 0 && w > 0 && n_labels >= 0, "region_convex_area: bad arguments");
    YAM_REQUIRE(h < (1 << 24) && w < (1 << 24), "region_convex_area: image side must be below 2^24");
    if (n_labels == 0) return YAM_OK;
    YAM_CUDA(cudaMemsetAsync(convex_area_dev, 0, (size_t)n_labels * sizeof(int64_t), ctx->stream));
    void* s2 = nullptr;
    if (int rc = yam_scratch2(ctx, (size_t)(n_labels + 1) * sizeof(int64_t), &s2)) return rc;
    int64_t* off = (int64_t*)s2;
    hull_offsets_kernel<<<1, 1024, 0, ctx->stream>>>(props_dev, n_labels, off);
    YAM_LAUNCHED(ctx);
    void* pin = nullptr;
    if (int rc = yam_pinned(ctx, 64, &pin)) return rc;
    int64_t* total_host = (int64_t*)pin;
    YAM_CUDA(cudaMemcpyAsync(total_host, off + n_labels, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    YAM_CUDA(cudaStreamSynchronize(ctx->stream));   // the point tables are sized by the sum of the region heights
    const int64_t total = *total_host;
    YAM_REQUIRE(total >= 0 && total <= n_labels * (2 * h + 1), "region_convex_area: props do not describe this label image");
    if (total == 0) return YAM_OK;
    void* s = nullptr;
    if (int rc = yam_scratch(ctx, (size_t)total * 4 * sizeof(int), &s)) return rc;
    int* vl = (int*)s;
    int* vr = vl + total;
    int* stack_l = vr + total;
    int* stack_r = stack_l + total;
    YAM_CUDA(cudaMemsetAsync(vl, 0x7f, (size_t)total * 2 * sizeof(int), ctx->stream));
    const int64_t groups = h * ((w + 7) / 8);
    int64_t bx = (groups + 255) / 256;
    const int64_t cap = (int64_t)ctx->num_sms * 16;
    if (bx > cap) bx = cap;
    hull_fill_kernel<<<(unsigned)bx, 256, 0, ctx->stream>>>(labels, (int)h, (int)w, n_labels, props_dev, off, vl, vr);
    YAM_LAUNCHED(ctx);
    const int64_t threads = 2 * n_labels;
    hull_chain_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(off, n_labels, vl, vr, stack_l, stack_r,
                                                                                 (unsigned long long*)convex_area_dev);
    YAM_LAUNCHED(ctx);
    return YAM_OK;
}

}  // extern "C"
