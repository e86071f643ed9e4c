/*
 * yamb200.h — C ABI of libyamb200.so, the B200 (sm_100a) backend for the per-pixel hot path of
 * GerryDoesStuff/YamImageProcessor (preprocessing -> segmentation -> extraction).
 *
 * The reference is pure Python; its "FFI" for this path is the set of cv2 / skimage calls made
 * by its step functions.  Each entry point below replaces one of those call sites (cited as
 * file:line relative to the reference checkout) and is bound from Python with ctypes
 * (yamimageprocessor_b200/_lib.py; the reference-side stub is shown in INTEGRATION.md).
 *
 * Conventions
 *  - plain C types only; no exceptions or aborts cross this boundary.  Every function returns
 *    YAM_OK (0) or a negative YAM_E* code; yam_last_error() returns the thread-local message.
 *  - images are dense row-major planes; a call processes a stack of `n` frames laid out
 *    back-to-back (n, h, w[, 3]).  The reference processes stacks plane by plane
 *    (processing/pipeline_manager.py:475-492), so every statistic (min/max, histogram, Otsu
 *    threshold, CLAHE tiles, labels) is per frame.
 *  - all `src`/`dst`/`labels`... pointers are DEVICE pointers on the context's GPU, 16-byte aligned.
 *    Work is enqueued on the context's stream; functions with host out-parameters synchronise
 *    that stream before returning.  A context is not re-entrant; use one per host thread.
 *  - there is no CPU fallback anywhere in this library.
 */
#ifndef YAMB200_H
#define YAMB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YAM_ABI_VERSION 1

/* error codes */
#define YAM_OK 0
#define YAM_EINVAL (-1)   /* bad argument (unsupported dtype / size / parameter) */
#define YAM_ECUDA (-2)    /* CUDA runtime error; message carries cudaGetErrorString */
#define YAM_ENOMEM (-3)   /* device scratch allocation failed */
#define YAM_ENODEV (-4)   /* no usable sm_100 device */

/* dtype codes */
#define YAM_U8 0
#define YAM_U16 1
#define YAM_F32 2
#define YAM_I32 3

/* border modes (cv2 names) */
#define YAM_BORDER_REFLECT101 0
#define YAM_BORDER_REPLICATE 1

/* morphology */
#define YAM_MORPH_ERODE 0
#define YAM_MORPH_DILATE 1
#define YAM_MORPH_OPEN 2
#define YAM_MORPH_CLOSE 3
#define YAM_MORPH_OPEN_CLOSE 4 /* open followed by close (bit-mask path) */
#define YAM_SHAPE_RECT 0
#define YAM_SHAPE_ELLIPSE 1
#define YAM_SHAPE_CROSS 2

typedef struct yam_ctx yam_ctx;

/* ---- context ------------------------------------------------------------------------------ */
int yam_abi_version(void);
const char* yam_last_error(void);
/* number of visible CUDA devices, or a negative error */
int yam_device_count(void);
int yam_ctx_create(int device, yam_ctx** out);
int yam_ctx_destroy(yam_ctx* ctx);
/* stream: a cudaStream_t; NULL is CUDA's legacy default stream. A new context starts on its own
 * non-blocking stream. */
int yam_ctx_set_stream(yam_ctx* ctx, void* stream);
int yam_ctx_synchronize(yam_ctx* ctx);
/* counts kernel launches issued through this context since the last reset (bench "gpu_launches") */
int64_t yam_ctx_launch_count(yam_ctx* ctx, int reset);

/* Host worker threads the host-only helper yam_otsu_from_hists may use (default: all hardware
 * threads, at most 32).  The device operators do not use host threads.  Returns the value in effect. */
int yam_set_host_threads(int threads);
/* raw device memory + copies for hosts that do not bring their own allocator */
int yam_malloc(yam_ctx* ctx, int64_t bytes, void** out);
int yam_free(yam_ctx* ctx, void* ptr);
int yam_memcpy_h2d(yam_ctx* ctx, void* dst_dev, const void* src_host, int64_t bytes);
int yam_memcpy_d2h(yam_ctx* ctx, void* dst_host, const void* src_dev, int64_t bytes);

/* ---- host-only helpers (no GPU needed; used by the CPU test-suite) ------------------------- */
/* cv2.getGaussianKernel(ksize, sigma) in double; sigma<=0 => 0.3*((k-1)*0.5-1)+0.8 / small tables */
int yam_gaussian_taps_f64(int ksize, double sigma, double* out);
/* cv2 fixed-point taps (bits = 8 | 16), edge->centre error diffusion */
int yam_gaussian_taps_fixed(int ksize, double sigma, int bits, int64_t* out);
/* cv2.getStructuringElement(shape,(k,k)) as k*k bytes of 0/1 */
int yam_structuring_element(int shape, int ksize, uint8_t* out);
/* cv2 Otsu recurrence on a histogram of `bins` 64-bit counts (bit-exact incl. >= 2^31 px) */
int yam_otsu_from_hist(const uint64_t* hist, int bins, int* out_threshold);
/* the same for n histograms (hists[n][bins], host memory) on the host worker pool */
int yam_otsu_from_hists(const uint64_t* hists, int bins, int64_t n, int32_t* out_thresholds);

/* ---- K1 colour -> gray: cv2.cvtColor(BGR2GRAY) --------------------------------------------
 * replaces modules/preprocessing.py:54, core/preprocessing.py:56, core/segmentation.py:48,
 * core/extraction.py:47.   src (n,h,w,3) interleaved BGR -> dst (n,h,w); dtype U8|U16|F32. */
int yam_bgr2gray(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w, int dtype);

/* ---- K2 elementwise --------------------------------------------------------------------------
 * yam_minmax: per-frame min and max -> out_dev[2*n] doubles on device (and host if out_host). */
int yam_minmax(yam_ctx* ctx, const void* src, int64_t n, int64_t h, int64_t w, int dtype,
               double* out_dev, double* out_host);
/* cv2.normalize(src,None,alpha,beta,NORM_MINMAX) — modules/preprocessing.py:123. dtype kept. */
int yam_normalize_minmax(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w,
                         int dtype, double alpha, double beta);
/* cv2.convertScaleAbs — modules/preprocessing.py:78, core/segmentation.py:52. dst is U8. */
int yam_convert_scale_abs(yam_ctx* ctx, const void* src, void* dst, int64_t count, int dtype,
                          double alpha, double beta);
/* cv2.LUT(u8, table[256]) — modules/preprocessing.py:102 (Gamma). table is a HOST pointer. */
int yam_lut_u8(yam_ctx* ctx, const void* src, void* dst, int64_t count, const uint8_t* table_host);
/* cv2.threshold(src,t,maxval,THRESH_BINARY) — core/segmentation.py:142, core/preprocessing.py */
int yam_threshold(yam_ctx* ctx, const void* src, void* dst, int64_t count, int dtype,
                  double thresh, double maxval);

/* ---- K3/K5 separable filters ---------------------------------------------------------------
 * cv2.GaussianBlur(src,(k,k),sigma) — modules/preprocessing.py:145,170; core/preprocessing.py:86.
 * U8/U16: cv2's fixed-point path, bit-exact. F32: cv2 4.13 vector-path summation order. */
int yam_gaussian(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w,
                 int dtype, int ksize, double sigma, int border);
/* cv2.blur(src,(k,k)) odd k, BORDER_REFLECT_101 (absent from the reference; north_star op) */
int yam_box(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w, int dtype,
            int ksize);
/* cv2.medianBlur(src,k) — modules/preprocessing.py:147. border REPLICATE; k in {3,5}, uint8 also odd k 7..15
 * (cv2 accepts ksize > 5 for CV_8U only; range ui/control_metadata.py:210-218) */
/* (uint8 also takes odd ksize 7..15, like cv2.medianBlur; uint16: 3 or 5) */
int yam_median(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w,
               int dtype, int ksize);

/* ---- K9 adaptive threshold -----------------------------------------------------------------
 * cv2.adaptiveThreshold(src,255,GAUSSIAN_C,THRESH_BINARY,block,C) — core/segmentation.py:93.
 * src U8 (reference semantics) or U16 (extension: same formula, mean saturated to u16); dst U8. */
int yam_adaptive_threshold(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h,
                           int64_t w, int dtype, int block_size, double C);

/* ---- K6 morphology -------------------------------------------------------------------------
 * cv2.erode/dilate/morphologyEx with getStructuringElement(shape,(k,k)), `iterations` —
 * core/segmentation.py:264-314, :102-103. border = identity element. dtype U8|U16. */
int yam_morph(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w, int dtype,
              int op, int shape, int ksize, int iterations);
/* open(k,it) followed by close(k,it), rectangular SE, fused in one pass (segmentation config) */
int yam_morph_open_close(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w,
                         int dtype, int ksize, int iterations);

/* ---- fused binary segmentation path ----------------------------------------------------------
 * A thresholded mask is binary, so adaptive threshold -> open -> close -> connected components can
 * run on 1-bit-per-pixel masks (row-major uint32 words, wpr = ceil(w/32) words per row, bit i of
 * word j = pixel 32*j+i, bits beyond the width are 0) without ever writing the 8-bit mask:
 * yam_adaptive_threshold_bits = yam_adaptive_threshold with packed output (block 3,5,7,11,15);
 * yam_bits_morph = cv2.erode/dilate/morphologyEx on a {0,255} mask, RECTANGULAR element;
 * yam_bits_unpack = packed bits -> uint8 {0,255}; yam_ccl_label_bits = yam_ccl_label on bits. */
int yam_adaptive_threshold_bits(yam_ctx* ctx, const void* src, uint32_t* bits_out, int64_t n, int64_t h,
                                int64_t w, int dtype, int block_size, double C);
/* The same, and in the same pass over src the GLOBAL threshold mask mask_out = src > thresh_dev[frame] ? maxval : 0
 * (source dtype; = yam_threshold_frames): a pipeline that needs both the Otsu mask (core/segmentation.py:147) and the
 * adaptive segmentation (:91-94) of one image reads it once.  thresh_dev: int32[n] on the device. */
int yam_adaptive_threshold_bits_mask(yam_ctx* ctx, const void* src, uint32_t* bits_out, int64_t n, int64_t h,
                                     int64_t w, int dtype, int block_size, double C, const int32_t* thresh_dev,
                                     void* mask_out, double maxval);
int yam_bits_morph(yam_ctx* ctx, const uint32_t* bits_in, uint32_t* bits_out, int64_t n, int64_t h,
                   int64_t w, int op, int ksize, int iterations);
int yam_bits_unpack(yam_ctx* ctx, const uint32_t* bits, void* mask_u8, int64_t n, int64_t h, int64_t w);
int yam_ccl_label_bits(yam_ctx* ctx, const uint32_t* bits, int32_t* labels, int64_t n, int64_t h,
                       int64_t w, int32_t* counts_dev, int32_t* counts_host);

/* ---- K7/K8 histogram family ----------------------------------------------------------------
 * per-frame histogram: hist_dev[n][bins] of uint64, bins = 256 (U8) | 65536 (U16) */
int yam_histogram(yam_ctx* ctx, const void* src, int64_t n, int64_t h, int64_t w, int dtype,
                  uint64_t* hist_dev);
/* Otsu threshold per frame from src (cv2.threshold(...THRESH_OTSU) — core/segmentation.py:147,
 * core/extraction.py:59,72): thresholds to thresh_dev[n] (int32, device) and thresh_host[n]
 * (optional), then dst = src > t ? maxval : 0 in the source dtype if dst != NULL. */
/* The scan runs on the device and nothing is read back: histogram -> certified parallel scan (one
 * 8-CTA cluster per frame proves the arg-max from exact integer prefix sums and a forward error bound of
 * cv2's fp64 recurrence) -> exact sequential chain kernels only for frames that could not be certified.
 * yam_otsu_from_hist_dev: the scan alone, for histograms that are already on the device (all-reduced
 * across ranks: core/segmentation.py:147 on a mosaic); hist_dev[n][bins] uint64, bins = 256 | 65536.
 * certified_dev (optional, int32[n]): 1 where the certificate decided the frame, 0 where the chain ran.
 * yam_otsu_set_force_chain(1): every frame goes through the exact chain (tests, comparisons); returns the old value.
 * yam_threshold_frames: dst = src > thresh_dev[frame] ? maxval : 0 per frame (thresholds int32 on the device). */
int yam_otsu_from_hist_dev(yam_ctx* ctx, const uint64_t* hist_dev, int bins, int64_t n, int32_t* thresh_dev,
                           int32_t* certified_dev);
int yam_otsu_set_force_chain(int on);
int yam_threshold_frames(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w, int dtype,
                         const int32_t* thresh_dev, double maxval);
int yam_otsu_threshold(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w,
                       int dtype, double maxval, int32_t* thresh_dev, int32_t* thresh_host);
/* cv2.equalizeHist — core/preprocessing.py:76. U8 only, like the reference. */
int yam_equalize_hist(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w);
/* cv2.createCLAHE(clip,(tiles_x,tiles_y)).apply (absent from the reference; north_star op) */
int yam_clahe(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w, int dtype,
              double clip_limit, int tiles_x, int tiles_y);

/* CLAHE split in two for row-strip sharding of a mosaic (each rank computes the LUTs of the tile
 * rows it owns, LUTs are all-gathered, every rank applies them to its rows with GLOBAL geometry):
 * yam_clahe_luts  -> luts_dev[tiles_y][tiles_x][bins] (source dtype) for the single image `src`;
 * yam_clahe_apply -> `rows` x w pixels whose first row is global row y_offset of an image tiled
 *                    tiles_x x tiles_y with tiles of tile_w x tile_h pixels. */
int yam_clahe_luts(yam_ctx* ctx, const void* src, int64_t h, int64_t w, int dtype, double clip_limit,
                   int tiles_x, int tiles_y, void* luts_dev);
int yam_clahe_apply(yam_ctx* ctx, const void* src, void* dst, int64_t rows, int64_t w, int dtype,
                    const void* luts_dev, int tiles_x, int tiles_y, int tile_w, int tile_h,
                    int64_t y_offset);

/* ---- K10 connected components --------------------------------------------------------------
 * 8-connectivity, background 0, labels 1..N per frame numbered in raster order of each
 * component's first pixel (skimage.measure.label order — core/extraction.py:60,73; same partition
 * as cv2.connectedComponents — core/segmentation.py:108).  mask U8 (non-zero = foreground),
 * labels I32.  counts_dev[n] int32 on device (optional), counts_host[n] optional. */
int yam_ccl_label(yam_ctx* ctx, const void* mask, int32_t* labels, int64_t n, int64_t h, int64_t w,
                  int32_t* counts_dev, int32_t* counts_host);

/* The same labelling in two steps, for callers that renumber before the labels are written (the
 * row-strip mosaic: strip-local labels become global labels after the cross-strip merge, SURVEY.md
 * 8e) -- the label image is then written ONCE instead of written, re-read and rewritten:
 *   resolve   everything up to the per-segment roots and their labels, kept in a caller-owned
 *             workspace of yam_ccl_workspace_bytes() bytes (256-byte aligned device memory);
 *             counts_dev[n] receives the component counts;
 *   emit_rows writes the labels of rows [row_begin, row_end) (rows of the n*h stacked rows) to
 *             `labels` (pointing at the first pixel of row_begin); remap_dev (optional, single
 *             frame): label = remap_dev[local label], remap_dev[0] = 0. */
int64_t yam_ccl_workspace_bytes(yam_ctx* ctx, int64_t n, int64_t h, int64_t w);
int yam_ccl_resolve_bits(yam_ctx* ctx, const uint32_t* bits, int64_t n, int64_t h, int64_t w, void* workspace,
                         int32_t* counts_dev);
int yam_ccl_emit_rows(yam_ctx* ctx, const uint32_t* bits, int64_t n, int64_t h, int64_t w,
                      const void* workspace, const int32_t* remap_dev, int64_t row_begin, int64_t row_end,
                      int32_t* labels);
/* the same with the table length stated (remap_size > 0): local labels >= remap_size map to 0 instead of
 * reading past the table (tables sized from bounds, yam_merge_strips_remap_bounded) */
int yam_ccl_emit_rows_bounded(yam_ctx* ctx, const uint32_t* bits, int64_t n, int64_t h, int64_t w,
                              const void* workspace, const int32_t* remap_dev, int64_t remap_size,
                              int64_t row_begin, int64_t row_end, int32_t* labels);

/* labels[i] = remap_dev[labels[i]] for labels in (0, remap_size); used by the cross-strip label merge */
int yam_relabel(yam_ctx* ctx, int32_t* labels, int64_t count, const int32_t* remap_dev,
                int64_t remap_size);

/* Cross-strip label merge for row-strip sharded mosaics: strip r holds canonical
 * labels 1..count_r; global id = offsets[r] + local label.  edges_dev is [world][2][w] int32: the
 * first and last label row of every strip.  Unites the ids that touch across a strip boundary
 * (8-connectivity) and writes, for every global id in [0, total], the smallest id of its merged
 * set to root_dev[total + 1] (root_dev[0] = 0).  offsets_dev: world + 1 int64 on the device. */
int yam_merge_strip_labels(yam_ctx* ctx, const int32_t* edges_dev, const int64_t* offsets_dev, int world,
                           int64_t w, int64_t total, int32_t* root_dev);

/* The whole cross-strip step in one call (merge + raster-first renumbering, no host arithmetic on
 * label arrays).  packed_dev is [world][stride] int32 (stride >= 2w): per strip its first label row
 * (w values) followed by its last label row (w values) -- what an all-gather of every rank's two
 * boundary rows delivers; anything after 2w in a row (e.g. the strip's count) is ignored.
 * offsets_host[world + 1] (HOST memory) is the exclusive prefix of the per-strip component counts.
 * Writes, for each of the strips rank .. rank + rank_count - 1 (the strips this process owns), a table
 * of count_r + 1 entries back to back into remap_dev: remap[l] = global label of local label l
 * (remap[0] = 0; labels number merged components by their first pixel in raster order, the same
 * numbering a dense run produces), and total_dev[0] = number of merged components.  `workspace`:
 * yam_merge_strips_workspace_bytes(offsets_host[world]) bytes of 256-byte aligned device memory. */
int64_t yam_merge_strips_workspace_bytes(int64_t total);
int yam_merge_strips_remap(yam_ctx* ctx, const int32_t* packed_dev, int64_t stride, int world, int64_t w,
                           const int64_t* offsets_host, int rank, int rank_count, void* workspace,
                           int32_t* remap_dev, int32_t* total_dev);
/* The same without the host ever learning the counts: offsets_bound_host are exclusive prefixes of per-strip UPPER
 * BOUNDS (e.g. from the previous run over the same source), the real counts stay on the device
 * (counts_dev[r * counts_stride], r < world).  Ids between a count and its bound take no rank, so remap / total are
 * the same as with exact offsets; *overflow_dev = 1 when a count exceeds its bound (results are then unusable and the
 * caller repeats the merge with exact offsets).  Tables hold bound_r + 1 entries per strip. */
int yam_merge_strips_remap_bounded(yam_ctx* ctx, const int32_t* packed_dev, int64_t stride, int world, int64_t w,
                                   const int64_t* offsets_bound_host, const int32_t* counts_dev, int64_t counts_stride,
                                   int rank, int rank_count, void* workspace, int32_t* remap_dev, int32_t* total_dev,
                                   int32_t* overflow_dev);

/* Interleaved (px, channels) <-> planar (channels, px) copies, channels <= 4: cv2's neighbourhood
 * filters (GaussianBlur, medianBlur, blur, erode / dilate, modules/preprocessing.py:140-150) treat the
 * channels of a colour image independently, so colour input runs as a stack of planes. */
int yam_split_channels(yam_ctx* ctx, const void* src, void* dst, int64_t px, int channels, int dtype);
int yam_merge_channels(yam_ctx* ctx, const void* src, void* dst, int64_t px, int channels, int dtype);

/* ---- watershed front half (core/segmentation.py:97-111) and second moments ---------
 * yam_threshold_inv: cv2.threshold(..., THRESH_BINARY_INV [+ OTSU]) (:100): dst = src > t ? 0 : maxval per frame;
 *   t = thresh_dev[frame] (device int32, e.g. from yam_otsu_threshold) when thresh_dev != NULL, else `thresh`.
 * yam_distance_transform: cv2.distanceTransform(mask, DIST_L2, 5) (:104): float32 chamfer distance
 *   (a = 1, b = 1.4, c = 2.1969) of every non-zero pixel to the nearest zero pixel, computed as the least
 *   fixed point of the chamfer relaxation; within 1e-5 relative of cv2's sequential two-pass scan (equal
 *   except for isolated one-ulp differences).  launches_out (optional, host) = relaxation launches used.
 * yam_watershed_combine: markers = labels + 1; markers[sure_bg == 255 and sure_fg == 0] = 0 (:109-110).
 * yam_region_moments: per label (1..n_labels) sum r^2, sum c^2, sum r*c as int64 [n_labels][3]: with the
 *   first-order sums of yam_region_props these give skimage's central moments / inertia tensor
 *   (orientation, eccentricity; core/extraction.py:81-85). */
int yam_threshold_inv(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w, int dtype,
                      const int32_t* thresh_dev, double thresh, double maxval);
int yam_distance_transform(yam_ctx* ctx, const uint8_t* mask, float* dist, int64_t n, int64_t h, int64_t w,
                           int* launches_out);
int yam_watershed_combine(yam_ctx* ctx, const int32_t* labels, const uint8_t* sure_bg, const uint8_t* sure_fg,
                          int32_t* markers, int64_t count);
int yam_region_moments(yam_ctx* ctx, const int32_t* labels, int64_t h, int64_t w, int64_t n_labels, int64_t* moments_dev);

/* ---- histogram equalisation of colour images (Preprocessor.histogram_equalization, core/preprocessing.py:74-79:
 * cvtColor(BGR2YCrCb) -> equalizeHist(Y) -> cvtColor(YCrCb2BGR), uint8) -------------------------------------
 * yam_bgr_luma_ycrcb: y[i] = the Y channel cv2.cvtColor(COLOR_BGR2YCrCb) gives pixel i of an interleaved
 *   BGR uint8 image (px pixels, any number of frames back to back); feed it to yam_equalize_hist.
 * yam_bgr_replace_luma_ycrcb: dst = cvtColor(YCrCb2BGR) of (y_new, Cr, Cb) with Cr / Cb of the source pixel
 *   as BGR2YCrCb stores them (saturated uint8); dst may alias bgr.  cv2's 14-bit fixed point, bit-exact. */
int yam_bgr_luma_ycrcb(yam_ctx* ctx, const uint8_t* bgr, uint8_t* y, int64_t px);
int yam_bgr_replace_luma_ycrcb(yam_ctx* ctx, const uint8_t* bgr, const uint8_t* y_new, uint8_t* dst, int64_t px);

/* ---- perimeter and solidity of labelled regions (core/extraction.py:80,83: prop.perimeter, prop.solidity;
 * skimage.measure.regionprops) ----------------------------------------------------------------------------
 * yam_region_perimeter: skimage.measure.perimeter(region image, neighborhood=4) decomposed into exact
 *   integer counts: counts_dev[n_labels][3] (int64) = border pixels of label i+1 whose class
 *   v = 1 + 2 * (border 4-neighbours of the same label) + 10 * (border diagonal neighbours) weighs
 *   1 ({5,7,15,17,25,27}), sqrt 2 ({21,33}) and (1 + sqrt 2) / 2 ({13,23}); a border pixel has a
 *   4-neighbour outside its region (the image edge counts as outside).  perimeter = the weighted sum.
 * yam_region_convex_area: RegionProperties.area_convex = pixels whose centre lies in the closed convex
 *   hull of the edge midpoints of the region's pixels (skimage.morphology.convex_hull_image with
 *   offset_coordinates=True, include_borders=True), exact integer arithmetic; convex_area_dev[n_labels]
 *   (int64), 0 for labels without pixels.  props_dev = the table yam_region_props wrote for the SAME
 *   label image (area and bounding rows are read from it).  solidity = area / area_convex.
 *   Synchronises the stream once (the per-region row tables are sized on the device). */
int yam_region_perimeter(yam_ctx* ctx, const int32_t* labels, int64_t h, int64_t w, int64_t n_labels, int64_t* counts_dev);
int yam_region_convex_area(yam_ctx* ctx, const int32_t* labels, int64_t h, int64_t w, int64_t n_labels, const int64_t* props_dev,
                           int64_t* convex_area_dev);

/* Order-independent 64-bit content checksum: ADDS to *sum_dev (device, caller zeroes it) the sum over
 * i < count of mix64((index_base + i) * 0x9E3779B97F4A7C15 + value_i) mod 2^64 (mix64 = splitmix64's
 * finalizer; values zero-extended from U8 | U16 | I32 bit patterns).  A row strip passes the linear
 * index of its first pixel as index_base, so the per-strip sums of a sharded run add up (all-reduce)
 * to the checksum of the dense image -- bench.py's `check` block and the mosaic parity tests. */
int yam_checksum64(yam_ctx* ctx, const void* src, int64_t count, int dtype, int64_t index_base, uint64_t* sum_dev);

/* ---- K11 region properties -----------------------------------------------------------------
 * skimage.measure.regionprops restated (core/extraction.py:61,74): for a single labelled frame
 * with labels 1..n_labels writes, per label (row i = label i+1), 8 int64 accumulators to
 * props_dev[n_labels][8]: area, sum_row, sum_col, sum_intensity, min_row, min_col, max_row+1,
 * max_col+1.  intensity may be NULL (sum_intensity = 0); intensity dtype U8|U16. */
#define YAM_PROPS_STRIDE 8
int yam_region_props(yam_ctx* ctx, const int32_t* labels, const void* intensity, int intensity_dtype,
                     int64_t h, int64_t w, int64_t n_labels, int64_t* props_dev);

/* The same for a stack of n labelled frames (time-lapse batches, processing/pipeline_manager.py
 * _apply_slice_wise:475-492 treats planes independently): frame f holds labels 1..count_f and its
 * rows start at props_dev[offsets_dev[f]]; offsets_dev = exclusive prefix of the per-frame counts
 * (n + 1 entries, device memory), total = offsets[n]. */
int yam_region_props_stack(yam_ctx* ctx, const int32_t* labels, const void* intensity, int intensity_dtype,
                           int64_t n, int64_t h, int64_t w, const int64_t* offsets_dev, int64_t total,
                           int64_t* props_dev);

/* ---- remaining per-pixel menu steps -----------------------------------------
 * cv2.addWeighted (SharpenModule, modules/preprocessing.py:167-171): scalars as float32,
 * dst = saturate(rint(fmaf(a, alpha, fmaf(b, beta, gamma)))); U8 | U16, dtype preserved. */
int yam_add_weighted(yam_ctx* ctx, const void* a, const void* b, void* dst, int64_t count, int dtype,
                     double alpha, double beta, double gamma);

/* SelectChannelModule (modules/preprocessing.py:188-209) on interleaved BGR (px pixels):
 * B / G / R pick (U8 | U16) or the truncated mean of two channels (U8 only). */
#define YAM_CHANNEL_B 0
#define YAM_CHANNEL_G 1
#define YAM_CHANNEL_R 2
#define YAM_CHANNEL_RG 3
#define YAM_CHANNEL_GB 4
#define YAM_CHANNEL_BR 5
int yam_select_channel(yam_ctx* ctx, const void* bgr, void* dst, int64_t px, int dtype, int mode);
/* cv2.cvtColor(GRAY2BGR): channel "All" on a single-channel image (modules/preprocessing.py:192-193) */
int yam_gray2bgr(yam_ctx* ctx, const void* gray, void* bgr, int64_t px, int dtype);

/* remove_border_regions (core/segmentation.py:316-325): elements closer than border_distance to an
 * image edge become 0; like the reference's [d:-d] slice, d == 0 or 2d >= side clears everything. */
int yam_border_clear(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w, int channels,
                     int dtype, int border_distance);

/* sobel_operator / prewitt_operator / laplacian_operator (core/segmentation.py:150-169): gray U8 |
 * U16 in, uint8 magnitude out (clip to 255, truncate), BORDER_REFLECT_101; ksize 1, 3, 5 or 7
 * (exact integer range; Prewitt is 3x3). */
#define YAM_EDGE_SOBEL 0
#define YAM_EDGE_PREWITT 1
#define YAM_EDGE_LAPLACIAN 2
int yam_edge_filter(yam_ctx* ctx, const void* src, void* dst_u8, int64_t n, int64_t h, int64_t w, int dtype,
                    int kind, int ksize);

/* Raw material of cv2.moments(mask) (hu_moments_data, core/extraction.py:100-105): for every row y
 * of a mask (non-zero = set) out_dev[y][4] = (count, sum x, sum x^2, sum x^3) over the set pixels,
 * exact int64.  m_pq = 255 * sum_y y^q * out[y][p] is formed on the host in exact arithmetic. */
int yam_mask_row_moments(yam_ctx* ctx, const void* mask, int64_t h, int64_t w, int dtype, int64_t* out_dev);

#ifdef __cplusplus
}
#endif
#endif /* YAMB200_H */
