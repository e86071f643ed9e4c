// Host-side parameter preparation shared by the .cu translation units.
#pragma once
#include <stddef.h>
#include <stdint.h>

#define YAM_MAX_TAPS 127  // adaptive threshold block size goes to 101 (ui/control_metadata.py:301-309)
#define YAM_MAX_SE 31     // morphology kernel_size 1..31 (ui/control_metadata.py:557-676)

void yam_host_gaussian_taps(int k, double sigma, double* out);
void yam_host_fixed_taps(const double* kf, int k, int bits, int64_t* out);
void yam_host_structuring_element(int shape, int k, uint8_t* out);
int yam_host_otsu(const uint64_t* h, int bins);
int yam_host_otsu32(const uint32_t* h, int bins);  // same recurrence on 32-bit counts (frames < 2^32 px)
// 1..4 histograms (stride_bytes apart; 32-bit counts if narrow) advanced in lock step by the calling
// thread: independent dependency chains fill the pipeline one chain leaves idle
void yam_host_otsu_group(const void* hists, size_t stride_bytes, int count, int bins, int narrow, int32_t* out);

// TMA fast path of the adaptive threshold -> packed bits (yam_adaptive.cu); *handled = 0 when the shape
// does not qualify and the caller must use the generic tiled kernel
struct yam_ctx;
int yam_adaptive_bits_tma(yam_ctx* ctx, const void* src, uint32_t* bits, int64_t n, int64_t h, int64_t w, int dtype,
                          int block_size, const float* taps_f, int idelta, int* handled, const int32_t* t_dev, void* mask_out,
                          double maxval);
// TMA fast path of the 16-bit fixed-point Gaussian (yam_gauss_tma.cu); *handled = 0 -> use sep_fixed_tiled
int yam_gauss16_tma(yam_ctx* ctx, const void* src, void* dst, int64_t n, int64_t h, int64_t w, int ksize,
                    const uint32_t* taps_q, int border, int* handled);
