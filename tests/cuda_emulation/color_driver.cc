// Driver for the host emulation of csrc/yam_color.cu's kernels (see emu_prelude.h).
// usage: color_emu <in.bin> <out.bin> <use_quads 0|1>;  in = int64 px; uint8 bgr[3 px]; uint8 y_new[px]
//                                                      out = uint8 y[px]; uint8 dst[3 px]
int main(int argc, char** argv) {
    if (argc != 4) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 3;
    int64_t px = 0;
    if (fread(&px, 8, 1, f) != 1) return 4;
    std::vector<uint32_t> src_w((size_t)(3 * px + 3) / 4 + 1), ynew_w((size_t)(px + 3) / 4 + 1), y_w((size_t)(px + 3) / 4 + 1),
        dst_w((size_t)(3 * px + 3) / 4 + 1);
    uint8_t* src = (uint8_t*)src_w.data();
    uint8_t* y_new = (uint8_t*)ynew_w.data();
    if (fread(src, 1, (size_t)(3 * px), f) != (size_t)(3 * px)) return 4;
    if (fread(y_new, 1, (size_t)px, f) != (size_t)px) return 4;
    fclose(f);
    const int64_t quads = atoi(argv[3]) ? px / 4 : 0;
    emu_launch(3, 256, [&]() { bgr_luma_kernel(src, (uint8_t*)y_w.data(), px, quads); });
    emu_launch(3, 256, [&]() { bgr_replace_luma_kernel(src, y_new, (uint8_t*)dst_w.data(), px, quads); });
    f = fopen(argv[2], "wb");
    if (!f) return 5;
    fwrite(y_w.data(), 1, (size_t)px, f);
    fwrite(dst_w.data(), 1, (size_t)(3 * px), f);
    fclose(f);
    return 0;
}
