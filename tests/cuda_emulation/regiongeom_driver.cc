// Driver for the host emulation of csrc/yam_regiongeom.cu's kernels (see emu_prelude.h): follows the launch
// sequence of yam_region_perimeter / yam_region_convex_area with small grids, so the grid-stride loops run.
// usage: regiongeom_emu <in.bin> <out.bin>;  in = int64 h, w, n; int32 labels[h*w]; int64 props[n*8]
//                                           out = int64 counts[n*3]; int64 convex_area[n]
int main(int argc, char** argv) {
    if (argc != 3) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 3;
    int64_t hdr[3];
    if (fread(hdr, 8, 3, f) != 3) return 4;
    const int64_t h = hdr[0], w = hdr[1], n = hdr[2];
    std::vector<int32_t> labels((size_t)(h * w));
    std::vector<int64_t> props((size_t)(n * 8));
    if (fread(labels.data(), 4, labels.size(), f) != labels.size()) return 4;
    if (fread(props.data(), 8, props.size(), f) != props.size()) return 4;
    fclose(f);
    std::vector<int64_t> counts((size_t)(n * 3), 0), convex((size_t)n, 0), off((size_t)(n + 1), -1);

    const int64_t tiles_x = (w + kPT - 1) / kPT, tiles_y = (h + kPT - 1) / kPT, tiles = tiles_x * tiles_y;
    emu_launch((unsigned)(tiles < 3 ? tiles : 3), 256, [&]() {
        region_perimeter_kernel(labels.data(), (int)h, (int)w, n, (int)tiles_x, tiles, (unsigned long long*)counts.data());
    });

    emu_launch(1, 1024, [&]() { hull_offsets_kernel(props.data(), n, off.data()); });
    const int64_t total = off[(size_t)n];
    std::vector<int> v((size_t)(4 * total + 4));
    memset(v.data(), 0x7f, (size_t)total * 2 * sizeof(int));
    int* vl = v.data();
    int* vr = vl + total;
    int* sl = vr + total;
    int* sr = sl + total;
    if (total > 0) {
        emu_launch(2, 256, [&]() { hull_fill_kernel(labels.data(), (int)h, (int)w, n, props.data(), off.data(), vl, vr); });
        emu_launch((unsigned)((2 * n + 127) / 128), 128,
                   [&]() { hull_chain_kernel(off.data(), n, vl, vr, sl, sr, (unsigned long long*)convex.data()); });
    }
    f = fopen(argv[2], "wb");
    if (!f) return 5;
    fwrite(counts.data(), 8, counts.size(), f);
    fwrite(convex.data(), 8, convex.size(), f);
    fclose(f);
    return 0;
}
