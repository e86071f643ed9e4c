#!/usr/bin/env python
"""CPU cross-check of a c4 bench line's `check` block against the reference's own library calls.

Rebuilds the same synthetic mosaic, runs the reference's cv2 call sites over it (oracle/cv2_path.py:
GaussianBlur -> CLAHE -> Otsu; adaptive threshold on the CLAHE output -> open -> close ->
cv2.connectedComponents), renumbers the components raster-first and compares Otsu t, the component count
and the two 64-bit content checksums (oracle/np_oracle.py: checksum64, the NumPy statement of
yam_checksum64) with the values the GPU run printed.  Test infrastructure; needs cv2.

    python tools/check_bench_check_block.py <bench_line.json> [--chunk-rows R] [--out result.json]

Everything that is local (adaptive threshold, masks, checksums, renumbering) runs in row chunks so that the
full 65536^2 mosaic fits in ~45 GB of RAM; the global steps (Gaussian, CLAHE, morphology, labelling) are single
cv2 calls on the whole image, exactly the calls the reference makes.  The 16-bit Otsu threshold comes from the
oracle's restatement on 64-bit counts (cv2's own getThreshVal overflows at >= 2^31 pixels);
below that size it is also compared with cv2.threshold(OTSU)."""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402

import bench  # noqa: E402
from oracle import cv2_path as R  # noqa: E402
from oracle import np_oracle as O  # noqa: E402

cv2 = R.cv2
M64 = (1 << 64) - 1


def chunks(n, step):
    return [(a, min(n, a + step)) for a in range(0, n, step)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("line")
    ap.add_argument("--chunk-rows", type=int, default=4096)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    line = json.loads(Path(args.line).read_text())
    size = int(line["config"]["mosaic"][0])
    W = size
    tile = min(4096, size // 8)
    t0 = time.time()
    log = lambda what: print(f"[{time.time() - t0:7.1f} s] {what}", file=sys.stderr, flush=True)  # noqa: E731

    frame = bench.mosaic_rows(size, 0, size, tile, bench.mosaic_tiles(tile))
    log("mosaic built")
    g = R.noise_reduction_gaussian(frame, 11)
    del frame
    g = R.clahe(g, 2.0, (8, 8))
    log("Gaussian + CLAHE (cv2)")
    hist = np.zeros(65536, np.int64)
    for a, b in chunks(size, args.chunk_rows):
        hist += np.bincount(g[a:b].ravel(), minlength=65536)
    t = O.otsu_from_hist(hist)
    if g.size < (1 << 31):
        assert t == int(cv2.threshold(g, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)[0]), "oracle Otsu != cv2 Otsu"
    log(f"Otsu t = {t}")

    # Otsu mask checksum and the adaptive threshold (uint16 extension of cv2.adaptiveThreshold, oracle/cv2_path.py:63-72)
    mask_sum = 0
    seg = np.empty((size, W), np.uint8)
    halo = 5
    for a, b in chunks(size, args.chunk_rows):
        mask_sum = (mask_sum + O.checksum64(O.threshold_binary(g[a:b], t, 255), index_base=a * W)) & M64
        a0, b0 = max(0, a - halo), min(size, b + halo)
        part = R.adaptive_threshold(g[a0:b0], 11, 2)       # BORDER_REPLICATE only acts at the true image border
        seg[a:b] = part[a - a0: a - a0 + (b - a)]
    del g
    log("Otsu mask checksum, adaptive threshold")
    seg = R.morphological_closing(R.morphological_opening(seg, "Rectangular", 5, 1), "Rectangular", 5, 1)
    log("open / close (cv2)")
    n_plus_1, lab = cv2.connectedComponents(seg)
    del seg
    n = int(n_plus_1) - 1
    log(f"cv2.connectedComponents: {n} components")

    # raster-first renumbering: first linear index of every cv2 label (numpy keeps the LAST value written to a
    # repeated index, so writing the reversed chunk leaves the smallest index), rank of the first indices
    first = np.full(n + 1, -1, np.int64)
    for a, b in chunks(size, args.chunk_rows):
        flat = lab[a:b].ravel()
        nz = np.flatnonzero(flat)
        vals = flat[nz]
        tmp = np.full(n + 1, -1, np.int64)
        tmp[vals[::-1]] = nz[::-1] + a * W
        new = (first < 0) & (tmp >= 0)
        first[new] = tmp[new]
    assert (first[1:] >= 0).all()
    remap = np.zeros(n + 1, np.int32)
    remap[np.argsort(first[1:], kind="stable") + 1] = np.arange(1, n + 1, dtype=np.int32)
    lab_sum = 0
    for a, b in chunks(size, args.chunk_rows):
        lab_sum = (lab_sum + O.checksum64(remap[lab[a:b]], index_base=a * W)) & M64
    log("labels renumbered raster-first, checksum")

    got = {"otsu_threshold": int(t), "components": n, "labels_checksum64": f"{lab_sum:016x}", "otsu_mask_checksum64": f"{mask_sum:016x}"}
    want = {k: line["check"][k] for k in got}
    result = {"mosaic": [size, size], f"cpu (cv2 {cv2.__version__} call sites, {cv2.getNumThreads()} threads)": got,
              "gpu bench line": want, "gpu line": {"n_gpus": line.get("n_gpus"), "ms_per_step": line.get("ms_per_step")},
              "equal": got == want, "cpu_seconds": round(time.time() - t0, 1)}
    print(json.dumps(result, indent=1))
    if args.out:
        Path(args.out).write_text(json.dumps(result, indent=1))
    sys.exit(0 if got == want else 1)


if __name__ == "__main__":
    main()
