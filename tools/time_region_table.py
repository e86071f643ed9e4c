#!/usr/bin/env python
"""CUDA-event timing of the region-table kernels (sums, second moments, perimeter classes, convex area) on the
c3 frame (8192^2, ~100 k nuclei) and of the colour equalisation at 4096^2 x 3:
    python tools/time_region_table.py [out.json]"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from yamimageprocessor_b200 import synth
from yamimageprocessor_b200.backend import get_backend

be = get_backend(0)
H = W = 8192
tile = synth.nuclei(4096, 4096, seed=2)
frame = np.tile(tile, (2, 2))
x = be.to_device(frame)
labels, counts = be.segment_fused(x, 11, 2, 5, 1)
n = int(be.to_host(counts)[0])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=be.device)
out = {"shape": [H, W], "regions": n, "hbm_peak_GBps": 6553.0, "ops": []}


def timeit(name, fn, bytes_per_px, px=H * W, reps=5):
    fn()
    fn()
    ms = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    m = float(np.median(ms))
    gbps = px * bytes_per_px / m / 1e6
    out["ops"].append({"op": name, "ms": round(m, 4), "bytes_per_px": bytes_per_px, "GBps": round(gbps, 1), "frac": round(gbps / 6553.0, 3)})
    print(f"{name:44s} {m:8.3f} ms {gbps:8.0f} GB/s {gbps / 6553.0:6.3f}")
    return r


props = timeit("region_props (sums, bbox, intensity)", lambda: be.region_props(labels, x, n), 6)
timeit("region_moments (second order)", lambda: be.region_moments(labels, n), 4)
timeit("region_perimeter (border classes)", lambda: be.region_perimeter_counts(labels, n), 4)
timeit("region_convex_area (hull chains, 1 sync)", lambda: be.region_convex_area(labels, n, props), 4)
bgr = be.to_device(np.random.default_rng(0).integers(0, 256, (4096, 4096, 3), dtype=np.uint8))
timeit("equalize_hist_bgr 4096^2 (3 launches + LUT)", lambda: be.equalize_hist_bgr(bgr), 9, px=4096 * 4096)
if len(sys.argv) > 1:
    Path(sys.argv[1]).write_text(json.dumps(out, indent=1))
