#!/usr/bin/env python
"""CUDA-event timing of every operator at a given shape: python tools/time_ops.py H W [reps]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from yamimageprocessor_b200 import synth
from yamimageprocessor_b200.backend import get_backend
H, W = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
be = get_backend(0)
tile = min(4096, H, W)
t = synth.nuclei(tile, tile, seed=100)
frame = np.tile(t, (H // tile, W // tile))
x = be.to_device(frame)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=be.device)
def timeit(name, fn, bpp):
    fn(); fn()
    ms = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
    m = min(ms)
    print(f"{name:28s} {m:9.3f} ms  {H*W*bpp/m/1e6:8.0f} GB/s  {H*W*bpp/m/1e6/6553:6.3f}")
    return out
g = timeit("gaussian k11", lambda: be.gaussian(x, 11, 0.0), 4)
luts = timeit("clahe_luts", lambda: be.clahe_luts(g, 2.0, (8, max(1, H // (W // 8)))), 2)
c = timeit("clahe (full)", lambda: be.clahe(g, 2.0, (8, max(1, H // (W // 8)))), 6)
timeit("histogram", lambda: be.histogram(c), 2)
timeit("otsu_threshold", lambda: be.otsu_threshold(c, 255), 6)
timeit("threshold", lambda: be.threshold(c, 20000.0, 255), 4)
m = timeit("adaptive", lambda: be.adaptive_threshold(c, 11, 2), 3)
m2 = timeit("open_close", lambda: be.morph_open_close(m, 5, 1), 4)
lab, cnt = timeit("ccl", lambda: be.ccl_label(m2), 5)
n = int(be.to_host(cnt)[0])
timeit("props", lambda: be.region_props(lab, c, n), 6)
timeit("median5", lambda: be.median(x, 5), 4)
timeit("normalize", lambda: be.normalize_minmax(x, 0, 255), 6)
print("components", n)
