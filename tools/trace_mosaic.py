#!/usr/bin/env python
"""Phase timing of the sharded mosaic pipeline on one GPU: python tools/trace_mosaic.py [size] [strips]"""
import sys, time, threading
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from yamimageprocessor_b200 import synth
from yamimageprocessor_b200.backend import get_backend
from yamimageprocessor_b200.host import mosaic
size = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
strips = int(sys.argv[2]) if len(sys.argv) > 2 else 8
be = get_backend(0)
tile = synth.nuclei(4096, 4096, seed=100)
marks, lock = [], threading.Lock()
def trace(name, rank):
    torch.cuda.synchronize()
    with lock:
        marks.append((time.perf_counter(), rank, name))
p = mosaic.MosaicParams(trace=trace)
class Src:
    shape = (size, size)
dev = []
for r in range(strips):
    r0, r1 = mosaic.input_rows(size, r, strips, p)
    rows = np.tile(np.roll(tile, -(r0 % 4096), axis=0), ((r1 - r0 + 4095) // 4096 + 1, size // 4096))[: r1 - r0]
    dev.append(be.to_device(np.ascontiguousarray(rows)))
for it in range(3):
    marks.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = mosaic.run_local_strips(be, Src, strips, False, p, device_sources=dev)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    del res
print(f"total {1e3 * (t1 - t0):.1f} ms for {strips} strips of {size}^2")
last = t0
for t, rank, name in sorted(marks):
    print(f"  +{1e3 * (t - last):7.2f} ms  strip {rank} {name}")
    last = t
