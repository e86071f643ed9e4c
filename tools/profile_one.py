#!/usr/bin/env python
"""ncu driver for one operator: python tools/profile_one.py adaptive|gauss|morph|clahe|otsu|ccl|props [size]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from yamimageprocessor_b200 import synth
from yamimageprocessor_b200.backend import get_backend
op = sys.argv[1]
size = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
be = get_backend(0)
x = be.to_device(synth.nuclei(size, size, seed=1000))
for _ in range(3):
    if op == "adaptive":
        y = be.adaptive_threshold(x, 11, 2)
    elif op == "gauss":
        y = be.gaussian(x, 11, 0.0)
    elif op == "morph":
        m = be.adaptive_threshold(x, 11, 2); y = be.morph_open_close(m, 5, 1)
    elif op == "clahe":
        y = be.clahe(x, 2.0, (8, 8))
    elif op == "otsu":
        y = be.otsu_threshold(x, 255)
    elif op == "fused":
        y = be.segment_fused(x, 11, 2, 5, 1)
    elif op == "ccl":
        m = be.morph_open_close(be.adaptive_threshold(x, 11, 2), 5, 1); y = be.ccl_label(m)
be.synchronize()
print("done")
