#!/usr/bin/env python
"""ncu driver: Otsu on a stack (staged device scan): python tools/profile_otsu_stack.py [frames] [size]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from yamimageprocessor_b200 import synth
from yamimageprocessor_b200.backend import get_backend
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
size = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
be = get_backend(0)
x = be.to_device(np.stack([synth.nuclei(size, size, seed=1000 + i) for i in range(n)]))
c = be.clahe(be.gaussian(x, 11, 0.0), 2.0, (8, 8))
for _ in range(2):
    t, m = be.otsu_threshold(c, 255)
be.synchronize()
print("thresholds", be.to_host(t)[:8].tolist())
