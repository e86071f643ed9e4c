#!/usr/bin/env python
"""Small driver for ncu: runs one workload's device-resident ops a few times.
usage: python tools/profile_ops.py c2 [reps]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from yamimageprocessor_b200 import synth  # noqa: E402
from yamimageprocessor_b200.backend import get_backend  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
be = get_backend(0)
if wl == "c2":
    x = be.to_device(synth.nuclei(8192, 8192, seed=1000))
    for _ in range(reps):
        m = be.adaptive_threshold(x, 11, 2)
        m2 = be.morph_open_close(m, 5, 1)
        labels, counts = be.ccl_label(m2)
        props = be.region_props(labels, x, int(be.to_host(counts)[0]))
elif wl == "c1":
    x = be.to_device(synth.nuclei(4096, 4096, seed=1000))
    for _ in range(reps):
        g = be.gaussian(x, 11, 0.0)
        c = be.clahe(g, 2.0, (8, 8))
        t, m = be.otsu_threshold(c, 255)
be.synchronize()
print("done", be.launch_count())
