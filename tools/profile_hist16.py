"""hist16 at 4096^2 (c1's Otsu histogram) for ncu source-level profiling."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yamimageprocessor_b200 import synth
from yamimageprocessor_b200.backend import get_backend
be = get_backend(0)
x1 = be.to_device(synth.nuclei(4096, 4096, seed=1000))
c1 = be.clahe(be.gaussian(x1, 11, 0.0), 2.0, (8, 8))
for _ in range(2):
    h = be.histogram(c1)
be.synchronize()
print("done")
