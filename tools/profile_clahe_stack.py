import sys
sys.path.insert(0, "/root/repo")
import numpy as np
from yamimageprocessor_b200 import synth
from yamimageprocessor_b200.backend import get_backend
be = get_backend(0)
x = be.to_device(np.stack([synth.nuclei(2048, 2048, seed=1000 + i) for i in range(32)]))
g = be.gaussian(x, 11, 0.0)
for _ in range(2):
    c = be.clahe(g, 2.0, (8, 8))
be.synchronize()
x1 = be.to_device(synth.nuclei(4096, 4096, seed=1))
g1 = be.gaussian(x1, 11, 0.0)
for _ in range(2):
    c1 = be.clahe(g1, 2.0, (8, 8))
be.synchronize()
print("done")
