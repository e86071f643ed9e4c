"""One launch set of the c1 Otsu operator (hist16 -> certify -> chain/sigma early exits -> threshold) and of
CLAHE at 4096^2 and on a 32 x 2048^2 stack, for `ncu --set full`."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yamimageprocessor_b200 import synth
from yamimageprocessor_b200.backend import get_backend
be = get_backend(0)
x1 = be.to_device(synth.nuclei(4096, 4096, seed=1000))
c1 = be.clahe(be.gaussian(x1, 11, 0.0), 2.0, (8, 8))
t, m = be.otsu_threshold(c1, 255)
xs = be.to_device(np.stack([synth.nuclei(2048, 2048, seed=1000 + i) for i in range(32)]))
gs = be.gaussian(xs, 11, 0.0)
cs = be.clahe(gs, 2.0, (8, 8))
ts, ms = be.otsu_threshold(cs, 255)
be.synchronize()
print("done", int(t[0].item()), ts.cpu().tolist()[:4])
