"""Otsu scan on SPARSE histograms (12-bit data in a 16-bit container: every 16th bin): the certificate cannot break the
exact ties between empty bins, so these frames take the exact chain kernels; their empty-bin runs are copied."""
import os, sys, statistics
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yamimageprocessor_b200 import synth
from yamimageprocessor_b200.backend import get_backend
from oracle import np_oracle as O
be = get_backend(0)
frames = np.stack([(synth.nuclei(2048, 2048, seed=1000 + i) >> 4) << 4 for i in range(32)]).astype(np.uint16)
x = be.to_device(frames)
h = be.histogram(x)
t, cert = be.otsu_from_histogram_device(h, want_certified=True)
want = [O.otsu_from_hist(hh) for hh in h.cpu().numpy()[:4]]
print("certified", int(cert.sum().item()), "of", cert.numel(), "| thresholds", t.cpu().tolist()[:4], "oracle", want)
def tm(name, fn, reps=5):
    fn(); evs = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    print(f"{name:48s} {statistics.mean(p.elapsed_time(q) for p, q in evs):8.3f} ms")
tm("scan, 32 sparse histograms (chain fall-back)", lambda: be.otsu_from_histogram_device(h))
tm("scan, 1 sparse histogram", lambda: be.otsu_from_histogram_device(h[:1]))
