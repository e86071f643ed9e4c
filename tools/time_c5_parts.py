"""CUDA-event times of the parts of the c5 CLAHE and Otsu operators on a 32 x 2048^2 stack."""
import os, sys, statistics
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yamimageprocessor_b200 import synth
from yamimageprocessor_b200.backend import get_backend
be = get_backend(0)
xs = be.to_device(np.stack([synth.nuclei(2048, 2048, seed=1000 + i) for i in range(32)]))
g = be.gaussian(xs, 11, 0.0)
c = be.clahe(g, 2.0, (8, 8))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=be.device)
def t(name, fn, reps=5):
    fn(); evs = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); o = fn(); b.record(); del o; evs.append((a, b))
    torch.cuda.synchronize()
    print(f"{name:40s} {statistics.mean(p.elapsed_time(q) for p, q in evs):8.3f} ms")
t("clahe (luts + apply) stack", lambda: be.clahe(g, 2.0, (8, 8)))
t("histogram stack", lambda: be.histogram(c))
h = be.histogram(c)
t("otsu scan from hist (certify)", lambda: be.otsu_from_histogram_device(h))
old = be.lib.yam_otsu_set_force_chain(1)
t("otsu scan from hist (forced chain)", lambda: be.otsu_from_histogram_device(h))
be.lib.yam_otsu_set_force_chain(old)
tt = be.otsu_from_histogram_device(h)
t("threshold_frames stack", lambda: be.threshold_frames(c, tt, 255))
t("otsu_threshold stack (all)", lambda: be.otsu_threshold(c, 255))
x1 = be.to_device(synth.nuclei(4096, 4096, seed=1000))
c1 = be.clahe(be.gaussian(x1, 11, 0.0), 2.0, (8, 8))
t("c1: histogram 4096^2", lambda: be.histogram(c1))
h1 = be.histogram(c1)
t("c1: otsu scan (certify)", lambda: be.otsu_from_histogram_device(h1))
t("c1: otsu_threshold", lambda: be.otsu_threshold(c1, 255))
t("c1: clahe", lambda: be.clahe(x1, 2.0, (8, 8)))
