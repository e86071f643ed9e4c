"""Per-operator times of the c4 chain on one 8192-row x 65536-px strip (CUDA events, mean of 3 after 1 warm-up)."""
import os, sys, statistics
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yamimageprocessor_b200 import synth
from yamimageprocessor_b200.backend import get_backend

be = get_backend(0)
rows, W = int(os.environ.get("ROWS", 8192)), int(os.environ.get("COLS", 65536))
tile = synth.nuclei(4096, 4096, seed=100)
x = be.to_device(np.tile(tile, (rows // 4096, W // 4096)))
px = rows * W
def t(name, bpp, fn):
    fn(); evs = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); o = fn(); b.record(); del o; evs.append((a, b))
    torch.cuda.synchronize()
    ms = statistics.mean(p.elapsed_time(q) for p, q in evs)
    print(f"{name:34s} {ms:8.3f} ms {px*bpp/ms/1e6:8.0f} GB/s  {px*bpp/ms/1e6/6553:.3f}")
g = be.gaussian(x, 11, 0.0)
luts = be.clahe_luts(g, 2.0, (8, 1))
c = be.clahe_apply(g, luts, (W // 8, rows), 0)
bits = be.adaptive_threshold_bits(c, 11, 2)
bits2 = be.bits_morph(bits, W, 4, 5, 1)
t("gaussian k11", 4, lambda: be.gaussian(x, 11, 0.0))
t("clahe_luts", 2, lambda: be.clahe_luts(g, 2.0, (8, 1)))
t("clahe_apply", 4, lambda: be.clahe_apply(g, luts, (W // 8, rows), 0))
t("histogram", 2, lambda: be.histogram(c))
t("threshold", 4, lambda: be.threshold(c, 30000.0, 255))
t("adaptive bits", 2.125, lambda: be.adaptive_threshold_bits(c, 11, 2))
t("bits morph open+close", 0.25, lambda: be.bits_morph(bits, W, 4, 5, 1))
def ccl():
    ws, cnt = be.ccl_resolve_bits(bits2, W); return be.ccl_emit(bits2, W, ws)
t("ccl resolve+emit", 4.125, ccl)
print("checks", hex(int(be.checksum64(c).item()) & (2**64-1)), hex(int(be.checksum64(be.histogram(c)[0].to(torch.int32)).item()) & (2**64-1)))
