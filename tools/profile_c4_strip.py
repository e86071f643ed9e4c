"""One launch of every c4 operator on one 8192-row x 65536-px strip (for `ncu --set full`; no timing here)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yamimageprocessor_b200 import synth
from yamimageprocessor_b200.backend import get_backend

be = get_backend(0)
rows, W = int(os.environ.get("ROWS", 8192)), int(os.environ.get("COLS", 65536))
tile = synth.nuclei(4096, 4096, seed=100)
x = be.to_device(np.tile(tile, (rows // 4096, W // 4096)))
g = be.gaussian(x, 11, 0.0)
luts = be.clahe_luts(g, 2.0, (8, 1))
c = be.clahe_apply(g, luts, (W // 8, rows), 0)
hist = be.histogram(c)
t_dev = be.otsu_from_histogram_device(hist)                       # certified scan (+ early-exit chain / sigma kernels)
if os.environ.get("FUSED_MASK", "1") == "1":
    bits, m = be.adaptive_threshold_bits(c, 11, 2, mask_thresh=t_dev, maxval=255)   # final schedule: the Otsu mask rides along
else:
    m = be.threshold_frames(c, t_dev, 255)
    bits = be.adaptive_threshold_bits(c, 11, 2)
bits2 = be.bits_morph(bits, W, 4, 5, 1)
# the labeller indexes pixels with 32 bits: strips of 2^31 px or more go through in sub-strips (as host/mosaic.py does)
sub = rows
while sub * W >= (1 << 31) or rows % sub:
    sub //= 2
lab = torch.empty((rows, W), dtype=torch.int32, device=be.device)
total = 0
for y in range(0, rows, sub):
    ws, cnt = be.ccl_resolve_bits(bits2[y:y + sub], W)
    be.ccl_emit(bits2[y:y + sub], W, ws, out=lab[y:y + sub])
    total += int(cnt.item()) if hasattr(cnt, "item") else int(cnt)
torch.cuda.synchronize()
print("done", total)
