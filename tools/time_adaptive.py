"""A/B of the adaptive-threshold -> bits kernels at 8192^2 (YAM_ADAPTIVE_LEGACY=1 selects sep_f32_tiled).
Run twice (once per setting); prints the CUDA-event time with a 256 MiB L2 flush between launches."""
import os, sys, statistics
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yamimageprocessor_b200 import synth
from yamimageprocessor_b200.backend import get_backend

be = get_backend(0)
size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
x = be.to_device(synth.nuclei(size, size, seed=2))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=be.device)
for blk in (11, 5, 15):
    for _ in range(3):
        bits = be.adaptive_threshold_bits(x, blk, 2)
    evs = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); bits = be.adaptive_threshold_bits(x, blk, 2); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    ms = statistics.mean(p.elapsed_time(q) for p, q in evs)
    print(f"legacy={os.environ.get('YAM_ADAPTIVE_LEGACY','0')} block {blk}: {ms*1e3:.1f} us  {size*size*2.125/ms/1e6:.0f} GB/s  checksum {int(be.checksum64(bits).item()) & (2**64-1):016x}")
