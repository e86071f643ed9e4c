import sys
sys.path.insert(0, "/root/repo")
import torch
from yamimageprocessor_b200 import synth
from yamimageprocessor_b200.backend import get_backend
be = get_backend(0)
x = be.to_device(synth.nuclei(8192, 8192, seed=1000))
bits = be.adaptive_threshold_bits(x, 11, 2)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=be.device)
for _ in range(3): be.bits_morph(bits, 8192, 4, 5, 1)
ts = []
for _ in range(10):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); be.bits_morph(bits, 8192, 4, 5, 1); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print("bits_morph open+close", min(ts) * 1e3, "us")
