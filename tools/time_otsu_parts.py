import sys, time, ctypes as C
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from yamimageprocessor_b200 import synth, _lib
from yamimageprocessor_b200.backend import get_backend
be = get_backend(0)
print("host_threads", be.host_threads)
x = be.to_device(np.stack([synth.nuclei(2048, 2048, seed=1000 + i) for i in range(32)]))
c = be.clahe(be.gaussian(x, 11, 0.0), 2.0, (8, 8))
def sync(): torch.cuda.synchronize()
def T(fn, reps=5):
    fn(); sync(); best = 1e9
    for _ in range(reps):
        sync(); t0 = time.perf_counter(); r = fn(); sync(); best = min(best, time.perf_counter() - t0)
    return best * 1e3, r
ms, hist = T(lambda: be.histogram(c)); print(f"histogram        {ms:.3f} ms")
host = torch.empty(hist.shape, dtype=hist.dtype, pin_memory=True)
ms, _ = T(lambda: host.copy_(hist, non_blocking=True)); print(f"D2H 16 MB u64    {ms:.3f} ms")
hn = host.numpy(); out = np.zeros(32, np.int32)
ms, _ = T(lambda: be.lib.yam_otsu_from_hists(hn.ctypes.data_as(C.c_void_p), 65536, 32, out.ctypes.data_as(C.c_void_p))); print(f"pooled scans x32 {ms:.3f} ms")
ms, _ = T(lambda: be.lib.yam_otsu_from_hists(hn.ctypes.data_as(C.c_void_p), 65536, 1, out.ctypes.data_as(C.c_void_p))); print(f"one scan         {ms:.3f} ms")
t = be.to_device(out)
ms, _ = T(lambda: be.threshold(c, 20000.0, 255)); print(f"threshold        {ms:.3f} ms")
ms, _ = T(lambda: be.otsu_threshold(c, 255)); print(f"otsu_threshold   {ms:.3f} ms")
be.lib.yam_set_host_threads(1)
ms, _ = T(lambda: be.otsu_threshold(c, 255)); print(f"otsu (device scan) {ms:.3f} ms")
