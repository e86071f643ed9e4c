#!/usr/bin/env python
"""Benchmark of the hot path (megapixels/s per pipeline, HBM-roofline fraction, CPU reference).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c5] [--impl reference]

A "step" is one pass of the named pipeline over one batch of synthetic frames on every rank.
Default workload = BASELINE.json configs[1]: segmentation (adaptive threshold 11/2 -> open 5x5 ->
close 5x5 -> connected components) on a synthetic 8192x8192 uint16 frame per GPU.
N > 1 (torchrun, one rank per GPU): frames are independent units, every rank processes its own
frame(s), no data-path collective -> weak scaling; time = max over ranks.

Prints ONE JSON line (rank 0).  See DESIGN.md §6 for the definition of every key.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from yamimageprocessor_b200 import synth  # noqa: E402

MP = 1.0e6

# algorithmic bytes per pixel; pipeline totals are the sums
WORKLOADS = {
    "c1": dict(h=4096, w=4096, frames=1, bpp=16.0,
               name="preprocess: Gaussian ksize=11 (sigma 2) -> CLAHE(2.0, 8x8) -> Otsu threshold, 4096x4096 uint16"),
    "c2": dict(h=8192, w=8192, frames=1, bpp=12.0,
               name="segmentation: adaptive threshold(11,2) -> open 5x5 -> close 5x5 -> connected components, 8192x8192 uint16"),
    "c3": dict(h=8192, w=8192, frames=1, bpp=6.0,
               name="extraction: per-region area/centroid/bbox/mean-intensity on the labelled 8192x8192 frame (~99k nuclei)"),
    "c4": dict(h=65536, w=65536, frames=1, bpp=28.0,
               name="mosaic: 65536x65536 uint16, preprocess (Gaussian k=11 -> CLAHE 8x8 -> Otsu) + segment (adaptive -> "
                    "open/close 5x5 -> connected components) in 8 row strips of whole CLAHE tile rows, halo over-fetch, "
                    "LUT all-gather, histogram all-reduce, cross-strip label merge"),
    "c5": dict(h=2048, w=2048, frames=32, bpp=34.0,
               name="time-lapse: preprocess+segment+extract on a batch of 2048x2048 uint16 frames, frame-sharded"),
}


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _traffic(op_name: str, px: int):
    """Measured DRAM bytes per launch of an operator (ncu --set full, profiles/r01_traffic_c2.json);
    None when no capture exists for this operator at this frame size."""
    p = ROOT / "profiles" / "r01_traffic_c2.json"
    try:
        d = json.loads(p.read_text())
    except Exception:
        return None
    if int(d.get("pixels", 0)) != int(px):
        return None
    for key, row in d.get("ops", {}).items():
        if op_name.startswith(key):
            return int(row["dram_read_bytes"]) + int(row["dram_write_bytes"])
    return None


# ---------------------------------------------------------------------------------------------
# CPU reference arm
def _cpu_pipeline(workload: str):
    from oracle import cv2_path as P

    if workload == "c1":
        return lambda fr, aux: [P.preprocess(f) for f in fr], P
    if workload == "c2":
        return lambda fr, aux: [P.segment(f) for f in fr], P
    if workload == "c3":
        return lambda fr, aux: [P.extract(l, f) for f, l in zip(fr, aux)], P
    return lambda fr, aux: [P.full_chain(f) for f in fr], P  # c4 / c5


def _cpu_inputs(workload: str, cfg, sample_frames: int, sample_hw):
    h, w = sample_hw
    frames = [synth.nuclei(h, w, seed=100 + i) for i in range(sample_frames)]
    aux = None
    if workload == "c3":
        from oracle import cv2_path as P

        aux = [P.segment(f) for f in frames]
    return frames, aux


def cpu_measure(workload: str, cfg, repeats: int, warmup: int, sample_hw, sample_frames: int = 1):
    fn, P = _cpu_pipeline(workload)
    frames, aux = _cpu_inputs(workload, cfg, sample_frames, sample_hw)
    px = sample_frames * sample_hw[0] * sample_hw[1]
    for _ in range(warmup):
        fn(frames, aux)
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        fn(frames, aux)
        times.append(time.perf_counter() - t0)
    best = min(times)
    return dict(
        value=px / MP / best,
        unit="megapixels/s",
        cores=int(P.THREADS),
        kind="port",
        sample=f"{sample_frames} frame(s) of {sample_hw[0]}x{sample_hw[1]} uint16 of the same synthetic workload, "
               f"best of {repeats}; {P.KIND}; host cores visible {os.cpu_count()}",
        ms=best * 1e3,
    ), statistics.mean(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = WORKLOADS[args.workload]
    sample_hw = (cfg["h"], cfg["w"]) if cfg["h"] <= 4096 else (4096, 4096)
    sample_frames = 1 if args.workload != "c5" else 2
    base, mean_s = cpu_measure(args.workload, cfg, max(1, args.steps), max(0, args.warmup), sample_hw, sample_frames)
    px = sample_frames * sample_hw[0] * sample_hw[1]
    value = px / MP / mean_s
    base["value"] = value
    line = {
        "impl": "reference",
        "metric": "megapixels/s per pipeline",
        "value": value,
        "unit": "megapixels/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": mean_s * 1e3,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u16",
        "data": "synthetic",
        "config": {"workload": cfg["name"], "sample": base["sample"]},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": "megapixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# GPU arm
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region by an NVML polling thread
    (about every 2 ms; `nvidia-smi -lms` cannot sample a region of a few milliseconds)."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self._stop = None
        self._thread = None
        self.max_mhz = None

    def start(self):
        import threading

        try:
            import pynvml

            pynvml.nvmlInit()
            # LOCAL_RANK indexes CUDA_VISIBLE_DEVICES; map through the UUID-free common case (identity)
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            return
        names = {
            getattr(pynvml, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(pynvml, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(
            pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")
        self._stop = threading.Event()

        def poll():
            while not self._stop.is_set():
                try:
                    self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                    mask = int(get_reasons(h))
                    for bit, name in names.items():
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
                self._stop.wait(0.002)

        self._thread = threading.Thread(target=poll, daemon=True)
        self._thread.start()

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.samples:
            out["sm_mhz"] = statistics.median(self.samples)
        return out


def build_gpu_workload(workload: str, cfg, be, frames_np):
    """Returns (device_inputs, run(device_inputs) -> outputs, ops for the per-op breakdown)."""
    import torch

    stack = np.stack(frames_np) if len(frames_np) > 1 else frames_np[0]
    x = be.to_device(stack)

    if workload == "c1":
        def run(inp):
            g = be.gaussian(inp, 11, 0.0)
            c = be.clahe(g, 2.0, (8, 8))
            t, m = be.otsu_threshold(c, 255)
            return m

        def ops(inp):
            g = be.gaussian(inp, 11, 0.0)
            c = be.clahe(g, 2.0, (8, 8))
            return [
                ("gaussian_fixed_u16_k11", 4.0, lambda: be.gaussian(inp, 11, 0.0)),
                ("clahe_u16 (lut + apply)", 6.0, lambda: be.clahe(g, 2.0, (8, 8))),
                ("otsu_threshold_u16 (hist + exact fp64 scan + threshold)", 6.0, lambda: be.otsu_threshold(c, 255)),
            ]
        return x, run, ops

    if workload == "c2":
        # The four reference steps dispatched through the executor's device-resident chain -- the very
        # schedule PipelineManager.apply (the e2e leg) runs between its upload and download.  The
        # thresholded mask is binary, so the executor keeps it 1 bit/pixel from the threshold kernel
        # through open/close into the labelling (same labels as the byte-mask chain, whose operators
        # are listed under roofline.ops_unfused for comparison).
        from yamimageprocessor_b200.host.executor import B200Executor
        from yamimageprocessor_b200.modules import b200_backend as plugin

        _mods = {cls().metadata.identifier: cls() for cls in plugin.MODULE_CLASSES}

        def _step(name, **params):
            st = _mods[name].create_pipeline_step()
            st.enabled = True
            st.params.update(params)
            return st

        _chain = [_step("Adaptive"), _step("Opening", kernel_size=5), _step("Closing", kernel_size=5),
                  _step("ConnectedComponents")]
        _ex = B200Executor(be)

        def run(inp):
            return _ex.run_chain_on_device(_chain, inp)

        def ops(inp):
            wd = int(inp.shape[-1])
            bits = be.adaptive_threshold_bits(inp, 11, 2)
            bits2 = be.bits_morph(bits, wd, 4, 5, 1)
            return [
                ("adaptive_threshold_bits_u16_b11 (sep_f32_tiled -> packed bits)", 2.125, lambda: be.adaptive_threshold_bits(inp, 11, 2)),
                ("bits_morph open+close 5x5 (bit_morph_reg_kernel)", 0.25, lambda: be.bits_morph(bits, wd, 4, 5, 1)),
                ("ccl_label_bits (scan, tile, border, rank, final_warp)", 4.125, lambda: be.ccl_label_bits(bits2, wd)),
            ]

        def ops_unfused(inp):
            m = be.adaptive_threshold(inp, 11, 2)
            m2 = be.morph_open_close(m, 5, 1)
            return [
                ("adaptive_threshold_u16_b11 (sep_f32_tiled)", 3.0, lambda: be.adaptive_threshold(inp, 11, 2)),
                ("morph_open_close_5x5_u8 (morph_fast_kernel)", 4.0, lambda: be.morph_open_close(m, 5, 1)),
                ("ccl_label (pack, scan, tile, border, rank, final_warp)", 5.0, lambda: be.ccl_label(m2)),
            ]
        ops.unfused = ops_unfused
        return x, run, ops

    if workload == "c3":
        m = be.morph_open_close(be.adaptive_threshold(x, 11, 2), 5, 1)
        labels, counts = be.ccl_label(m)
        n_labels = int(be.to_host(counts)[0])
        inp = (labels, x, n_labels)

        def run(inp_):
            return be.region_props(inp_[0], inp_[1], inp_[2])

        def ops(inp_):
            return [("region_props (props_kernel)", 6.0, lambda: be.region_props(inp_[0], inp_[1], inp_[2]))]
        return inp, run, ops

    # c5: per-frame full chain on a stack (n, h, w)
    def run(inp):
        g = be.gaussian(inp, 11, 0.0)
        c = be.clahe(g, 2.0, (8, 8))
        t, otsu_mask = be.otsu_threshold(c, 255)
        labels, counts = be.segment_fused(c, 11, 2, 5, 1)
        tables, offsets = be.region_props_stack(labels, c, counts)
        return otsu_mask, labels, tables

    def ops(inp):
        g = be.gaussian(inp, 11, 0.0)
        c = be.clahe(g, 2.0, (8, 8))
        wd = int(inp.shape[-1])
        bits = be.adaptive_threshold_bits(c, 11, 2)
        bits2 = be.bits_morph(bits, wd, 4, 5, 1)
        labels, counts = be.ccl_label_bits(bits2, wd)
        return [
            ("gaussian_fixed_u16_k11", 4.0, lambda: be.gaussian(inp, 11, 0.0)),
            ("clahe_u16 (lut + apply)", 6.0, lambda: be.clahe(g, 2.0, (8, 8))),
            ("otsu_threshold_u16 (hist + exact fp64 scan + threshold)", 6.0, lambda: be.otsu_threshold(c, 255)),
            ("adaptive_threshold_bits_u16_b11 (sep_f32_tiled -> packed bits)", 2.125, lambda: be.adaptive_threshold_bits(c, 11, 2)),
            ("bits_morph open+close 5x5 (bit_morph_reg_kernel)", 0.25, lambda: be.bits_morph(bits, wd, 4, 5, 1)),
            ("ccl_label_bits (scan, tile, border, rank, frame_offsets, final_warp)", 4.125, lambda: be.ccl_label_bits(bits2, wd)),
            ("region_props_stack (props_init + props_kernel)", 6.0, lambda: be.region_props_stack(labels, c, counts)),
        ]
    return x, run, ops


def e2e_callable(workload: str, be, frames_np):
    """The same pipeline through the reference-facing API with HOST buffers (H2D + D2H inside)."""
    from yamimageprocessor_b200.host.executor import B200Executor
    from yamimageprocessor_b200.host.pipeline import PipelineManager
    from yamimageprocessor_b200.modules import b200_backend as plugin

    mods = {cls().metadata.identifier: cls() for cls in plugin.MODULE_CLASSES}

    def step(name, **params):
        s = mods[name].create_pipeline_step()
        s.enabled = True
        s.params.update(params)
        return s

    ex = B200Executor(be)
    stack = np.stack(frames_np) if len(frames_np) > 1 else frames_np[0]
    host_in = be.pinned_empty(stack.shape, stack.dtype)
    host_in[...] = stack
    if workload == "c1":
        pm = PipelineManager([step("NoiseReduction", method="Gaussian", ksize=11), step("CLAHE"), step("Otsu")],
                             gpu_executor=ex)
    elif workload == "c2":
        pm = PipelineManager([step("Adaptive"), step("Opening", kernel_size=5), step("Closing", kernel_size=5),
                              step("ConnectedComponents")], gpu_executor=ex)
    elif workload == "c3":
        # labels + intensity in, table out: the extraction entry point of the plugin
        def call():
            t = plugin.region_properties_data(host_in)
            return t["area"]
        return call, host_in.nbytes, None
    else:
        pm = PipelineManager([step("NoiseReduction", method="Gaussian", ksize=11), step("CLAHE"), step("Adaptive"),
                              step("Opening", kernel_size=5), step("Closing", kernel_size=5),
                              step("ConnectedComponents")], gpu_executor=ex)

    def call():
        return pm.apply(host_in)
    return call, host_in.nbytes, pm


class _Shape:
    def __init__(self, h, w):
        self.shape = (h, w)


def mosaic_rows(size: int, r0: int, r1: int, tile: int, tiles: list) -> np.ndarray:
    """Rows [r0, r1) of the synthetic mosaic: a size x size grid of `tile`-sized frames drawn from a
    small set of distinct seeded frames (pattern (7*ty + 3*tx) % len(tiles))."""
    out = np.empty((r1 - r0, size), np.uint16)
    y = r0
    while y < r1:
        ty = y // tile
        y_end = min(r1, (ty + 1) * tile)
        for tx in range(size // tile):
            t = tiles[(7 * ty + 3 * tx) % len(tiles)]
            out[y - r0: y_end - r0, tx * tile: (tx + 1) * tile] = t[y - ty * tile: y_end - ty * tile]
        y = y_end
    return out


def run_gpu_mosaic(args):
    import torch
    import torch.distributed as dist

    from yamimageprocessor_b200.backend import get_backend
    from yamimageprocessor_b200.host import ingest, mosaic

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    be = get_backend(local_rank)
    cfg = WORKLOADS["c4"]
    size = int(args.mosaic_size)
    strips_total = 8
    if strips_total % world:
        raise SystemExit("c4 needs 1, 2, 4 or 8 ranks")
    local = strips_total // world
    tile = min(4096, size // 8)
    tiles = [synth.nuclei(tile, tile, seed=100 + i) for i in range(4)]
    shape = _Shape(size, size)
    p = mosaic.MosaicParams()
    trace_marks = []
    if os.environ.get("YAM_MOSAIC_TRACE") and rank == 0:
        # diagnostics: phase end times of rank 0's strips (adds a device sync per phase; not for reported numbers)
        def _trace(name, strip):
            torch.cuda.synchronize()
            trace_marks.append((time.perf_counter(), strip, name))
        p.trace = _trace
    host_strips, dev_strips = [], []
    for li in range(local):
        s = rank * local + li
        r0, r1 = mosaic.input_rows(size, s, strips_total, p)
        hs = mosaic_rows(size, r0, r1, tile, tiles)
        host_strips.append(hs)
        dev_strips.append(be.to_device(hs))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(device_sources):
        return mosaic.run_local_strips(be, shape, local, world > 1, p, with_props=False, device_sources=device_sources)

    for _ in range(max(3, args.warmup)):
        res = step(dev_strips)
    n_components = res[0].n_components
    del res
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    be.launch_count(reset=True)
    times = []
    for _ in range(args.steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        t_step = time.perf_counter()
        trace_marks.clear()
        res = step(dev_strips)
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
        if p.trace is not None:
            last = t_step
            for tm, strip, name in sorted(trace_marks):
                sys.stderr.write(f"[trace] +{1e3 * (tm - last):7.2f} ms strip {strip} {name}\n")
                last = tm
            sys.stderr.write(f"[trace] step total {1e3 * (time.perf_counter() - t_step):.2f} ms\n")
        del res
    barrier()
    launches = be.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = float(sum(times))
    # end to end: host strips in, labels + Otsu mask out (pageable host memory of this size is not pinned)
    barrier()
    t0 = time.perf_counter()
    res = mosaic.run_local_strips(be, shape, local, world > 1, p, with_props=False,
                                  device_sources=[ingest.upload_rows(be, h, 0, h.shape[0]) for h in host_strips])
    outs = []
    for r in res:  # results land in fresh pageable arrays, streamed through the pinned ring
        lab = np.empty(tuple(r.labels.shape), np.int32)
        msk = np.empty(tuple(r.otsu_mask.shape), np.uint16)
        outs.append((ingest.download_into(be, r.labels, lab), ingest.download_into(be, r.otsu_mask, msk)))
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    h2d = sum(h.nbytes for h in host_strips)
    d2h = sum(a_.nbytes + b_.nbytes for a_, b_ in outs)
    del outs, res
    if world > 1:
        t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=be.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s = float(t[0]), float(t[1])
        v = torch.tensor([h2d, d2h], dtype=torch.int64, device=be.device)
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
        h2d, d2h = int(v[0]), int(v[1])
    if rank == 0:
        peak, peak_src = _peaks()
        px = size * size
        ms_per_step = total_ms / args.steps
        value = px / MP / (ms_per_step / 1e3)
        achieved = px * cfg["bpp"] / (ms_per_step / 1e3) / 1e9
        cpu, _ = cpu_measure("c5", cfg, 2, 1, (4096, 4096), 1)
        line = {
            "metric": "megapixels/s per pipeline", "value": value, "unit": "megapixels/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
            "config": {"workload": cfg["name"] if size == 65536 else cfg["name"].replace("65536x65536", f"{size}x{size}"),
                       "mosaic": [size, size], "strips": strips_total, "strips_per_gpu": local,
                       "components": int(n_components), "algorithmic_bytes_per_px": cfg["bpp"],
                       "l2": "every strip (>= 1 GiB at 65536^2) exceeds the 126 MB L2", "seed": "tiles 100..103"},
            "roofline": {"bound": "hbm", "kernel": "whole mosaic pipeline (all kernels of all strips of the slowest rank)",
                         "achieved": achieved / world, "peak": peak, "unit": "GB/s", "frac": achieved / world / peak,
                         "traffic": None, "peak_source": peak_src, "per_gpu": True},
            "cpu_baseline": cpu,
            "e2e": {"value": px / MP / e2e_s, "unit": "megapixels/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h),
                    "api": "host.ingest.upload_rows(pageable strips) -> host.mosaic.run_local_strips -> host.ingest.download_into(fresh pageable arrays): labels + Otsu mask"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_gpu(args):
    import torch
    import torch.distributed as dist

    if args.workload == "c4":
        return run_gpu_mosaic(args)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from yamimageprocessor_b200.backend import get_backend

    be = get_backend(local_rank)
    cfg = WORKLOADS[args.workload]
    h, w, nfr = cfg["h"], cfg["w"], cfg["frames"]
    frames_np = [synth.nuclei(h, w, seed=1000 + rank * nfr + i) for i in range(nfr)]
    inp, run, ops = build_gpu_workload(args.workload, cfg, be, frames_np)
    px_per_rank = nfr * h * w

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=be.device)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        run(inp)
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    be.launch_count(reset=True)
    barrier()
    for i in range(args.steps):
        flush.zero_()
        starts[i].record()
        run(inp)
        stops[i].record()
    barrier()
    launches = be.launch_count()
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]
    total_ms = float(sum(step_ms))
    clocks = sampler.stop() if rank == 0 else None

    # per-op breakdown for the roofline of the dominant op (device-resident, L2 flushed)
    def measure_ops(op_list):
        rows = []
        for name, bpp, fn in op_list:
            for _ in range(2):
                fn()
            reps = max(3, min(args.steps, 10))
            evs = []
            for _ in range(reps):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                evs.append((a, b))
            torch.cuda.synchronize()
            rows.append((name, bpp, statistics.mean(a.elapsed_time(b) for a, b in evs)))
        return rows

    breakdown = measure_ops(ops(inp))
    breakdown_unfused = measure_ops(ops.unfused(inp)) if hasattr(ops, "unfused") else []

    # end to end through the reference-facing API, host buffers, wall clock
    call, h2d_bytes, _pm = e2e_callable(args.workload, be, frames_np)
    out = call()
    d2h_bytes = int(getattr(out, "nbytes", 0))
    for _ in range(2):
        call()
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        call()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0)

    # max over ranks
    if world > 1:
        t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=be.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s = float(t[0]), float(t[1])

    if rank == 0:
        peak, peak_src = _peaks()
        ms_per_step = total_ms / args.steps
        value = world * px_per_rank / MP / (ms_per_step / 1e3)
        e2e_value = world * px_per_rank / MP / (e2e_s / e2e_steps)
        dom = max(breakdown, key=lambda r: r[2])
        achieved = px_per_rank * dom[1] / (dom[2] / 1e3) / 1e9
        pipe_achieved = px_per_rank * cfg["bpp"] / (ms_per_step / 1e3) / 1e9
        cpu_hw = (h, w) if h <= 4096 else (4096, 4096)
        cpu, _ = cpu_measure(args.workload, cfg, 2, 1, cpu_hw, 1 if args.workload != "c5" else 2)
        line = {
            "metric": "megapixels/s per pipeline",
            "value": value,
            "unit": "megapixels/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "u16",
            "data": "synthetic",
            "config": {
                "workload": cfg["name"],
                "frames_per_gpu": nfr,
                "frame": [h, w],
                "algorithmic_bytes_per_px": cfg["bpp"],
                "l2": "256 MiB flush write between timed steps (not timed)",
                "seed": "1000 + rank*frames + i",
            },
            "roofline": {
                "bound": "hbm",
                "kernel": dom[0],
                "achieved": achieved,
                "peak": peak,
                "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": _traffic(dom[0], px_per_rank),
                "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                "algorithmic_bytes": int(px_per_rank * dom[1]),
                "peak_source": peak_src,
                "pipeline_achieved": pipe_achieved,
                "pipeline_frac": pipe_achieved / peak,
                "ops": [{"op": n, "bytes_per_px": b, "ms": m, "GBps": px_per_rank * b / (m / 1e3) / 1e9,
                         "frac": px_per_rank * b / (m / 1e3) / 1e9 / peak} for n, b, m in breakdown],
                "ops_unfused": [{"op": n, "bytes_per_px": b, "ms": m, "GBps": px_per_rank * b / (m / 1e3) / 1e9,
                                 "frac": px_per_rank * b / (m / 1e3) / 1e9 / peak} for n, b, m in breakdown_unfused],
            },
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "megapixels/s", "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": d2h_bytes, "api": "PipelineManager(gpu_executor=B200Executor).apply(ndarray)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="c2")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--mosaic-size", type=int, default=65536, help="side of the c4 mosaic (multiple of 64)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"),
               str(Path(__file__).resolve()), "--gpus", str(args.gpus), "--steps", str(args.steps),
               "--warmup", str(args.warmup), "--workload", args.workload, "--mosaic-size", str(args.mosaic_size)]
        return subprocess.call(cmd)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
