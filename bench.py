#!/usr/bin/env python
"""Benchmark of the hot path (megapixels/s per pipeline, HBM-roofline fraction, CPU reference).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4|c5] [--impl reference]

A "step" is one pass of the named pipeline over the whole synthetic job.
Default workload = c4, BASELINE.json configs[3]: the 65536x65536 uint16 mosaic, full preprocess +
segment, as ONE row strip per GPU (whole CLAHE tile rows, over-fetched halos, LUT all-gather,
histogram all-reduce, cross-strip label merge) -> STRONG scaling over N = 1, 2, 4, 8 (the config the
north_star's 8-GPU target is quoted on; it fits one B200, so N = 1 runs it as a single strip with no
collective).  c5 = 1024 frames of 2048x2048 split N ways (strong).  c1 / c2 / c3 are the single-frame
configs (replicas only: every rank runs its own frame, weak).  Time = max over ranks.

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
                    "open/close 5x5 -> connected components), one row strip of whole CLAHE tile rows per GPU, halo "
                    "over-fetch, LUT all-gather, histogram all-reduce, cross-strip label merge"),
    "c5": dict(h=2048, w=2048, frames=1024, bpp=34.0,
               name="time-lapse: 1024 frames of 2048x2048 uint16 through preprocess+segment+extract, frame-sharded "
                    "(1024/N frames per GPU, batches of 32 frames per launch)"),
}
C5_BATCH = 32


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
    """Measured DRAM bytes per launch of an operator (ncu --set full; profiles/r0*_traffic_*.json);
    None when no capture exists for this operator at this pixel count."""
    for name in ("r02_traffic_c4.json", "r02_traffic_c4_n1.json", "r02_traffic_c2.json", "r01_traffic_c2.json"):
        try:
            d = json.loads((ROOT / "profiles" / name).read_text())
        except Exception:
            continue
        if int(d.get("pixels", 0)) != int(px):
            continue
        for key, row in d.get("ops", {}).items():
            if op_name.startswith(key):
                return int(row["dram_read_bytes"]) + int(row["dram_write_bytes"])
    return None


# ---------------------------------------------------------------------------------------------
# CPU reference arm: the unmodified reference when its checkout is importable (build container),
# else its call sites restated on the same cv2 (oracle/ref_path.py says which) -- all host threads.
def _cpu_pipeline(workload: str):
    from oracle import ref_path as R

    if workload == "c1":
        return lambda fr, aux: [R.preprocess(f) for f in fr], R
    if workload == "c2":
        return lambda fr, aux: [R.segment(f) for f in fr], R
    if workload == "c3":
        return lambda fr, aux: [R.extract(l, f) for f, l in zip(fr, aux)], R
    if workload == "c4":
        return lambda fr, aux: [R.mosaic_chain(f) for f in fr], R
    return lambda fr, aux: [R.full_chain(f) for f in fr], R  # c5


def cpu_sample(workload: str):
    """(frames, aux, description): the bounded sample of the workload one CPU step processes."""
    cfg = WORKLOADS[workload]
    if workload == "c4":
        # an 8192 x 8192 crop of the mosaic (two by two of its 4096^2 source frames): ~3 s of CPU work per step
        tiles = mosaic_tiles(4096)
        frame = mosaic_rows(8192, 0, 8192, 4096, tiles)
        return [frame], None, "rows 0..8191 x cols 0..8191 of the same synthetic 65536^2 mosaic (1/64 of the job)"
    if workload == "c5":
        frames = [synth.nuclei(cfg["h"], cfg["w"], seed=1000 + i) for i in range(8)]
        return frames, None, "frames 0..7 of the 1024-frame job (1/128 of the job)"
    frames = [synth.nuclei(cfg["h"], cfg["w"], seed=1000)]
    aux = None
    if workload == "c3":
        from oracle import ref_path as R

        aux = [R.segment(f) for f in frames]
    return frames, aux, f"the whole job: 1 frame of {cfg['h']}x{cfg['w']} uint16 (seed 1000)"


def cpu_measure(workload: str, repeats: int, warmup: int):
    """Mean over `repeats` timed passes of the CPU path on the bounded sample (same statistic as the GPU arm)."""
    fn, R = _cpu_pipeline(workload)
    frames, aux, what = cpu_sample(workload)
    px = sum(int(f.shape[0]) * int(f.shape[1]) for f in frames)
    for _ in range(warmup):
        fn(frames, aux)
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        fn(frames, aux)
        times.append(time.perf_counter() - t0)
    mean_s = statistics.mean(times)
    kind, detail = R.describe()
    return dict(
        value=px / MP / mean_s,
        unit="megapixels/s",
        cores=int(R.THREADS),
        kind=kind,
        sample=f"{what}; mean of {repeats} pass(es) after {warmup} warm-up; {detail}; host cores visible {os.cpu_count()}",
        ms=mean_s * 1e3,
    ), mean_s


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = WORKLOADS[args.workload]
    base, mean_s = cpu_measure(args.workload, max(1, args.steps), max(0, args.warmup))
    value = base["value"]
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
        "scaling": "strong" if args.workload in ("c4", "c5") else "weak",
        "vs_baseline": None,
        "dtype": "u16",
        "data": "synthetic",
        "config": {"workload": cfg["name"], "sample": base["sample"],
                   "note": "CPU arm: each step is the bounded sample named in `sample`; value = sample megapixels / mean step time "
                           "(the CPU path does not speed up with --gpus)"},
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
    if workload == "c5" and cfg.get("frames_rank", len(frames_np)) > len(frames_np):
        # this rank's share of the 1024 frames, resident in HBM: the distinct synthetic frames, cycled
        x = x.repeat(cfg["frames_rank"] // len(frames_np), 1, 1)

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
                ("otsu_threshold_u16 (hist + certified scan + threshold)", 6.0, lambda: be.otsu_threshold(c, 255)),
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
                ("adaptive_threshold_bits_u16_b11 (adaptive_bits_tma_kernel -> packed bits)", 2.125, lambda: be.adaptive_threshold_bits(inp, 11, 2)),
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

    # c5: per-frame full chain on stacks (n, h, w), C5_BATCH frames per launch, over this rank's share
    def run_batch(inp):
        g = be.gaussian(inp, 11, 0.0)
        c = be.clahe(g, 2.0, (8, 8))
        # histogram + certified parallel Otsu scan on the device (thresholds never leave HBM); the Otsu mask is
        # written by the adaptive-threshold kernel, which has the same pixels staged in shared memory
        t, _ = be.otsu_threshold(c, want_image=False)
        labels, counts, otsu_mask = be.segment_fused(c, 11, 2, 5, 1, mask_thresh=t, maxval=255)
        tables, offsets = be.region_props_stack(labels, c, counts)
        return otsu_mask, labels, tables

    def run(inp):
        out = None
        for b0 in range(0, int(inp.shape[0]), C5_BATCH):
            out = run_batch(inp[b0:b0 + C5_BATCH])
        return out

    def ops(inp):
        inp = inp[:C5_BATCH]
        g = be.gaussian(inp, 11, 0.0)
        c = be.clahe(g, 2.0, (8, 8))
        wd = int(inp.shape[-1])
        bits = be.adaptive_threshold_bits(c, 11, 2)
        bits2 = be.bits_morph(bits, wd, 4, 5, 1)
        labels, counts = be.ccl_label_bits(bits2, wd)
        t_probe, _ = be.otsu_threshold(c, want_image=False)
        return [
            ("gaussian_fixed_u16_k11", 4.0, lambda: be.gaussian(inp, 11, 0.0)),
            ("clahe_u16 (lut + apply)", 6.0, lambda: be.clahe(g, 2.0, (8, 8))),
            ("otsu_u16 (hist + certified scan)", 2.0, lambda: be.otsu_threshold(c, want_image=False)),
            ("adaptive_threshold_bits_u16_b11 + Otsu mask (one pass: packed bits + u16 mask)", 4.125,
             lambda: be.adaptive_threshold_bits(c, 11, 2, mask_thresh=t_probe, maxval=255)),
            ("bits_morph open+close 5x5 (bit_morph_reg_kernel)", 0.25, lambda: be.bits_morph(bits, wd, 4, 5, 1)),
            ("ccl_label_bits (scan, tile, border, rank, frame_offsets, final_warp)", 4.125, lambda: be.ccl_label_bits(bits2, wd)),
            ("region_props_stack (props_init + props_kernel)", 6.0, lambda: be.region_props_stack(labels, c, counts)),
        ]
    return x, run, ops


def e2e_callable(workload: str, be, frames_np, pinned: bool = False):
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
    if pinned:
        host_in = be.pinned_empty(stack.shape, stack.dtype)
        host_in[...] = stack
    else:
        host_in = np.array(stack, copy=True)   # ordinary pageable memory, what the reference's callers hold
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


def mosaic_tiles(tile: int):
    """The distinct seeded source frames the synthetic mosaic is assembled from (seeds 100..103)."""
    return [synth.nuclei(tile, tile, seed=100 + i) for i in range(4)]


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


class RowWindowRecord:
    """Lazy-handle stand-in with the TiledImageRecord surface (core/tiled_image.py:52): a 65536^2
    mosaic of which THIS process keeps only the rows it will read (its strip + halo) in ordinary
    pageable memory, served as zero-copy views like slices of a warmed np.memmap.  Reading anything
    else, or densifying, is an error -- which is the point of the tiled route."""

    def __init__(self, size: int, r0: int, rows: np.ndarray):
        self.shape, self.dtype, self.size = (size, size), rows.dtype, None
        self._r0, self._rows = r0, rows

    def read_region(self, box):
        left, top, right, bottom = box
        if top < self._r0 or bottom > self._r0 + self._rows.shape[0]:
            raise IndexError(f"rows [{top}, {bottom}) are not resident in this process")
        return self._rows[top - self._r0: bottom - self._r0, left:right]

    def to_array(self):
        raise RuntimeError("the mosaic handle must not be densified")

    def close(self):
        pass


def cv2_reference_check(size: int, check: dict):
    """The `check` values the reference's cv2 call sites give for the same synthetic mosaic, computed on the CPU by
    tools/check_bench_check_block.py and committed under profiles/ (65536^2 and 16384^2), next to a verdict on this
    run's values.  None when no committed result covers `size`; never raises (evidence only, not on the timed path)."""
    try:
        path = ROOT / "profiles" / f"r02_check_c4_{int(size)}_vs_cv2.json"
        ref = json.loads(path.read_text())
        cpu = next(v for k, v in ref.items() if k.startswith("cpu (cv2"))
        lib = next(k for k in ref if k.startswith("cpu (cv2"))
        return {"source": f"profiles/{path.name}: {lib} over the same mosaic (tools/check_bench_check_block.py)",
                "values": cpu, "equal": all(check.get(k) == v for k, v in cpu.items())}
    except Exception:  # noqa: BLE001 - a missing / unreadable file only drops the extra key
        return None


def _u64hex(v: int) -> str:
    return f"{int(v) & 0xFFFFFFFFFFFFFFFF:016x}"


def run_gpu_mosaic(args):
    import torch
    import torch.distributed as dist

    from yamimageprocessor_b200.backend import get_backend
    from yamimageprocessor_b200.host import ingest, mosaic
    from yamimageprocessor_b200.host.pipeline import PipelineManager
    from yamimageprocessor_b200.host.tiles import TiledPipelineImage
    from yamimageprocessor_b200.modules import b200_backend as plugin

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    be = get_backend(local_rank)
    cfg = WORKLOADS["c4"]
    size = int(args.mosaic_size)
    if 8 % world:
        raise SystemExit("c4 needs 1, 2, 4 or 8 ranks (strips are whole rows of the 8x8 CLAHE grid)")
    tile = min(4096, size // 8)
    tiles = mosaic_tiles(tile)
    shape = _Shape(size, size)
    p = mosaic.MosaicParams()
    trace_marks = []
    if os.environ.get("YAM_MOSAIC_TRACE") and rank == 0:
        # diagnostics: phase end times of rank 0 (adds a device sync per phase; not for reported numbers)
        # YAM_MOSAIC_TRACE=events: CUDA events at the phase ends, no synchronisation (host and device clocks side by side)
        trace_events = os.environ.get("YAM_MOSAIC_TRACE") == "events"

        def _trace(name, strip):
            if trace_events:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                trace_marks.append((time.perf_counter(), strip, name, ev))
            else:
                torch.cuda.synchronize()
                trace_marks.append((time.perf_counter(), strip, name, None))
        p.trace = _trace
    r0, r1 = mosaic.input_rows(size, rank, world, p)
    c0, c1 = mosaic.strip_rows(size, p.tile_grid[1], rank, world)
    host_rows = mosaic_rows(size, r0, r1, tile, tiles)          # pageable host memory
    dev_rows = ingest.upload_rows(be, host_rows, 0, host_rows.shape[0])
    comm = mosaic.TorchComm() if world > 1 else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return mosaic.run_strip(be, shape, rank, world, p, device_source=dev_rows, comm=comm)

    warmup = max(3, args.warmup)
    for _ in range(warmup):
        res = step()
    del res
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    be.launch_count(reset=True)
    times = []
    res = None
    for _ in range(args.steps):
        del res
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        t_step = time.perf_counter()
        trace_marks.clear()
        res = step()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
        if p.trace is not None:
            last = t_step
            prev_ev = a
            for tm, strip, name, ev in trace_marks:
                if ev is not None:
                    sys.stderr.write(f"[trace] host +{1e3 * (tm - last):7.2f} ms | device +{prev_ev.elapsed_time(ev):7.3f} ms  {name}\n")
                    prev_ev = ev
                else:
                    sys.stderr.write(f"[trace] +{1e3 * (tm - last):7.2f} ms strip {strip} {name}\n")
                last = tm
            if prev_ev is not a:
                sys.stderr.write(f"[trace] device tail +{prev_ev.elapsed_time(b):7.3f} ms (after the last mark)\n")
            sys.stderr.write(f"[trace] step total {1e3 * (time.perf_counter() - t_step):.2f} ms\n")
    barrier()
    launches = be.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = float(sum(times))

    # ---- parity evidence: identical for every N (and equal to the dense single-strip run of
    # tests/test_gpu_mosaic.py at reduced size): Otsu t, component count, content checksums
    W = size
    sums = torch.zeros((2,), dtype=torch.int64, device=be.device)
    be.checksum64(res.labels, index_base=c0 * W, accumulate=sums[0:1])
    be.checksum64(res.otsu_mask, index_base=c0 * W, accumulate=sums[1:2])
    if world > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)       # int64 addition wraps mod 2^64, like the checksum
    sums_host = sums.cpu().tolist()
    check = {"otsu_threshold": int(res.otsu_threshold), "components": int(res.n_components),
             "labels_checksum64": _u64hex(sums_host[0]), "otsu_mask_checksum64": _u64hex(sums_host[1]),
             "definition": "yam_checksum64 (include/yamb200.h): sum of mix64(linear index * golden + value) mod 2^64, "
                           "per-strip sums all-reduced; must be identical for N = 1, 2, 4, 8"}
    reference_check = cv2_reference_check(size, check)
    if reference_check is not None:
        check["cv2_cpu_chain"] = reference_check
    n_components = int(res.n_components)
    del res

    # ---- per-operator times on this rank's strip (CUDA events, inputs > L2), rank 0 reports
    strip_px = (c1 - c0) * W

    def measure(fn, reps=3):
        fn()
        evs = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = fn()
            b.record()
            del out
            evs.append((a, b))
        torch.cuda.synchronize()
        return statistics.mean(x.elapsed_time(y) for x, y in evs)

    g = be.gaussian(dev_rows, p.gauss_ksize, 0.0)
    g_core = g[c0 - r0: c1 - r0]
    luts = be.clahe_luts(g_core, p.clip_limit, (p.tile_grid[0], p.tile_grid[1] // world))
    th, tw = size // p.tile_grid[1], size // p.tile_grid[0]
    c = be.clahe_apply(g_core, luts, (tw, th), y_offset=0)
    bits = be.adaptive_threshold_bits(c, p.block_size, p.C)
    bits2 = be.bits_morph(bits, W, 4, p.morph_ksize, 1)
    t_dev = torch.full((1,), 30000, dtype=torch.int32, device=be.device)

    def otsu_op():     # histogram + scan; the Otsu MASK is written by the adaptive-threshold kernel (same pass over c)
        h = be.histogram(c)
        return be.otsu_from_histogram_device(h), h

    _, cert = be.otsu_from_histogram_device(be.histogram(c), want_certified=True)
    otsu_where = ("device: certified parallel scan (one 8-CTA cluster; " +
                  ("certificate decided this histogram, no sequential chain)" if int(cert[0].item())
                   else "NOT certified for this histogram: exact fp64 chain kernels ran)"))

    sub_rows = (c1 - c0) // max(1, -(-strip_px // mosaic._CCL_MAX_PX))

    def ccl_op():   # in sub-strips of < 2^31 pixels like run_strip (32-bit pixel indices in the labeller)
        lab = torch.empty((c1 - c0, W), dtype=torch.int32, device=be.device)
        for y in range(0, c1 - c0, sub_rows):
            ws, cnt = be.ccl_resolve_bits(bits2[y:y + sub_rows], W)
            be.ccl_emit(bits2[y:y + sub_rows], W, ws, out=lab[y:y + sub_rows])
        return lab

    op_rows = [
        ("gaussian_fixed_u16_k11 (gauss16_tma_kernel)", 4.0, measure(lambda: be.gaussian(g_core, p.gauss_ksize, 0.0))),
        ("clahe_u16 (tile LUTs + apply)", 6.0, measure(lambda: be.clahe_apply(
            g_core, be.clahe_luts(g_core, p.clip_limit, (p.tile_grid[0], p.tile_grid[1] // world)), (tw, th), 0))),
        ("otsu_u16 (histogram + certified scan)", 2.0, measure(otsu_op)),
        ("adaptive_threshold_bits_u16_b11 + Otsu mask (one pass: packed bits + u16 mask)", 4.125,
         measure(lambda: be.adaptive_threshold_bits(c, p.block_size, p.C, mask_thresh=t_dev, maxval=255))),
        ("bits_morph open+close 5x5 (bit_morph_reg_kernel)", 0.25, measure(lambda: be.bits_morph(bits, W, 4, p.morph_ksize, 1))),
        ("ccl from bits (resolve: scan, tile, border, rank; emit: final_warp)", 4.125, measure(ccl_op)),
    ]
    del g, g_core, luts, c, bits, bits2

    # ---- N = 1 only: the same mosaic as 8 lock-step strips on this GPU (round 1's schedule), for comparison
    alt = None
    if world == 1 and not args.no_alt:
        dev8 = []
        for s8 in range(8):
            a0, a1 = mosaic.input_rows(size, s8, 8, p)
            dev8.append(dev_rows[a0:a1])   # views of the resident mosaic
        for _ in range(2):
            r8 = mosaic.run_local_strips(be, shape, 8, False, p, device_sources=dev8)
            del r8
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r8 = mosaic.run_local_strips(be, shape, 8, False, p, device_sources=dev8)
        b.record()
        torch.cuda.synchronize()
        alt = {"schedule": "8 row strips in lock-step host threads on one GPU (in-process collectives)",
               "ms_per_step": a.elapsed_time(b), "components": int(r8[0].n_components)}
        del r8, dev8

    # ---- end to end through the reference-facing API: PipelineManager.apply(TiledPipelineImage) ->
    # supports_tiled_input step (MosaicModule.process): pageable host rows -> pinned ring -> HBM ->
    # kernels + collectives -> labels back into a fresh pageable array
    del dev_rows
    torch.cuda.empty_cache()
    step_mod = plugin.MosaicModule().create_pipeline_step()
    step_mod.enabled = True
    pm = PipelineManager([step_mod])
    handle = TiledPipelineImage(RowWindowRecord(size, r0, host_rows), tile_size=(tile, tile))
    e2e_times = []
    out = pm.apply(handle)     # warm-up pass: first use of the page-locked result pool (cudaHostAlloc is slow once)
    for it in range(2):
        del out
        barrier()
        t0 = time.perf_counter()
        out = pm.apply(handle)
        torch.cuda.synchronize()
        e2e_times.append(time.perf_counter() - t0)
    e2e_s = min(e2e_times) if args.e2e_best else statistics.mean(e2e_times)
    h2d, d2h = int(host_rows.nbytes), int(out.nbytes)
    e2e_labels_ok = bool(out.shape == (c1 - c0, W) and out.dtype == np.int32)
    del out
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
        pipe_achieved = strip_px * cfg["bpp"] / (ms_per_step / 1e3) / 1e9
        dom = max(op_rows, key=lambda r: r[2])
        achieved = strip_px * dom[1] / (dom[2] / 1e3) / 1e9
        cpu, _ = cpu_measure("c4", 2, 1)
        line = {
            "metric": "megapixels/s per pipeline", "value": value, "unit": "megapixels/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
            "config": {"workload": cfg["name"] if size == 65536 else cfg["name"].replace("65536x65536", f"{size}x{size}"),
                       "mosaic": [size, size], "strips": world, "strips_per_gpu": 1, "rows_per_gpu": c1 - c0,
                       "halo_rows_overfetched": [c0 - r0, r1 - c1],
                       "chain": "Gaussian k=11 -> CLAHE(2.0, 8x8) -> [Otsu mask of the CLAHE output] ; adaptive(11, 2) on the CLAHE "
                                "output -> open 5x5 -> close 5x5 -> 8-connected labels (SURVEY.md 8(d) chain definition)",
                       "algorithmic_bytes_per_px": cfg["bpp"],
                       "l2": "every strip (>= 1 GiB at 65536^2) exceeds the 126 MB L2; no flush needed",
                       "seed": "source frames 100..103, pattern (7*ty + 3*tx) % 4",
                       "otsu_scan": otsu_where},
            "check": check,
            "roofline": {"bound": "hbm", "kernel": dom[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": _traffic(dom[0], strip_px),
                         "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                         "algorithmic_bytes": int(strip_px * dom[1]), "peak_source": peak_src,
                         "scope": f"rank 0's strip ({c1 - c0} rows x {W} px), each operator timed alone with CUDA events",
                         "pipeline_achieved": pipe_achieved, "pipeline_frac": pipe_achieved / peak, "per_gpu": True,
                         "ops": [{"op": n, "bytes_per_px": bpp, "ms": m, "GBps": strip_px * bpp / (m / 1e3) / 1e9,
                                  "frac": strip_px * bpp / (m / 1e3) / 1e9 / peak} for n, bpp, m in op_rows]},
            "cpu_baseline": cpu,
            "e2e": {"value": px / MP / e2e_s, "unit": "megapixels/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "host_memory": "pageable input rows (staged through a pinned ring); result = fresh array from the page-locked host pool (Backend.to_host semantics)",
                    "statistic": "mean of 2 passes after 1 warm-up pass", "labels_shape_ok": e2e_labels_ok,
                    "api": "PipelineManager([Mosaic step]).apply(TiledPipelineImage) -> MosaicModule.process "
                           "(supports_tiled_input route, processing/pipeline_manager.py:412-416): int32 label rows"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if alt is not None:
            line["n1_schedules"] = {"one_strip_ms": ms_per_step, "eight_strips_lockstep": alt}
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
    cfg = dict(WORKLOADS[args.workload])
    h, w, nfr = cfg["h"], cfg["w"], cfg["frames"]
    strong = args.workload == "c5"
    if strong:
        # 1024 frames split N ways: this rank's contiguous block (host/sharding.frame_block), resident in HBM
        from yamimageprocessor_b200.host.sharding import frame_block

        nfr = int(args.frames) if args.frames else nfr
        f0, f1 = frame_block(nfr, rank, world)
        if (f1 - f0) % C5_BATCH:
            raise SystemExit(f"c5: {nfr} frames over {world} ranks must give whole batches of {C5_BATCH}")
        cfg["frames_rank"] = f1 - f0
        frames_np = [synth.nuclei(h, w, seed=1000 + i) for i in range(C5_BATCH)]   # distinct frames, cycled
        frames_rank = f1 - f0
    else:
        frames_np = [synth.nuclei(h, w, seed=1000 + rank * nfr + i) for i in range(nfr)]
        frames_rank = nfr
    inp, run, ops = build_gpu_workload(args.workload, cfg, be, frames_np)
    px_per_rank = frames_rank * h * w
    px_ops = (C5_BATCH if strong else nfr) * h * w   # the per-operator table is measured on one batch

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=be.device)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        run(inp)
    barrier()

    otsu_where = None
    if args.workload in ("c1", "c5"):     # where the Otsu scan runs for this workload's histograms (untimed probe)
        probe = inp[:C5_BATCH] if strong else inp
        _, cert = be.otsu_from_histogram_device(be.histogram(be.clahe(be.gaussian(probe, 11, 0.0), 2.0, (8, 8))), want_certified=True)
        ncert, nall = int(cert.sum().item()), int(cert.numel())
        otsu_where = (f"device: certified parallel scan, one 8-CTA cluster per frame; {ncert} of {nall} frames decided by the "
                      "certificate, the rest by the exact fp64 chain kernels; no host scan, nothing read back")
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

    # end to end through the reference-facing API, HOST buffers, wall clock.  Headline: ordinary pageable
    # input (what the reference's callers hold); the page-locked variant is reported next to it.
    # c5: one call per 32-frame batch over this rank's whole share of the job.
    calls_per_step = frames_rank // C5_BATCH if strong else 1

    def time_e2e(pinned: bool):
        call, h2d_b, _pm = e2e_callable(args.workload, be, frames_np, pinned=pinned)
        out = call()
        d2h_b = int(getattr(out, "nbytes", 0))
        del out
        call()
        barrier()
        reps = 1 if strong else max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(reps * calls_per_step):
            call()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps, h2d_b * calls_per_step, d2h_b * calls_per_step

    e2e_s, h2d_bytes, d2h_bytes = time_e2e(pinned=False)
    e2e_pinned_s, _, _ = time_e2e(pinned=True)

    # max over ranks
    if world > 1:
        t = torch.tensor([total_ms, e2e_s, e2e_pinned_s], dtype=torch.float64, device=be.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s, e2e_pinned_s = float(t[0]), float(t[1]), float(t[2])

    if rank == 0:
        peak, peak_src = _peaks()
        ms_per_step = total_ms / args.steps
        job_px = (nfr if strong else world * nfr) * h * w
        value = job_px / MP / (ms_per_step / 1e3)
        e2e_value = job_px / MP / e2e_s
        dom = max(breakdown, key=lambda r: r[2])
        achieved = px_ops * dom[1] / (dom[2] / 1e3) / 1e9
        pipe_achieved = px_per_rank * cfg["bpp"] / (ms_per_step / 1e3) / 1e9
        cpu, _ = cpu_measure(args.workload, 2, 1)
        line = {
            "metric": "megapixels/s per pipeline",
            "value": value,
            "unit": "megapixels/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None,
            "dtype": "u16",
            "data": "synthetic",
            "config": {
                "workload": cfg["name"],
                "frames_per_gpu": frames_rank,
                "frames_total": nfr if strong else world * nfr,
                "frame": [h, w],
                "algorithmic_bytes_per_px": cfg["bpp"],
                "l2": "256 MiB flush write between timed steps (not timed)",
                "seed": "1000 + i for 32 distinct frames, cycled over the job" if strong else "1000 + rank*frames + i",
                "otsu_scan": otsu_where if args.workload in ("c1", "c5") else None,
                "multi_gpu": ("contiguous frame blocks per rank, no data-path collective" if strong
                              else "replicas only: every rank runs its own frame (SURVEY.md 8e)"),
            },
            "roofline": {
                "bound": "hbm",
                "kernel": dom[0],
                "achieved": achieved,
                "peak": peak,
                "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": _traffic(dom[0], px_ops),
                "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                "algorithmic_bytes": int(px_ops * dom[1]),
                "peak_source": peak_src,
                "pipeline_achieved": pipe_achieved,
                "pipeline_frac": pipe_achieved / peak,
                "ops": [{"op": n, "bytes_per_px": b, "ms": m, "GBps": px_ops * b / (m / 1e3) / 1e9,
                         "frac": px_ops * b / (m / 1e3) / 1e9 / peak} for n, b, m in breakdown],
                "ops_unfused": [{"op": n, "bytes_per_px": b, "ms": m, "GBps": px_ops * b / (m / 1e3) / 1e9,
                                 "frac": px_ops * b / (m / 1e3) / 1e9 / peak} for n, b, m in breakdown_unfused],
            },
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "megapixels/s", "h2d_bytes_per_step": int(h2d_bytes) * (world if strong else world),
                    "d2h_bytes_per_step": int(d2h_bytes) * world, "host_memory": "pageable in, fresh array out",
                    "value_pinned_input": job_px / MP / e2e_pinned_s,
                    "api": "PipelineManager(gpu_executor=B200Executor).apply(ndarray)"},
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
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="c4")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--mosaic-size", type=int, default=65536, help="side of the c4 mosaic (multiple of 64)")
    ap.add_argument("--frames", type=int, default=0, help="c5: total frames of the job (default 1024)")
    ap.add_argument("--no-alt", action="store_true", help="c4, N=1: skip the 8-lock-step-strips comparison run")
    ap.add_argument("--e2e-best", action="store_true", help="c4: report the better of the two e2e passes instead of the mean")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"),
               str(Path(__file__).resolve()), "--gpus", str(args.gpus), "--steps", str(args.steps),
               "--warmup", str(args.warmup), "--workload", args.workload, "--mosaic-size", str(args.mosaic_size),
               "--frames", str(args.frames)] + (["--no-alt"] if args.no_alt else []) + (["--e2e-best"] if args.e2e_best else [])
        return subprocess.call(cmd)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
