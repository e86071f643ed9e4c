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
    return lambda fr, aux: [P.full_chain(f) for f in fr], P


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
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = Path(f"/tmp/yam_clocks_{os.getpid()}.csv")

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx = [], []
        reasons = set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.path.read_text().splitlines():
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        try:
            self.path.unlink()
        except OSError:
            pass
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
                ("otsu_threshold_u16 (hist + scan + threshold)", 6.0, lambda: be.otsu_threshold(c, 255)),
            ]
        return x, run, ops

    if workload == "c2":
        def run(inp):
            m = be.adaptive_threshold(inp, 11, 2)
            m = be.morph_open_close(m, 5, 1)
            return be.ccl_label(m)[0]

        def ops(inp):
            m = be.adaptive_threshold(inp, 11, 2)
            m2 = be.morph_open_close(m, 5, 1)
            return [
                ("adaptive_threshold_u16_b11 (sep_f32_kernel)", 3.0, lambda: be.adaptive_threshold(inp, 11, 2)),
                ("morph_open_close_5x5_u8 (morph_rect_chain_kernel)", 4.0, lambda: be.morph_open_close(m, 5, 1)),
                ("ccl_label (pack, union, flatten, scan, prefix, final)", 5.0, lambda: be.ccl_label(m2)),
            ]
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
        m = be.adaptive_threshold(c, 11, 2)
        m = be.morph_open_close(m, 5, 1)
        labels, counts = be.ccl_label(m)
        cnt = be.to_host(counts)
        tables = []
        for i in range(labels.shape[0]):
            tables.append(be.region_props(labels[i], c[i], int(cnt[i])))
        return otsu_mask, labels, tables

    def ops(inp):
        g = be.gaussian(inp, 11, 0.0)
        c = be.clahe(g, 2.0, (8, 8))
        m = be.adaptive_threshold(c, 11, 2)
        m2 = be.morph_open_close(m, 5, 1)
        return [
            ("gaussian_fixed_u16_k11", 4.0, lambda: be.gaussian(inp, 11, 0.0)),
            ("clahe_u16 (lut + apply)", 6.0, lambda: be.clahe(g, 2.0, (8, 8))),
            ("otsu_threshold_u16 (hist + scan + threshold)", 6.0, lambda: be.otsu_threshold(c, 255)),
            ("adaptive_threshold_u16_b11 (sep_f32_kernel)", 3.0, lambda: be.adaptive_threshold(c, 11, 2)),
            ("morph_open_close_5x5_u8 (morph_rect_chain_kernel)", 4.0, lambda: be.morph_open_close(m, 5, 1)),
            ("ccl_label (pack, union, flatten, scan, prefix, final)", 5.0, lambda: be.ccl_label(m2)),
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


def run_gpu(args):
    import torch
    import torch.distributed as dist

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
    breakdown = []
    for name, bpp, fn in ops(inp):
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
        ms = statistics.mean(a.elapsed_time(b) for a, b in evs)
        breakdown.append((name, bpp, ms))

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
                "traffic": None,
                "peak_source": peak_src,
                "pipeline_achieved": pipe_achieved,
                "pipeline_frac": pipe_achieved / peak,
                "ops": [{"op": n, "bytes_per_px": b, "ms": m, "GBps": px_per_rank * b / (m / 1e3) / 1e9,
                         "frac": px_per_rank * b / (m / 1e3) / 1e9 / peak} for n, b, m in breakdown],
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
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"),
               str(Path(__file__).resolve()), "--gpus", str(args.gpus), "--steps", str(args.steps),
               "--warmup", str(args.warmup), "--workload", args.workload]
        return subprocess.call(cmd)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
