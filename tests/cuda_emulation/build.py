"""Compile the anonymous namespace of a .cu file (its kernels and device helpers) for the HOST, on top of
emu_prelude.h, together with a driver that launches them: logic checks of CUDA source without a GPU.
Test infrastructure only."""
from __future__ import annotations

import shutil
import subprocess
from pathlib import Path

import pytest

EMU = Path(__file__).resolve().parent


def kernel_namespace(cu: Path) -> str:
    text = cu.read_text()
    start = text.index("namespace {") + len("namespace {")
    end = text.index("}  // namespace")
    return text[start:end]


def build_emulator(cu: Path, driver: str, workdir: Path, name: str) -> Path:
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    src = workdir / f"{name}.cc"
    src.write_text('#include "emu_prelude.h"\n' + kernel_namespace(cu) + "\n" + (EMU / driver).read_text())
    exe = workdir / name
    # AddressSanitizer + UBSan: the emulated kernels run on exactly-sized host buffers (std::vector) and static
    # "shared memory" arrays, so an out-of-bounds index or a signed overflow in the kernel source aborts the run
    # (compute-sanitizer is not available on the GPU pool; this is the memcheck of the kernel logic)
    base = ["g++", "-std=c++17", "-O1", "-g", "-pthread", "-Wno-unknown-pragmas", f"-I{EMU}", str(src), "-o", str(exe)]
    proc = subprocess.run(base[:5] + ["-fsanitize=address,undefined", "-fno-sanitize-recover=all"] + base[5:], capture_output=True, text=True)
    if proc.returncode != 0:   # a toolchain without the sanitizer runtimes: plain build
        proc = subprocess.run(base, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr[-3000:]
    return exe
