"""Colour histogram equalisation (core/preprocessing.py:74-79) on the CPU: the oracle's YCrCb fixed point
against live cv2 over ALL 2^24 colours and against the reference's own outputs (tests/golden/
make_golden_color.py), and the kernels' SOURCE (csrc/yam_color.cu) on the host emulation against the oracle."""
from __future__ import annotations

import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import np_oracle as O

ROOT = Path(__file__).resolve().parents[1]
CU = ROOT / "yamimageprocessor_b200" / "csrc" / "yam_color.cu"
GOLD = Path(__file__).resolve().parent / "golden" / "reference_outputs_color.npz"


def all_colours():
    r = np.arange(256, dtype=np.uint8)
    a, b, c = np.meshgrid(r, r, r, indexing="ij")
    return np.stack([a, b, c], axis=-1).reshape(4096, 4096, 3)


def test_ycrcb_fixed_point_matches_live_cv2_on_every_colour():
    cv2 = pytest.importorskip("cv2")
    img = all_colours()
    assert np.array_equal(O.bgr2ycrcb_u8(img), cv2.cvtColor(img, cv2.COLOR_BGR2YCrCb))
    assert np.array_equal(O.ycrcb2bgr_u8(img), cv2.cvtColor(img, cv2.COLOR_YCrCb2BGR))


def test_equalize_hist_bgr_matches_reference_outputs():
    g = np.load(GOLD)
    names = [k[3:] for k in g.files if k.startswith("in_")]
    assert len(names) >= 5
    for name in names:
        assert np.array_equal(O.equalize_hist_bgr(g[f"in_{name}"]), g[f"equalized_{name}"]), name


@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    from cuda_emulation.build import build_emulator

    work = tmp_path_factory.mktemp("color_emu")
    return build_emulator(CU, "color_driver.cc", work, "color_emu"), work


def run_emulator(emulator, bgr, y_new, quads):
    exe, work = emulator
    fin, fout = work / "in.bin", work / "out.bin"
    px = bgr.shape[0]
    with open(fin, "wb") as f:
        f.write(np.array([px], np.int64).tobytes())
        f.write(np.ascontiguousarray(bgr, np.uint8).tobytes())
        f.write(np.ascontiguousarray(y_new, np.uint8).tobytes())
    proc = subprocess.run([str(exe), str(fin), str(fout), "1" if quads else "0"], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, (proc.returncode, proc.stderr[-2000:])
    raw = np.fromfile(fout, np.uint8)
    return raw[:px], raw[px:].reshape(px, 3)


@pytest.mark.parametrize("quads", [True, False])
def test_kernel_source_on_host_emulation(emulator, rng, quads):
    # extremes of every channel plus random colours; px % 4 != 0 so the scalar tail runs too
    r = np.array([0, 1, 2, 63, 64, 127, 128, 129, 200, 254, 255], np.uint8)
    a, b, c = np.meshgrid(r, r, r, indexing="ij")
    bgr = np.concatenate([np.stack([a, b, c], -1).reshape(-1, 3), rng.integers(0, 256, (70001, 3), dtype=np.uint8)])
    bgr = bgr[: bgr.shape[0] - (1 if bgr.shape[0] % 4 == 0 else 0)]
    assert bgr.shape[0] % 4 != 0
    y_new = rng.integers(0, 256, bgr.shape[0], dtype=np.uint8)
    y_new[:64] = np.tile(np.array([0, 255], np.uint8), 32)
    y, dst = run_emulator(emulator, bgr, y_new, quads)
    ycc = O.bgr2ycrcb_u8(bgr[None])
    assert np.array_equal(y, ycc[0, :, 0])
    ycc[0, :, 0] = y_new
    assert np.array_equal(dst, O.ycrcb2bgr_u8(ycc)[0])
