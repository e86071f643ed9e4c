"""perimeter / solidity columns (core/extraction.py:80,83) on the CPU:
* the oracle's literal restatement of skimage's perimeter and convex_hull_image on analytic shapes;
* the integer model of the device algorithm (tests/region_geometry_model.py) against that oracle;
* the kernels' SOURCE (csrc/yam_regiongeom.cu) compiled for the host through a small CUDA emulation
  (tests/cuda_emulation/) against both -- logic check of the code the GPU runs, no GPU needed."""
from __future__ import annotations

import math
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest
from scipy import ndimage as ndi

import region_geometry_model as M
from oracle import np_oracle as O

ROOT = Path(__file__).resolve().parents[1]
CU = ROOT / "yamimageprocessor_b200" / "csrc" / "yam_regiongeom.cu"
EMU = Path(__file__).resolve().parent / "cuda_emulation"


def props_of(labels, n):
    p = O.region_props(labels, None, n)
    out = np.zeros((n, 8), np.int64)
    out[:, 0] = p["area"]
    out[:, 1], out[:, 2] = p["sum_y"], p["sum_x"]
    out[:, 4:8] = np.where(p["area"][:, None] > 0, p["bbox"], 0)
    return out


def label_cases(rng, count=40, lo=5, hi=70):
    for it in range(count):
        h, w = (int(v) for v in rng.integers(lo, hi, 2))
        dens = float(rng.choice([0.2, 0.4, 0.5, 0.6, 0.8, 0.95]))
        m = rng.random((h, w)) < dens
        if it % 3 == 0:
            m = ndi.binary_opening(m) | (rng.random((h, w)) < 0.02)
        elif it % 3 == 1:
            m = ndi.binary_dilation(rng.random((h, w)) < 0.03, iterations=int(rng.integers(1, 5)))
        lab, n = ndi.label(m, structure=np.ones((3, 3)))
        if n:
            yield lab.astype(np.int32), int(n)
    for _ in range(6):   # arbitrary label images: regions need not be connected, rows may be missing
        shape = (int(rng.integers(3, 30)), int(rng.integers(3, 30)))
        lab = rng.integers(0, 4, shape).astype(np.int32)
        yield lab, 3 + int(rng.integers(0, 3))     # n beyond the largest label: empty table rows


def test_oracle_analytic_shapes():
    sq = np.ones((5, 5), np.uint8)
    assert O.perimeter4(sq) == 16.0 and O.convex_area(sq) == 25           # 4 (n - 1) for a filled square
    r = np.ones((4, 9), np.uint8)
    assert O.perimeter4(r) == 2 * 3 + 2 * 8 and O.convex_area(r) == 36
    yy, xx = np.mgrid[0:7, 0:7]
    diamond = (abs(yy - 3) + abs(xx - 3) <= 3).astype(np.uint8)
    assert math.isclose(O.perimeter4(diamond), 12 * math.sqrt(2)) and O.convex_area(diamond) == 25
    assert O.perimeter4(np.ones((1, 1))) == 0.0 and O.convex_area(np.ones((1, 1))) == 1
    line = np.ones((1, 6), np.uint8)
    assert O.perimeter4(line) == 4.0 and O.convex_area(line) == 6          # end pixels class 3 weigh nothing
    ell = np.array([[1, 1, 1], [1, 0, 0], [1, 0, 0]], np.uint8)
    assert O.convex_area(ell) == 6                                         # the anti-diagonal is on the hull edge
    two = np.array([[1, 0, 0], [0, 0, 0], [0, 0, 1]], np.uint8)
    assert O.convex_area(two) == 3


def test_model_matches_oracle(rng):
    regions = 0
    for lab, n in label_cases(rng):
        props = props_of(lab, n)
        ref = O.region_perimeter_solidity(lab, n)
        counts = M.perimeter_counts(lab, n)
        assert np.allclose(M.perimeter_from_counts(counts), ref["perimeter"], rtol=1e-13, atol=1e-12)
        assert np.array_equal(M.convex_area_model(lab, n, props), ref["area_convex"])
        regions += n
    assert regions > 500


def test_kernel_class_table_matches_model():
    text = CU.read_text()
    m = re.search(r"kPerimeterClass\[50\]\s*=\s*\{(.*?)\};", text, flags=re.S)
    body = re.sub(r"/\*.*?\*/", " ", m.group(1))
    table = np.array([int(v) for v in body.replace("\n", " ").split(",") if v.strip()], np.int64)
    assert table.shape == (50,) and np.array_equal(table, M.PERIMETER_CLASS)


@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    from cuda_emulation.build import build_emulator

    work = tmp_path_factory.mktemp("regiongeom_emu")
    return build_emulator(CU, "regiongeom_driver.cc", work, "regiongeom_emu"), work


def run_emulator(emulator, lab, n, props):
    exe, work = emulator
    fin, fout = work / "in.bin", work / "out.bin"
    with open(fin, "wb") as f:
        f.write(np.array([lab.shape[0], lab.shape[1], n], np.int64).tobytes())
        f.write(np.ascontiguousarray(lab, np.int32).tobytes())
        f.write(np.ascontiguousarray(props, np.int64).tobytes())
    proc = subprocess.run([str(exe), str(fin), str(fout)], capture_output=True, text=True, timeout=120)
    assert proc.returncode == 0, (proc.returncode, proc.stderr[-2000:])
    raw = np.fromfile(fout, np.int64)
    return raw[:3 * n].reshape(n, 3), raw[3 * n:]


def test_kernel_source_on_host_emulation(emulator, rng):
    regions = 0
    for lab, n in label_cases(rng, count=12, lo=20, hi=100):
        props = props_of(lab, n)
        counts, convex = run_emulator(emulator, lab, n, props)
        assert np.array_equal(counts, M.perimeter_counts(lab, n))
        ref = O.region_perimeter_solidity(lab, n)
        assert np.allclose(M.perimeter_from_counts(counts), ref["perimeter"], rtol=1e-13, atol=1e-12)
        assert np.array_equal(convex, ref["area_convex"])
        regions += n
    assert regions > 150


def test_host_contour_columns_from_integer_tables(rng):
    """backend.contour_columns (the float64 host step behind region_table) fed with the model's integer tables"""
    from yamimageprocessor_b200.backend import contour_columns

    lab, n = next(iter(label_cases(rng, count=3, lo=40, hi=80)))
    props = props_of(lab, n)
    cols = contour_columns(props, M.perimeter_counts(lab, n), M.convex_area_model(lab, n, props))
    ref = O.region_perimeter_solidity(lab, n)
    assert np.allclose(cols["perimeter"], ref["perimeter"], rtol=1e-13, atol=1e-12)
    assert np.array_equal(cols["area_convex"], ref["area_convex"])
    assert np.allclose(cols["solidity"], ref["solidity"], rtol=1e-15)
