"""CPU checks of the C-ABI library: it loads, exports every symbol the header declares, and its
host-only helpers (tap generation, structuring elements, Otsu scan, median networks) agree with
the oracle.  No compute call needs a GPU here."""
from __future__ import annotations

import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import np_oracle as O
from yamimageprocessor_b200 import _lib

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


def test_header_parses_and_library_exports_every_symbol(lib):
    protos = _lib.parse_header()
    assert len(protos) >= 30
    for name in protos:
        assert hasattr(lib, name), f"libyamb200.so does not export {name}"
    assert lib.yam_abi_version() == 1
    # and nothing exported with the yam_ prefix is undeclared
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T yam_" in ln}
    assert exported == set(protos), exported ^ set(protos)


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    archs = {ln.split(".")[-2] for ln in out.splitlines() if "sm_" in ln}
    assert archs == {"sm_100a"}, archs


def test_context_creation_fails_loudly_without_gpu(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert lib.yam_ctx_create(0, C.byref(h)) != 0
    assert "no CPU fallback" in _lib.last_error()
    from yamimageprocessor_b200.backend import Backend, BackendUnavailable

    with pytest.raises(BackendUnavailable):
        Backend(0)


def test_gaussian_taps(lib):
    for k in range(1, 32, 2):
        for sigma in (0.0, 0.5, 1.0, 2.0, 2.5, 3.3):
            kf = (C.c_double * k)()
            assert lib.yam_gaussian_taps_f64(k, sigma, C.cast(kf, C.c_void_p)) == 0
            want = O.gaussian_kernel(k, sigma)
            assert np.array_equal(np.array(kf, np.float64).astype(np.float32), want.astype(np.float32))
            for bits in (8, 16):
                kq = (C.c_int64 * k)()
                assert lib.yam_gaussian_taps_fixed(k, sigma, bits, C.cast(kq, C.c_void_p)) == 0
                assert list(kq) == O.fixed_kernel(want, bits).tolist()
                assert sum(kq) == 1 << bits
    assert lib.yam_gaussian_taps_f64(4, 0.0, None) != 0
    assert "odd" in _lib.last_error()


def test_structuring_elements(lib):
    for code, name in ((0, "Rectangular"), (1, "Elliptical"), (2, "Cross")):
        for k in range(1, 32):
            buf = (C.c_uint8 * (k * k))()
            assert lib.yam_structuring_element(code, k, C.cast(buf, C.c_void_p)) == 0
            assert np.array_equal(np.array(buf, np.uint8).reshape(k, k), O.structuring_element(name, k))


def test_otsu_scan(lib, rng):
    for bins in (256, 65536):
        for _ in range(4):
            centres = rng.integers(0, bins, 3)
            vals = np.clip(np.concatenate([rng.normal(c, bins / 20, 4000) for c in centres]), 0, bins - 1).astype(np.int64)
            h = np.bincount(vals, minlength=bins).astype(np.uint64)
            t = C.c_int()
            assert lib.yam_otsu_from_hist(h.ctypes.data_as(C.c_void_p), bins, C.cast(C.byref(t), C.c_void_p)) == 0
            assert t.value == O.otsu_from_hist(h.astype(np.int64))
    # counts beyond 2^31 (cv2 itself overflows there): 64-bit recurrence stays finite and ordered
    h = np.zeros(65536, np.uint64)
    h[1000], h[50000] = 3_000_000_000, 2_500_000_000
    t = C.c_int()
    assert lib.yam_otsu_from_hist(h.ctypes.data_as(C.c_void_p), 65536, C.cast(C.byref(t), C.c_void_p)) == 0
    assert t.value == O.otsu_from_hist(h.astype(np.int64)) == 1000


def test_median_networks_exhaustive(tmp_path):
    """zero-one principle: a comparator network selects the median iff it does on all 0/1 inputs."""
    src = tmp_path / "mednet.cpp"
    src.write_text(
        '#include <cstdio>\n#include <cstdint>\n#include "yam_median_net.h"\n'
        "int main(){long bad=0;"
        "for(uint32_t m=0;m<(1u<<9);m++){uint32_t p[9];int c=0;for(int i=0;i<9;i++){p[i]=(m>>i)&1;c+=p[i];}"
        "if(median9(p)!=(uint32_t)(c>=5))bad++;}"
        "for(uint32_t m=0;m<(1u<<25);m++){uint32_t p[25];int c=0;for(int i=0;i<25;i++){p[i]=(m>>i)&1;c+=p[i];}"
        "if(median25(p)!=(uint32_t)(c>=13))bad++;}"
        'printf("%ld\\n",bad);return bad!=0;}\n'
    )
    exe = tmp_path / "mednet"
    subprocess.run(["g++", "-O2", "-I", str(ROOT / "yamimageprocessor_b200" / "csrc"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip()
    assert out == "0"
