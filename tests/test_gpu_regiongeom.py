"""perimeter / solidity columns of region_properties_data (core/extraction.py:80,83) on the GPU:
``yam_region_perimeter`` and ``yam_region_convex_area`` against the oracle's literal restatement of
skimage.measure.perimeter / convex_hull_image (oracle/np_oracle.py) and against the integer model of the
device algorithm (tests/region_geometry_model.py).  Integer tables are bit-exact; the float64 columns are
formed from them on the host (tolerance 1e-12 relative: summation order only)."""
from __future__ import annotations

import numpy as np
import pytest
from scipy import ndimage as ndi

import region_geometry_model as M
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu


def device_tables(backend, lab, n):
    labels = backend.to_device(np.ascontiguousarray(lab, np.int32))
    props = backend.region_props(labels, None, n)
    counts = backend.to_host(backend.region_perimeter_counts(labels, n))
    convex = backend.to_host(backend.region_convex_area(labels, n, props))
    return backend.to_host(props), counts, convex


def test_perimeter_and_convex_area_small_cases(backend, rng):
    cases = []
    for it in range(12):
        h, w = (int(v) for v in rng.integers(5, 120, 2))
        m = rng.random((h, w)) < float(rng.choice([0.3, 0.5, 0.7, 0.95]))
        if it % 2:
            m = ndi.binary_dilation(rng.random((h, w)) < 0.02, iterations=int(rng.integers(1, 6)))
        lab, n = ndi.label(m, structure=np.ones((3, 3)))
        if n:
            cases.append((lab.astype(np.int32), int(n)))
    cases.append((rng.integers(0, 4, (23, 31)).astype(np.int32), 5))      # unconnected labels, empty table rows
    cases.append((np.ones((1, 1), np.int32), 1))
    cases.append((np.ones((300, 7), np.int32), 1))                        # one tall region: a long chain
    for lab, n in cases:
        props, counts, convex = device_tables(backend, lab, n)
        assert np.array_equal(counts, M.perimeter_counts(lab, n))
        assert np.array_equal(convex, M.convex_area_model(lab, n, props))
        ref = O.region_perimeter_solidity(lab, n)
        assert np.array_equal(convex, ref["area_convex"])
        assert np.allclose(M.perimeter_from_counts(counts), ref["perimeter"], rtol=1e-12, atol=1e-12)


def test_region_table_perimeter_solidity_nuclei(backend):
    from yamimageprocessor_b200 import synth
    from yamimageprocessor_b200.host.steps import region_table

    frame = synth.nuclei(1040, 1300, seed=9)
    m = O.morph_close(O.morph_open(O.adaptive_threshold(frame, 11, 2), "Rectangular", 5, 1), "Rectangular", 5, 1)
    m[100:104, 50:900] = 255          # elongated regions crossing many perimeter tiles
    m[200:900, 300:302] = 255
    n, lab = O.ccl_label(m)
    assert n > 1500
    table = region_table(backend, backend.to_device(lab), backend.to_device(frame), n)
    ref = O.region_perimeter_solidity(lab, n)
    assert np.array_equal(table["area_convex"], ref["area_convex"])
    assert np.allclose(table["perimeter"], ref["perimeter"], rtol=1e-12, atol=1e-12)
    assert np.allclose(table["solidity"], ref["solidity"], rtol=1e-12, atol=0)
    assert {"perimeter", "solidity", "extent", "eccentricity", "orientation", "area", "centroid"} <= set(table)


def test_convex_area_rejects_foreign_props(backend):
    import torch

    labels = backend.to_device(np.ones((8, 8), np.int32))
    with pytest.raises(TypeError):
        backend.region_convex_area(labels, 1, torch.zeros((1, 4), dtype=torch.int64, device=backend.device))
