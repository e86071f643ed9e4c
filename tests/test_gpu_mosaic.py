"""Row-strip sharded mosaic pipeline vs the dense single-GPU run (bit-exact), with the ranks
emulated on one GPU (host/mosaic.py LocalComm): halo over-fetch, CLAHE LUT gather with global
geometry, global Otsu histogram, cross-strip label merge, reduced region tables."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import np_oracle as O
from yamimageprocessor_b200 import synth
from yamimageprocessor_b200.host import mosaic

pytestmark = pytest.mark.gpu


def dense(backend, frame, p):
    x = backend.to_device(frame)
    g = backend.gaussian(x, p.gauss_ksize, 0.0)
    c = backend.clahe(g, p.clip_limit, p.tile_grid)
    t, om = backend.otsu_threshold(c, 255)
    m = backend.morph_open_close(backend.adaptive_threshold(c, p.block_size, p.C), p.morph_ksize, 1)
    labels, counts = backend.ccl_label(m)
    n = int(backend.to_host(counts)[0])
    props = backend.region_props(labels, c, n)
    return (backend.to_host(c), int(backend.to_host(t)[0]), backend.to_host(om), backend.to_host(labels), n,
            backend.to_host(props))


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_sharded_mosaic_equals_dense(backend, world):
    p = mosaic.MosaicParams()
    frame = synth.nuclei(1024, 768, seed=11)
    # long vertical structures so that components cross several strip boundaries
    frame[:, 100:104] = 60000
    frame[300:800, 400:403] = 55000
    want_c, want_t, want_om, want_lab, want_n, want_props = dense(backend, frame, p)
    res = mosaic.run_emulated(backend, frame, world, p, with_props=True)
    assert [r.rows for r in res] == [mosaic.strip_rows(1024, 8, r, world) for r in range(world)]
    got_c = np.concatenate([backend.to_host(r.clahe) for r in res])
    got_om = np.concatenate([backend.to_host(r.otsu_mask) for r in res])
    got_lab = np.concatenate([backend.to_host(r.labels) for r in res])
    assert all(r.otsu_threshold == want_t for r in res)
    assert all(r.n_components == want_n for r in res)
    assert np.array_equal(got_c, want_c), "CLAHE output differs from the dense run"
    assert np.array_equal(got_om, want_om), "Otsu mask differs"
    assert np.array_equal(got_lab, want_lab), "labels differ"
    for r in res:
        assert np.array_equal(backend.to_host(r.props), want_props), "reduced region table differs"


def test_dense_chain_matches_oracle(backend):
    """anchor: the dense chain the sharded path is compared with equals the CPU oracle"""
    p = mosaic.MosaicParams()
    frame = synth.nuclei(512, 512, seed=12)
    c, t, om, lab, n, props = dense(backend, frame, p)
    oc = O.clahe(O.gaussian_fixed(frame, 11, 0.0), 2.0, (8, 8))
    ot, oom = O.otsu_threshold(oc, 255)
    olab = O.ccl_label(O.morph_close(O.morph_open(O.adaptive_threshold(oc, 11, 2), "Rectangular", 5, 1), "Rectangular", 5, 1))
    assert np.array_equal(c, oc) and t == ot and np.array_equal(om, oom)
    assert n == olab[0] and np.array_equal(lab, olab[1])


def test_strip_geometry_errors():
    with pytest.raises(ValueError):
        mosaic.strip_rows(1000, 8, 0, 3)
    with pytest.raises(ValueError):
        mosaic.strip_rows(1001, 8, 0, 2)
    assert mosaic.input_rows(1024, 1, 2) == (512 - 18, 1024)
