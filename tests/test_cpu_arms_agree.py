"""The two CPU arms of bench.py compute the same thing: ``oracle/ref_path.py`` driving the UNMODIFIED reference
(``PipelineManager.apply`` over its own step functions; build container only) and ``oracle/cv2_path.py``, the port
that is timed on the GPU box where ``/root/reference`` does not exist -- and both equal the NumPy oracle the CUDA
path is checked against.  So `cpu_baseline.kind` "port" and "reference" time the same arithmetic."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import np_oracle as O
from yamimageprocessor_b200 import synth

cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def frame():
    f = synth.nuclei(520, 640, seed=77)
    f[:, 300:303] = 60000       # a component that spans the frame
    return f


def test_port_equals_numpy_oracle(frame):
    from oracle import cv2_path as P

    g, otsu_mask = P.preprocess(frame)
    og = O.clahe(O.gaussian_fixed(frame, 11, 0.0), 2.0, (8, 8))
    assert np.array_equal(g, og) and np.array_equal(otsu_mask, O.otsu_threshold(og, 255)[1])
    lab = P.segment(g)
    on, olab = O.ccl_label(O.morph_close(O.morph_open(O.adaptive_threshold(og, 11, 2), "Rectangular", 5, 1), "Rectangular", 5, 1))
    assert int(lab.max()) == on and np.array_equal(O.canonicalise_labels(lab), olab)
    table = P.extract(olab, g)
    assert np.array_equal(table["area"], np.bincount(olab.ravel())[1:])


def test_unmodified_reference_arm_equals_port(frame):
    from oracle import cv2_path as P
    from oracle import ref_path as R

    kind, detail = R.describe()
    if kind != "reference":
        pytest.skip(f"reference checkout not importable here: {detail}")
    g_ref, mask_ref = R.preprocess(frame)
    g_port, mask_port = P.preprocess(frame)
    assert g_ref.dtype == g_port.dtype and np.array_equal(g_ref, g_port) and np.array_equal(mask_ref, mask_port)
    assert np.array_equal(R.segment(g_ref), P.segment(g_port))
    for a, b in zip(R.mosaic_chain(frame), P.mosaic_chain(frame)):
        assert np.array_equal(a, b)
    ref_full, port_full = R.full_chain(frame), P.full_chain(frame)
    assert np.array_equal(ref_full[0], port_full[0]) and np.array_equal(ref_full[1], port_full[1])
    for key in ("area", "bbox", "sum_intensity"):
        assert np.array_equal(ref_full[2][key], port_full[2][key]), key
