"""Plugin contract (register_module / ModuleBase), parameter sanitising and cache-key parity."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import pytest

from yamimageprocessor_b200.host import cache_keys
from yamimageprocessor_b200.host.executor import B200Executor
from yamimageprocessor_b200.host.params import MODULE_PARAMS, ensure_odd
from yamimageprocessor_b200.host.pipeline import PipelineStep
from yamimageprocessor_b200.host.plugin import ModuleBase, ModuleStage
from yamimageprocessor_b200.host.steps import DEVICE_STEPS
from yamimageprocessor_b200.modules import b200_backend as plugin

GOLD = Path(__file__).resolve().parent / "golden"


class FakeAppCore:
    def __init__(self):
        self.registered = []

    def register_module(self, cls):
        if not (isinstance(cls, type) and issubclass(cls, ModuleBase)):
            raise TypeError("module_cls must be a ModuleBase subclass")  # core/app_core.py:753-757
        self.registered.append(cls)


def test_register_module_registers_every_operator():
    core = FakeAppCore()
    plugin.register_module(core)
    ids = [cls().metadata.identifier for cls in core.registered]
    assert len(ids) == len(set(ids)) == len(plugin.MODULE_CLASSES)
    for expected in ("Grayscale", "NoiseReduction", "IntensityNormalization", "BrightnessContrast", "Gamma",
                     "Otsu", "Adaptive", "Opening", "Closing", "Dilation", "Erosion", "RegionLabels"):
        assert expected in ids
    assert set(ids) == set(DEVICE_STEPS)  # every registered step has a kernel, and vice versa
    for i in ids:
        assert B200Executor.supports(i)


def test_pipeline_steps_keep_reference_names_params_and_are_gpu_marked():
    mods = {cls().metadata.identifier: cls() for cls in plugin.MODULE_CLASSES}
    step = mods["NoiseReduction"].create_pipeline_step()
    assert step.name == "NoiseReduction" and step.params == {"method": "Gaussian", "ksize": 5}
    assert step.execution.requires_gpu and not step.execution.supports_inplace and step.enabled is False
    assert step.stage is ModuleStage.PREPROCESSING and step.function == mods["NoiseReduction"].process
    assert mods["Opening"].create_pipeline_step().params == {"kernel_shape": "Rectangular", "kernel_size": 3, "iterations": 1}
    assert mods["Adaptive"].create_pipeline_step().params == {"block_size": 11, "C": 2}
    assert mods["Adaptive"].metadata.stage is ModuleStage.SEGMENTATION
    assert mods["RegionLabels"].metadata.stage is ModuleStage.ANALYSIS
    assert "Region Properties" not in mods  # the reference step of that name returns an annotated image


def test_sanitize_parameters_matches_reference_registry():
    mods = {cls().metadata.identifier: cls() for cls in plugin.MODULE_CLASSES}
    nr = mods["NoiseReduction"]
    assert nr.sanitize_parameters({"ksize": 4}) == {"method": "Gaussian", "ksize": 5}      # _ensure_odd
    assert nr.sanitize_parameters({"ksize": 99})["ksize"] == 15                             # clamp 1..15
    assert nr.sanitize_parameters({"method": "Nope"})["method"] == "Gaussian"               # choices
    assert nr.sanitize_parameters({"ksize": "abc"})["ksize"] == 5                           # bad value -> default
    assert mods["Adaptive"].sanitize_parameters({"block_size": 200, "C": -99}) == {"block_size": 101, "C": -10}
    assert mods["BrightnessContrast"].sanitize_parameters({"alpha": 1.23456}) == {"alpha": 1.23, "beta": 0}
    assert ensure_odd(6) == 7 and ensure_odd(7.2) == 7
    assert set(MODULE_PARAMS) == set(DEVICE_STEPS)


def test_cache_keys_match_reference_golden():
    meta = json.loads((GOLD / "reference_meta.json").read_text())
    a = np.arange(16, dtype=np.uint16).reshape(4, 4)
    sid = cache_keys.source_id(a)
    assert sid == meta["cache"]["source_id"]
    mods = {cls().metadata.identifier: cls() for cls in plugin.MODULE_CLASSES}
    s0 = mods["Grayscale"].create_pipeline_step(); s0.enabled = True
    s1 = mods["NoiseReduction"].create_pipeline_step(); s1.enabled = True
    s2 = mods["IntensityNormalization"].create_pipeline_step(); s2.enabled = False
    final, recs = cache_keys.predict(sid, [s0, s1, s2])   # GPU-backed steps, default params
    assert [r.signature for r in recs] == meta["cache"]["signatures"] and final == meta["cache"]["final"]
    op = mods["Opening"].create_pipeline_step(); op.params["kernel_size"] = 5; op.enabled = True
    other = PipelineStep("X", lambda x: x, params={"t": (1, 2), "m": {"b": 1, "a": [2.5, None]}})
    final2, recs2 = cache_keys.predict(sid, [op, other])
    assert [r.signature for r in recs2] == meta["cache2"]["signatures"] and final2 == meta["cache2"]["final"]
    assert cache_keys.dense_cache_filename(sid, final) == f"{sid}_{final}.npy"


def test_unknown_step_has_no_fallback():
    ex = B200Executor(backend=object())
    with pytest.raises(KeyError, match="no CPU fallback"):
        ex.run_on_device("Watershed", None, {})


def test_n3_step_names_and_parameter_names_are_the_references():
    """the names / parameter keys that feed the cache signatures.  Names from
    processing/segmentation_pipeline.py:114-122,178-180 and modules/preprocessing.py:161,182; keys and
    defaults from ui/control_metadata.py:220-245,381-402,677-686."""
    expected = {
        "Sharpen": {"strength": 1.0},
        "SelectChannel": {"channel": "All"},
        "Sobel": {"ksize": 3},
        "Prewitt": {},
        "Laplacian": {"ksize": 3},
        "Border Removal": {"border_distance": 100},
    }
    mods = {cls().metadata.identifier: cls() for cls in plugin.MODULE_CLASSES}
    for name, defaults in expected.items():
        assert name in DEVICE_STEPS and name in mods
        step = mods[name].create_pipeline_step()
        assert step.name == name and dict(step.params) == defaults
    # sanitising follows the reference's coercions (odd kernel sizes, clamped ranges, choices)
    assert mods["Sobel"].sanitize_parameters({"ksize": 4})["ksize"] == 5
    assert mods["Laplacian"].sanitize_parameters({"ksize": 99})["ksize"] == 31
    assert mods["Sharpen"].sanitize_parameters({"strength": 9.0})["strength"] == 5.0
    assert mods["SelectChannel"].sanitize_parameters({"channel": "XYZ"})["channel"] == "All"
    assert mods["Border Removal"].sanitize_parameters({"border_distance": -3})["border_distance"] == 0


def test_executor_bit_run_detection_is_pure_host_logic():
    """_bit_run_length only looks at names, params and tensor metadata (no GPU needed)."""
    import torch

    ex = B200Executor(backend=object())
    mods = {cls().metadata.identifier: cls() for cls in plugin.MODULE_CLASSES}

    def step(name, **params):
        s = mods[name].create_pipeline_step()
        s.params.update(params)
        return s

    plane = torch.empty((8, 8), dtype=torch.uint16)
    colour = torch.empty((8, 8, 3), dtype=torch.uint8)
    chain = [step("Adaptive"), step("Opening", kernel_size=5), step("Closing", kernel_size=5), step("ConnectedComponents"),
             step("BoxFilter")]
    assert ex._bit_run_length(chain, 0, plane) == 4
    assert ex._bit_run_length(chain, 1, plane) == 0                       # a run starts at Adaptive
    assert ex._bit_run_length(chain, 0, colour) == 0                      # colour input: Adaptive converts to gray first
    assert ex._bit_run_length([step("Adaptive", block_size=9), step("Opening")], 0, plane) == 0
    assert ex._bit_run_length([step("Adaptive"), step("Opening", kernel_shape="Elliptical")], 0, plane) == 0
    assert ex._bit_run_length([step("Adaptive"), step("Erosion"), step("Dilation", kernel_shape="Cross")], 0, plane) == 2
    assert ex._bit_run_length([step("Adaptive")], 0, plane) == 0          # nothing to fuse with
    assert ex._bit_run_length([step("Adaptive"), step("ConnectedComponents")], 0, torch.empty((2, 8, 8), dtype=torch.uint8)) == 2
