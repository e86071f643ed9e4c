"""Conformance against the REAL reference AppCore (core/app_core.py:680-793), not a fake.

Needs the reference checkout (this container: /root/reference; absent on the GPU box -> skipped).
The probe runs in a subprocess: the plugin module chooses its ModuleBase at import time, so the
reference packages must be importable first (exactly the situation inside the application)."""
from __future__ import annotations

import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

REF = Path(os.environ.get("YAM_REFERENCE", "/root/reference"))
HERE = Path(__file__).resolve().parent
pytestmark = pytest.mark.skipif(not (REF / "core" / "app_core.py").exists(), reason="reference checkout not present")

# identifiers of the hot path that exist in the reference's own plugin package too (modules/preprocessing.py:46-182)
SHARED = ("Grayscale", "BrightnessContrast", "Gamma", "IntensityNormalization", "NoiseReduction", "Sharpen", "SelectChannel")


def _probe(order: str) -> dict:
    proc = subprocess.run([sys.executable, str(HERE / "ref_appcore_probe.py"), str(REF), order],
                          capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stderr[-3000:]
    line = [l for l in proc.stdout.splitlines() if l.startswith("PROBE ")][-1]
    return json.loads(line[6:])


def test_real_appcore_resolves_every_hot_path_identifier_to_the_gpu_module():
    from yamimageprocessor_b200.host.steps import DEVICE_STEPS

    got = _probe("b200_first")   # the order INTEGRATION.md documents
    cat = got["catalogue"]
    for ident in DEVICE_STEPS:
        assert ident in cat, ident
        assert cat[ident]["module"] == "yamimageprocessor_b200.modules.b200_backend", (ident, cat[ident])
        # Mosaic consumes the lazy handle itself (supports_tiled_input); _run_step would densify it for an
        # executor (processing/pipeline_manager.py:449-454), so it alone is not executor-routed
        assert cat[ident]["requires_gpu"] is (ident != "Mosaic")
        assert cat[ident]["tiled"] is (ident == "Mosaic")
    assert cat["Crop"]["module"] == "modules.preprocessing"           # the reference's own module is still there
    # get_pipeline_manager() builds, every hot-path step is GPU-marked, stage order preprocessing -> segmentation -> analysis
    steps = got["steps"]
    assert set(DEVICE_STEPS) <= set(steps)
    assert all(got["step_requires_gpu"][n] for n in DEVICE_STEPS if n != "Mosaic")
    assert steps.index("NoiseReduction") < steps.index("Adaptive") < steps.index("RegionLabels")
    # PipelineCache.predict over the GPU-backed steps reproduces SURVEY.md App. B
    meta = json.loads((HERE / "golden" / "reference_meta.json").read_text())
    assert got["cache"]["source_id"] == meta["cache"]["source_id"]
    assert got["cache"]["signatures"] == meta["cache"]["signatures"] and got["cache"]["final"] == meta["cache"]["final"]


def test_real_pipeline_manager_routes_gpu_steps_to_the_executor():
    """The reference's own PipelineManager (built by its AppCore from this package's modules) hands every enabled
    GPU-marked step to B200Executor.execute(step, ndarray), in order, and takes the returned array as the
    pipeline image (processing/pipeline_manager.py:448-454); set_gpu_executor / clone keep the executor."""
    got = _probe("b200_first")["executor"]
    assert got["calls"] == got["enabled_in_order"]        # the manager's own step order
    assert sorted(got["calls"]) == sorted(["Grayscale", "NoiseReduction", "Otsu", "Opening", "ConnectedComponents"])
    assert got["result_sum"] == 5 * 6 * 7        # five steps, each +1 on a 6 x 7 zero image
    assert got["input_untouched"] and got["clone_keeps_executor"]


def test_real_appcore_duplicate_rule_is_why_the_order_matters():
    """core/app_core.py:762-771: the FIRST registration of an identifier wins, later ones are dropped
    with a warning.  With the reference package listed first its seven CPU modules shadow ours."""
    got = _probe("modules_first")
    for ident in SHARED:
        assert got["catalogue"][ident]["module"] == "modules.preprocessing"
        assert got["catalogue"][ident]["requires_gpu"] is False
    assert got["catalogue"]["Adaptive"]["requires_gpu"] is True        # identifiers the reference lacks still register


def test_fake_appcore_with_the_reference_duplicate_rule():
    """Same rule on a stand-in (runs everywhere, also without the checkout present in-process)."""
    from yamimageprocessor_b200.host.plugin import ModuleBase
    from yamimageprocessor_b200.modules import b200_backend as plugin

    class Core:
        def __init__(self):
            self.catalog = {}

        def register_module(self, cls):
            if not (isinstance(cls, type) and issubclass(cls, ModuleBase)):
                raise TypeError
            m = cls()
            stage = self.catalog.setdefault(m.metadata.stage, {})
            if m.metadata.identifier in stage:
                return                                                 # duplicate ignored
            stage[m.metadata.identifier] = m

    core = Core()
    plugin.register_module(core)
    n = sum(len(v) for v in core.catalog.values())
    plugin.register_module(core)                                       # registering twice changes nothing
    assert sum(len(v) for v in core.catalog.values()) == n == len(plugin.MODULE_CLASSES)
