#!/usr/bin/env python
"""Bootstraps the UNMODIFIED reference AppCore with this package's plugin (run as a subprocess by
tests/test_reference_appcore.py so the reference's own ModuleBase is importable BEFORE the plugin
module is imported -- host/plugin.py:binding() picks the base class at import time).

    python tests/ref_appcore_probe.py /root/reference <order>      order = b200_first | b200_only | modules_first

Prints one JSON object: the module catalogue (identifier -> defining module, requires_gpu, stage),
the unified pipeline's step names, and PipelineCache.predict signatures for SURVEY.md App. B's chain.
"""
import json
import sys
import tempfile
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent


def main() -> None:
    ref, order = sys.argv[1], sys.argv[2]
    sys.path.insert(0, str(REPO))
    sys.path.insert(0, ref)
    from oracle.ref_stubs import install_stubs

    install_stubs()  # PyQt5 / skimage stubs of SURVEY.md App. C

    import numpy as np
    from core.app_core import AppConfiguration, AppCore  # the reference's own classes
    from plugins.module_base import ModuleStage

    packages = {"b200_first": ["yamimageprocessor_b200.modules", "modules"],
                "b200_only": ["yamimageprocessor_b200.modules"],
                "modules_first": ["modules", "yamimageprocessor_b200.modules"]}[order]
    tmp = Path(tempfile.mkdtemp(prefix="yam_appcore_"))
    core = AppCore(AppConfiguration(plugin_packages=packages, log_directory=tmp / "logs", autosave_directory=tmp / "auto",
                                    allowed_roots=[tmp], session_temp_parent=tmp))
    core.bootstrap()
    catalogue = {}
    for stage in ModuleStage:
        for ident, mod in core._module_catalog.get(stage, {}).items():
            catalogue[ident] = {"module": type(mod).__module__, "stage": stage.value,
                                "requires_gpu": bool(mod.pipeline_execution_metadata().requires_gpu),
                                "tiled": bool(mod.supports_tiled_input())}
    pm = core.get_pipeline_manager()
    steps = {s.name: s for s in pm.steps}
    out = {"catalogue": catalogue, "steps": [s.name for s in pm.steps],
           "step_requires_gpu": {s.name: bool(s.execution.requires_gpu) for s in pm.steps}}
    if all(n in steps for n in ("Grayscale", "NoiseReduction", "IntensityNormalization")):
        cache = core.pipeline_cache if hasattr(core, "pipeline_cache") else None
        from processing.pipeline_cache import PipelineCache

        cache = PipelineCache()
        a = np.arange(16, dtype=np.uint16).reshape(4, 4)
        sid = cache.register_source(a)
        chain = [steps["Grayscale"].clone(), steps["NoiseReduction"].clone(), steps["IntensityNormalization"].clone()]
        chain[0].enabled, chain[1].enabled, chain[2].enabled = True, True, False
        final, records = cache.predict(sid, chain)
        out["cache"] = {"source_id": sid, "final": final, "signatures": [r.signature for r in records]}
    # The GpuExecutor protocol through the reference's OWN manager (processing/pipeline_manager.py:448-454): every
    # enabled requires_gpu step must reach execute(step, ndarray) under a name DEVICE_STEPS knows, in order, and the
    # returned array must become the pipeline image.  No GPU in this container: the dispatch is recorded, the
    # kernels themselves are covered by the -m gpu suite.
    from yamimageprocessor_b200.host.executor import B200Executor

    class RecordingExecutor(B200Executor):
        def execute(self, step, image):
            assert isinstance(image, np.ndarray) and self.supports(step.name)
            self.calls.append(step.name)
            return np.ascontiguousarray(image) + 1

    chain_names = ["Grayscale", "NoiseReduction", "Otsu", "Opening", "ConnectedComponents"]
    if all(n in steps for n in chain_names) and all(steps[n].execution.requires_gpu for n in chain_names):
        pm2 = core.get_pipeline_manager()
        for s in pm2.steps:
            s.enabled = s.name in chain_names
        ex = RecordingExecutor()
        pm2.set_gpu_executor(ex)
        src = np.zeros((6, 7), np.uint8)
        res = pm2.apply(src)
        clone_ex = pm2.clone()._gpu_executor if hasattr(pm2.clone(), "_gpu_executor") else None
        out["executor"] = {"calls": ex.calls, "enabled_in_order": [s.name for s in pm2.steps if s.enabled],
                           "result_sum": int(np.asarray(res).sum()), "input_untouched": bool((src == 0).all()),
                           "clone_keeps_executor": clone_ex is ex}
    print("PROBE " + json.dumps(out))


if __name__ == "__main__":
    main()
