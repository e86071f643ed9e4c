"""Run the REFERENCE'S OWN test files against this repository's mirror of its interfaces.

    python tests/ref_conformance_runner.py <reference checkout> <test file> [<test file> ...]

The reference's boundary tests (the driver-semantics suite) import
``processing.pipeline_manager``, ``processing.tiled_records``, ``core.tiled_image`` and
``plugins.module_base``.  This runner puts the reference checkout on ``sys.path`` (so everything else --
``processing.pipeline_cache``, ``core.*`` -- is the reference's own code), then binds those four module names
to ``yamimageprocessor_b200.host.pipeline`` / ``host.tiles`` / ``host.plugin`` before pytest imports the test
files.  The unmodified tests then exercise the mirror classes, including through the reference's own
``PipelineCache``.  Test infrastructure (build container only: the reference does not travel to the GPU box)."""
from __future__ import annotations

import importlib
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def main() -> int:
    ref = Path(sys.argv[1]).resolve()
    files = sys.argv[2:]
    sys.path.insert(0, str(ROOT))
    from oracle.ref_stubs import install_stubs

    install_stubs()                                  # PyQt5 / skimage stand-ins
    sys.path.insert(0, str(ref))
    import yamimageprocessor_b200.host.plugin as our_plugin

    # host/plugin.py subclasses the reference's ModuleBase when it is importable (so AppCore accepts the
    # modules); this run wants the stand-alone mirror classes themselves under test
    mirror = {
        "processing.pipeline_manager": "yamimageprocessor_b200.host.pipeline",
        "processing.tiled_records": "yamimageprocessor_b200.host.tiles",
        "core.tiled_image": "yamimageprocessor_b200.host.tiles",
        "plugins.module_base": "yamimageprocessor_b200.host.plugin",
    }
    for pkg in ("processing", "core", "plugins"):
        importlib.import_module(pkg)                  # the reference's packages stay the parents
    for name, ours in mirror.items():
        mod = importlib.import_module(ours)
        sys.modules[name] = mod
        parent, _, leaf = name.rpartition(".")
        setattr(sys.modules[parent], leaf, mod)
    del our_plugin
    import json

    import pytest

    class Bindings:
        """report which module every interface class used by the collected test modules comes from"""

        names = ("PipelineManager", "PipelineStep", "StepExecutionMetadata", "TiledPipelineImage", "TiledImageRecord",
                 "ModuleBase", "ModuleMetadata", "ModuleStage", "PipelineCache")

        def pytest_collection_finish(self, session):
            seen = {}
            for item in session.items:
                mod = item.module
                row = seen.setdefault(Path(mod.__file__).name, {})
                for name in self.names:
                    obj = getattr(mod, name, None)
                    if obj is not None:
                        row[name] = getattr(obj, "__module__", "?")
            print("BINDINGS " + json.dumps(seen, sort_keys=True))

    return int(pytest.main(["-q", "-p", "no:cacheprovider", "--rootdir", str(ref), "-o", "addopts=", *files], plugins=[Bindings()]))


if __name__ == "__main__":
    raise SystemExit(main())
