"""PyQt5 / skimage / skfuzzy / sklearn stand-ins that let the UNMODIFIED reference import without its
GUI and optional dependencies (recipe of SURVEY.md App. C).

TEST / BENCHMARK INFRASTRUCTURE ONLY: used by tests/golden/make_golden*.py, tests/ref_appcore_probe.py
and the reference arm of bench.py; never imported by ``yamimageprocessor_b200``.  No reference source is
copied: the stubs only provide the NAMES the reference's import statements ask for."""
from __future__ import annotations

import sys
import types


def install_stubs() -> None:
    class _Sig:
        def __init__(self, *a, **k): pass
        def connect(self, *a, **k): pass
        def emit(self, *a, **k): pass

    class _Any:
        def __init__(self, *a, **k): pass
        def __getattr__(self, n): return _Any()
        def __call__(self, *a, **k): return _Any()

    class QObject:
        def __init__(self, *a, **k): pass

    class QRunnable:
        def __init__(self, *a, **k): pass
        def setAutoDelete(self, *a): pass

    class QThreadPool:
        @staticmethod
        def globalInstance(): return QThreadPool()
        def setMaxThreadCount(self, n): pass
        def start(self, r): r.run()
        def waitForDone(self): pass

    qtcore = types.ModuleType("PyQt5.QtCore")
    for name, val in dict(QObject=QObject, QRunnable=QRunnable, pyqtSignal=lambda *a, **k: _Sig(),
                          pyqtSlot=lambda *a, **k: (lambda f: f), QThreadPool=QThreadPool, QTranslator=_Any,
                          QLocale=_Any(), QCoreApplication=_Any()).items():
        setattr(qtcore, name, val)
    qtw = types.ModuleType("PyQt5.QtWidgets")
    for name in ("QApplication", "QWidget", "QMainWindow"):
        setattr(qtw, name, _Any)
    pyqt = types.ModuleType("PyQt5")
    pyqt.QtCore, pyqt.QtWidgets = qtcore, qtw
    sys.modules.update({"PyQt5": pyqt, "PyQt5.QtCore": qtcore, "PyQt5.QtWidgets": qtw})
    for name, attrs in {
        "skimage": (), "skimage.io": ("imread",), "skimage.feature": ("hog", "local_binary_pattern"),
        "skimage.measure": ("label", "regionprops"), "skimage.segmentation": ("active_contour",),
        "skimage.filters": ("gaussian",), "skfuzzy": ("cmeans",), "sklearn": (), "sklearn.mixture": ("GaussianMixture",),
    }.items():
        if name in sys.modules:
            continue
        m = types.ModuleType(name)
        for a in attrs:
            setattr(m, a, None)
        sys.modules[name] = m
    sys.modules["skimage"].io = sys.modules["skimage.io"]
