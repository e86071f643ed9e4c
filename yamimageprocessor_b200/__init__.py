"""yamimageprocessor_b200 — B200-native backend for YamImageProcessor's per-pixel hot path.

Importing the package never touches CUDA; ``backend.get_backend()`` creates the
device context and raises ``BackendUnavailable`` when libyamb200.so or a GPU is
missing (there is no CPU fallback).
"""
from ._lib import BackendUnavailable, YamError  # noqa: F401

__version__ = "0.1.0"
