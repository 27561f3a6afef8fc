"""feature_detector_fast_b200 -- B200 (sm_100a) FAST-n corner detector.

Host-side mirror of the reference crate's public API for the detection path
(/root/reference/src/lib.rs:15-64): ``Point``, ``NonMaximalSuppression``, ``Config``,
``Config.detect`` and ``detect``, with the same names, argument meaning and error behaviour
(the reference *panics* for ``count`` outside 9..=16 -- fast_simd.rs:302-305, :797-801 -- here
that is a ``FdfPanic`` exception).  All compute happens in ``lib/libfdf_cuda.so`` (hand-written
CUDA behind the C ABI of ``include/fdf.h``); there is no CPU fallback: importing works anywhere,
but detecting without the library or without a B200 raises.
"""
from .api import (  # noqa: F401
    Config,
    Detector,
    FdfError,
    FdfPanic,
    NonMaximalSuppression,
    Pipe,
    Point,
    default_detector,
    detect,
    detect_array,
)
from ._lib import LIB_PATH, build_library, library_exports, load_library  # noqa: F401

__all__ = [
    "Config",
    "Detector",
    "FdfError",
    "FdfPanic",
    "NonMaximalSuppression",
    "Pipe",
    "Point",
    "default_detector",
    "detect",
    "detect_array",
    "LIB_PATH",
    "build_library",
    "library_exports",
    "load_library",
]
