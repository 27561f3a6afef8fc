import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    """The reference's shipped input image and the keypoint sets painted into its golden renders."""
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "shipped_300x200.npz")))


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle

    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def emulator():
    """tests/host/strip_emulator.cpp: the kernel's per-thread phase bodies executed on the CPU."""
    import __graft_entry__ as entry

    lib = C.CDLL(entry.build_host_emulator())
    lib.fdf_emulate_detect.restype = C.c_int64
    lib.fdf_emulate_detect.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint8, C.c_uint8,
                                       C.c_uint8, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.fdf_core_check.restype = C.c_int64
    lib.fdf_core_check.argtypes = [C.c_uint64, C.c_uint64]

    def run(img, t, n, nms, sr, want_fallbacks=False):
        """Runs the kernel's phase bodies on the CPU.  want_fallbacks: also return how many chunks took the
        queue-overflow (row group) path and the dense-NMS path."""
        img = np.ascontiguousarray(img)
        h, w = img.shape
        cap = max(1, w * h)
        out = np.zeros((cap, 2), np.uint32)
        fb = np.zeros(2, np.int32)
        k = lib.fdf_emulate_detect(img.ctypes.data, w, h, w, t, n, nms, sr, out.ctypes.data, cap, fb.ctypes.data)
        assert k >= 0, k
        return (out[:k].copy(), fb.tolist()) if want_fallbacks else out[:k].copy()

    run.core_check = lib.fdf_core_check
    return run


@pytest.fixture(scope="session")
def detector():
    """A live fdf context on cuda:0 (gpu tests only).  Fails loudly if the CUDA library is missing."""
    import feature_detector_fast_b200 as fdf

    det = fdf.Detector(0)
    yield det
    det.close()


def same_points(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return a.shape == b.shape and np.array_equal(a, b)
