"""The C-ABI shared library loads and exports everything include/fdf.h declares.  No compute calls."""
import ctypes as C
import os

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def fdf():
    import feature_detector_fast_b200 as fdf

    if not os.path.exists(fdf.LIB_PATH):
        fdf.build_library()
    return fdf


def test_library_exports_every_declared_symbol(fdf):
    exports = fdf.library_exports()
    assert set(exports) >= {"fdf_create", "fdf_destroy", "fdf_detect", "fdf_detect_batch", "fdf_detect_device",
                            "fdf_synth_frames_device", "fdf_kernel_launches", "fdf_check_device_flags",
                            "fdf_last_error", "fdf_status_string", "fdf_version"}
    assert all(exports.values()), [k for k, v in exports.items() if not v]


def test_status_strings_and_version(fdf):
    lib = fdf.load_library()
    assert b"sm_100a" in lib.fdf_version()
    assert lib.fdf_status_string(0) == b"ok"
    assert b"9..=16" in lib.fdf_status_string(1)


def test_library_is_sm100a_only_and_uses_tma():
    """cuobjdump: the only device code is sm_100a SASS, and the detection kernel contains UTMALDG (TMA)."""
    import shutil
    import subprocess

    import feature_detector_fast_b200 as fdf

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    elf = subprocess.run([cuobjdump, "-lelf", fdf.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in elf and "sm_90" not in elf and "sm_80" not in elf
    sass = subprocess.run([cuobjdump, "-sass", fdf.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTMALDG" in sass and "VABSDIFF4" in sass and "SYNCS" in sass


def test_no_context_without_gpu_is_an_error_not_a_fallback(fdf):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(fdf.FdfError):
        fdf.Detector(0)
    import numpy as np

    with pytest.raises(fdf.FdfError):
        fdf.detect(np.zeros((32, 32), np.uint8), fdf.Config(16, 9, fdf.NonMaximalSuppression.Off))


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing in the package or the C sources may reference it."""
    pkg = os.path.join(ROOT, "feature_detector_fast_b200")
    for dirpath, _, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, name)).read()
                assert "import oracle" not in text and "from oracle" not in text, name
                assert "fdf_oracle_detect" not in text and "fdf_avx2_port" not in text, name


def test_api_types_mirror_the_crate(fdf):
    # lib.rs:15-52
    assert [m.name for m in fdf.NonMaximalSuppression] == ["Off", "MaxThreshold", "SumAbsolute"]
    assert [int(m) for m in fdf.NonMaximalSuppression] == [0, 1, 2]  # fast_simd.rs:74-76
    p = fdf.Point(3, 4)
    assert (p.x, p.y) == (3, 4) and fdf.Point() == fdf.Point(0, 0) and hash(p) == hash(fdf.Point(3, 4))
    cfg = fdf.Config(threshold=16, count=9, non_maximal_supression=fdf.NonMaximalSuppression.MaxThreshold)
    assert cfg.threshold == 16 and cfg.count == 9 and callable(cfg.detect)
