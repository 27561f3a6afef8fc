"""Loads lib/libfdf_cuda.so (the C ABI of include/fdf.h) with ctypes.  Fails loudly when absent."""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FDF_LIB") or os.path.join(_PKG, "lib", "libfdf_cuda.so")  # FDF_LIB: what-if builds (tools/)
HEADER_PATH = os.path.join(os.path.dirname(_PKG), "include", "fdf.h")
_lib = None


class fdf_point(C.Structure):  # include/fdf.h: fdf_point  (lib.rs:15-20)
    _fields_ = [("x", C.c_uint32), ("y", C.c_uint32)]


def build_library(force: bool = False, verbose: bool = False) -> str:
    """nvcc-compiles the library for sm_100a (cross-compiles without a GPU).  Returns its path."""
    csrc = os.path.join(_PKG, "csrc")
    cmd = ["make", "-C", csrc] + (["-B"] if force else [])
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("building libfdf_cuda.so failed (see output above)")
    return LIB_PATH


def header_symbols() -> list:
    """Every function name include/fdf.h declares."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fdf_[a-z_0-9]+)\s*\(", text)))


def load_library() -> C.CDLL:
    """dlopen the CUDA library and declare its prototypes.  No fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C feature_detector_fast_b200/csrc`.  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    u8, u32, u64, sz, vp = C.c_uint8, C.c_uint32, C.c_uint64, C.c_size_t, C.c_void_p
    lib.fdf_create.restype = C.c_int
    lib.fdf_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.fdf_destroy.restype = None
    lib.fdf_destroy.argtypes = [vp]
    lib.fdf_detect.restype = C.c_int
    lib.fdf_detect.argtypes = [vp, vp, u32, u32, u32, u8, u8, u8, vp, sz, C.POINTER(sz)]
    lib.fdf_detect_batch.restype = C.c_int
    lib.fdf_detect_batch.argtypes = [vp, vp, u32, u32, u32, u32, u64, u8, u8, u8, vp, sz, vp]
    lib.fdf_detect_device.restype = C.c_int
    lib.fdf_detect_device.argtypes = [vp, vp, u32, u32, u32, u32, u64, u8, u8, u8, vp, sz, vp, vp]
    if hasattr(lib, "fdf_shard_push"):  # (what-if builds of older sources loaded through FDF_LIB lack these)
        lib.fdf_shard_push.restype = C.c_int
        lib.fdf_shard_push.argtypes = [vp, vp, u32, u32, u32, u32, vp, vp, sz, vp, vp]
        lib.fdf_shared_alloc.restype = C.c_int
        lib.fdf_shared_alloc.argtypes = [vp, sz, C.POINTER(vp), vp]
        lib.fdf_shared_open.restype = C.c_int
        lib.fdf_shared_open.argtypes = [vp, vp, C.POINTER(vp)]
        lib.fdf_shared_close.restype = C.c_int
        lib.fdf_shared_close.argtypes = [vp, vp]
    if hasattr(lib, "fdf_pipe_create"):
        lib.fdf_pipe_create.restype = C.c_int
        lib.fdf_pipe_create.argtypes = [vp, u32, u32, u32, sz, C.POINTER(vp)]
        lib.fdf_pipe_destroy.restype = None
        lib.fdf_pipe_destroy.argtypes = [vp]
        lib.fdf_pipe_submit.restype = C.c_int
        lib.fdf_pipe_submit.argtypes = [vp, vp, u32, u32, u32, u8, u8, u8]
        lib.fdf_pipe_collect.restype = C.c_int
        lib.fdf_pipe_collect.argtypes = [vp, vp, sz, C.POINTER(sz)]
        lib.fdf_pipe_in_flight.restype = u32
        lib.fdf_pipe_in_flight.argtypes = [vp]
    lib.fdf_rgb8_to_luma8_device.restype = C.c_int
    lib.fdf_rgb8_to_luma8_device.argtypes = [vp, vp, u32, u32, u32, u32, u64, vp, u32, u64, vp]
    lib.fdf_rgb8_to_grey_sum3_device.restype = C.c_int
    lib.fdf_rgb8_to_grey_sum3_device.argtypes = [vp, vp, u32, u32, u32, u32, u64, vp, u32, u64, vp]
    lib.fdf_detect_rgb8.restype = C.c_int
    lib.fdf_detect_rgb8.argtypes = [vp, vp, u32, u32, u32, u8, u8, u8, vp, sz, C.POINTER(sz)]
    lib.fdf_synth_frames_device.restype = C.c_int
    lib.fdf_synth_frames_device.argtypes = [vp, vp, u32, u32, u32, u32, u64, u64, u32, u32, u32, vp]
    lib.fdf_kernel_launches.restype = u64
    lib.fdf_kernel_launches.argtypes = [vp]
    if hasattr(lib, "fdf_set_tuning"):
        lib.fdf_set_tuning.restype = C.c_int
        lib.fdf_set_tuning.argtypes = [vp, C.c_int, u32]
    if hasattr(lib, "fdf_set_item_parts"):
        lib.fdf_set_item_parts.restype = C.c_int
        lib.fdf_set_item_parts.argtypes = [vp, u32]
    if hasattr(lib, "fdf_set_idle_sms"):
        lib.fdf_set_idle_sms.restype = C.c_int
        lib.fdf_set_idle_sms.argtypes = [vp, u32]
    lib.fdf_set_timing.restype = C.c_int
    lib.fdf_set_timing.argtypes = [vp, u32]
    lib.fdf_get_timing.restype = C.c_int
    lib.fdf_get_timing.argtypes = [vp, u32, C.POINTER(C.c_float)]
    lib.fdf_check_device_flags.restype = C.c_int
    lib.fdf_check_device_flags.argtypes = [vp, C.POINTER(u32)]
    lib.fdf_last_error.restype = C.c_char_p
    lib.fdf_last_error.argtypes = [vp]
    lib.fdf_status_string.restype = C.c_char_p
    lib.fdf_status_string.argtypes = [C.c_int]
    lib.fdf_version.restype = C.c_char_p
    lib.fdf_version.argtypes = []
    _lib = lib
    return lib


def library_exports() -> dict:
    """{symbol: exported?} for every function include/fdf.h declares (no compute is called)."""
    lib = load_library()
    return {name: hasattr(lib, name) for name in header_symbols()}
