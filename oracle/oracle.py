"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

NMS_OFF, NMS_MAX_THRESHOLD, NMS_SUM_ABSOLUTE = 0, 1, 2

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_libs: dict = {}


def build(force: bool = False) -> None:
    """Compile oracle/_build/*.so with the committed Makefile (gcc/g++ only, no CUDA)."""
    want = [os.path.join(_BUILD, n) for n in ("libfdf_oracle.so", "libfdf_avx2_port.so")]
    srcs = [os.path.join(_HERE, n) for n in ("fdf_oracle.c", "fdf_oracle.h", "fdf_avx2_port.cpp", "Makefile")]
    stale = force or any(
        not os.path.exists(w) or os.path.getmtime(w) < max(os.path.getmtime(s) for s in srcs) for w in want
    )
    if stale:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)


def _lib(name: str) -> C.CDLL:
    if name not in _libs:
        path = os.path.join(_BUILD, name)
        if not os.path.exists(path):
            build()
        lib = C.CDLL(path)
        _declare(lib, name)
        _libs[name] = lib
    return _libs[name]


_u8p = C.POINTER(C.c_uint8)


def _declare(lib: C.CDLL, name: str) -> None:
    if name == "libfdf_oracle.so":
        lib.fdf_oracle_detect.restype = C.c_int64
        lib.fdf_oracle_detect.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint8, C.c_uint8,
                                          C.c_uint8, C.c_void_p, C.c_size_t, C.c_void_p]
        lib.fdf_oracle_is_keypoint.restype = C.c_int
        lib.fdf_oracle_is_keypoint.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint8, C.c_uint8]
        lib.fdf_oracle_consecutive.restype = C.c_int
        lib.fdf_oracle_consecutive.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.fdf_oracle_score_max_threshold_px.restype = C.c_uint16
        lib.fdf_oracle_score_max_threshold_px.argtypes = [C.c_uint8, C.c_void_p, C.c_uint8]
        lib.fdf_oracle_score_sum_abs_px.restype = C.c_uint16
        lib.fdf_oracle_score_sum_abs_px.argtypes = [C.c_uint8, C.c_void_p, C.c_uint8]
        lib.fdf_oracle_hash_points.restype = C.c_uint64
        lib.fdf_oracle_hash_points.argtypes = [C.c_void_p, C.c_size_t]
        lib.fdf_oracle_siphash13.restype = C.c_uint64
        lib.fdf_oracle_siphash13.argtypes = [C.c_void_p, C.c_size_t]
        lib.fdf_oracle_synth_frame.restype = None
        lib.fdf_oracle_synth_frame.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64,
                                               C.c_uint32, C.c_uint32, C.c_uint32]
        lib.fdf_oracle_circle.restype = None
        lib.fdf_oracle_circle.argtypes = [C.c_void_p]
    else:
        lib.fdf_avx2_port_detect.restype = C.c_int64
        lib.fdf_avx2_port_detect.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint8, C.c_uint8,
                                             C.c_uint8, C.c_void_p, C.c_size_t]
        lib.fdf_avx2_port_detect_batch.restype = C.c_int
        lib.fdf_avx2_port_detect_batch.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                                   C.c_uint64, C.c_uint8, C.c_uint8, C.c_uint8, C.c_void_p,
                                                   C.c_void_p, C.c_uint32]
        lib.fdf_avx2_port_score_max_threshold.restype = C.c_uint16
        lib.fdf_avx2_port_score_max_threshold.argtypes = [C.c_uint8, C.c_void_p, C.c_uint8]
        lib.fdf_avx2_port_score_sum_abs.restype = C.c_uint16
        lib.fdf_avx2_port_score_sum_abs.argtypes = [C.c_uint8, C.c_void_p, C.c_uint8]
        lib.fdf_kat_random_max_threshold.restype = C.c_int64
        lib.fdf_kat_random_max_threshold.argtypes = [C.c_uint64, C.c_uint8]
        lib.fdf_kat_random_sum_abs.restype = C.c_int64
        lib.fdf_kat_random_sum_abs.argtypes = [C.c_uint64]


def _as_image(img: np.ndarray) -> np.ndarray:
    a = np.asarray(img)
    if a.dtype != np.uint8 or a.ndim != 2:
        raise TypeError("image must be a 2-D uint8 array (rows x columns)")
    return np.ascontiguousarray(a)


def circle() -> np.ndarray:
    """(16, 2) int32 array of (dx, dy), index 0 = north, clockwise (opencv_compat.rs:42-61)."""
    out = np.zeros(32, np.int32)
    _lib("libfdf_oracle.so").fdf_oracle_circle(out.ctypes.data)
    return out.reshape(16, 2)


def detect(img: np.ndarray, threshold: int, count: int, nms: int, return_scores: bool = False):
    """Scalar oracle: ordered (K, 2) uint32 array of (x, y).  Raises ValueError where the reference panics."""
    a = _as_image(img)
    h, w = a.shape
    lib = _lib("libfdf_oracle.so")
    cap = max(16, (w * h) // 8)
    while True:
        pts = np.zeros((cap, 2), np.uint32)
        scores = np.zeros(cap, np.uint16)
        n = lib.fdf_oracle_detect(a.ctypes.data, w, h, w, threshold, count, nms, pts.ctypes.data, cap,
                                  scores.ctypes.data)
        if n < 0:
            raise ValueError(f"oracle rejected the configuration (code {n}): count must be 9..=16, nms 0..2")
        if n <= cap:
            break
        cap = int(n)
    if return_scores:
        return pts[:n].copy(), scores[:n].copy()
    return pts[:n].copy()


def port_detect(img: np.ndarray, threshold: int, count: int, nms: int) -> np.ndarray:
    """AVX2 port of fast_simd.rs (the timed CPU baseline).  Same output contract as detect()."""
    a = _as_image(img)
    h, w = a.shape
    padded = np.zeros(a.size + 64, np.uint8)  # the dword gathers over-read by <= 3 bytes (S17)
    padded[: a.size] = a.reshape(-1)
    lib = _lib("libfdf_avx2_port.so")
    cap = max(16, (w * h) // 8)
    while True:
        pts = np.zeros((cap, 2), np.uint32)
        n = lib.fdf_avx2_port_detect(padded.ctypes.data, w, h, w, threshold, count, nms, pts.ctypes.data, cap)
        if n < 0:
            raise ValueError(f"port rejected the configuration (code {n})")
        if n <= cap:
            break
        cap = int(n)
    return pts[:n].copy()


def port_detect_many(frames: np.ndarray, threshold: int, count: int, nms: int, n_threads: int = 1,
                     want_hashes: bool = True) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """AVX2 port over a (F, H, W) uint8 batch, one frame per worker thread; returns (counts, hashes).

    `frames` must own >= 4 bytes of slack after the last frame; pass an array made by `padded_batch`.
    """
    if frames.dtype != np.uint8 or frames.ndim != 3 or not frames.flags.c_contiguous:
        raise TypeError("frames must be a C-contiguous (F, H, W) uint8 array")
    f, h, w = frames.shape
    # the gathers over-read by <= 3 bytes (S17): always hand the port a buffer with slack after the last frame
    end = frames.ctypes.data + frames.nbytes
    owner = frames.base
    has_slack = isinstance(owner, np.ndarray) and owner.flags.c_contiguous and \
        owner.ctypes.data + owner.nbytes >= end + 4
    if has_slack:
        buf = frames
        ptr = frames.ctypes.data
    else:
        buf = np.zeros(frames.size + 64, np.uint8)
        buf[: frames.size] = frames.reshape(-1)
        ptr = buf.ctypes.data
    counts = np.zeros(f, np.int64)
    hashes = np.zeros(f, np.uint64) if want_hashes else None
    rc = _lib("libfdf_avx2_port.so").fdf_avx2_port_detect_batch(
        ptr, f, w, h, w, w * h, threshold, count, nms, counts.ctypes.data,
        hashes.ctypes.data if want_hashes else None, n_threads)
    if rc != 0:
        raise ValueError(f"port rejected the configuration (code {rc})")
    del buf
    return counts, hashes


def is_keypoint(img: np.ndarray, x: int, y: int, threshold: int, count: int) -> bool:
    a = _as_image(img)
    return bool(_lib("libfdf_oracle.so").fdf_oracle_is_keypoint(a.ctypes.data, a.shape[1], x, y, threshold, count))


def consecutive(flags, n: int) -> bool:
    f = np.ascontiguousarray(np.asarray(flags, dtype=np.uint8))
    return bool(_lib("libfdf_oracle.so").fdf_oracle_consecutive(f.ctypes.data, len(f), n))


def _ring(ring) -> np.ndarray:
    r = np.ascontiguousarray(np.asarray(ring, dtype=np.uint8))
    if r.shape != (16,):
        raise ValueError("ring must hold 16 values")
    return r


def score_max_threshold_px(centre: int, ring, count: int) -> int:
    r = _ring(ring)
    return int(_lib("libfdf_oracle.so").fdf_oracle_score_max_threshold_px(centre, r.ctypes.data, count))


def score_sum_abs_px(centre: int, ring, threshold: int) -> int:
    r = _ring(ring)
    return int(_lib("libfdf_oracle.so").fdf_oracle_score_sum_abs_px(centre, r.ctypes.data, threshold))


def port_score_max_threshold_px(centre: int, ring, count: int) -> int:
    r = _ring(ring)
    return int(_lib("libfdf_avx2_port.so").fdf_avx2_port_score_max_threshold(centre, r.ctypes.data, count))


def port_score_sum_abs_px(centre: int, ring, threshold: int) -> int:
    r = _ring(ring)
    return int(_lib("libfdf_avx2_port.so").fdf_avx2_port_score_sum_abs(centre, r.ctypes.data, threshold))


def kat_random_max_threshold(n_seeds: int, count: int) -> int:
    """Mismatches port-vs-scalar over the reference's randomised MaxThreshold test (fast_simd.rs:939-945)."""
    return int(_lib("libfdf_avx2_port.so").fdf_kat_random_max_threshold(n_seeds, count))


def kat_random_sum_abs(iterations: int) -> int:
    """Mismatches port-vs-scalar over the reference's randomised SAD test (fast_simd.rs:1198-1236)."""
    return int(_lib("libfdf_avx2_port.so").fdf_kat_random_sum_abs(iterations))


def hash_points(points: np.ndarray) -> int:
    """SipHash-1-3 of a point list in tests/compare.rs:5-12 format."""
    p = np.ascontiguousarray(np.asarray(points, dtype=np.uint32).reshape(-1, 2))
    return int(_lib("libfdf_oracle.so").fdf_oracle_hash_points(p.ctypes.data, len(p)))


def siphash13(data: bytes) -> int:
    b = np.frombuffer(bytes(data), np.uint8)
    return int(_lib("libfdf_oracle.so").fdf_oracle_siphash13(b.ctypes.data if len(b) else None, len(b)))


def synth_frame(w: int, h: int, seed: int, frame: int, kind: int = 0, amp: int = 4) -> np.ndarray:
    """Counter-based synthetic frame (bit-identical to the CUDA generator fdf_synth_frames_device)."""
    out = np.zeros((h, w), np.uint8)
    _lib("libfdf_oracle.so").fdf_oracle_synth_frame(out.ctypes.data, w, h, w, seed, frame, kind, amp)
    return out
