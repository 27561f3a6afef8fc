"""Host-side mirror of the reference crate API (lib.rs:15-64) over the C ABI (include/fdf.h)."""
from __future__ import annotations

import ctypes as C
import enum
import threading
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from . import _lib


class FdfError(RuntimeError):
    """A C-ABI call returned a non-zero fdf_status."""

    def __init__(self, status: int, message: str):
        super().__init__(f"fdf status {status}: {message}")
        self.status = status


class FdfPanic(FdfError):
    """Where the reference panics: count < 9 (assert, fast_simd.rs:302-305) or count > 16 (:797-801)."""


class NonMaximalSuppression(enum.IntEnum):
    """lib.rs:25-36; numbered like fast_simd.rs:74-76."""

    Off = 0
    MaxThreshold = 1
    SumAbsolute = 2


@dataclass(frozen=True, order=True)
class Point:
    """lib.rs:15-20: a feature point at an image position."""

    x: int = 0
    y: int = 0


@dataclass(frozen=True, order=True)
class Config:
    """lib.rs:38-52.  (`non_maximal_supression` is spelled as in the reference.)"""

    threshold: int
    count: int
    non_maximal_supression: NonMaximalSuppression

    def detect(self, img) -> List[Point]:
        """lib.rs:54-59: method access to run the detector."""
        return detect(img, self)


def _raise(lib, ctx, status: int):
    msg = lib.fdf_last_error(ctx).decode() if ctx else ""
    if not msg:
        msg = lib.fdf_status_string(status).decode()
    if status == 1:
        raise FdfPanic(status, msg)
    raise FdfError(status, msg)


def _as_gray(img) -> np.ndarray:
    a = np.asarray(img)
    if a.dtype != np.uint8 or a.ndim != 2:
        raise TypeError("expected a grayscale image: 2-D uint8 array (rows x columns), like image::GrayImage")
    if a.strides[1] != 1 or a.strides[0] < a.shape[1]:
        a = np.ascontiguousarray(a)
    return a


class Detector:
    """One fdf_ctx: a device, a stream, the scan workspace and the staging buffers.

    Not thread-safe (one per thread), exactly like the C context it wraps.
    """

    def __init__(self, device: int = 0):
        self._lib = _lib.load_library()
        self._ctx = C.c_void_p()
        st = self._lib.fdf_create(device, C.byref(self._ctx))
        if st != 0:
            raise FdfError(st, self._lib.fdf_status_string(st).decode() +
                           " (fdf_create: needs a visible sm_100 GPU; there is no CPU fallback)")
        self.device = device
        self._points = np.zeros((0, 2), np.uint32)

    def close(self) -> None:
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.fdf_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- host-memory entry points -------------------------------------------------------------
    def _scratch(self, n: int) -> np.ndarray:
        if len(self._points) < n:
            self._points = np.zeros((n, 2), np.uint32)
        return self._points

    def detect_array(self, img, config: Config, cap: Optional[int] = None) -> np.ndarray:
        """fdf_detect: ordered (K, 2) uint32 array of (x, y)."""
        a = _as_gray(img)
        h, w = a.shape
        worst = max(0, w - 6) * max(0, h - 6)
        # Without a caller's capacity: room for 1 keypoint in 16 pixels (real images have < 2 %), grown to the exact need
        # and retried when the library reports FDF_ERR_CAPACITY (it returns the number found) -- the result is the same,
        # the buffers are 16 times smaller than "every pixel a keypoint".
        grow = cap is None
        cap = min(worst, max(4096, worst // 16)) if cap is None else int(cap)
        while True:
            buf = self._scratch(max(cap, 1))
            n = C.c_size_t(0)
            st = self._lib.fdf_detect(self._ctx, a.ctypes.data, w, h, a.strides[0], int(config.threshold),
                                      int(config.count), int(config.non_maximal_supression), buf.ctypes.data, cap,
                                      C.byref(n))
            if st == 4 and grow and n.value > cap:  # FDF_ERR_CAPACITY
                cap = n.value
                continue
            if st != 0:
                _raise(self._lib, self._ctx, st)
            return buf[: n.value].copy()

    def detect_rgb8_array(self, rgb, config: Config, cap: Optional[int] = None) -> np.ndarray:
        """fdf_detect_rgb8 (main.rs:53-67: to_rgb8 -> to_luma8 -> detect) on an (H, W, 3) uint8 array."""
        a = np.ascontiguousarray(np.asarray(rgb))
        if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
            raise TypeError("rgb must be an (H, W, 3) uint8 array")
        h, w, _ = a.shape
        worst = max(0, w - 6) * max(0, h - 6)
        cap = worst if cap is None else int(cap)
        buf = self._scratch(max(cap, 1))
        n = C.c_size_t(0)
        st = self._lib.fdf_detect_rgb8(self._ctx, a.ctypes.data, w, h, a.strides[0], int(config.threshold),
                                       int(config.count), int(config.non_maximal_supression), buf.ctypes.data, cap,
                                       C.byref(n))
        if st != 0:
            _raise(self._lib, self._ctx, st)
        return buf[: n.value].copy()

    def detect_batch(self, frames: np.ndarray, config: Config, cap: Optional[int] = None,
                     out: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
        """fdf_detect_batch over a C-contiguous (F, H, W) uint8 array: (points (K, 2), offsets (F+1,))."""
        if frames.dtype != np.uint8 or frames.ndim != 3 or not frames.flags.c_contiguous:
            raise TypeError("frames must be a C-contiguous (F, H, W) uint8 array")
        f, h, w = frames.shape
        worst = f * max(0, w - 6) * max(0, h - 6)
        # as in detect_array: without a caller's capacity start at 1 keypoint per 16 pixels and grow on FDF_ERR_CAPACITY
        grow = cap is None and out is None
        if cap is None:
            cap = min(worst, max(4096, worst // 16)) if out is None else len(out)
        offsets = np.zeros(f + 1, np.uint64)
        while True:
            buf = out if out is not None else self._scratch(max(cap, 1))
            st = self._lib.fdf_detect_batch(self._ctx, frames.ctypes.data, f, w, h, w, w * h, int(config.threshold),
                                            int(config.count), int(config.non_maximal_supression), buf.ctypes.data,
                                            int(cap), offsets.ctypes.data)
            if st == 4 and grow and int(offsets[f]) > cap:  # FDF_ERR_CAPACITY: offsets[F] is the number found
                cap = int(offsets[f])
                continue
            if st != 0:
                _raise(self._lib, self._ctx, st)
            k = int(offsets[f])
            return (buf[:k] if out is not None else buf[:k].copy()), offsets

    def detect_batch_pinned(self, frames_ptr: int, n_frames: int, w: int, h: int, config: Config, out_ptr: int,
                            cap: int, offsets_ptr: int) -> None:
        """fdf_detect_batch on raw HOST pointers (bench.py: pinned torch tensors); contiguous frames."""
        st = self._lib.fdf_detect_batch(self._ctx, frames_ptr, n_frames, w, h, w, w * h, int(config.threshold),
                                        int(config.count), int(config.non_maximal_supression), out_ptr, int(cap),
                                        offsets_ptr)
        if st != 0:
            _raise(self._lib, self._ctx, st)

    # ---- device-resident entry points (torch tensors carry the device memory) -------------------
    def detect_device(self, frames, config: Config, points=None, offsets=None, stream=None):
        """fdf_detect_device on a CUDA uint8 tensor (F, H, W) (row stride % 16 == 0, frame stride % 16 == 0).

        Enqueues on torch's current stream (or `stream`) and returns (points int32 (cap, 2), offsets int64 (F+1,))
        device tensors without synchronising; the points of frame f are points[offsets[f]:offsets[f+1]].
        """
        import torch

        if frames.dtype != torch.uint8 or frames.dim() != 3 or not frames.is_cuda or frames.stride(2) != 1:
            raise TypeError("frames must be a CUDA uint8 tensor of shape (F, H, W) with unit column stride")
        f, h, w = frames.shape
        if points is None:
            points = torch.empty((max(1, f * max(0, w - 6) * max(0, h - 6)), 2), dtype=torch.int32,
                                 device=frames.device)
        if offsets is None:
            offsets = torch.empty(f + 1, dtype=torch.int64, device=frames.device)
        s = stream if stream is not None else torch.cuda.current_stream(frames.device)
        st = self._lib.fdf_detect_device(self._ctx, frames.data_ptr(), f, w, h, frames.stride(1),
                                         frames.stride(0), int(config.threshold), int(config.count),
                                         int(config.non_maximal_supression), points.data_ptr(), points.shape[0],
                                         offsets.data_ptr(), s.cuda_stream)
        if st != 0:
            _raise(self._lib, self._ctx, st)
        return points, offsets

    def rgb8_to_luma8_device(self, rgb, out=None, stream=None, sum3: bool = False):
        """fdf_rgb8_to_luma8_device: CUDA uint8 (F, H, W, 3) -> (F, H, Wp) luma view of width W, row stride a multiple
        of 16 bytes (ready for detect_device).  sum3=True: fdf_rgb8_to_grey_sum3_device, util.rs's (r + g + b) / 3."""
        import torch

        if rgb.dtype != torch.uint8 or rgb.dim() != 4 or rgb.shape[3] != 3 or not rgb.is_cuda or not rgb.is_contiguous():
            raise TypeError("rgb must be a contiguous CUDA uint8 tensor of shape (F, H, W, 3)")
        f, h, w, _ = rgb.shape
        if out is None:
            out = torch.empty((f, h, (w + 15) // 16 * 16), dtype=torch.uint8, device=rgb.device)[:, :, :w]
        s = stream if stream is not None else torch.cuda.current_stream(rgb.device)
        fn = self._lib.fdf_rgb8_to_grey_sum3_device if sum3 else self._lib.fdf_rgb8_to_luma8_device
        st = fn(self._ctx, rgb.data_ptr(), f, w, h, rgb.stride(1), rgb.stride(0), out.data_ptr(), out.stride(1),
                out.stride(0), s.cuda_stream)
        if st != 0:
            _raise(self._lib, self._ctx, st)
        return out

    def synth_frames(self, n_frames: int, w: int, h: int, seed: int, first_frame: int = 0, kind: int = 0,
                     amp: int = 4, out=None, device=None):
        """fdf_synth_frames_device: (F, H, W) uint8 CUDA tensor of synthetic frames (pitch == w)."""
        import torch

        dev = torch.device("cuda", self.device) if device is None else device
        if out is None:
            out = torch.empty((n_frames, h, w), dtype=torch.uint8, device=dev)
        s = torch.cuda.current_stream(dev)
        st = self._lib.fdf_synth_frames_device(self._ctx, out.data_ptr(), n_frames, w, h, out.stride(1),
                                               out.stride(0), seed, first_frame, kind, amp, s.cuda_stream)
        if st != 0:
            _raise(self._lib, self._ctx, st)
        return out

    def set_tuning(self, strip_rows: int = 0, sub_batch_mb: int = 0) -> None:
        """fdf_set_tuning: scored rows per strip (0 = automatic, 32, 48, 64) and the host path's sub-batch size in MB
        (0 = default).  Results never depend on either."""
        st = self._lib.fdf_set_tuning(self._ctx, int(strip_rows), int(sub_batch_mb))
        if st != 0:
            _raise(self._lib, self._ctx, st)

    def set_timing(self, slots: int) -> None:
        """Record CUDA events around the three launches of every following detect_device call (0 = off)."""
        st = self._lib.fdf_set_timing(self._ctx, int(slots))
        if st != 0:
            _raise(self._lib, self._ctx, st)

    def get_timing(self, slot: int):
        """(detection, scan, gather) kernel milliseconds of timing slot `slot`."""
        ms = (C.c_float * 3)()
        st = self._lib.fdf_get_timing(self._ctx, int(slot), ms)
        if st != 0:
            _raise(self._lib, self._ctx, st)
        return float(ms[0]), float(ms[1]), float(ms[2])

    def set_item_parts(self, parts: int = 0) -> None:
        """fdf_set_item_parts: work items per strip of the detection kernel (0 = automatic, 1, 2, 4, 8)."""
        st = self._lib.fdf_set_item_parts(self._ctx, int(parts))
        if st != 0:
            _raise(self._lib, self._ctx, st)

    def set_idle_sms(self, sm_stride: int = 0) -> None:
        """fdf_set_idle_sms: the detection kernel leaves every sm_stride-th SM to other kernels (0 = uses all)."""
        st = self._lib.fdf_set_idle_sms(self._ctx, int(sm_stride))
        if st != 0:
            _raise(self._lib, self._ctx, st)

    def pipe(self, depth: int, max_w: int, max_h: int, cap: Optional[int] = None) -> "Pipe":
        """fdf_pipe_create: a streaming detector with up to `depth` images of at most max_w x max_h in flight."""
        return Pipe(self, depth, max_w, max_h, cap)

    def device_flags(self) -> int:
        flags = C.c_uint32(0)
        st = self._lib.fdf_check_device_flags(self._ctx, C.byref(flags))
        if st != 0:
            _raise(self._lib, self._ctx, st)
        return flags.value

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.fdf_kernel_launches(self._ctx))


class Pipe:
    """fdf_pipe_*: the streaming form of `detect` -- up to `depth` images in flight, results first in, first out.

        pipe = detector.pipe(depth=4, max_w=1920, max_h=1080)
        for img in frames:
            if pipe.in_flight == pipe.depth:
                handle(pipe.collect())
            pipe.submit(img, cfg)
        while pipe.in_flight:
            handle(pipe.collect())

    A pageable image may be reused as soon as submit returns; an image in pinned memory is read by the DMA engine
    directly and must stay untouched until its collect."""

    def __init__(self, detector: "Detector", depth: int, max_w: int, max_h: int, cap: Optional[int] = None):
        self._det, self._lib = detector, detector._lib
        self.depth = int(depth)
        worst = max(0, max_w - 6) * max(0, max_h - 6)
        self.cap = int(cap) if cap is not None else min(worst, max(4096, worst // 16))
        self._pipe = C.c_void_p()
        st = self._lib.fdf_pipe_create(detector._ctx, self.depth, int(max_w), int(max_h), self.cap, C.byref(self._pipe))
        if st != 0:
            _raise(self._lib, detector._ctx, st)
        self._buf = np.zeros((max(self.cap, 1), 2), np.uint32)
        self._keep = []  # the submitted arrays, until collected (pinned images are read asynchronously)

    @property
    def in_flight(self) -> int:
        return int(self._lib.fdf_pipe_in_flight(self._pipe)) if self._pipe.value else 0

    def submit(self, img, config: Config) -> None:
        a = _as_gray(img)
        h, w = a.shape
        st = self._lib.fdf_pipe_submit(self._pipe, a.ctypes.data, w, h, a.strides[0], int(config.threshold),
                                       int(config.count), int(config.non_maximal_supression))
        if st != 0:
            _raise(self._lib, self._det._ctx, st)
        self._keep.append(a)

    def collect(self) -> np.ndarray:
        """Keypoints of the oldest image in flight, (K, 2) uint32 (x, y) in the reference's order."""
        n = C.c_size_t(0)
        st = self._lib.fdf_pipe_collect(self._pipe, self._buf.ctypes.data, self.cap, C.byref(n))
        if st != 3 and self._keep:  # (status 3 = nothing was in flight)
            self._keep.pop(0)
        if st != 0:
            _raise(self._lib, self._det._ctx, st)
        return self._buf[: n.value].copy()

    def close(self) -> None:
        if getattr(self, "_pipe", None) is not None and self._pipe.value:
            self._lib.fdf_pipe_destroy(self._pipe)
            self._pipe = C.c_void_p()
            self._keep = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_tls = threading.local()


def default_detector(device: int = 0) -> Detector:
    """A per-thread Detector (the reference function is re-entrant; the context is not shared)."""
    cache = getattr(_tls, "detectors", None)
    if cache is None:
        cache = _tls.detectors = {}
    if device not in cache:
        cache[device] = Detector(device)
    return cache[device]


def detect_array(img, config: Config, device: int = 0) -> np.ndarray:
    """`detect` returning an ordered (K, 2) uint32 array of (x, y) instead of a list of Points."""
    return default_detector(device).detect_array(img, config)


def detect(img, config: Config) -> List[Point]:
    """lib.rs:62-64: `pub fn detect(img: &image::GrayImage, config: &Config) -> Vec<Point>`."""
    pts = detect_array(img, config)
    return [Point(int(x), int(y)) for x, y in pts]
