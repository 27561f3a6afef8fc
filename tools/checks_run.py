#!/usr/bin/env python3
"""Runs the small all-paths exercise (tools/sanitize_small.py's shapes) and a 4K batch against a library built with
-DFDF_CHECKS (tools/build_variant.sh checks -DFDF_CHECKS; FDF_LIB=build/variants/libfdf_checks.so): every shared-memory
index of the detection kernel's phases is range-checked on the device and the first failing source line is reported.
This stands in for compute-sanitizer, which is closed on the development pool."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_detector_fast_b200 as fdf  # noqa: E402

det = fdf.Detector(0)
fn = det._lib.fdf_debug_check_failure
fn.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
total, line = 0, C.c_int(0)
for (w, h, kind) in [(700, 150, 0), (333, 77, 0), (520, 80, 1), (64, 40, 1), (3840, 300, 0), (4100, 70, 1)]:
    frames = det.synth_frames(2, (w + 15) // 16 * 16, h, seed=3, kind=kind)[:, :, :w]
    for nms in (0, 1, 2):
        for n in (9, 12, 16):
            for sr in (32, 48, 64):
                det.set_tuning(strip_rows=sr)
                t = 3 if kind == 1 else 16
                pts, offs = det.detect_device(frames, fdf.Config(t, n, fdf.NonMaximalSuppression(nms)))
                torch.cuda.synchronize()
                total += int(offs[-1])
                assert det.device_flags() == 0
det.set_tuning()
frames = det.synth_frames(16, 3840, 2160, seed=20240, kind=0)
pts = torch.empty((16 * 100000, 2), dtype=torch.int32, device="cuda")
offs = torch.empty(17, dtype=torch.int64, device="cuda")
det.detect_device(frames, fdf.Config(20, 9, fdf.NonMaximalSuppression.MaxThreshold), points=pts, offsets=offs)
torch.cuda.synchronize()
total += int(offs[-1])
assert fn(det._ctx, C.byref(line)) == 0
print("checks_run: keypoints", total, "first failing check at fdf_strip.cuh line", line.value, "(0 = none)")
sys.exit(1 if line.value else 0)
