#!/usr/bin/env python3
"""Records the per-warp, per-chunk phase timeline of the CTAs on SM 0 (library built with -DFDF_TRACE:
tools/build_variant.sh trace -DFDF_TRACE) and saves it as gpurun_out/trace_<tag>.npy
([cta 4][warp 16][chunk 200][slot 12] clock64 values; analysed offline by tools/trace_report.py)."""
import argparse, ctypes as C, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_detector_fast_b200 as fdf

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=128)
ap.add_argument("--nms", type=int, default=1)
ap.add_argument("--tag", default="t")
a = ap.parse_args()
det = fdf.Detector(0)
frames = det.synth_frames(a.frames, 3840, 2160, seed=20240, kind=0)
cfg = fdf.Config(20, 9, fdf.NonMaximalSuppression(a.nms))
pts = torch.empty((a.frames * 100000, 2), dtype=torch.int32, device="cuda")
offs = torch.empty(a.frames + 1, dtype=torch.int64, device="cuda")
fn = det._lib.fdf_debug_trace
fn.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
buf = np.zeros((4, 16, 200, 12), dtype=np.int64)
for _ in range(2):
    det.detect_device(frames, cfg, points=pts, offsets=offs)
    torch.cuda.synchronize()
    fn(det._ctx, buf.ctypes.data, buf.nbytes)
buf[:] = 0
det.detect_device(frames, cfg, points=pts, offsets=offs)
torch.cuda.synchronize()
assert fn(det._ctx, buf.ctypes.data, buf.nbytes) == 0
os.makedirs("gpurun_out", exist_ok=True)
np.save(f"gpurun_out/trace_{a.tag}.npy", buf)
print("trace saved", a.tag, "found", int(offs[-1]), "nonzero", int((buf != 0).sum()))
