#!/usr/bin/env python3
"""Per-phase cycle counts of the detection kernel (library built with -DFDF_PHASE_CLOCKS, see README in profiles/).
Build here:  make -C feature_detector_fast_b200/csrc clean all EXTRA=-DFDF_PHASE_CLOCKS ; run on the GPU box; rebuild."""
import argparse, ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_detector_fast_b200 as fdf
from feature_detector_fast_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=128)
ap.add_argument("--nms", type=int, default=1)
a = ap.parse_args()
det = fdf.Detector(0)
frames = det.synth_frames(a.frames, 3840, 2160, seed=20240, kind=0)
cfg = fdf.Config(20, 9, fdf.NonMaximalSuppression(a.nms))
pts = torch.empty((a.frames * 100000, 2), dtype=torch.int32, device="cuda")
offs = torch.empty(a.frames + 1, dtype=torch.int64, device="cuda")
lib = det._lib
fn = lib.fdf_debug_phase_clocks
fn.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
buf = (C.c_uint64 * 256)()
for _ in range(2):
    det.detect_device(frames, cfg, points=pts, offsets=offs)
torch.cuda.synchronize()
fn(det._ctx, buf)
det.detect_device(frames, cfg, points=pts, offsets=offs)
torch.cuda.synchronize()
fn(det._ctx, buf)
names = {0: "wait tile", 1: "phase A", 2: "wait next strip", 4: "wait queue", 5: "phase B", 6: "barrier 1",
         3: "housekeeping (t0)", 7: "nms", 8: "barrier 2", 9: "stage"}
chunks = a.frames * 35 * 16
print("cycles per chunk:  " + " ".join(f"warp{w:d}" for w in range(10)))
for k, v in names.items():
    print(f"{v:18s} " + " ".join(f"{buf[w * 16 + k] / chunks:6.0f}" for w in range(10)))
print("tile latency when a filter warp had to wait: %.0f cycles (%.0f %% of the chunks)" % (
    sum(buf[w * 16 + 8] for w in range(4)) / max(1, sum(buf[w * 16 + 9] for w in range(4))),
    100.0 * sum(buf[w * 16 + 9] for w in range(4)) / 4 / chunks))
