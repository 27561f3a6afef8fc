#!/usr/bin/env python3
"""Times fdf_detect_device on the benchmark workload without validating the result (for what-if builds)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_detector_fast_b200 as fdf  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=256)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--nms", type=int, default=1)
ap.add_argument("--w", type=int, default=3840)
ap.add_argument("--h", type=int, default=2160)
ap.add_argument("--t", type=int, default=20)
ap.add_argument("--kind", type=int, default=0)
a = ap.parse_args()
det = fdf.Detector(0)
frames = det.synth_frames(a.frames, a.w, a.h, seed=20240, kind=a.kind)
cfg = fdf.Config(a.t, 9, fdf.NonMaximalSuppression(a.nms))
pts = torch.empty((a.frames * (100000 if a.kind == 0 else a.w * a.h // 3), 2), dtype=torch.int32, device="cuda")
offs = torch.empty(a.frames + 1, dtype=torch.int64, device="cuda")
for _ in range(3):
    det.detect_device(frames, cfg, points=pts, offsets=offs)
torch.cuda.synchronize()
det.set_timing(a.steps)
for _ in range(a.steps):
    det.detect_device(frames, cfg, points=pts, offsets=offs)
torch.cuda.synchronize()
ms = [det.get_timing(i) for i in range(a.steps)]
k = sum(m[0] for m in ms) / a.steps
n_found = min(int(offs[-1]), pts.shape[0])
pl = pts[:n_found].long()
idx = torch.arange(n_found, device="cuda", dtype=torch.int64)
chk = int(((pl[:, 0] * 7919 + pl[:, 1] * 104729 + 1) * (idx % 65521 + 1)).sum().item()) & 0xFFFFFFFFFFFF  # order-sensitive
tag = os.environ.get("FDF_LIB", "default").split("/")[-1]
print(f"[{tag}] checksum {chk:012x} ", end="")
print(f"frames {a.frames} nms {a.nms}: detect {k:.4f} ms  scan {sum(m[1] for m in ms) / a.steps:.4f}  gather "
      f"{sum(m[2] for m in ms) / a.steps:.4f}  -> {a.frames * a.w * a.h / k / 1e6:.1f} Gpix/s  found {int(offs[-1])}"
      f"  flags {det.device_flags()}")
